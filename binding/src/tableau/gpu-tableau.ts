/**
 * @file src/tableau/gpu-tableau.ts  (drop into the reference tree next to tableau.ts)
 * @description `Tableau` with storage, the pivot loop and branch-and-cut on a B200 behind libjslp_b200.so.
 *
 * Overrides exactly the numerical members of the reference's Tableau seam (src/tableau/tableau.ts:103-258,
 * SURVEY.md 8b); every override is one call of the N-API addon (binding/jslp_addon.cc), which is one call of the
 * C ABI (include/jslp_b200.h).  Everything the front end and result shaping read afterwards -- `matrix` column 0,
 * `varIndexByRow/Col`, `rowByVarIndex/colByVarIndex`, `feasible`, `bounded`, `evaluation`, `bestPossibleEval`,
 * `simplexIters`, `unboundedVarIndex`, `optionalObjectives[].reducedCosts`, `model.messages`, and after
 * branchAndCut() `branchAndCutIterations`, `__isIntegral`, `height`, `nVars`, `lastElementIndex`,
 * `variablesPerIndex[newSlack]`, `model.solutions` -- is synchronised back (the "state contract after simplex()").
 * The authoritative matrix lives in HBM; `materializeMatrix()` fetches all of it for callers that need more than
 * the right-hand-side column (MIR cuts on the host, the dynamic-modification API).
 *
 * Not compiled in this repository (no Node.js / tsc in the build image); the same call sequence is exercised end
 * to end by the Python mirror jslpsolver_b200/tableau.py::GpuTableau through the identical C entry points.
 */
import Tableau from "./tableau";
import { SlackVariable } from "../expressions";
import type Model from "../model";
import type { Constraint, Variable } from "../expressions";
import type { BranchCut, VariableValue } from "./types";
import type { BranchAndCutService } from "./branch-and-cut";

/* eslint-disable @typescript-eslint/no-var-requires */
const addon: GpuAddon = require("../../binding/build/Release/jslp_b200.node");

interface LpStatus {
    feasible: boolean; bounded: boolean; cycled: number; cycleStart: number; cycleLength: number;
    phase1Pivots: number; phase2Pivots: number; unboundedVarIndex: number; simplexIters: number;
    width: number; height: number; evaluation: number; bestPossibleEval: number; gpuMs: number;
}
interface Downloaded {
    width: number; height: number; matrix?: Float64Array; rhs?: Float64Array; cost?: Float64Array;
    varIndexByRow?: Int32Array; varIndexByCol?: Int32Array; opt?: Float64Array;
}
interface BnbResult {
    feasible: boolean; bounded: boolean; isIntegral: boolean; timedOut: boolean; iterations: number;
    evaluation: number; bestPossibleEval: number; bestCuts: BranchCut[];
    solutions: Array<{ evaluation: number; varIndexByRow: Int32Array; rhs: Float64Array }>;
    rounds: number; nodeLps: number; pivots: number; gpuMs: number;
}
interface GpuTab {
    upload(matrix: Float64Array, varIndexByRow: Int32Array, varIndexByCol: Int32Array, unrestricted: Uint8Array | null,
           intVarIndices: Int32Array | null, optCosts: Float64Array | null, nOpt: number): void;
    setOption(key: number, value: number): void;
    simplex(checkCycles: boolean): LpStatus; phase1(checkCycles: boolean): LpStatus; phase2(checkCycles: boolean): LpStatus;
    pivot(row: number, col: number): void; save(): void; restore(): void;
    addCuts(cuts: BranchCut[]): void; applyCuts(cuts: BranchCut[], checkCycles: boolean): LpStatus;
    isIntegral(): boolean; mostFractional(): VariableValue;
    download(what: { matrix?: boolean; rhs?: boolean; cost?: boolean; maps?: boolean; opt?: number }): Downloaded;
    pivotLog(): Int32Array; createComm(id: Uint8Array, rank: number, nRanks: number): void;
    putInBase(varIndex: number): number; takeOutOfBase(varIndex: number): number;
    updateRhs(constraintIndex: number, difference: number): void;
    updateCoefficient(constraintIndex: number, varIndex: number, difference: number): void;
    updateCost(varIndex: number, optSlot: number, difference: number): void;
    addConstraint(isUpperBound: boolean, rhs: number, slackIndex: number, termVars: Int32Array, termCoefs: Float64Array): void;
    removeConstraint(slackIndex: number): void;
    addVariable(varIndex: number, costEntry: number, optSlot: number, isInteger: boolean, isUnrestricted: boolean): void;
    removeVariable(varIndex: number): void;
    info(): { width: number; height: number; nVars: number; lastElementIndex: number };
    branchAndCut(opts: Record<string, unknown>): BnbResult; destroy(): void;
}
interface GpuAddon {
    Tab: { new (width: number, height: number, rowCapacity: number, precision: number, device?: number): GpuTab; uniqueId(): Uint8Array };
    abiVersion(): number;
}

/** true when the addon loaded and reports the ABI this file was written against */
export function gpuAvailable(): boolean {
    try {
        return addon.abiVersion() === 2;
    } catch {
        return false;
    }
}

export default class GpuTableau extends Tableau {
    private dev: GpuTab | null = null;
    /** the host copy of `matrix` holds only column 0 (+ row 0 entry 0) until materializeMatrix() */
    private hostMatrixPartial = false;
    device = 0;

    constructor(precision = 1e-8, branchAndCutService?: BranchAndCutService) {
        super(precision, branchAndCutService);
    }

    // ---- tableau.ts:382-391: the front end builds the initial tableau on the host, it is uploaded once
    setModel(model: Model): this {
        super.setModel(model);
        this.dev?.destroy();
        this.dev = new addon.Tab(this.width, this.height, this.height + 64, this.precision, this.device);
        const nIndex = this.width + this.height - 2;
        const unres = new Uint8Array(nIndex);
        for (const k of Object.keys(this.unrestrictedVars)) if (+k < nIndex && this.unrestrictedVars[+k]) unres[+k] = 1;
        const ints = Int32Array.from(model.integerVariables.map((v) => v.index));
        const nOpt = this.optionalObjectives.length;   // already sorted by priority (tableau.ts:278-290)
        let opt: Float64Array | null = null;
        if (nOpt > 0) {
            opt = new Float64Array(nOpt * this.width);
            this.optionalObjectives.forEach((o, k) => opt!.set(o.reducedCosts.slice(0, this.width), k * this.width));
        }
        this.dev.upload(this.matrix, Int32Array.from(this.varIndexByRow), Int32Array.from(this.varIndexByCol),
                        unres, ints.length > 0 ? ints : null, opt, nOpt);
        this.hostMatrixPartial = false;
        return this;
    }

    private tab(): GpuTab {
        if (this.dev === null) throw new Error("GpuTableau: setModel() has not been called");
        return this.dev;
    }
    private checkCycles(): boolean { return this.model?.checkForCycles ?? true; }

    // ---- tableau.ts:103-123
    simplex(): this { this.absorb(this.tab().simplex(this.checkCycles())); return this; }
    phase1(): number { const s = this.tab().phase1(this.checkCycles()); this.absorb(s); return s.phase1Pivots; }
    phase2(): number { const s = this.tab().phase2(this.checkCycles()); this.absorb(s); return s.phase2Pivots; }
    pivot(pivotRowIndex: number, pivotColumnIndex: number): void {
        this.tab().pivot(pivotRowIndex, pivotColumnIndex);
        this.syncFromDevice();
    }

    // ---- backup.ts:49-105 (device-side snapshot; savedState stays null on the host)
    save(): void { this.tab().save(); }
    restore(): void { this.tab().restore(); this.syncFromDevice(); }

    // ---- cutting-strategies.ts:16-72 / branch-and-cut.ts:33-52
    addCutConstraints(branchingCuts: BranchCut[]): void {
        const first = this.lastElementIndex;
        this.tab().addCuts(branchingCuts);
        for (let h = 0; h < branchingCuts.length; h++) {          // cutting-strategies.ts:64-70
            const index = first + h;
            this.variablesPerIndex[index] = new SlackVariable("s" + index, index);
        }
        this.lastElementIndex = first + branchingCuts.length;
        this.syncFromDevice();
        this.nVars = this.width + this.height - 2 + branchingCuts.length;  // cutting-strategies.ts:34,70 (kept as is)
    }
    applyCuts(branchingCuts: BranchCut[]): void {
        this.tab().setOption(13 /* JSLP_OPT_USE_MIR_CUTS */, this.model?.useMIRCuts ? 1 : 0);  // MIR loop runs on the device
        const first = this.lastElementIndex;
        this.absorb(this.tab().applyCuts(branchingCuts, this.checkCycles()));
        for (let h = 0; h < branchingCuts.length; h++) this.variablesPerIndex[first + h] = new SlackVariable("s" + (first + h), first + h);
    }

    // ---- mip-utils.ts:43-61,100-126
    isIntegral(): boolean { return this.tab().isIntegral(); }
    getMostFractionalVar(): VariableValue { return this.tab().mostFractional(); }

    // ---- tableau.ts:244-246: the whole loop of branch-and-cut.ts:54-199 runs behind one call
    branchAndCut(): void { this.runBranchAndCut({}); }

    /** `strategy` = {nodeSelection?, branching?, strongBranchingCandidates?} runs the enhanced service's loop
     *  (enhanced-branch-and-cut.ts:223-434) instead; see createGpuEnhancedBranchAndCutService. */
    runBranchAndCut(strategy: Record<string, unknown>): void {
        const model = this.model!;
        if (model.useMIRCuts) this.tab().setOption(13 /* JSLP_OPT_USE_MIR_CUTS */, 1);
        const r = this.tab().branchAndCut({
            tolerance: model.tolerance ?? 0, isMinimization: model.isMinimization, checkCycles: model.checkForCycles,
            keepSolutions: model.keep_solutions === true, timeout: model.timeout ?? 0, ...strategy,
        });
        this.feasible = r.feasible; this.bounded = r.bounded; this.evaluation = r.evaluation;
        this.bestPossibleEval = r.bestPossibleEval; this.branchAndCutIterations = r.iterations;
        if (r.isIntegral) this.__isIntegral = true;
        const first = this.lastElementIndex;             // the winner's cut rows stay appended (branch-and-cut.ts:195-197)
        for (let h = 0; h < r.bestCuts.length; h++) this.variablesPerIndex[first + h] = new SlackVariable("s" + (first + h), first + h);
        this.syncFromDevice();
        if (model.keep_solutions) {                     // branch-and-cut.ts:143-153
            const rounding = Math.round(1 / this.precision);
            for (const s of r.solutions) {
                const store: Record<string, number> = {};
                for (let row = 1; row < s.rhs.length; row++) {
                    const variable = this.variablesPerIndex[s.varIndexByRow[row]];
                    if (variable === undefined || variable.isSlack === true) continue;
                    store[variable.id] = Math.round((Number.EPSILON + s.rhs[row]) * rounding) / rounding;
                }
                store.result = model.isMinimization ? s.evaluation : -s.evaluation;
                (model.solutions ??= []).push(store as never);
            }
        }
    }

    // ---- dynamic-modification.ts:16-55,78-316: edits go to the device tableau, no rebuild / re-upload
    private optSlot(priority: number): number {
        if (priority === 0) return -1;
        const k = this.optionalObjectives.findIndex((o) => o.priority === priority);
        if (k < 0) throw new Error("GpuTableau: an optional objective with a new priority needs setModel() again");
        return k;
    }
    private resync(): void {
        const i = this.tab().info();
        this.width = i.width; this.height = i.height; this.nVars = i.nVars; this.lastElementIndex = i.lastElementIndex;
        this.syncFromDevice();
    }
    getNewElementIndex(): number {
        if (this.availableIndexes.length > 0) return this.availableIndexes.pop() as number;
        this.lastElementIndex = this.tab().info().lastElementIndex;
        return this.lastElementIndex++;
    }
    putInBase(varIndex: number): number { const r = this.tab().putInBase(varIndex); this.syncFromDevice(); return r; }
    takeOutOfBase(varIndex: number): number { const c = this.tab().takeOutOfBase(varIndex); this.syncFromDevice(); return c; }
    updateRightHandSide(constraint: Constraint, difference: number): void {
        this.tab().updateRhs(constraint.index, difference); this.syncFromDevice();
    }
    updateConstraintCoefficient(constraint: Constraint, variable: Variable, difference: number): void {
        if (constraint.index === variable.index) {
            throw new Error("[Tableau.updateConstraintCoefficient] constraint index should not be equal to variable index !");
        }
        this.tab().updateCoefficient(constraint.index, variable.index, difference); this.syncFromDevice();
    }
    updateCost(variable: Variable, difference: number): void {
        this.tab().updateCost(variable.index, this.optSlot(variable.priority), difference); this.syncFromDevice();
    }
    addConstraint(constraint: Constraint): void {
        this.tab().addConstraint(constraint.isUpperBound, constraint.rhs, constraint.index,
                                 Int32Array.from(constraint.terms.map((t) => t.variable.index)),
                                 Float64Array.from(constraint.terms.map((t) => t.coefficient)));
        this.resync();
    }
    removeConstraint(constraint: Constraint): void {
        this.tab().removeConstraint(constraint.index);
        this.availableIndexes.push(constraint.index);       // dynamic-modification.ts:246
        constraint.slack.index = -1;                        // :248
        this.resync();
    }
    addVariable(variable: Variable): void {
        const cost = this.model!.isMinimization === true ? -variable.cost : variable.cost;   // :258
        this.tab().addVariable(variable.index, cost, this.optSlot(variable.priority), variable.isInteger === true,
                               this.unrestrictedVars[variable.index] === true);
        this.resync();
    }
    removeVariable(variable: Variable): void {
        this.tab().removeVariable(variable.index);
        this.availableIndexes.push(variable.index);         // :313
        this.resync();
    }

    /** Whole matrix on the host (MIR cuts, dynamic-modification API, debugging): one D2H copy of H*W doubles. */
    materializeMatrix(): Float64Array {
        const d = this.tab().download({ matrix: true });
        this.matrix = d.matrix!;
        this.hostMatrixPartial = false;
        return this.matrix;
    }

    private absorb(s: LpStatus): void {
        this.feasible = s.feasible; this.bounded = s.bounded; this.evaluation = s.evaluation;
        this.bestPossibleEval = s.bestPossibleEval; this.simplexIters = s.simplexIters;
        this.unboundedVarIndex = s.unboundedVarIndex < 0 ? null : s.unboundedVarIndex;
        if (s.cycled) this.model?.messages.push("Cycle in phase " + s.cycled, "Start :" + s.cycleStart, "Length :" + s.cycleLength);
        this.syncFromDevice();
    }

    /** updateVariableValues / generateSolutionSet read matrix[r * width] and the index maps: fetch just those. */
    private syncFromDevice(): void {
        const nOpt = this.optionalObjectives.length;
        const d = this.tab().download({ rhs: true, cost: true, maps: true, opt: nOpt });
        this.height = d.height;
        if (this.matrix.length < this.height * this.width) this.matrix = new Float64Array(this.height * this.width);
        for (let r = 0; r < this.height; r++) this.matrix[r * this.width] = d.rhs![r];
        this.matrix.set(d.cost!, 0);                                    // cost row: reduced costs after the solve
        this.hostMatrixPartial = true;
        this.varIndexByRow = Array.from(d.varIndexByRow!);
        this.varIndexByCol = Array.from(d.varIndexByCol!);
        const n = Math.max(this.rowByVarIndex.length, this.width + this.height);
        this.rowByVarIndex = new Array<number>(n).fill(-1);
        this.colByVarIndex = new Array<number>(n).fill(-1);
        for (let r = 1; r < this.height; r++) this.rowByVarIndex[this.varIndexByRow[r]] = r;
        for (let c = 1; c < this.width; c++) this.colByVarIndex[this.varIndexByCol[c]] = c;
        for (let o = 0; o < nOpt; o++) {
            const rc = this.optionalObjectives[o].reducedCosts;
            for (let c = 0; c < this.width; c++) rc[c] = d.opt![o * this.width + c];
        }
    }
}
