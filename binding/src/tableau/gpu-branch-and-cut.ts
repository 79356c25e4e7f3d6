/**
 * @file src/tableau/gpu-branch-and-cut.ts  (drop into the reference tree next to branch-and-cut.ts)
 * @description BranchAndCutService (branch-and-cut.ts:19-22) backed by the B200 frontier manager.
 *
 * `solver.branchAndCutService` / `selectBranchAndCutService` (src/main.ts:48-49,62-83) may return this service for
 * models without nodeSelection / branching / useIncremental options; both members delegate to GpuTableau, so a
 * plain host Tableau passed in by other callers falls back to the reference service.
 */
import type Tableau from "./tableau";
import type { BranchCut } from "./types";
import type { BranchAndCutService } from "./branch-and-cut";
import { createBranchAndCutService } from "./branch-and-cut";
import GpuTableau from "./gpu-tableau";

/** The enhanced service (enhanced-branch-and-cut.ts) on the device tableau: same options object as
 *  createEnhancedBranchAndCutService; main.ts:62-83 returns it when options.nodeSelection / options.branching are set. */
export function createGpuEnhancedBranchAndCutService(options: {
    nodeSelection?: "best-first" | "depth-first" | "hybrid";
    branching?: "most-fractional" | "pseudocost" | "strong";
    strongBranchingCandidates?: number;
} = {}): BranchAndCutService {
    const strategy = { enhanced: true, nodeSelection: options.nodeSelection ?? "hybrid", branching: options.branching ?? "pseudocost",
                       strongBranchingCandidates: options.strongBranchingCandidates ?? 5 };
    return {
        applyCuts(tableau: Tableau, cuts: BranchCut[]): void { (tableau as GpuTableau).applyCuts(cuts); },
        branchAndCut(tableau: Tableau): void { (tableau as GpuTableau).runBranchAndCut(strategy); },
    };
}

/** The incremental service (incremental-branch-and-cut.ts, options.useIncremental) on the device tableau: the enhanced
 *  loop with parent checkpoints kept in HBM. */
export function createGpuIncrementalBranchAndCutService(options: {
    nodeSelection?: "best-first" | "depth-first" | "hybrid";
    branching?: "most-fractional" | "pseudocost" | "strong";
} = {}): BranchAndCutService {
    const strategy = { useIncremental: true, nodeSelection: options.nodeSelection ?? "hybrid", branching: options.branching ?? "pseudocost" };
    return {
        applyCuts(tableau: Tableau, cuts: BranchCut[]): void { (tableau as GpuTableau).applyCuts(cuts); },
        branchAndCut(tableau: Tableau): void { (tableau as GpuTableau).runBranchAndCut(strategy); },
    };
}

export function createGpuBranchAndCutService(): BranchAndCutService {
    const host = createBranchAndCutService();
    return {
        applyCuts(tableau: Tableau, cuts: BranchCut[]): void {
            if (tableau instanceof GpuTableau) GpuTableau.prototype.applyCuts.call(tableau, cuts);
            else host.applyCuts(tableau, cuts);
        },
        branchAndCut(tableau: Tableau): void {
            if (tableau instanceof GpuTableau) GpuTableau.prototype.branchAndCut.call(tableau);
            else host.branchAndCut(tableau);
        },
    };
}
