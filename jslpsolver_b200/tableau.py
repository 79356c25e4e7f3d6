"""GpuTableau: the reference's `Tableau` seam (src/tableau/tableau.ts) with storage, pivot loop
and branch-and-cut on the B200 behind the C ABI (include/jslp_b200.h).

Method names, flags and read-back state follow the reference 1:1:
  simplex/phase1/phase2/pivot      tableau.ts:103-123   -> jslp_simplex/phase1/phase2/pivot
  save/restore                     tableau.ts:223-229   -> jslp_save/restore
  addCutConstraints / applyCuts    tableau.ts:145,240   -> jslp_add_cuts / jslp_apply_cuts
  isIntegral/getMostFractionalVar  tableau.ts:135,163   -> jslp_is_integral / jslp_most_fractional
  branchAndCut                     tableau.ts:244-246   -> jslp_branch_and_cut (or an injected service)
  solve/getSolution                tableau.ts:250-274
`matrix` is materialised lazily from the device (the authoritative copy lives in HBM).
There is no CPU path: every numerical method raises JslpError without the CUDA library + a GPU.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np

from . import _lib
from ._lib import BnbOpts, BnbStatus, Cut, JslpError, LpStatus

EPSILON = 2.220446049250313e-16

_CTX: dict[int, "DeviceContext"] = {}


class DeviceContext:
    """One jslp_ctx per GPU (launches on torch's current stream when torch is initialised)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self.lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self.lib.jslp_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h)))
        self.handle = h
        self.device = device

    @property
    def launches(self) -> int:
        return int(self.lib.jslp_ctx_launches(self.handle))

    def sync(self) -> None:
        _lib.check(self.lib.jslp_ctx_sync(self.handle))

    def close(self) -> None:
        if self.handle:
            self.lib.jslp_ctx_destroy(self.handle)
            self.handle = None


def default_context(device: Optional[int] = None) -> DeviceContext:
    import os
    if device is None:
        device = int(os.environ.get("JSLP_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    if device not in _CTX:
        _CTX[device] = DeviceContext(device)
    return _CTX[device]


def js_round(x: float) -> float:
    if x != x or math.isinf(x):
        return x
    f = math.floor(x)
    return float(f + 1) if x - f >= 0.5 else float(f)


class Solution:  # solution.ts:16-61
    def __init__(self, tableau: "GpuTableau", evaluation: float, feasible: bool, bounded: bool):
        self.feasible, self.evaluation, self.bounded = feasible, evaluation, bounded
        self._tableau = tableau
        self.solutionSet: dict = {}

    def generateSolutionSet(self) -> dict:
        t = self._tableau
        out: dict = {}
        if t.handle is None:
            return out
        rhs, vrow = t.rhs_column(), t.varIndexByRow
        rounding = js_round(1 / t.precision)
        vpi = t.variablesPerIndex
        for r in range(1, t.height):
            var = vpi.get(int(vrow[r]))
            if var is None or var.isSlack:
                continue
            out[var.id] = js_round((EPSILON + float(rhs[r])) * rounding) / rounding
        return out


class MilpSolution(Solution):  # solution.ts:66-80
    def __init__(self, tableau, evaluation, feasible, bounded, branchAndCutIterations):
        super().__init__(tableau, evaluation, feasible, bounded)
        self.iter = branchAndCutIterations


class GpuTableau:
    def __init__(self, precision: float = 1e-8, branchAndCutService=None, context: Optional[DeviceContext] = None):
        self.precision = precision
        self.branchAndCutService = branchAndCutService
        self.context = context
        self.model = None
        self.handle = None
        self.width = self.height = 0
        self.feasible = True
        self.bounded = True
        self.evaluation = 0.0
        self.bestPossibleEval = 0.0
        self.simplexIters = 0
        self.unboundedVarIndex = None
        self.branchAndCutIterations = 0
        self.__isIntegral = None
        self.nVars = 0
        self.lastElementIndex = 0
        self.variablesPerIndex: dict = {}
        self.optionalPriorities: list = []
        self.lastStatus: Optional[LpStatus] = None
        self.lastBnbStatus: Optional[BnbStatus] = None
        self.bestCuts: list = []
        self.engine = 0
        self.max_spec_batch = 0
        self.distributed = False  # True: branchAndCut shards node rounds over torch.distributed ranks
        self.shard_policy = 0     # jslp_bnb_opts.shard_policy: 0 = shard only rounds whose node LPs run in HBM, 1 = always
        self.node_slots = None    # JSLP_OPT_NODE_SLOTS (None = library default: auto)
        self.slot_steps = None    # JSLP_OPT_SLOT_STEPS
        self.options: dict = {}   # JSLP_OPT_* -> value, applied after every upload (tuning aids)
        self.availableIndexes: list = []   # tableau.ts:77
        self._cache: dict = {}
        self._cache_log = None

    # name-mangling-free accessor used by Solve (main.ts:180)
    @property
    def isIntegralFlag(self):
        return self.__isIntegral

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def close(self) -> None:
        if self.handle is not None and self.context is not None and self.context.handle:
            self.context.lib.jslp_tab_destroy(self.handle)
        self.handle = None

    # ------------------------------------------------------------------ setup
    def upload(self, matrix, varIndexByRow, varIndexByCol, unrestricted=None, integerIndices=None,
               optionalCosts=None, row_capacity: Optional[int] = None) -> "GpuTableau":
        """Tableau.initialize + _resetMatrix result -> device (tableau.ts:292-380)."""
        if self.context is None:
            self.context = default_context()
        L = self.context.lib
        M = np.ascontiguousarray(matrix, dtype=np.float64)
        H, W = M.shape
        self.close()
        h = C.c_void_p()
        cap = row_capacity if row_capacity else H + 64
        _lib.check(L.jslp_tab_create(self.context.handle, W, H, cap, float(self.precision), C.byref(h)))
        self.handle = h
        vrow = np.ascontiguousarray(varIndexByRow, dtype=np.int32)
        vcol = np.ascontiguousarray(varIndexByCol, dtype=np.int32)
        n_index = W + H - 2
        unres = None
        if unrestricted is not None and len(unrestricted):
            unres = np.zeros(n_index, dtype=np.uint8)
            unres[:len(unrestricted)] = np.asarray(unrestricted, dtype=np.uint8)[:n_index]
        ints = None if integerIndices is None else np.ascontiguousarray(integerIndices, dtype=np.int32)
        opt = None if optionalCosts is None or len(optionalCosts) == 0 else np.ascontiguousarray(optionalCosts, dtype=np.float64)
        self.nOpt = 0 if opt is None else opt.shape[0]
        _lib.check(L.jslp_tab_upload(
            self.handle, M.ctypes.data, vrow.ctypes.data, vcol.ctypes.data,
            None if unres is None else unres.ctypes.data, n_index,
            None if ints is None or len(ints) == 0 else ints.ctypes.data, 0 if ints is None else len(ints),
            self.nOpt, None if opt is None else opt.ctypes.data))
        self.width, self.height = W, H
        self.nVars = n_index
        self.lastElementIndex = n_index
        self.nInts = 0 if ints is None else len(ints)
        if self.engine:
            self.set_option(_lib.OPT_ENGINE, self.engine)
        for k, v in self.options.items():
            self.set_option(k, v)
        self._cache.clear()
        self._cache_log = None
        return self

    def setModel(self, model) -> "GpuTableau":  # tableau.ts:382-391
        self.model = model
        it = model.initial_tableau()
        self.variablesPerIndex = model.variablesPerIndex
        self.optionalPriorities = it.optionalPriorities
        return self.upload(it.matrix, it.varIndexByRow, it.varIndexByCol, it.unrestricted, it.integerIndices,
                           it.optionalCosts)

    def set_option(self, key: int, value: float) -> None:
        _lib.check(self.context.lib.jslp_tab_set_option(self._h(), key, float(value)))

    def _h(self):
        if self.handle is None:
            raise JslpError("tableau has no device state: call setModel()/upload() first")
        return self.handle

    def _check_cycles(self) -> int:
        return int(bool(getattr(self.model, "checkForCycles", True)))

    def _absorb(self, st: LpStatus) -> None:
        self.lastStatus = st
        self.feasible = bool(st.feasible)
        self.bounded = bool(st.bounded)
        self.evaluation = st.evaluation
        self.bestPossibleEval = st.best_possible_eval
        self.simplexIters = st.simplex_iters
        self.unboundedVarIndex = None if st.unbounded_var_index < 0 else st.unbounded_var_index
        self.width, self.height = st.width, st.height
        if st.cycled and self.model is not None:  # simplex.ts:86-88 / 313-315
            self.model.messages += [f"Cycle in phase {st.cycled}", f"Start :{st.cycle_start}",
                                    f"Length :{st.cycle_length}"]
        self._cache.clear()

    # ------------------------------------------------------------------ simplex (tableau.ts:103-123)
    def simplex(self) -> "GpuTableau":
        st = LpStatus()
        _lib.check(self.context.lib.jslp_simplex(self._h(), self._check_cycles(), C.byref(st)))
        self._absorb(st)
        return self

    def phase1(self) -> int:
        st = LpStatus()
        _lib.check(self.context.lib.jslp_phase1(self._h(), self._check_cycles(), C.byref(st)))
        self._absorb(st)
        return st.phase1_pivots

    def phase2(self) -> int:
        st = LpStatus()
        _lib.check(self.context.lib.jslp_phase2(self._h(), self._check_cycles(), C.byref(st)))
        self._absorb(st)
        return st.phase2_pivots

    def pivot(self, pivotRowIndex: int, pivotColumnIndex: int) -> None:
        _lib.check(self.context.lib.jslp_pivot(self._h(), pivotRowIndex, pivotColumnIndex))
        self._cache.clear()

    # ------------------------------------------------------------------ backup (tableau.ts:219-229)
    def save(self) -> None:
        _lib.check(self.context.lib.jslp_save(self._h()))

    def restore(self) -> None:
        _lib.check(self.context.lib.jslp_restore(self._h()))
        self._refresh_dims()

    def _refresh_dims(self) -> None:
        info = (C.c_int32 * 6)()
        _lib.check(self.context.lib.jslp_tab_info(self._h(), info))
        self.width, self.height, self.nVars, self.lastElementIndex = info[0], info[1], info[2], info[3]
        self._cache.clear()

    def getNewElementIndex(self) -> int:  # tableau.ts:393-401
        if self.availableIndexes:
            return self.availableIndexes.pop()
        self._refresh_dims()
        index = self.lastElementIndex
        self.lastElementIndex += 1
        return index

    # ------------------------------------------------------------------ dynamic modification (dynamic-modification.ts)
    def _opt_slot(self, priority) -> int:
        if not priority:
            return -1
        if priority not in self.optionalPriorities:
            raise JslpError("an optional objective with a new priority needs setModel() again (the uploaded set is fixed)")
        return self.optionalPriorities.index(priority)

    def putInBase(self, varIndex: int) -> int:
        r = C.c_int()
        _lib.check(self.context.lib.jslp_put_in_base(self._h(), int(varIndex), C.byref(r)))
        self._cache.clear()
        return r.value

    def takeOutOfBase(self, varIndex: int) -> int:
        c = C.c_int()
        _lib.check(self.context.lib.jslp_take_out_of_base(self._h(), int(varIndex), C.byref(c)))
        self._cache.clear()
        return c.value

    def updateRightHandSide(self, constraint, difference: float) -> None:
        idx = constraint if isinstance(constraint, int) else constraint.index
        _lib.check(self.context.lib.jslp_update_rhs(self._h(), int(idx), float(difference)))
        self._cache.clear()

    def updateConstraintCoefficient(self, constraint, variable, difference: float) -> None:
        ci = constraint if isinstance(constraint, int) else constraint.index
        vi = variable if isinstance(variable, int) else variable.index
        if ci == vi:  # dynamic-modification.ts:114-118
            raise ValueError("[Tableau.updateConstraintCoefficient] constraint index should not be equal to variable index !")
        _lib.check(self.context.lib.jslp_update_coefficient(self._h(), int(ci), int(vi), float(difference)))
        self._cache.clear()

    def updateCost(self, variable, difference: float, priority: int = 0) -> None:
        vi = variable if isinstance(variable, int) else variable.index
        pr = priority if isinstance(variable, int) else variable.priority
        _lib.check(self.context.lib.jslp_update_cost(self._h(), int(vi), self._opt_slot(pr), float(difference)))
        self._cache.clear()

    def addConstraint(self, constraint=None, *, isUpperBound=None, rhs=None, index=None, terms=None) -> None:
        if constraint is not None:
            isUpperBound, rhs, index = constraint.isUpperBound, constraint.rhs, constraint.index
            terms = [(t.variable.index, t.coefficient) for t in constraint.terms]
        terms = terms or []
        tv = np.ascontiguousarray([v for v, _ in terms], dtype=np.int32)
        tc = np.ascontiguousarray([c for _, c in terms], dtype=np.float64)
        _lib.check(self.context.lib.jslp_add_constraint(self._h(), int(bool(isUpperBound)), float(rhs), int(index),
                                                        tv.ctypes.data if len(terms) else None,
                                                        tc.ctypes.data if len(terms) else None, len(terms)))
        self._refresh_dims()

    def removeConstraint(self, constraint) -> None:
        idx = constraint if isinstance(constraint, int) else constraint.index
        _lib.check(self.context.lib.jslp_remove_constraint(self._h(), int(idx)))
        self.availableIndexes.append(idx)          # dynamic-modification.ts:246
        if not isinstance(constraint, int):
            constraint.slack.index = -1            # :248
        self._refresh_dims()

    def addVariable(self, variable=None, *, index=None, cost=0.0, priority=0, isInteger=False, isUnrestricted=False) -> None:
        if variable is not None:
            index, cost, priority, isInteger = variable.index, variable.cost, variable.priority, variable.isInteger
            isUnrestricted = bool(self.model is not None and self.model.unrestrictedVariables.get(index))
        is_min = True if self.model is None else self.model.isMinimization
        entry = -cost if is_min else cost                      # dynamic-modification.ts:258
        _lib.check(self.context.lib.jslp_add_variable(self._h(), int(index), float(entry), self._opt_slot(priority),
                                                      int(bool(isInteger)), int(bool(isUnrestricted))))
        self._refresh_dims()

    def removeVariable(self, variable) -> None:
        idx = variable if isinstance(variable, int) else variable.index
        _lib.check(self.context.lib.jslp_remove_variable(self._h(), int(idx)))
        self.availableIndexes.append(idx)          # dynamic-modification.ts:313
        self._refresh_dims()

    # ------------------------------------------------------------------ cuts / MIP helpers
    @staticmethod
    def _cut_array(cuts):
        arr = (Cut * max(1, len(cuts)))()
        for i, c in enumerate(cuts):
            if isinstance(c, dict):
                ty, vi, val = c["type"], c["varIndex"], c["value"]
            else:
                ty, vi, val = c
            arr[i].type = 0 if ty in (0, "min") else 1
            arr[i].var_index = int(vi)
            arr[i].value = float(val)
        return arr

    def addCutConstraints(self, branchingCuts) -> None:
        _lib.check(self.context.lib.jslp_add_cuts(self._h(), self._cut_array(branchingCuts), len(branchingCuts)))
        self._refresh_dims()

    def addLowerBoundMIRCut(self, rowIndex: int) -> bool:   # cutting-strategies.ts:74-134
        a = C.c_int()
        _lib.check(self.context.lib.jslp_add_mir_cut(self._h(), int(rowIndex), 0, C.byref(a)))
        self._refresh_dims()
        return bool(a.value)

    def addUpperBoundMIRCut(self, rowIndex: int) -> bool:   # cutting-strategies.ts:136-196
        a = C.c_int()
        _lib.check(self.context.lib.jslp_add_mir_cut(self._h(), int(rowIndex), 1, C.byref(a)))
        self._refresh_dims()
        return bool(a.value)

    def applyMIRCuts(self) -> None:                         # cutting-strategies.ts:198-212
        n = C.c_int()
        _lib.check(self.context.lib.jslp_apply_mir_cuts(self._h(), C.byref(n)))
        self._refresh_dims()

    def computeFractionalVolume(self, ignoreIntegerValues: bool = False) -> float:  # mip-utils.ts:67-98
        v = C.c_double()
        _lib.check(self.context.lib.jslp_fractional_volume(self._h(), int(bool(ignoreIntegerValues)), C.byref(v)))
        return v.value

    def _sync_mir_option(self) -> None:
        self.set_option(_lib.OPT_USE_MIR_CUTS, 1 if getattr(self.model, "useMIRCuts", False) else 0)

    def applyCuts(self, branchingCuts) -> None:
        if self.branchAndCutService is not None:
            self.branchAndCutService.applyCuts(self, branchingCuts)
            return
        self._sync_mir_option()
        st = LpStatus()
        _lib.check(self.context.lib.jslp_apply_cuts(self._h(), self._cut_array(branchingCuts), len(branchingCuts),
                                                    self._check_cycles(), C.byref(st)))
        self._absorb(st)

    def isIntegral(self) -> bool:
        v = C.c_int()
        _lib.check(self.context.lib.jslp_is_integral(self._h(), C.byref(v)))
        return bool(v.value)

    def getMostFractionalVar(self) -> dict:
        i, v = C.c_int32(), C.c_double()
        _lib.check(self.context.lib.jslp_most_fractional(self._h(), C.byref(i), C.byref(v)))
        return {"index": None if i.value < 0 else i.value, "value": v.value}

    def branchAndCut(self) -> None:  # tableau.ts:244-246
        if self.branchAndCutService is not None:
            self.branchAndCutService.branchAndCut(self)
            return
        m = self.model
        self._sync_mir_option()
        opts = BnbOpts()
        opts.tolerance = float(getattr(m, "tolerance", 0) or 0)
        opts.is_minimization = int(bool(getattr(m, "isMinimization", True)))
        opts.check_cycles = self._check_cycles()
        opts.max_spec_batch = int(self.max_spec_batch)
        opts.rank, opts.n_ranks = 0, 1
        keep = None
        if self.distributed:  # shard each round's nodes over ranks (one process per GPU)
            from . import distributed as D
            if D.is_active():
                opts.rank, opts.n_ranks = D.rank_and_world()
                comm = D.nccl_communicator(self.context)  # in-library NCCL; None under gloo (CPU tests, shared GPU)
                if comm is not None:
                    opts.comm = comm
                else:
                    opts.all_gather, keep = D.make_all_gather_hook()
        opts.shard_policy = int(self.shard_policy)
        sel = getattr(m, "branchAndCutOptions", None)  # main.ts:62-83: options.nodeSelection / options.branching
        if sel:
            opts.service = 2 if sel.get("useIncremental") else 1
            opts.node_selection = {"best-first": 1, "depth-first": 2, "hybrid": 3}[sel.get("nodeSelection") or "hybrid"]
            opts.branching = {"most-fractional": 1, "pseudocost": 2, "strong": 3}[sel.get("branching") or "pseudocost"]
        opts.max_nodes = int(getattr(m, "max_nodes", 0) or 0)
        # model.timeout [ms] and options.keep_solutions (model.ts:338-374; branch-and-cut.ts:61-63,76,143-153)
        opts.timeout_ms = float(getattr(m, "timeout", 0) or 0)
        opts.keep_solutions = int(bool(getattr(m, "keep_solutions", False)))
        if self.node_slots is not None:
            self.set_option(_lib.OPT_NODE_SLOTS, self.node_slots)
        if self.slot_steps is not None:
            self.set_option(_lib.OPT_SLOT_STEPS, self.slot_steps)
        st = BnbStatus()
        cap = 4096
        best = (Cut * cap)()
        _lib.check(self.context.lib.jslp_branch_and_cut(self._h(), C.byref(opts), C.byref(st), best, cap))
        self.lastBnbStatus = st
        self.feasible, self.bounded = bool(st.feasible), bool(st.bounded)
        self.evaluation, self.bestPossibleEval = st.evaluation, st.best_possible_eval
        self.branchAndCutIterations = st.iterations
        if st.is_integral:
            self.__isIntegral = True
        self.bestCuts = [(best[i].type, best[i].var_index, best[i].value) for i in range(min(cap, st.n_best_cuts))]
        self._refresh_dims()
        if opts.keep_solutions and m is not None:  # branch-and-cut.ts:143-153
            for i in range(st.n_solutions):
                m.solutions = (m.solutions or []) + [self._stored_solution(i)]

    def _stored_solution(self, i: int) -> dict:
        """model.solutions[i]: generateSolutionSet() of incumbent i plus `result` (branch-and-cut.ts:144-152)."""
        L = self.context.lib
        ev, h = C.c_double(), C.c_int32()
        _lib.check(L.jslp_bnb_solution(self._h(), i, C.byref(ev), C.byref(h), None, None, 0))
        vrow, rhs = np.empty(h.value, dtype=np.int32), np.empty(h.value, dtype=np.float64)
        _lib.check(L.jslp_bnb_solution(self._h(), i, None, None, vrow.ctypes.data, rhs.ctypes.data, h.value))
        rounding = js_round(1 / self.precision)
        store: dict = {}
        for r in range(1, h.value):
            var = self.variablesPerIndex.get(int(vrow[r]))
            if var is None or var.isSlack:
                continue
            store[var.id] = js_round((EPSILON + float(rhs[r])) * rounding) / rounding
        is_min = True if self.model is None else self.model.isMinimization
        store["result"] = ev.value if is_min else -ev.value
        return store

    def node_log(self) -> np.ndarray:
        n = C.c_int64()
        L = self.context.lib
        _lib.check(L.jslp_bnb_node_log(self._h(), None, 0, C.byref(n)))
        out = np.empty((n.value, 8), dtype=np.float64)
        if n.value:
            _lib.check(L.jslp_bnb_node_log(self._h(), out.ctypes.data, n.value, C.byref(n)))
        return out

    # ------------------------------------------------------------------ read-back
    def _download(self, what: str):
        if what in self._cache:
            return self._cache[what]
        L = self.context.lib
        H, W = self.height, self.width
        if what == "matrix":
            out = np.empty((H, W), dtype=np.float64)
            _lib.check(L.jslp_download(self._h(), out.ctypes.data, None, None, None, None, None, None, None))
        elif what == "rhs":
            out = np.empty(H, dtype=np.float64)
            _lib.check(L.jslp_download(self._h(), None, out.ctypes.data, None, None, None, None, None, None))
        elif what == "cost":
            out = np.empty(W, dtype=np.float64)
            _lib.check(L.jslp_download(self._h(), None, None, out.ctypes.data, None, None, None, None, None))
        elif what == "maps":
            vr, vc = np.empty(H, dtype=np.int32), np.empty(W, dtype=np.int32)
            _lib.check(L.jslp_download(self._h(), None, None, None, vr.ctypes.data, vc.ctypes.data, None, None, None))
            out = (vr, vc)
        elif what == "opt":
            out = np.empty((self.nOpt, W), dtype=np.float64)
            if self.nOpt:
                _lib.check(L.jslp_download(self._h(), None, None, None, None, None, out.ctypes.data, None, None))
        else:
            raise KeyError(what)
        self._cache[what] = out
        return out

    @property
    def matrix(self) -> np.ndarray:
        """Flat row-major Float64Array view, stride == width (tableau.ts:49,304)."""
        return self._download("matrix").reshape(-1)

    def matrix2d(self) -> np.ndarray:
        return self._download("matrix")

    def rhs_column(self) -> np.ndarray:
        return self._download("rhs")

    def cost_row(self) -> np.ndarray:
        return self._download("cost")

    @property
    def varIndexByRow(self) -> np.ndarray:
        return self._download("maps")[0]

    @property
    def varIndexByCol(self) -> np.ndarray:
        return self._download("maps")[1]

    @property
    def rowByVarIndex(self) -> np.ndarray:
        vr, vc = self._download("maps")
        n = max(int(vr.max(initial=-1)), int(vc.max(initial=-1))) + 1
        out = np.full(n, -1, dtype=np.int32)
        for r in range(1, len(vr)):
            out[vr[r]] = r
        return out

    @property
    def colByVarIndex(self) -> np.ndarray:
        vr, vc = self._download("maps")
        n = max(int(vr.max(initial=-1)), int(vc.max(initial=-1))) + 1
        out = np.full(n, -1, dtype=np.int32)
        for c in range(1, len(vc)):
            out[vc[c]] = c
        return out

    def optional_reduced_costs(self) -> np.ndarray:
        return self._download("opt")

    def pivot_log(self) -> np.ndarray:
        """All pivots executed on this tableau since upload (row, col, leaving var, entering var);
        needs JSLP_OPT_PIVOT_LOG_CAP.  The library hands the log over in drains; they are
        accumulated here."""
        n = C.c_int()
        L = self.context.lib
        cap = 1 << 20
        buf = np.empty((cap, 4), dtype=np.int32)
        _lib.check(L.jslp_pivot_log(self._h(), buf.ctypes.data, cap, C.byref(n)))
        new = buf[:min(cap, n.value)].copy()
        old = self._cache_log if self._cache_log is not None else np.empty((0, 4), dtype=np.int32)
        self._cache_log = np.concatenate([old, new]) if len(new) else old
        return self._cache_log

    # ------------------------------------------------------------------ solve (tableau.ts:250-274)
    def updateVariableValues(self) -> None:  # dynamic-modification.ts:57-76
        if self.model is None or self.handle is None:
            return
        rhs = self.rhs_column()
        row_of = self.rowByVarIndex
        rounding = js_round(1 / self.precision)
        for var in self.model.variables:
            r = row_of[var.index] if var.index < len(row_of) else -1
            var.value = 0 if r == -1 else js_round((float(rhs[r]) + EPSILON) * rounding) / rounding

    def solve(self):
        if self.model is not None and self.model.getNumberOfIntegerVariables() > 0:
            self.branchAndCut()
        else:
            self.simplex()
        self.updateVariableValues()
        return self.getSolution()

    def getSolution(self):
        is_min = True if self.model is None else self.model.isMinimization
        evaluation = self.evaluation if is_min else -self.evaluation
        if self.model is not None and self.model.getNumberOfIntegerVariables() > 0:
            return MilpSolution(self, evaluation, self.feasible, self.bounded, self.branchAndCutIterations)
        return Solution(self, evaluation, self.feasible, self.bounded)
