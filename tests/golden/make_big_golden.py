#!/usr/bin/env python
"""Generates the oracle's answers for BASELINE.json's two large configurations, so that the `-m gpu`
parity tests can compare the CUDA path against them without paying minutes of CPU on the GPU box:

  config3_dense2000.npz   dense packing LP 2000x2000, seed 12345 (bench.py's workload): full pivot log,
                          final basis arrays, right-hand-side column, cost row, SHA-256 of the final tableau
  config5_knapsack.npz    0/1 knapsack 1024 binaries x 512 constraints, seed 12345: root LP pivot log and
                          final root state, then the first N_NODES committed branch-and-cut nodes (node log),
                          final tableau hash after the capped run

Run from the repo root (CPU only, several minutes):  python tests/golden/make_big_golden.py [3] [5]
The inputs are regenerated from the seeded generators in jslpsolver_b200/problems.py (numpy's PCG64 stream
is stable across numpy versions for integers()), never stored.
"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))
N_NODES = 64  # committed nodes of the capped config-5 run


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def highs_objective(c, A, b, upper):
    """Optimal objective of max c.x, A x <= b, 0 <= x <= upper by an independent solver (SciPy's HiGHS)."""
    from scipy.optimize import linprog
    res = linprog(-np.asarray(c, dtype=float), A_ub=np.asarray(A, dtype=float), b_ub=np.asarray(b, dtype=float),
                  bounds=(0, upper), method="highs")
    assert res.status == 0
    return -res.fun


def config3():
    from jslpsolver_b200 import problems
    from oracle import ref_model
    it = problems.dense_packing_lp_tableau(2000, 2000, seed=12345)
    A, b, c = problems.dense_packing_lp_arrays(2000, 2000, seed=12345)
    t0 = time.time()
    o = ref_model.OracleTableau(it.matrix, it.varIndexByRow, it.varIndexByCol, fast_cycles=True, pivot_log=1 << 20)
    st = o.simplex()
    M = o.matrix()
    vrow, vcol = o.maps()
    np.savez_compressed(os.path.join(OUT, "config3_dense2000.npz"),
                        input_sha=np.array(sha(it.matrix)), pivot_log=o.pivot_log(), vrow=vrow, vcol=vcol,
                        rhs=M[:, 0].copy(), cost=M[0].copy(), matrix_sha=np.array(sha(M)),
                        flags=np.array([st.feasible, st.bounded, st.lastP1, st.lastP2, st.simplexIters], dtype=np.int64),
                        evaluation=np.array([st.evaluation, M[0, 0]]), highs_objective=np.array(highs_objective(c, A, b, None)))
    print(f"config 3: {st.lastP1}+{st.lastP2} pivots, evaluation {st.evaluation}, {time.time() - t0:.0f} s")


def knapsack_tableau():
    import jslpsolver_b200 as J
    from jslpsolver_b200 import problems
    model = problems.knapsack_mip_model(1024, 512, seed=12345)
    inst = J.Model().loadJson(model)
    return model, inst


def config5():
    from jslpsolver_b200 import problems
    from oracle import ref_model
    model = problems.knapsack_mip_model(1024, 512, seed=12345)
    t0 = time.time()
    # root LP alone, with its pivot log (the branch-and-cut run below repeats it as node 1)
    rm = ref_model.RefModel(None).loadJson(model)
    infeasible, fixed = ref_model.presolve(rm)
    assert not infeasible
    for var in fixed:
        var.value = fixed[var]
        var.cost = 0
    M, vrow, vcol, prios, rc = rm.build_tableau()
    ints = [v.index for v in rm.integerVariables]
    o = ref_model.OracleTableau(M, vrow, vcol, fast_cycles=True, integers=ints, is_min=rm.isMinimization,
                                pivot_log=1 << 20)
    st = o.simplex()
    Mr = o.matrix()
    rvrow, rvcol = o.maps()
    rng = np.random.default_rng(12345)  # the generator's arrays again (problems.knapsack_mip_model), for the independent solver
    Ak = rng.integers(1, 51, size=(512, 1024))
    ck = rng.integers(1, 51, size=1024)
    root = dict(root_highs_objective=np.array(highs_objective(ck, Ak, np.floor(0.5 * Ak.sum(axis=1)), 1)),
                root_input_sha=np.array(sha(M)), root_pivot_log=o.pivot_log(), root_vrow=rvrow, root_vcol=rvcol,
                root_rhs=Mr[:, 0].copy(), root_cost=Mr[0].copy(), root_matrix_sha=np.array(sha(Mr)),
                root_flags=np.array([st.feasible, st.bounded, st.lastP1, st.lastP2], dtype=np.int64),
                root_evaluation=np.array([st.evaluation, Mr[0, 0]]))
    print(f"config 5 root: {st.lastP1}+{st.lastP2} pivots, evaluation {st.evaluation}, {time.time() - t0:.0f} s", flush=True)
    del o
    sol = ref_model.solve_full(model, fast_cycles=True, node_log=1 << 16, max_nodes=N_NODES)
    t = sol.tableau
    Mf = t.matrix()
    fvrow, fvcol = t.maps()
    s = sol.state
    np.savez_compressed(os.path.join(OUT, "config5_knapsack.npz"), node_log=t.node_log(), max_nodes=np.array(N_NODES),
                        final_vrow=fvrow, final_vcol=fvcol, final_rhs=Mf[:, 0].copy(), final_matrix_sha=np.array(sha(Mf)),
                        final_flags=np.array([s.feasible, s.bounded, s.bncIterations, s.totalPivots, s.height], dtype=np.int64),
                        final_evaluation=np.array([s.evaluation, s.bestPossibleEval]), **root)
    print(f"config 5: {s.bncIterations} nodes, {s.totalPivots} pivots, {time.time() - t0:.0f} s")


if __name__ == "__main__":
    which = sys.argv[1:] or ["3", "5"]
    if "3" in which:
        config3()
    if "5" in which:
        config5()
