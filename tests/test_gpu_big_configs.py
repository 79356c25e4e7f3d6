"""GPU parity at BASELINE.json's two large configurations, against the oracle's answers cached in
tests/golden/config3_dense2000.npz and config5_knapsack.npz (made by tests/golden/make_big_golden.py, minutes
of CPU): full pivot log, basis index arrays, flags, right-hand-side column, cost row bit for bit, and the
SHA-256 of the whole final tableau.  Inputs are regenerated from the seeded generators and their hash is
checked against the one the fixture was made from."""
import hashlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def first_diff(glog, olog):
    n = min(len(glog), len(olog))
    bad = np.nonzero((glog[:n] != olog[:n]).any(axis=1))[0]
    return int(bad[0]) if len(bad) else n


def check_lp(g, st, z, prefix, what):
    glog, olog = g.pivot_log(), z[prefix + "pivot_log"]
    if len(glog) != len(olog) or not np.array_equal(glog, olog):
        i = first_diff(glog, olog)
        raise AssertionError(f"{what}: pivot sequence differs at pivot {i} of {len(olog)}: gpu={glog[i:i + 2].tolist()} "
                             f"oracle={olog[i:i + 2].tolist()} (gpu has {len(glog)})")
    fl = z[prefix + "flags"]
    assert (st.feasible, st.bounded, st.phase1_pivots, st.phase2_pivots) == tuple(int(x) for x in fl[:4]), what
    assert np.array_equal(g.varIndexByRow, z[prefix + "vrow"]) and np.array_equal(g.varIndexByCol, z[prefix + "vcol"]), what
    M = g.matrix2d()
    assert np.array_equal(bits(M[:, 0]), bits(z[prefix + "rhs"])), f"{what}: right-hand-side column bits differ"
    assert np.array_equal(bits(M[0]), bits(z[prefix + "cost"])), f"{what}: cost row bits differ"
    assert sha(M) == str(z[prefix + "matrix_sha"]), f"{what}: tableau hash differs"
    ev = z[prefix + "evaluation"]
    assert bits(st.evaluation) == bits(ev[0]) and bits(st.evaluation_raw) == bits(ev[1]), what
    # and against a solver that shares nothing with the reference or the oracle (SciPy's HiGHS, cached by the fixture
    # script): north_star's "objective within 1e-9 relative"
    highs = float(z[prefix + "highs_objective"])
    assert abs(-st.evaluation - highs) <= 1e-9 * abs(highs), (what, st.evaluation, highs)


# (engine, look-ahead, step variant[, pdl, pingpong]) -- the benched default first
CONFIG3_ENGINES = {"default": None, "fused_v0": (2, 1, 0), "fused_inplace": (2, 1, 1, 0, 0)}


@pytest.mark.parametrize("engine", list(CONFIG3_ENGINES))
def test_config3_dense_2000x2000_matches_oracle(engine):
    """bench.py's workload: 2001x2001 tableau, 6-7 rows per row CTA, two RC=4 passes, prefetch path."""
    from jslpsolver_b200 import _lib, problems
    from jslpsolver_b200.tableau import GpuTableau
    z = np.load(os.path.join(GOLD, "config3_dense2000.npz"))
    it = problems.dense_packing_lp_tableau(2000, 2000, seed=12345)
    assert sha(it.matrix) == str(z["input_sha"]), "generator drifted from the fixture's input"
    g = GpuTableau(1e-8)
    g.upload(it.matrix, it.varIndexByRow, it.varIndexByCol)
    g.set_option(_lib.OPT_PIVOT_LOG_CAP, 1 << 20)
    spec = CONFIG3_ENGINES[engine]
    if spec is not None:
        g.set_option(_lib.OPT_ENGINE, spec[0])
        g.set_option(_lib.OPT_LOOKAHEAD, spec[1])
        g.set_option(_lib.OPT_STEP_VARIANT, spec[2])
        if len(spec) > 3:
            g.set_option(_lib.OPT_PDL, spec[3])
            g.set_option(_lib.OPT_PINGPONG, spec[4])
    g.simplex()
    check_lp(g, g.lastStatus, z, "", f"config 3 [{engine}]")
    g.close()


def knapsack_instance():
    import jslpsolver_b200 as J
    from jslpsolver_b200 import problems
    from jslpsolver_b200.model import presolve
    model = problems.knapsack_mip_model(1024, 512, seed=12345)
    m = J.Model().loadJson(model)
    pr = presolve(m)
    assert not pr.isInfeasible
    for v in pr.fixedVariables:
        v.cost = 0
    return model, m


def test_config5_root_lp_matches_oracle():
    """Knapsack root relaxation: 1537x1025, ~85 k pivots, the mid-size (selector-chain bound) regime."""
    from jslpsolver_b200 import _lib
    from jslpsolver_b200.tableau import GpuTableau
    z = np.load(os.path.join(GOLD, "config5_knapsack.npz"))
    model, m = knapsack_instance()
    it = m.initial_tableau()
    assert sha(it.matrix) == str(z["root_input_sha"]), "generator / front end drifted from the fixture's input"
    g = GpuTableau(1e-8)
    g.upload(it.matrix, it.varIndexByRow, it.varIndexByCol, it.unrestricted, it.integerIndices, it.optionalCosts)
    g.set_option(_lib.OPT_PIVOT_LOG_CAP, 1 << 20)
    g.simplex()
    check_lp(g, g.lastStatus, z, "root_", "config 5 root")
    g.close()


@pytest.mark.parametrize("mode", ["spec1", "spec16", "spec64"])
def test_config5_first_nodes_match_oracle(mode):
    """Root + the first committed branch-and-cut nodes (HBM-path node LPs: restore, cut rows, phase-1 ping-pong
    steps), whatever the speculation width: same node log, same final tableau."""
    import jslpsolver_b200 as J
    z = np.load(os.path.join(GOLD, "config5_knapsack.npz"))
    model, _ = knapsack_instance()
    inst = J.Model().loadJson(model)
    inst.max_nodes = int(z["max_nodes"])
    inst.tableau.max_spec_batch = int(mode[4:])
    inst.solve()
    gt = inst.tableau
    gnl, onl = gt.node_log(), z["node_log"]
    assert gnl.shape == onl.shape, (gnl.shape, onl.shape)
    for i in range(len(onl)):
        a, b = gnl[i], onl[i]
        ok = all(a[k] == b[k] for k in (0, 1, 2, 4, 5, 7)) and bits(a[6]) == bits(b[6]) and (not b[2] or bits(a[3]) == bits(b[3]))
        assert ok, f"node {i}: gpu={a.tolist()} oracle={b.tolist()}"
    fl = z["final_flags"]
    assert (int(gt.feasible), int(gt.bounded), gt.branchAndCutIterations, gt.height) == (int(fl[0]), int(fl[1]), int(fl[2]), int(fl[4]))
    assert gt.lastBnbStatus.pivots == int(fl[3]), (gt.lastBnbStatus.pivots, int(fl[3]))
    assert np.array_equal(gt.varIndexByRow, z["final_vrow"]) and np.array_equal(gt.varIndexByCol, z["final_vcol"])
    M = gt.matrix2d()
    assert np.array_equal(bits(M[:, 0]), bits(z["final_rhs"]))
    assert sha(M) == str(z["final_matrix_sha"])
    gt.close()
