"""Pins the oracle (oracle/) against every golden vector the reference holds for the LP/MIP
path: the `expects` blocks of the 47 test/test-sanity fixtures (compared exactly as
src/solver.integration.test.ts:60-100 does) and the README / integration-test known answers.
CPU only."""
import numpy as np
import pytest

from conftest import load_bundle
from helpers import compare_solutions, strip_timeouts
from oracle import ref_model

BUNDLE = load_bundle()


@pytest.mark.parametrize("fx", BUNDLE["fixtures"] + BUNDLE["readme"], ids=lambda f: f["file"])
def test_fixture_expects(fx):
    res = ref_model.Solve(strip_timeouts(fx["model"]), fast_cycles=True)
    bad = compare_solutions(res, fx["expects"])
    assert not bad, bad     # every fixture on every listed key, no exceptions


def test_readme_berlin_pivot_sequence():
    """SURVEY.md appendix A worked example: pivots (2,2) then (1,1), exact tie on row 1."""
    fx = [f for f in BUNDLE["readme"] if f["file"] == "README Berlin Airlift"][0]
    sol = ref_model.solve_full(fx["model"], pivot_log=16)
    log = sol.tableau.pivot_log()
    assert log[:, :2].tolist() == [[2, 2], [1, 1]]
    vrow, vcol = sol.tableau.maps()
    assert vrow.tolist() == [-1, 3, 4, 2] and vcol.tolist() == [-1, 0, 1]
    assert sol.tableau.matrix()[0, 0] == -1080000


def test_cycle_detectors_agree():
    """cycles_fast is only valid under the calling discipline of simplex.ts (checked after
    every push, stop at first hit); under that discipline it must equal the literal scan."""
    import ctypes
    L = ref_model.lib()
    rng = np.random.default_rng(7)
    for trial in range(300):
        n_sym = int(rng.integers(1, 5))
        seq = rng.integers(0, n_sym, size=(40, 2)).astype(np.int32)
        for n in range(1, len(seq) + 1):
            cur = np.ascontiguousarray(seq[:n])
            s1, l1, s2, l2 = (ctypes.c_long() for _ in range(4))
            h1 = L.orc_cycles_ref(cur.ctypes.data, n, ctypes.byref(s1), ctypes.byref(l1))
            h2 = L.orc_cycles_fast(cur.ctypes.data, n, ctypes.byref(s2), ctypes.byref(l2))
            assert h1 == h2, (trial, n)
            if h1:
                assert (s1.value, l1.value) == (s2.value, l2.value)
                break


def test_js_semantics_helpers():
    assert ref_model.js_keys({"b": 1, "21": 1, "a": 1, "3": 1, "03": 1}) == ["3", "21", "b", "a", "03"]
    assert ref_model.js_round(2.5) == 3 and ref_model.js_round(-2.5) == -2
    assert ref_model.js_round(0.49999999999999994) == 0


def test_cut_rows_known_answers():
    """addCutConstraints RHS rule (cutting-strategies.ts:36-62; the reference pins the same
    values in cutting-strategies.test.ts:73-101): basic var at value 2.5, cut x >= 3 -> RHS
    -(3-2.5) = -0.5; cut x <= 2 -> RHS 2-2.5 = -0.5; non-basic var -> RHS sign*value."""
    M = np.array([[0.0, 1.0, 2.0], [2.5, 0.5, -1.0], [4.0, 1.0, 1.0]])
    vrow = np.array([-1, 2, 3], dtype=np.int32)   # var 2 basic in row 1
    vcol = np.array([-1, 0, 1], dtype=np.int32)   # vars 0, 1 non-basic
    t = ref_model.OracleTableau(M, vrow, vcol)
    t.add_cuts([("min", 2, 3.0), ("max", 2, 2.0), ("min", 0, 1.0), ("max", 1, 7.0)])
    X = t.matrix()
    assert X.shape == (7, 3)
    assert X[3].tolist() == [-0.5, 0.5, -1.0]
    assert X[4].tolist() == [-0.5, -0.5, 1.0]
    assert X[5].tolist() == [-1.0, -1.0, 0.0]
    assert X[6].tolist() == [7.0, 0.0, 1.0]
    vrow2, _ = t.maps()
    assert vrow2.tolist() == [-1, 2, 3, 4, 5, 6, 7]


def _mip_tableau(rhs, basic_vars, integers, precision=1e-9):
    """Tableau whose row r >= 1 holds variable basic_vars[r-1] at value rhs[r-1] (RHS = column 0, as in the
    reference's real tableau; its unit tests put it in column 2 of a mock)."""
    H = len(rhs) + 1
    M = np.zeros((H, 3))
    M[1:, 0] = rhs
    vrow = np.array([-1] + list(basic_vars), dtype=np.int32)
    n_vars = 3 + H - 2  # the oracle sizes its inverse maps by width + height - 2, like the reference
    free = [v for v in range(n_vars) if v not in basic_vars]
    assert len(free) == 2
    vcol = np.array([-1] + free, dtype=np.int32)
    return ref_model.OracleTableau(M, vrow, vcol, precision=precision, integers=list(integers))


def test_is_integral_known_answers():
    """The cases of the reference's mip-utils.test.ts:173-300 (isIntegral, mip-utils.ts:43-61)."""
    assert _mip_tableau([5.0, 3.0], [1, 2], [1, 2]).is_integral() is True
    assert _mip_tableau([5.0, 3.7], [1, 2], [1, 2]).is_integral() is False
    assert _mip_tableau([5.5], [1], []).is_integral() is True                 # no integer variables
    assert _mip_tableau([5.0], [1], [1, 2]).is_integral() is True             # variable 2 is not basic
    assert _mip_tableau([4.9999999], [1], [1], precision=1e-6).is_integral() is True


def test_most_fractional_known_answers():
    """The cases of mip-utils.test.ts:431-565 (getMostFractionalVar, mip-utils.ts:100-126): largest distance
    to the nearest integer, the first integer variable wins ties, non-basic variables are skipped."""
    assert _mip_tableau([5.3, 3.7], [1, 2], [1, 2]).most_fractional() == (1, 5.3)   # both 0.3 away: first wins
    assert _mip_tableau([5.1, 3.5], [1, 2], [1, 2]).most_fractional() == (2, 3.5)
    assert _mip_tableau([5.3], [1], [1, 2]).most_fractional()[0] == 1               # variable 2 not basic
    i, v = _mip_tableau([5.0], [1], [1]).most_fractional()                          # nothing fractional
    assert i < 0 and v == 0.0
    i, _ = _mip_tableau([5.5], [1], []).most_fractional()                           # no integer variables
    assert i < 0


def test_solver_options_known_answers():
    """Known answers of the reference's solver.options.test.ts:68-136 that run through the default
    branch-and-cut service (timeout removed: it is wall-clock)."""
    r = ref_model.simplify(ref_model.solve_full({
        "optimize": "profit", "opType": "max", "constraints": {"capacity": {"max": 10}},
        "variables": {"x": {"profit": 5, "capacity": 2}}, "ints": {"x": 1}}))
    assert r["feasible"] is True and r["x"] == 5
    r = ref_model.simplify(ref_model.solve_full({
        "optimize": "profit", "opType": "max", "constraints": {"budget": {"max": 100}},
        "variables": {"a": {"profit": 10, "budget": 10}, "b": {"profit": 15, "budget": 15}, "c": {"profit": 20, "budget": 20}},
        "ints": {"a": 1, "b": 1, "c": 1}, "tolerance": 0.1}))
    assert r["feasible"] is True and r["result"] >= 90
    r = ref_model.simplify(ref_model.solve_full({
        "optimize": "profit", "opType": "max", "constraints": {"budget": {"max": 100}},
        "variables": {"a": {"profit": 10, "budget": 10}, "b": {"profit": 8, "budget": 10}},
        "ints": {"a": 1, "b": 1}, "tolerance": 0}))
    assert r["feasible"] is True and r["result"] == 100
