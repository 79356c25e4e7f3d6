"""Shared test helpers: the reference's own comparison rule, restated.

compare_solutions follows src/solver.integration.test.ts:60-100 of the reference: both
infeasible short-circuits; otherwise every key listed in `expects` (except feasible/_timeout/
isIntegral/bounded) is compared after Number(v.toFixed(6)); a missing actual key counts as 0.
"""
import copy
import math
from decimal import ROUND_HALF_UP, Decimal


def to_fixed6(value):
    if isinstance(value, str):
        try:
            value = float(value)
        except ValueError:
            return value
    if isinstance(value, bool):
        return value
    if isinstance(value, (int, float)):
        if math.isfinite(value):
            d = Decimal(float(value)).quantize(Decimal("0.000001"), rounding=ROUND_HALF_UP)
            return float(d) + 0.0
        return value
    return 0 if value is None else value


def compare_solutions(actual: dict, expected: dict):
    """Returns a list of mismatch strings (empty = pass)."""
    if not actual.get("feasible") and not expected.get("feasible"):
        return []
    bad = []
    if bool(actual.get("feasible")) != bool(expected.get("feasible")):
        bad.append(f"feasible: {actual.get('feasible')} != {expected.get('feasible')}")
    for key, ev in expected.items():
        if key in ("feasible", "_timeout", "isIntegral", "bounded"):
            continue
        a, e = to_fixed6(actual.get(key)), to_fixed6(ev)
        if a != e and not (a != a and e != e):
            bad.append(f"{key}: {a} != {e}")
    return bad


def strip_timeouts(model: dict) -> dict:
    """Wall-clock `timeout` is inherently non-reproducible (branch-and-cut.ts:61-63,76)."""
    m = copy.deepcopy(model)
    m.pop("timeout", None)
    if isinstance(m.get("options"), dict):
        m["options"].pop("timeout", None)
    return m


def highs_solve(model: dict):
    """A jsLPSolver JSON model (single objective; `min`/`max`/`equal` rows, `ints`, `binaries`) solved by SciPy's
    HiGHS, which shares no code or pivot rule with the reference, the oracle or the CUDA path.
    Returns (status, objective): status 0 optimal, 2 infeasible, 3 unbounded, 4 unbounded-or-infeasible (scipy.optimize.milp's codes)."""
    import numpy as np
    from scipy.optimize import Bounds, LinearConstraint, milp
    names = list(model["variables"])
    key = model["optimize"]
    sign = -1.0 if model["opType"] == "max" else 1.0
    c = sign * np.array([model["variables"][v].get(key, 0.0) for v in names], dtype=float)
    rows, lo, hi = [], [], []
    for cname, spec in model["constraints"].items():
        rows.append([model["variables"][v].get(cname, 0.0) for v in names])
        if "equal" in spec:
            lo.append(spec["equal"]); hi.append(spec["equal"])
        else:
            lo.append(spec.get("min", -np.inf)); hi.append(spec.get("max", np.inf))
    ints, bins = model.get("ints", {}), model.get("binaries", {})
    integrality = np.array([1 if (v in ints or v in bins) else 0 for v in names])
    upper = np.array([1.0 if v in bins else np.inf for v in names])
    res = milp(c, constraints=LinearConstraint(np.array(rows, dtype=float), lo, hi), integrality=integrality,
               bounds=Bounds(0, upper))
    return res.status, (sign * res.fun if res.status == 0 else None)
