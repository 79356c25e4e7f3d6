"""Synthetic workloads for BASELINE.json's configs (SURVEY.md 8d "Concrete inputs").

Distributions follow the reference's seeded generators (src/test-utils/problem-generator.ts:
generateResourceAllocation 297-340, generateKnapsack); the random stream itself is numpy's
(the reference's Mulberry32 accumulates its seed as a double and is not reproducible across
languages for millions of draws -- SURVEY.md 8c), so one materialised instance is fed to every
implementation that is compared.
"""
from __future__ import annotations

import numpy as np

from .model import InitialTableau


def dense_packing_lp_arrays(n_vars: int, n_cons: int, seed: int = 12345):
    """max c.x  s.t.  A x <= b, x >= 0 with a_ij ~ U{1..20}, b_i ~ U{100..500}, c_j ~ U{1..50}."""
    rng = np.random.default_rng(seed)
    A = rng.integers(1, 21, size=(n_cons, n_vars)).astype(np.float64)
    b = rng.integers(100, 501, size=n_cons).astype(np.float64)
    c = rng.integers(1, 51, size=n_vars).astype(np.float64)
    return A, b, c


def dense_packing_lp_tableau(n_vars: int, n_cons: int, seed: int = 12345) -> InitialTableau:
    """The tableau `Model.loadJson(dense_packing_lp_model(...))` + `Tableau.setModel` would build
    (tableau.ts:319-380), emitted directly: constraints take element indices 0..m-1, variables
    m..m+n-1 (model.ts:288-332,382-416); opType max -> cost row +c (tableau.ts:332-339)."""
    A, b, c = dense_packing_lp_arrays(n_vars, n_cons, seed)
    H, W = n_cons + 1, n_vars + 1
    M = np.zeros((H, W), dtype=np.float64)
    M[0, 1:] = c
    M[1:, 0] = b
    M[1:, 1:] = A
    vrow = np.concatenate([[-1], np.arange(n_cons)]).astype(np.int32)
    vcol = np.concatenate([[-1], n_cons + np.arange(n_vars)]).astype(np.int32)
    return InitialTableau(M, vrow, vcol, np.zeros(W + H - 2, dtype=np.uint8), np.zeros(0, dtype=np.int32), [],
                          np.zeros((0, W), dtype=np.float64))


def dense_packing_lp_model(n_vars: int, n_cons: int, seed: int = 12345) -> dict:
    """Same instance as a jsLPSolver JSON model (use for moderate sizes)."""
    A, b, c = dense_packing_lp_arrays(n_vars, n_cons, seed)
    constraints = {f"resource{i}": {"max": float(b[i])} for i in range(n_cons)}
    variables = {}
    for j in range(n_vars):
        v = {"profit": float(c[j])}
        for i in range(n_cons):
            v[f"resource{i}"] = float(A[i, j])
        variables[f"activity{j}"] = v
    return {"name": f"DensePacking_{n_vars}x{n_cons}_seed{seed}", "optimize": "profit", "opType": "max",
            "constraints": constraints, "variables": variables}


def mixed_lp_model(n_vars: int, n_cons: int, seed: int = 7, density: float = 0.7, frac_min: float = 0.3) -> dict:
    """generateRandomLP-like instance (problem-generator.ts:54-108) with a share of `min` rows so
    that phase 1 has work; rhs of `min` rows is kept small so most instances stay feasible."""
    rng = np.random.default_rng(seed)
    constraints, variables = {}, {f"x{j}": {"objective": float(rng.integers(1, 101))} for j in range(n_vars)}
    for i in range(n_cons):
        is_min = rng.random() < frac_min
        for j in range(n_vars):
            if rng.random() < density:
                variables[f"x{j}"][f"c{i}"] = float(rng.integers(1, 101))
        constraints[f"c{i}"] = {"min": float(rng.integers(10, 200))} if is_min else {"max": float(rng.integers(2000, 20000))}
    return {"name": f"MixedLP_{n_vars}x{n_cons}_seed{seed}", "optimize": "objective",
            "opType": "max" if rng.random() < 0.5 else "min", "constraints": constraints, "variables": variables}


def knapsack_mip_model(n_items: int, n_cons: int, seed: int = 12345, tolerance: float = 0.0) -> dict:
    """Multi-dimensional 0/1 knapsack: a_ij ~ U{1..50} dense,
    b_i = floor(0.5 * sum_j a_ij), c_j ~ U{1..50}; every binary adds an `x <= 1` row
    (model.ts:392-395), so the root tableau is (n_cons + n_items + 1) x (n_items + 1)."""
    rng = np.random.default_rng(seed)
    A = rng.integers(1, 51, size=(n_cons, n_items))
    c = rng.integers(1, 51, size=n_items)
    constraints = {f"k{i}": {"max": float(np.floor(0.5 * A[i].sum()))} for i in range(n_cons)}
    variables, binaries = {}, {}
    for j in range(n_items):
        v = {"value": float(c[j])}
        for i in range(n_cons):
            v[f"k{i}"] = float(A[i, j])
        variables[f"item{j}"] = v
        binaries[f"item{j}"] = 1
    m = {"name": f"Knapsack_{n_items}x{n_cons}_seed{seed}", "optimize": "value", "opType": "max",
         "constraints": constraints, "variables": variables, "binaries": binaries}
    if tolerance:
        m["tolerance"] = tolerance
    return m


# ---- the reference's stress families (solver.stress.test.ts:41-215 solves them at 10..50 variables) ----------
# Same shapes and distributions as problem-generator.ts's families, drawn from numpy's stream; every model is a
# plain jsLPSolver JSON dict, so every implementation that is compared (and an outside solver) sees the same instance.

def _sparse_columns(rng, names, rows, density, lo, hi):
    """variables[name][row] = U{lo..hi} wherever a Bernoulli(density) mask is set."""
    mask = rng.random((len(names), len(rows))) < density
    coef = rng.integers(lo, hi + 1, size=mask.shape)
    return [{rows[i]: float(coef[j, i]) for i in np.nonzero(mask[j])[0]} for j in range(len(names))]


def random_lp_model(n_vars: int, n_cons: int, seed: int = 12345, density: float = 0.6) -> dict:
    """generateRandomLP (problem-generator.ts:54-108): objective and entries U{1..100}, right-hand sides U{10..1000},
    each row `max` or `min` with equal odds, sense max or min with equal odds.  Often infeasible or unbounded: the
    reference's test only asks the solver to finish with a verdict."""
    rng = np.random.default_rng(seed)
    names, rows = [f"x{j}" for j in range(n_vars)], [f"c{i}" for i in range(n_cons)]
    cols = _sparse_columns(rng, names, rows, density, 1, 100)
    obj = rng.integers(1, 101, size=n_vars)
    rhs, side = rng.integers(10, 1001, size=n_cons), rng.random(n_cons) < 0.5
    return {"name": f"RandomLP_{n_vars}x{n_cons}_seed{seed}", "optimize": "objective",
            "opType": "max" if rng.random() < 0.5 else "min",
            "constraints": {rows[i]: {"max" if side[i] else "min": float(rhs[i])} for i in range(n_cons)},
            "variables": {names[j]: {"objective": float(obj[j]), **cols[j]} for j in range(n_vars)}}


def random_mip_model(n_vars: int, n_cons: int, seed: int = 12345, density: float = 0.5, integer_fraction: float = 0.3) -> dict:
    """generateRandomMIP (problem-generator.ts:113-145): the random LP with a Bernoulli(integer_fraction) share of
    its variables declared `ints`."""
    m = random_lp_model(n_vars, n_cons, seed, density)
    pick = np.random.default_rng(seed + 1).random(n_vars) < integer_fraction
    m["name"] = m["name"].replace("RandomLP", "RandomMIP")
    if pick.any():
        m["ints"] = {f"x{j}": 1 for j in np.nonzero(pick)[0]}
    return m


def single_knapsack_model(n_items: int, seed: int = 12345) -> dict:
    """generateKnapsack (problem-generator.ts:147-186): value, weight ~ U{1..50}, one capacity row U{100..500}."""
    rng = np.random.default_rng(seed)
    vw = rng.integers(1, 51, size=(n_items, 2))
    return {"name": f"Knapsack_{n_items}_seed{seed}", "optimize": "value", "opType": "max",
            "constraints": {"capacity": {"max": float(rng.integers(100, 501))}},
            "variables": {f"item{j}": {"value": float(vw[j, 0]), "weight": float(vw[j, 1])} for j in range(n_items)},
            "binaries": {f"item{j}": 1 for j in range(n_items)}}


def set_cover_model(n_sets: int, n_elements: int, seed: int = 12345, density: float = 0.4) -> dict:
    """generateSetCover (problem-generator.ts:188-236): binary sets with cost U{1..20}, every element covered at
    least once (`min: 1` rows, so phase 1 runs at the root and after most branches); may be infeasible."""
    rng = np.random.default_rng(seed)
    cost = rng.integers(1, 21, size=n_sets)
    member = rng.random((n_sets, n_elements)) < density
    return {"name": f"SetCover_{n_sets}x{n_elements}_seed{seed}", "optimize": "cost", "opType": "min",
            "constraints": {f"element{e}": {"min": 1.0} for e in range(n_elements)},
            "variables": {f"set{s}": {"cost": float(cost[s]), **{f"element{e}": 1.0 for e in np.nonzero(member[s])[0]}}
                          for s in range(n_sets)},
            "binaries": {f"set{s}": 1 for s in range(n_sets)}}


def transportation_model(n_sources: int, n_destinations: int, seed: int = 12345) -> dict:
    """generateTransportation (problem-generator.ts:238-295): supplies U{50..200} as `max` rows, equal demands
    floor(total / destinations) as `min` rows, unit costs U{1..100}; always feasible."""
    rng = np.random.default_rng(seed)
    supply = rng.integers(50, 201, size=n_sources)
    cost = rng.integers(1, 101, size=(n_sources, n_destinations))
    demand = float(int(supply.sum()) // n_destinations)
    constraints = {f"supply{s}": {"max": float(supply[s])} for s in range(n_sources)}
    constraints.update({f"demand{d}": {"min": demand} for d in range(n_destinations)})
    return {"name": f"Transportation_{n_sources}x{n_destinations}_seed{seed}", "optimize": "cost", "opType": "min",
            "constraints": constraints,
            "variables": {f"ship_{s}_to_{d}": {"cost": float(cost[s, d]), f"supply{s}": 1.0, f"demand{d}": 1.0}
                          for s in range(n_sources) for d in range(n_destinations)}}


def resource_allocation_model(n_activities: int, n_resources: int, seed: int = 12345, density: float = 0.6) -> dict:
    """generateResourceAllocation (problem-generator.ts:297-340): profit U{1..50}, usage U{1..20} at the given
    density, capacities U{100..500}; the sparse cousin of dense_packing_lp_model."""
    rng = np.random.default_rng(seed)
    names, rows = [f"activity{a}" for a in range(n_activities)], [f"resource{r}" for r in range(n_resources)]
    profit = rng.integers(1, 51, size=n_activities)
    cols = _sparse_columns(rng, names, rows, density, 1, 20)
    cap = rng.integers(100, 501, size=n_resources)
    return {"name": f"ResourceAllocation_{n_activities}x{n_resources}_seed{seed}", "optimize": "profit", "opType": "max",
            "constraints": {rows[r]: {"max": float(cap[r])} for r in range(n_resources)},
            "variables": {names[a]: {"profit": float(profit[a]), **cols[a]} for a in range(n_activities)}}


def stress_suite(seed: int = 12345):
    """The instances of solver.stress.test.ts:41-132 (sizes and options as there), as (label, model) pairs."""
    import math
    out = []
    for n, m in ((15, 8), (30, 15), (50, 25)):
        out.append((f"random_lp_{n}x{m}", random_lp_model(n, m, seed, 0.6)))
        out.append((f"resource_allocation_{n}x{m}", resource_allocation_model(n, m, seed)))
        k = math.ceil(math.sqrt(n))
        out.append((f"transportation_{k}x{k}", transportation_model(k, k, seed)))
    for n, m in ((10, 5), (20, 10), (30, 15)):
        out.append((f"random_mip_{n}x{m}", random_mip_model(n, m, seed, 0.5, 0.3)))
        out.append((f"knapsack_{n}", single_knapsack_model(n, seed)))
    for n, m in ((10, 6), (15, 10), (20, 12)):
        out.append((f"set_cover_{n}x{m}", set_cover_model(n, m, seed, 0.4)))
    return out
