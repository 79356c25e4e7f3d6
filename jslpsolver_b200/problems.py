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
