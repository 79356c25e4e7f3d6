// jslpsolver_b200/csrc/jslp_hostmath.h -- host-only JS number semantics used when results are rounded
// (included by jslp_api.cu; also compiled on its own by tests/test_host_cpu.py with g++).
#pragma once

#include <cmath>

// JS Math.round: nearest integer, ties toward +infinity (NaN and infinities pass through).
static inline double js_round_h(double x) {
    if (!(x == x) || std::isinf(x)) return x;
    const double f = std::floor(x);
    return (x - f >= 0.5) ? f + 1.0 : f;
}

// Tableau.setEvaluation (tableau.ts:420-430): round matrix[0][0] to the model precision, Number.EPSILON added
// before rounding.
static inline double jslp_round_evaluation(double raw, double precision) {
    const double roundingCoeff = js_round_h(1 / precision);
    return js_round_h((2.220446049250313e-16 + raw) * roundingCoeff) / roundingCoeff;
}
