// Prints js_round_h(x) and jslp_round_evaluation(x, precision) of the product (jslpsolver_b200/csrc/jslp_hostmath.h)
// for every "x precision" pair on stdin, as raw 64-bit patterns; tests/test_host_cpu.py compares them with the
// oracle's Python restatement of Math.round / Tableau.setEvaluation.
#include <cstdio>
#include <cstring>
#include "jslp_hostmath.h"

int main() {
    unsigned long long xb, pb;
    while (std::scanf("%llx %llx", &xb, &pb) == 2) {
        double x, p;
        std::memcpy(&x, &xb, 8);
        std::memcpy(&p, &pb, 8);
        const double r = js_round_h(x), e = jslp_round_evaluation(x, p);
        unsigned long long rb, eb;
        std::memcpy(&rb, &r, 8);
        std::memcpy(&eb, &e, 8);
        std::printf("%llx %llx\n", rb, eb);
    }
    return 0;
}
