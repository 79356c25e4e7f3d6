// jslpsolver_b200/csrc/jslp_cycles.h -- host-only cycle detection over the drained pivot log
// (included by jslp_api.cu; also compiled on its own by tests/test_host_cpu.py with g++).
#pragma once

#include <unordered_map>
#include <vector>

// checkForCycles (simplex.ts:415-440) under its calling discipline (called after every push,
// stops at the first hit): a repeated block must end at the newest element, so only suffix
// squares are examined; the literal scan reports the smallest start = the longest block.
static bool cycle_hit(const std::vector<long long> &h, int *start, int *len) {
    const long n = (long)h.size();
    for (long L = n / 2; L >= 1; L--) {
        const long e1 = n - 2 * L, e2 = n - L;
        if (h[e1] != h[e2]) continue;
        bool eq = true;
        for (long i = 1; i < L; i++)
            if (h[e1 + i] != h[e2 + i]) { eq = false; break; }
        if (eq) { *start = (int)e1; *len = (int)L; return true; }
    }
    return false;
}

// The same test in O(occurrences of the newest pair) per push instead of O(n): a block of length L
// that repeats up to the newest element needs h[n-1-L] == h[n-1], so only earlier positions of the
// newest pair are candidate block ends (largest L first, as the literal scan reports it).
struct CycleHist {
    std::vector<long long> h;
    std::unordered_map<long long, std::vector<int>> pos;
    bool push_and_check(long long v, int *start, int *len) {
        h.push_back(v);
        const long n = (long)h.size();
        std::vector<int> &p = pos[v];
        bool hit = false;
        for (size_t k = 0; k < p.size() && !hit; k++) {
            const long L = (n - 1) - p[k];
            if (2 * L > n) continue;
            const long e1 = n - 2 * L, e2 = n - L;
            bool eq = true;
            for (long i = 0; i < L; i++)
                if (h[e1 + i] != h[e2 + i]) { eq = false; break; }
            if (eq) { *start = (int)e1; *len = (int)L; hit = true; }
        }
        p.push_back((int)(n - 1));
        return hit;
    }
};
