// The product's cycle detectors (jslpsolver_b200/csrc/jslp_cycles.h: the suffix-square scan cycle_hit and the
// hashed CycleHist) against a literal restatement of the reference's checkForCycles (simplex.ts:415-440), under
// the reference's calling discipline: the detector runs after every push and the solve stops at the first hit
// (simplex.ts:27-36,102-112).  Built and run by tests/test_host_cpu.py with g++; prints CYCLES OK.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "jslp_cycles.h"

// simplex.ts:415-440, pair (leaving, entering) packed into one 64-bit key
static bool literal(const std::vector<long long> &v, int *start, int *len) {
    const long n = (long)v.size();
    for (long e1 = 0; e1 < n - 1; e1++)
        for (long e2 = e1 + 1; e2 < n; e2++) {
            if (v[e1] != v[e2]) continue;
            if (e2 - e1 > n - e2) break;
            bool found = true;
            for (long i = 1; i < e2 - e1; i++)
                if (v[e1 + i] != v[e2 + i]) { found = false; break; }
            if (found) { *start = (int)e1; *len = (int)(e2 - e1); return true; }
        }
    return false;
}

static unsigned long long rng_state = 88172645463325252ull;
static unsigned int rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (unsigned int)(rng_state >> 11); }

int main() {
    long hits = 0, pushes = 0, long_hits = 0;
    for (int trial = 0; trial < 4000; trial++) {
        const int alphabet = 2 + (int)(rnd() % 9);           // few distinct pairs: repeats are frequent
        const int period = 1 + (int)(rnd() % 7), lead = (int)(rnd() % 12);
        const bool periodic = (rnd() & 3) == 0;                // some sequences fall into an exact cycle
        const bool no_stutter = (rnd() & 1) == 0;              // half of the trials never repeat the previous pair,
                                                               // so their first hit is a block of length >= 2
        std::vector<long long> h;
        CycleHist hist;
        for (int k = 0; k < 200; k++) {
            long long v;
            if (periodic && k >= lead) v = 1000 + ((k - lead) % period) * 7 + (long long)(((k - lead) % period) % 3) * (1ll << 32);
            else {
                do v = (long long)(rnd() % alphabet) | ((long long)(rnd() % 2) << 32);
                while (no_stutter && !h.empty() && v == h.back());
            }
            h.push_back(v);
            pushes++;
            int s0 = -1, l0 = -1, s1 = -1, l1 = -1, s2 = -1, l2 = -1;
            const bool r0 = literal(h, &s0, &l0);
            const bool r1 = cycle_hit(h, &s1, &l1);
            const bool r2 = hist.push_and_check(v, &s2, &l2);
            if (r0 != r1 || r0 != r2 || (r0 && (s0 != s1 || l0 != l1 || s0 != s2 || l0 != l2))) {
                std::printf("FAILED trial %d push %d: literal %d (%d,%d) suffix %d (%d,%d) hashed %d (%d,%d)\n", trial, k, r0, s0, l0,
                            r1, s1, l1, r2, s2, l2);
                return 1;
            }
            if (r0) { hits++; if (l0 >= 3) long_hits++; break; }   // the reference returns at the first hit
        }
    }
    if (hits < 1000 || long_hits < 200) {
        std::printf("FAILED: %ld hits, %ld of length >= 3: the generator does not exercise the detectors\n", hits, long_hits);
        return 1;
    }
    std::printf("CYCLES OK (%ld hits, %ld of block length >= 3, %ld pushes)\n", hits, long_hits, pushes);
    return 0;
}
