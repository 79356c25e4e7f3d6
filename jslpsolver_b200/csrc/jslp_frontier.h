// jslpsolver_b200/csrc/jslp_frontier.h -- host-only data structures of the branch-and-cut frontier
// (included by jslp_bnb.cuh; also compiled on its own by tests/test_host_cpu.py with g++).
//
// Frontier replaces BranchMinHeap (min-heap.ts:18-119): best-first on relaxedEvaluation, most recently
// pushed first on ties (min-heap.ts:43-49).  Branch / NodeEval replace Branch / BranchCut (types.ts:17-26)
// plus the cached result of a speculative evaluation; to_wire / from_wire are the 128-byte record the
// ranks all-gather per node.
#pragma once

#include <cstring>
#include <memory>
#include <utility>
#include <vector>

namespace jslp_bnb {

struct NodeEval {  // what the commit loop needs from one node LP
    bool valid = false;
    int feasible = 0, bounded = 1, optimal = 0, is_integral = 0, branch_var = -1, pivots = 0;
    double evaluation = 0, branch_value = 0;
    // useMIRCuts only (never on the wire: such models are not sharded): simplex() calls of this node that ended
    // optimal (Tableau.simplexIters) and the evaluation of the first one (Tableau.bestPossibleEval at the root)
    int n_optimal = -1;  // -1 = one solve: use `optimal`
    double first_eval = 0;
    double opt0[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // optionalObjectives[o].reducedCosts[0]
};

struct Branch {
    double relaxedEvaluation;
    std::vector<jslp_cut> cuts;
    NodeEval ev;  // cached speculative result
};

// Total order of min-heap.ts:43-49: lower relaxedEvaluation first, then higher seq (LIFO).
struct Frontier {
    struct Entry {
        std::unique_ptr<Branch> b;
        long seq;
    };
    std::vector<Entry> h;
    long seqCounter = 0;
    static bool before(const Entry &a, const Entry &b) {
        if (a.b->relaxedEvaluation != b.b->relaxedEvaluation) return a.b->relaxedEvaluation < b.b->relaxedEvaluation;
        return a.seq > b.seq;
    }
    bool empty() const { return h.empty(); }
    void push_entry(Entry e) {
        h.push_back(std::move(e));
        size_t i = h.size() - 1;
        while (i > 0) {
            const size_t p = (i - 1) / 2;
            if (!before(h[i], h[p])) break;
            std::swap(h[i], h[p]);
            i = p;
        }
    }
    void push(std::unique_ptr<Branch> br) { push_entry(Entry{std::move(br), seqCounter++}); }
    Entry pop_entry() {
        Entry top = std::move(h[0]);
        if (h.size() > 1) h[0] = std::move(h.back());
        h.pop_back();
        size_t i = 0;
        const size_t n = h.size();
        for (;;) {
            size_t l = 2 * i + 1, r = l + 1, m = i;
            if (l < n && before(h[l], h[m])) m = l;
            if (r < n && before(h[r], h[m])) m = r;
            if (m == i) break;
            std::swap(h[i], h[m]);
            i = m;
        }
        return top;
    }
};

static const int WIRE_DOUBLES = 16;  // all-gather record per node (128 bytes)
static void to_wire(const NodeEval &e, double *w) {
    w[0] = e.valid; w[1] = e.feasible; w[2] = e.bounded; w[3] = e.optimal; w[4] = e.is_integral;
    w[5] = e.branch_var; w[6] = e.pivots; w[7] = e.evaluation;
    memcpy(&w[7], &e.evaluation, 8);
    memcpy(&w[8], &e.branch_value, 8);
    for (int o = 0; o < 7; o++) memcpy(&w[9 + o], &e.opt0[o], 8);
}
static void from_wire(NodeEval &e, const double *w) {
    e.valid = w[0] != 0; e.feasible = (int)w[1]; e.bounded = (int)w[2]; e.optimal = (int)w[3];
    e.is_integral = (int)w[4]; e.branch_var = (int)w[5]; e.pivots = (int)w[6];
    memcpy(&e.evaluation, &w[7], 8);
    memcpy(&e.branch_value, &w[8], 8);
    for (int o = 0; o < 7; o++) memcpy(&e.opt0[o], &w[9 + o], 8);
}

}  // namespace jslp_bnb
