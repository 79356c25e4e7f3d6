// jslpsolver_b200/csrc/jslp_bnb.cuh -- branch-and-cut frontier manager (included by jslp_api.cu).
//
// Replaces BranchAndCutService.branchAndCut (branch-and-cut.ts:54-199) and BranchMinHeap
// (min-heap.ts).  The node order is the reference's: best-first on relaxedEvaluation, most
// recently pushed first on ties.  Each node LP is a pure function of (root snapshot, cut list)
// (SURVEY.md 3.8), evaluated on the device by restore -> add cuts -> simplex; the frontier and the
// commit decisions stay on the host because they are a few hundred bytes per node.
#pragma once

#include <memory>

namespace jslp_bnb {

struct Branch {
    double relaxedEvaluation;
    std::vector<jslp_cut> cuts;
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
    void push(std::unique_ptr<Branch> br) {
        h.push_back(Entry{std::move(br), seqCounter++});
        size_t i = h.size() - 1;
        while (i > 0) {
            const size_t p = (i - 1) / 2;
            if (!before(h[i], h[p])) break;
            std::swap(h[i], h[p]);
            i = p;
        }
    }
    std::unique_ptr<Branch> pop() {
        std::unique_ptr<Branch> top = std::move(h[0].b);
        h[0] = std::move(h.back());
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

}  // namespace jslp_bnb

extern "C" int jslp_branch_and_cut(jslp_tab *t, const jslp_bnb_opts *opts, jslp_bnb_status *out,
                                   jslp_cut *best_cuts, int best_cuts_cap) {
    using namespace jslp_bnb;
    if (!t || !opts) return fail(JSLP_E_INVALID, "NULL argument");
    if (t->n_int <= 0) return fail(JSLP_E_INVALID, "branch_and_cut needs integer variables (upload int_var_indices)");
    jslp_ctx *ctx = t->ctx;
    CK(cudaSetDevice(ctx->device));
    const int64_t launches0 = ctx->launches;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));

    Frontier branches;
    int iterations = 0;
    const double tolerance = opts->tolerance;
    bool toleranceFlag = true;
    double bestEvaluation = INFINITY;
    std::unique_ptr<Branch> bestBranch;
    std::vector<double> bestOpt((size_t)t->nOpt, INFINITY);
    std::vector<double> optNow((size_t)std::max(1, t->nOpt) * t->W);
    int64_t pivots = 0, nodes = 0;
    t->node_log.clear();
    t->saved.valid = false;
    bool early_return = false;

    branches.push(std::unique_ptr<Branch>(new Branch{-INFINITY, {}}));
    while (!branches.empty() && toleranceFlag) {
        if (opts->max_nodes > 0 && iterations >= opts->max_nodes) break;
        const double acceptableThreshold =
            opts->is_minimization ? t->bestPossibleEval * (1 + tolerance) : t->bestPossibleEval * (1 - tolerance);
        if (tolerance > 0 && bestEvaluation < acceptableThreshold) toleranceFlag = false;

        std::unique_ptr<Branch> active = branches.pop();
        if (active->relaxedEvaluation > bestEvaluation) continue;

        jslp_lp_status st;
        int rc = jslp_restore(t);
        if (rc) return rc;
        rc = jslp_add_cuts(t, active->cuts.data(), (int)active->cuts.size());
        if (rc) return rc;
        rc = run_lp(t, 0, opts->check_cycles, &st, false);
        if (rc) return rc;
        iterations++;
        nodes++;
        pivots += st.phase1_pivots + st.phase2_pivots;

        NodeLogEntry nl;
        nl.v[0] = iterations; nl.v[1] = (double)active->cuts.size(); nl.v[2] = t->feasible;
        nl.v[3] = t->evaluation; nl.v[4] = -1; nl.v[5] = -1; nl.v[6] = 0;
        nl.v[7] = st.phase1_pivots + st.phase2_pivots;

        if (!t->feasible) { t->node_log.push_back(nl); continue; }
        const double evaluation = t->evaluation;
        if (evaluation > bestEvaluation) { t->node_log.push_back(nl); continue; }

        if (t->nOpt > 0) {
            rc = jslp_download(t, nullptr, nullptr, nullptr, nullptr, nullptr, optNow.data(), nullptr, nullptr);
            if (rc) return rc;
        }
        if (evaluation == bestEvaluation) {  // branch-and-cut.ts:107-127
            bool worse = true;
            for (int o = 0; o < t->nOpt; o++) {
                const double v = optNow[(size_t)o * t->W];
                if (v > bestOpt[o]) break;
                if (v < bestOpt[o]) { worse = false; break; }
            }
            if (worse) { t->node_log.push_back(nl); continue; }
        }

        MipOut mo;
        rc = mip_scan(t, &mo);
        if (rc) return rc;
        if (mo.is_integral) {
            nl.v[4] = 1;
            t->node_log.push_back(nl);
            t->isIntegralFlag = 1;
            if (iterations == 1) { early_return = true; break; }
            bestEvaluation = evaluation;
            for (int o = 0; o < t->nOpt; o++) bestOpt[o] = optNow[(size_t)o * t->W];
            bestBranch = std::move(active);
        } else {
            if (iterations == 1) {
                rc = jslp_save(t);
                if (rc) return rc;
            }
            const int varIndex = mo.var_index;
            const double value = mo.value;
            nl.v[4] = 0; nl.v[5] = varIndex; nl.v[6] = value;
            t->node_log.push_back(nl);
            std::unique_ptr<Branch> high(new Branch{evaluation, {}}), low(new Branch{evaluation, {}});
            for (const jslp_cut &cut : active->cuts) {  // branch-and-cut.ts:166-179
                if (cut.var_index == varIndex) {
                    if (cut.type == 0) low->cuts.push_back(cut); else high->cuts.push_back(cut);
                } else {
                    high->cuts.push_back(cut);
                    low->cuts.push_back(cut);
                }
            }
            high->cuts.push_back(jslp_cut{0, varIndex, std::ceil(value)});
            low->cuts.push_back(jslp_cut{1, varIndex, std::floor(value)});
            branches.push(std::move(high));
            branches.push(std::move(low));
        }
    }

    int n_best = 0;
    if (!early_return && bestBranch) {  // branch-and-cut.ts:195-197
        jslp_lp_status st;
        int rc = jslp_restore(t);
        if (rc) return rc;
        rc = jslp_add_cuts(t, bestBranch->cuts.data(), (int)bestBranch->cuts.size());
        if (rc) return rc;
        rc = run_lp(t, 0, opts->check_cycles, &st, false);
        if (rc) return rc;
        nodes++;
        pivots += st.phase1_pivots + st.phase2_pivots;
        n_best = (int)bestBranch->cuts.size();
        if (best_cuts)
            for (int i = 0; i < n_best && i < best_cuts_cap; i++) best_cuts[i] = bestBranch->cuts[i];
    }
    t->bncIterations = iterations;

    float ms = 0.f;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev1));
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (out) {
        memset(out, 0, sizeof(*out));
        out->feasible = t->feasible; out->bounded = t->bounded; out->is_integral = t->isIntegralFlag;
        out->iterations = iterations; out->n_best_cuts = n_best; out->rounds = iterations;
        out->nodes_evaluated = nodes; out->pivots = pivots; out->evaluation = t->evaluation;
        out->best_possible_eval = t->bestPossibleEval; out->gpu_ms = ms;
        out->kernel_launches = ctx->launches - launches0;
    }
    return JSLP_OK;
}

extern "C" int jslp_bnb_node_log(jslp_tab *t, double *entries, int64_t cap, int64_t *n) {
    if (!t || !n) return fail(JSLP_E_INVALID, "NULL argument");
    const int64_t m = std::min<int64_t>((int64_t)t->node_log.size(), cap);
    for (int64_t i = 0; i < m && entries; i++) memcpy(entries + 8 * i, t->node_log[i].v, sizeof(double) * 8);
    *n = (int64_t)t->node_log.size();
    return JSLP_OK;
}
