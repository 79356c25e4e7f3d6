// jslpsolver_b200/csrc/jslp_bnb_enhanced.cuh -- the reference's opt-in "enhanced" branch-and-cut service on the
// device tableau (included by jslp_api.cu).
//
// Replaces createEnhancedBranchAndCutService (enhanced-branch-and-cut.ts:54-437), which main.ts:62-83 selects when a
// model sets options.nodeSelection or options.branching: depth-first / hybrid node selection next to best-first,
// pseudocost and "strong" (pseudocost-estimated) branching next to most-fractional, MIR rounds capped at three.
// The pseudocost table is updated from every evaluated node IN ORDER and decides later branching variables, so this
// service is sequential by construction: one node LP at a time on the single-tableau HBM path (restore, cut rows,
// fused pivot steps), the policy on the host.  Same kernels as the default service, different frontier.
//
// The same loop also replaces createIncrementalBranchAndCutService (incremental-branch-and-cut.ts:128-499,
// options.useIncremental, "experimental" in the reference): depth-first children restore a CHECKPOINT of their parent's
// solved tableau (device-to-device copy, at most maxCheckpoints = 50 per call) and add only their one new cut instead
// of restoring the root and re-adding every cut; pseudocosts are fed by `newCut`; "strong" means pseudocost.
#pragma once

#include <chrono>
#include <memory>

#include "jslp_frontier.h"

struct PseudoCost {
    double upSum = 0, downSum = 0;
    long upCount = 0, downCount = 0;
};

// enhanced-branch-and-cut.ts:94-108
static double pc_score(const PseudoCost &d, double fraction) {
    const double upPseudo = d.upCount > 0 ? d.upSum / d.upCount : 1;
    const double downPseudo = d.downCount > 0 ? d.downSum / d.downCount : 1;
    const double upEstimate = upPseudo * (1 - fraction);
    const double downEstimate = downPseudo * fraction;
    auto max6 = [](double x) { return x > 1e-6 ? x : (x != x ? x : 1e-6); };  // Math.max(x, 1e-6)
    return max6(upEstimate) * max6(downEstimate);
}

// StateCheckpoint (incremental-branch-and-cut.ts:28-107): the solved parent tableau, kept on the device
struct DevCheckpoint {
    double *M = nullptr;
    int *vrow = nullptr, *vcol = nullptr;
    int H = 0, W = 0, stride = 0, nVars = 0, lastElementIndex = 0, feasible = 1;
    double evaluation = 0;
    ~DevCheckpoint() { cudaFree(M); cudaFree(vrow); cudaFree(vcol); }
};
static int checkpoint_create(jslp_tab *t, std::shared_ptr<DevCheckpoint> *out) {
    std::shared_ptr<DevCheckpoint> c(new DevCheckpoint());
    cudaStream_t s = t->ctx->stream;
    c->H = t->H; c->W = t->W; c->stride = t->stride; c->nVars = t->nVars; c->lastElementIndex = t->lastElementIndex;
    c->feasible = t->feasible; c->evaluation = t->evaluation;
    CK(cudaMalloc(&c->M, sizeof(double) * (size_t)t->H * t->stride));
    CK(cudaMalloc(&c->vrow, sizeof(int) * (size_t)t->H));
    CK(cudaMalloc(&c->vcol, sizeof(int) * (size_t)t->W));
    CK(cudaMemcpyAsync(c->M, t->hd.M, sizeof(double) * (size_t)t->H * t->stride, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(c->vrow, t->hd.vrow, sizeof(int) * (size_t)t->H, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(c->vcol, t->hd.vcol, sizeof(int) * (size_t)t->W, cudaMemcpyDeviceToDevice, s));
    *out = c;
    return JSLP_OK;
}
static int checkpoint_restore(jslp_tab *t, const DevCheckpoint &c) {
    if (c.stride != t->stride || c.W != t->W) return fail(JSLP_E_INVALID, "checkpoint of another layout");
    int rc = grow_rows(t, c.H);
    if (rc) return rc;
    cudaStream_t s = t->ctx->stream;
    CK(cudaMemcpyAsync(t->hd.M, c.M, sizeof(double) * (size_t)c.H * t->stride, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(t->hd.vrow, c.vrow, sizeof(int) * (size_t)c.H, cudaMemcpyDeviceToDevice, s));
    CK(cudaMemcpyAsync(t->hd.vcol, c.vcol, sizeof(int) * (size_t)c.W, cudaMemcpyDeviceToDevice, s));
    const int grid_before = step_grid(t);
    t->H = c.H; t->nVars = c.nVars; t->lastElementIndex = c.lastElementIndex;
    t->evaluation = c.evaluation; t->feasible = c.feasible;
    rc = push_desc(t);
    if (rc) return rc;
    if (step_grid(t) != grid_before) drop_graphs(t);
    return JSLP_OK;
}
struct IncBranchInfo {  // IncrementalBranch (incremental-branch-and-cut.ts:49-53), keyed by the Branch it extends
    std::shared_ptr<DevCheckpoint> parent;
    jslp_cut newCut{0, -1, 0.0};
    bool hasNewCut = false;
};

struct FracCand {
    int index;
    double value, fraction;
};

// enhanced-branch-and-cut.ts:110-195 over the downloaded right-hand-side column; branching 1 = most-fractional,
// 2 = pseudocost, 3 = strong.  Returns the variable index or -1.
static int enh_select_variable(jslp_tab *t, std::unordered_map<int, PseudoCost> &pc, int branching, int strong_candidates, double *value) {
    std::vector<double> rhs((size_t)t->H);
    std::vector<int> vrow((size_t)t->H);
    if (jslp_download(t, nullptr, rhs.data(), nullptr, vrow.data(), nullptr, nullptr, nullptr, nullptr)) return -2;
    std::vector<double> val_at((size_t)t->n_int, 0.0);
    std::vector<int> var_at((size_t)t->n_int, -1);
    for (int r = 1; r < t->H; r++) {  // integerVars order == position order
        const int v = vrow[r];
        if (v >= 0 && v < (int)t->h_intpos.size() && t->h_intpos[v] >= 0) { var_at[t->h_intpos[v]] = v; val_at[t->h_intpos[v]] = rhs[r]; }
    }
    std::vector<FracCand> cand;
    for (int p = 0; p < t->n_int; p++) {
        if (var_at[p] < 0) continue;
        const double fraction = std::fabs(val_at[p] - js_round_h(val_at[p]));
        if (fraction > t->precision) cand.push_back(FracCand{var_at[p], val_at[p], fraction});
    }
    if (cand.empty()) return -1;
    auto by_fraction_desc = [](const FracCand &a, const FracCand &b) { return a.fraction > b.fraction; };
    FracCand best = cand[0];
    if (branching == 1) {
        std::stable_sort(cand.begin(), cand.end(), by_fraction_desc);
        best = cand[0];
    } else if (branching == 2) {
        double bestScore = -INFINITY;
        for (const FracCand &c : cand) {
            const double score = pc_score(pc[c.index], c.fraction);
            if (score > bestScore) { bestScore = score; best = c; }
        }
    } else if (branching == 3) {
        std::stable_sort(cand.begin(), cand.end(), by_fraction_desc);
        if ((int)cand.size() > strong_candidates) cand.resize((size_t)strong_candidates);
        double bestScore = -INFINITY;
        best = cand[0];
        for (const FracCand &c : cand) {
            const PseudoCost &d = pc[c.index];
            const double score = (d.upCount >= 2 && d.downCount >= 2) ? pc_score(d, c.fraction) : c.fraction * (1 - c.fraction);
            if (score > bestScore) { bestScore = score; best = c; }
        }
    }
    *value = best.value;
    return best.index;
}

static int enh_mir_rounds(jslp_tab *t, int check_cycles, int *pivots);

// applyCuts of the enhanced service (enhanced-branch-and-cut.ts:197-221)
static int enh_apply_cuts(jslp_tab *t, const std::vector<jslp_cut> &cuts, int check_cycles, int *pivots) {
    jslp_lp_status st;
    int rc = jslp_restore(t);
    if (rc) return rc;
    rc = jslp_add_cuts(t, cuts.data(), (int)cuts.size());
    if (rc) return rc;
    rc = run_lp(t, 0, check_cycles, &st, false);
    if (rc) return rc;
    *pivots = st.phase1_pivots + st.phase2_pivots;
    return enh_mir_rounds(t, check_cycles, pivots);
}

// the MIR rounds shared by applyCuts / applyIncrementalCuts (enhanced :204-220, incremental :263-280)
static int enh_mir_rounds(jslp_tab *t, int check_cycles, int *pivots) {
    if (!(t->use_mir && t->feasible)) return JSLP_OK;
    jslp_lp_status st;
    bool improved = true;
    int mirIterations = 0, rc;
    while (improved && mirIterations < 3) {
        double before = 0, after = 0;
        if ((rc = jslp_fractional_volume(t, 1, &before))) return rc;
        if ((rc = mir_cuts(t, -1, 0, nullptr))) return rc;
        if ((rc = run_lp(t, 0, check_cycles, &st, false))) return rc;
        *pivots += st.phase1_pivots + st.phase2_pivots;
        if ((rc = jslp_fractional_volume(t, 1, &after))) return rc;
        mirIterations++;
        if (after >= 0.9 * before) improved = false;
    }
    return JSLP_OK;
}

// applyIncrementalCuts fast path (incremental-branch-and-cut.ts:253-258): parent checkpoint + the one new cut
static int inc_apply_cuts(jslp_tab *t, const DevCheckpoint &parent, const jslp_cut &newCut, int check_cycles, int *pivots) {
    jslp_lp_status st;
    int rc = checkpoint_restore(t, parent);
    if (rc) return rc;
    if ((rc = jslp_add_cuts(t, &newCut, 1))) return rc;
    if ((rc = run_lp(t, 0, check_cycles, &st, false))) return rc;
    *pivots = st.phase1_pivots + st.phase2_pivots;
    return enh_mir_rounds(t, check_cycles, pivots);
}

static int bnb_enhanced(jslp_tab *t, const jslp_bnb_opts *opts, jslp_bnb_status *out, jslp_cut *best_cuts, int best_cuts_cap) {
    using namespace jslp_bnb;
    const bool incremental = opts->service == 2;
    const int maxCheckpoints = 50;  // incremental-branch-and-cut.ts:135
    int checkpointCount = 0;
    std::unordered_map<const Branch *, IncBranchInfo> inc;
    jslp_ctx *ctx = t->ctx;
    CK(cudaSetDevice(ctx->device));
    const int64_t launches0 = ctx->launches;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    const auto t_start = std::chrono::steady_clock::now();
    auto time_up = [&] {
        return opts->timeout_ms > 0 &&
               std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count() >= opts->timeout_ms;
    };
    const int nodeSelection = opts->node_selection > 0 ? opts->node_selection : 3;  // default "hybrid"
    int branching = opts->branching > 0 ? opts->branching : 2;                     // default "pseudocost"
    if (incremental && branching == 3) branching = 2;                              // incremental :203-219: no strong variant
    const int strongCandidates = opts->strong_candidates > 0 ? opts->strong_candidates : 5;
    const int check_cycles = opts->check_cycles;
    const double tolerance = opts->tolerance;

    Frontier branches;                                // BranchMinHeap
    std::vector<std::unique_ptr<Branch>> stack;       // depthFirstStack
    std::unordered_map<int, PseudoCost> pseudoCosts;  // one service instance per Solve (main.ts:62-83)
    int iterations = 0, solutionsFound = 0;
    bool toleranceFlag = true, timed_out = false, early_return = false;
    double bestEvaluation = INFINITY;
    std::unique_ptr<Branch> bestBranch;
    std::vector<double> bestOpt((size_t)t->nOpt, INFINITY);
    bool useDepthFirst = nodeSelection == 2 || nodeSelection == 3;
    int64_t pivots = 0, nodes = 0;
    t->node_log.clear();
    t->solutions.clear();
    t->saved.valid = false;
    t->isIntegralFlag = 0;
    {
        std::unique_ptr<Branch> root(new Branch{-INFINITY, {}, NodeEval()});
        if (useDepthFirst) stack.push_back(std::move(root)); else branches.push(std::move(root));
    }
    while ((useDepthFirst ? !stack.empty() : !branches.empty()) && toleranceFlag) {
        if (opts->max_nodes > 0 && iterations >= opts->max_nodes) break;
        if (time_up()) { timed_out = true; break; }
        const double acceptableThreshold = opts->is_minimization ? t->bestPossibleEval * (1 + tolerance) : t->bestPossibleEval * (1 - tolerance);
        if (tolerance > 0 && bestEvaluation < acceptableThreshold) toleranceFlag = false;
        std::unique_ptr<Branch> active;
        if (useDepthFirst && !stack.empty()) { active = std::move(stack.back()); stack.pop_back(); }
        else if (!branches.empty()) active = std::move(branches.pop_entry().b);
        else break;
        IncBranchInfo info;
        if (incremental) {  // taken out of the table whatever happens next: the Branch's address may be reused
            auto it = inc.find(active.get());
            if (it != inc.end()) { info = it->second; inc.erase(it); }
        }
        if (active->relaxedEvaluation > bestEvaluation) continue;
        const double parentEval = t->evaluation;
        int node_pivots = 0;
        int rc = (incremental && info.parent && info.hasNewCut) ? inc_apply_cuts(t, *info.parent, info.newCut, check_cycles, &node_pivots)
                                                                : enh_apply_cuts(t, active->cuts, check_cycles, &node_pivots);
        if (rc) return rc;
        iterations++; nodes++; pivots += node_pivots;
        NodeLogEntry nl;
        nl.v[0] = iterations; nl.v[1] = (double)active->cuts.size(); nl.v[2] = t->feasible;
        nl.v[3] = t->evaluation; nl.v[4] = -1; nl.v[5] = -1; nl.v[6] = 0; nl.v[7] = node_pivots;
        if (!t->feasible) { t->node_log.push_back(nl); continue; }
        const double evaluation = t->evaluation;
        if (evaluation > bestEvaluation) { t->node_log.push_back(nl); continue; }
        if ((incremental ? info.hasNewCut : !active->cuts.empty()) && parentEval != 0) {  // enhanced :281-294, incremental :353-362
            const jslp_cut lastCut = incremental ? info.newCut : active->cuts.back();
            const double improvement = std::fabs(evaluation - parentEval);
            const double fraction = 0.5;
            PseudoCost &d = pseudoCosts[lastCut.var_index];
            const bool up = lastCut.type == 0;
            const double normalizedImprovement = improvement / (up ? 1 - fraction : fraction);
            if (up) { d.upSum += normalizedImprovement; d.upCount++; } else { d.downSum += normalizedImprovement; d.downCount++; }
        }
        if (evaluation == bestEvaluation) {
            bool worse = true;
            if (t->nOpt > 0) {
                std::vector<double> opt((size_t)t->nOpt * t->W);
                rc = jslp_download(t, nullptr, nullptr, nullptr, nullptr, nullptr, opt.data(), nullptr, nullptr);
                if (rc) return rc;
                for (int o = 0; o < t->nOpt; o++) {
                    const double v = opt[(size_t)o * t->W];
                    if (v > bestOpt[o]) break;
                    if (v < bestOpt[o]) { worse = false; break; }
                }
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
            solutionsFound++;
            if (iterations == 1) { early_return = true; break; }
            bestEvaluation = evaluation;
            if (t->nOpt > 0) {
                std::vector<double> opt((size_t)t->nOpt * t->W);
                rc = jslp_download(t, nullptr, nullptr, nullptr, nullptr, nullptr, opt.data(), nullptr, nullptr);
                if (rc) return rc;
                for (int o = 0; o < t->nOpt; o++) bestOpt[o] = opt[(size_t)o * t->W];
            }
            if (opts->keep_solutions) {  // the tableau holds this node: store it as it stands
                jslp_tab::StoredSolution sol;
                sol.evaluation = t->evaluation;
                sol.vrow.resize((size_t)t->H);
                sol.rhs.resize((size_t)t->H);
                rc = jslp_download(t, nullptr, sol.rhs.data(), nullptr, sol.vrow.data(), nullptr, nullptr, nullptr, nullptr);
                if (rc) return rc;
                t->solutions.push_back(std::move(sol));
            }
            bestBranch = std::move(active);
            if (nodeSelection == 3 && solutionsFound >= 1) {  // hybrid: best-first from the first incumbent on
                useDepthFirst = false;
                while (!stack.empty()) { branches.push(std::move(stack.back())); stack.pop_back(); }
            }
        } else {
            nl.v[4] = 0;
            if (iterations == 1) {
                rc = jslp_save(t);
                if (rc) return rc;
            }
            double varValue = 0;
            const int varIndex = enh_select_variable(t, pseudoCosts, branching, strongCandidates, &varValue);
            if (varIndex == -2) return fail(JSLP_E_CUDA, "enhanced branch and cut: read-back failed");
            if (varIndex < 0) { t->node_log.push_back(nl); continue; }
            nl.v[5] = varIndex; nl.v[6] = varValue;
            t->node_log.push_back(nl);
            std::unique_ptr<Branch> high(new Branch{evaluation, {}, NodeEval()}), low(new Branch{evaluation, {}, NodeEval()});
            for (const jslp_cut &cut : active->cuts) {
                if (cut.var_index == varIndex) {
                    if (cut.type == 0) low->cuts.push_back(cut); else high->cuts.push_back(cut);
                } else {
                    high->cuts.push_back(cut);
                    low->cuts.push_back(cut);
                }
            }
            high->cuts.push_back(jslp_cut{0, varIndex, std::ceil(varValue)});
            low->cuts.push_back(jslp_cut{1, varIndex, std::floor(varValue)});
            if (incremental && useDepthFirst) {  // incremental :435-438,476-483: children carry the parent checkpoint + their cut
                std::shared_ptr<DevCheckpoint> cp;
                if (checkpointCount < maxCheckpoints) {
                    rc = checkpoint_create(t, &cp);
                    if (rc) return rc;
                    checkpointCount++;
                }
                inc[low.get()] = IncBranchInfo{cp, low->cuts.back(), true};
                inc[high.get()] = IncBranchInfo{cp, high->cuts.back(), true};
            }
            if (useDepthFirst) { stack.push_back(std::move(low)); stack.push_back(std::move(high)); }  // 'up' branch first
            else { branches.push(std::move(high)); branches.push(std::move(low)); }
        }
    }
    int n_best = 0;
    if (!early_return && bestBranch) {
        int p = 0;
        int rc = enh_apply_cuts(t, bestBranch->cuts, check_cycles, &p);
        if (rc) return rc;
        pivots += p; nodes++;
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
        out->timed_out = timed_out ? 1 : 0; out->n_solutions = (int)t->solutions.size();
    }
    return JSLP_OK;
}
