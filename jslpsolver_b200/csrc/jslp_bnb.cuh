// jslpsolver_b200/csrc/jslp_bnb.cuh -- branch-and-cut frontier manager (included by jslp_api.cu).
//
// Replaces BranchAndCutService.branchAndCut (branch-and-cut.ts:54-199) and BranchMinHeap
// (min-heap.ts).  The commit order is the reference's: best-first on relaxedEvaluation, most
// recently pushed first on ties.  A node's LP result is a pure function of (root snapshot, cut
// list), so the frontier is evaluated speculatively: every round takes the next K
// open nodes in exact pop order, evaluates them together on the device (one CTA per node when the
// tableau fits shared memory, else one after the other on the HBM path; sharded round-robin over
// ranks with an all-gather of the 128-byte summaries when n_ranks > 1), then COMMITS results
// sequentially in pop order exactly as the reference loop would.  A popped node without a cached
// result ends the round.  Speculation that is never popped is wasted work, never wrong work.
#pragma once

#include <chrono>
#include <map>
#include <memory>

#include "jslp_frontier.h"

// ---- node evaluation back-ends ---------------------------------------------------------------
// HBM path: applyCuts (branch-and-cut.ts:33-52) in place, exactly as the reference does it.
static int eval_node_streaming(jslp_tab *t, const jslp_bnb::Branch &b, int check_cycles, jslp_bnb::NodeEval &ev) {
    jslp_lp_status st;
    static const bool dbg = getenv("JSLP_DEBUG") != nullptr;
    auto tnow = [] { return std::chrono::steady_clock::now(); };
    auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b2) {
        return std::chrono::duration<double, std::micro>(b2 - a).count();
    };
    const auto t_0 = tnow();
    int rc = jslp_restore(t);
    if (rc) return rc;
    if (dbg) cudaStreamSynchronize(t->ctx->stream);
    const auto t_1 = tnow();
    rc = jslp_add_cuts(t, b.cuts.data(), (int)b.cuts.size());
    if (rc) return rc;
    if (dbg) cudaStreamSynchronize(t->ctx->stream);
    const auto t_2 = tnow();
    const double prevBest = t->bestPossibleEval;
    const int prevIters = t->simplexIters;
    int n_optimal = 0;
    double first_eval = 0;
    rc = simplex_with_mir(t, check_cycles, &st, false, &n_optimal, &first_eval);
    if (rc) return rc;
    const auto t_3 = tnow();
    // simplexIters / bestPossibleEval are frontier state: the commit loop owns them
    ev.optimal = n_optimal > 0;
    ev.n_optimal = t->use_mir ? n_optimal : -1;
    ev.first_eval = first_eval;
    t->simplexIters = prevIters;
    t->bestPossibleEval = prevBest;
    ev.valid = true;
    ev.feasible = t->feasible; ev.bounded = t->bounded; ev.evaluation = t->evaluation;
    ev.pivots = st.phase1_pivots + st.phase2_pivots;
    ev.is_integral = 0; ev.branch_var = -1; ev.branch_value = 0;
    if (ev.feasible) {
        MipOut mo;
        rc = mip_scan(t, &mo);
        if (rc) return rc;
        ev.is_integral = mo.is_integral; ev.branch_var = mo.var_index; ev.branch_value = mo.value;
        if (t->nOpt > 0) {
            std::vector<double> opt((size_t)t->nOpt * t->W);
            rc = jslp_download(t, nullptr, nullptr, nullptr, nullptr, nullptr, opt.data(), nullptr, nullptr);
            if (rc) return rc;
            for (int o = 0; o < t->nOpt && o < 7; o++) ev.opt0[o] = opt[(size_t)o * t->W];
        }
    }
    if (dbg)
        fprintf(stderr, "stream node: cuts %d pivots %d+%d restore %.0f us add_cuts %.0f us run_lp %.0f us rest %.0f us\n",
                (int)b.cuts.size(), st.phase1_pivots, st.phase2_pivots, us(t_0, t_1), us(t_1, t_2), us(t_2, t_3), us(t_3, tnow()));
    return JSLP_OK;
}

// ---- K3: HBM-resident node batch (jslp_slots.cuh) ---------------------------------------------
// Largest number of slots for which every slot still runs the ping-pong step (at most 32 rows per row CTA,
// all B * (G + 2) CTAs co-resident), capped by the option and by what keeps the buffers being written in L2.
static int slots_for(const jslp_tab *t, int rowcap, int want) {
    if (t->node_slots == 0 || want < 2) return 0;
    if (!(t->pingpong && t->lookahead && t->nOpt == 0) || t->engine == 1) return 0;
    const StepVariant &sv = STEP_VARIANTS[t->slot_variant];
    int per_sm = t->grid_per_sm > 0 ? t->grid_per_sm : sv.ctas_per_sm;
    {   // every CTA of the slot batch must be co-resident (selectors wait for row CTAs of the same launch)
        int nb = 0;
        cudaFuncSetAttribute(sv.fn_pp, cudaFuncAttributeMaxDynamicSharedMemorySize, t->stride * 8);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sv.fn_pp, sv.threads, (size_t)t->stride * 8) != cudaSuccess) nb = 0;
        per_sm = std::min(per_sm, nb);
        if (per_sm < 1) return 0;
    }
    const int C = t->ctx->num_sms * per_sm;
    int geom = 0;
    for (int b = 1; b <= 64; b++) {
        const int G = C / b - 2;
        if (G < 1 || rowcap / G + 1 > 32) break;
        geom = b;
    }
    int cap = geom;
    if (t->node_slots > 0) cap = std::min(cap, t->node_slots);
    else {  // auto: the buffers being WRITTEN by all slots (half of each pair; the dead-load hint lets L2 drop the other
            // half first) should stay in L2: each pivot re-reads what the last one wrote
        const double live_bytes = 8.0 * (double)rowcap * t->stride;
        const int l2 = (int)std::floor(0.55 * (double)t->ctx->l2_bytes / live_bytes);
        cap = std::min(cap, std::max(2, l2));
    }
    return std::min(cap, want) >= 2 ? std::min(cap, want) : 0;
}

static int ensure_slots(jslp_tab *t, int B, int need_rowcap) {
    NodeSlots &ns = t->slots;
    jslp_ctx *ctx = t->ctx;
    cudaStream_t s = ctx->stream;
    const StepVariant &sv = STEP_VARIANTS[t->slot_variant];
    int per_sm = t->grid_per_sm > 0 ? t->grid_per_sm : sv.ctas_per_sm;
    {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sv.fn_pp, sv.threads, (size_t)t->stride * 8) == cudaSuccess) per_sm = std::min(per_sm, std::max(1, nb));
    }
    const int G = ctx->num_sms * per_sm / B - 2;
    const int S = t->slot_steps;
    const int stride = t->stride;
    if (ns.B != B || ns.rowcap < need_rowcap) {
        CK(cudaStreamSynchronize(s));
        const int rowcap = std::max(need_rowcap + 32, std::max(ns.rowcap, t->saved.H + 64));
        ns.release();
        ns.B = B; ns.rowcap = rowcap;
        ns.plog_cap = 1024;
        const size_t tab = (size_t)rowcap * stride;
        CK(cudaMalloc(&ns.d_T, sizeof(TabDev) * B));
        CK(cudaMalloc(&ns.d_rec, sizeof(Rec) * B));
        CK(cudaMalloc(&ns.M, sizeof(double) * tab * B));
        CK(cudaMalloc(&ns.M2, sizeof(double) * tab * B));
        CK(cudaMalloc(&ns.prow, sizeof(double) * (size_t)stride * B));
        CK(cudaMalloc(&ns.crow, sizeof(double) * (size_t)stride * B));
        CK(cudaMalloc(&ns.pcol, sizeof(double) * (size_t)rowcap * B));
        CK(cudaMalloc(&ns.vrow, sizeof(int) * (size_t)rowcap * B));
        CK(cudaMalloc(&ns.vcol, sizeof(int) * (size_t)t->W * B));
        CK(cudaMalloc(&ns.part, sizeof(Part) * (size_t)(ctx->num_sms * 16) * B));
        CK(cudaMalloc(&ns.plog, sizeof(int4) * (size_t)ns.plog_cap * B));
        CK(cudaMallocHost(&ns.h_logs, sizeof(int4) * (size_t)ns.plog_cap * B));
        CK(cudaMemsetAsync(ns.M, 0, sizeof(double) * tab * B, s));
        CK(cudaMemsetAsync(ns.M2, 0, sizeof(double) * tab * B, s));
        CK(cudaMemsetAsync(ns.prow, 0, sizeof(double) * (size_t)stride * B, s));
        CK(cudaMemsetAsync(ns.pcol, 0, sizeof(double) * (size_t)rowcap * B, s));
        CK(cudaMemsetAsync(ns.part, 0xff, sizeof(Part) * (size_t)(ctx->num_sms * 16) * B, s));
        int rc;
        if ((rc = ResidentBufs::mapped(&ns.h_ctl, &ns.dv_ctl, sizeof(SlotCtl) * B))) return rc;
        if ((rc = ResidentBufs::mapped(&ns.h_out, &ns.dv_out, sizeof(SlotOut) * B))) return rc;
        ns.cuts_cap = B * (rowcap - t->saved.H);
        if ((rc = ResidentBufs::mapped(&ns.h_cuts, &ns.dv_cuts, sizeof(CutDev) * (size_t)std::max(1, ns.cuts_cap)))) return rc;
        std::vector<TabDev> hd((size_t)B, t->hd);
        std::vector<Rec> hr((size_t)B);
        for (int b = 0; b < B; b++) {
            TabDev &d = hd[b];
            d.M = ns.M + tab * b; d.M2 = ns.M2 + tab * b;
            d.prow = ns.prow + (size_t)stride * b; d.crow = ns.crow + (size_t)stride * b;
            d.pcol = ns.pcol + (size_t)rowcap * b;
            d.vrow = ns.vrow + (size_t)rowcap * b; d.vcol = ns.vcol + (size_t)t->W * b;
            d.part = ns.part + (size_t)(ctx->num_sms * 16) * b;
            d.plog = ns.plog + (size_t)ns.plog_cap * b; d.plog_cap = ns.plog_cap;
            d.opt = nullptr; d.optcoef = nullptr; d.dbg = nullptr; d.dbg_cap = 0; d.dbg_grid = 0; d.nOpt = 0;
            d.W = t->W; d.H = t->saved.H; d.stride = stride; d.rowcap = rowcap;
            d.n_index = t->n_index; d.prec = t->precision;
            memset(&hr[b], 0, sizeof(Rec));
            hr[b].status = ST_P1_DONE;
            ns.h_ctl[b] = SlotCtl{SLOT_IDLE, 0, 0, 0};
        }
        CK(cudaMemcpyAsync(ns.d_T, hd.data(), sizeof(TabDev) * B, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ns.d_rec, hr.data(), sizeof(Rec) * B, cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
        ns.key = -1;
    }
    // the graph bakes in the launch geometry and the snapshot it restores from
    const int key = (int)((((size_t)t->saved.M >> 4) * 2654435761u) ^ (size_t)(t->slot_variant * 131 + G * 7 + S * 1009 + t->saved.H * 31 +
                          t->saved.lastElementIndex * 17 + B)) & 0x7fffffff;
    if (ns.graph && ns.key == key && ns.G == G && ns.steps == S) return JSLP_OK;
    if (ns.graph) { cudaGraphExecDestroy(ns.graph); ns.graph = nullptr; }
    if (S + 2 > ns.plog_cap) return fail(JSLP_E_INVALID, "slot steps exceed the slot pivot-log capacity");
    if (G + 2 > ctx->num_sms * 16) return fail(JSLP_E_INVALID, "slot grid exceeds the partial-message buffer");
    const int smem = stride * 8;
    CK(cudaFuncSetAttribute(sv.fn_pp, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    SlotBatchDev sb;
    memset(&sb, 0, sizeof(sb));
    sb.T = ns.d_T; sb.rec = ns.d_rec; sb.ctl = ns.dv_ctl; sb.cuts = ns.dv_cuts; sb.out = ns.dv_out;
    sb.rootM = t->saved.M; sb.root_vrow = t->saved.vrow; sb.root_vcol = t->saved.vcol;
    sb.H0 = t->saved.H; sb.first_index = t->saved.lastElementIndex; sb.part_n = G + 2; sb.lookahead = 1;
    cudaGraph_t g;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    k_slot_begin<<<dim3(std::max(8, G + 2), B), 256, 0, s>>>(sb);
    k_select<<<B, 512, 0, s>>>(ns.d_T, ns.d_rec, -1, -1);
    for (int i = 0; i < S; i++)
        sv.fn_pp<<<dim3(G + 2, B), sv.threads, smem, s>>>(ns.d_T, ns.d_rec, 2, ns.prow, stride);
    k_slot_end<<<B, 256, 0, s>>>(sb);
    cudaMemcpy2DAsync(ns.h_logs, sizeof(int4) * (size_t)(S + 2), ns.plog, sizeof(int4) * (size_t)ns.plog_cap,
                      sizeof(int4) * (size_t)(S + 2), B, cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamEndCapture(s, &g);
    if (e != cudaSuccess) return fail(JSLP_E_CUDA, std::string("slot graph capture: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(&ns.graph, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(JSLP_E_CUDA, std::string("slot graph instantiate: ") + cudaGetErrorString(e));
    ns.key = key; ns.G = G; ns.steps = S;
    return JSLP_OK;
}

// What a round shares between ranks while its node LPs are running (n_ranks > 1): element-wise min.
struct RoundExchange {
    const jslp_bnb_opts *opts;
    int n_ranks;
    std::vector<double> scratch;
    int64_t collectives = 0;
    int min(double *v, int n) {
        collectives++;
        if (opts->comm) return jslp_comm_all_reduce_min(opts->comm, v, n);
        scratch.assign((size_t)n * n_ranks, 0.0);
        memcpy(&scratch[(size_t)opts->rank * n], v, sizeof(double) * n);
        if (opts->all_gather(opts->user, scratch.data(), (int64_t)sizeof(double) * n)) return fail(JSLP_E_INVALID, "all_gather hook failed");
        for (int r = 0; r < n_ranks; r++)
            for (int k = 0; k < n; k++) v[k] = r == 0 ? scratch[k] : std::min(v[k], scratch[(size_t)r * n + k]);
        return JSLP_OK;
    }
    int gather(void *buf, int64_t bytes_per_rank) {
        collectives++;
        if (opts->comm) return jslp_comm_all_gather(opts->comm, buf, bytes_per_rank);
        if (opts->all_gather(opts->user, buf, bytes_per_rank)) return fail(JSLP_E_INVALID, "all_gather hook failed");
        return JSLP_OK;
    }
};

// Evaluates nodes[0..n) in B slots side by side.  Nodes whose pivot log shows a cycle are re-evaluated on the
// single-tableau HBM path (exact stop-before-the-repeat semantics there).
//
// *U is the incumbent bound of the round: the smallest evaluation of an integral feasible node known so far
// (committed incumbent, cached results, nodes finished in this call -- on any rank when `ex` is set: one
// all-reduce(min) per poll keeps the ranks' poll loops in lockstep and shares the bound).  A speculative node
// whose relaxedEvaluation exceeds it is dropped or aborted: it sorts after the node that set the bound, so by
// the time the reference's loop pops it bestEvaluation <= U < relaxedEvaluation and it is skipped unevaluated
// (branch-and-cut.ts:90-92).  Pruned nodes keep ev.valid == false.
static int eval_nodes_slots(jslp_tab *t, jslp_bnb::Branch *const *nodes, int n, int B, int check_cycles, double *U,
                            RoundExchange *ex, int64_t *pruned) {
    jslp_ctx *ctx = t->ctx;
    cudaStream_t s = ctx->stream;
    int maxc = 0;
    for (int i = 0; i < n; i++) maxc = std::max(maxc, (int)nodes[i]->cuts.size());
    int rc = ensure_slots(t, B, t->saved.H + maxc);
    if (rc) return rc;
    NodeSlots &ns = t->slots;
    const int S = ns.steps;
    static const bool dbg = getenv("JSLP_DEBUG") != nullptr;
    std::vector<int> node_of((size_t)B, -1), redo;
    std::vector<CycleHist> hist((size_t)2 * B);
    int next = 0, done = 0;
    auto account = [&](int i, int pivots) {
        t->slot_pivots += pivots;
        t->slot_bytes += 16.0 * (double)(t->saved.H + (int)nodes[i]->cuts.size()) * t->stride * pivots;
    };
    bool any = n > 0 || ex != nullptr;
    while (any) {
        if (done < n) {
            int tot = 0;
            for (int b = 0; b < B; b++) {
                if (node_of[b] >= 0 && nodes[node_of[b]]->relaxedEvaluation > *U) {  // abort: the reference skips it
                    account(node_of[b], ns.h_out[b].rec.p1 + ns.h_out[b].rec.p2);
                    node_of[b] = -1; done++; (*pruned)++;
                }
                while (node_of[b] < 0 && next < n && nodes[next]->relaxedEvaluation > *U) { next++; done++; (*pruned)++; }
                if (node_of[b] < 0 && next < n) {
                    const std::vector<jslp_cut> &cuts = nodes[next]->cuts;
                    if (tot + (int)cuts.size() > ns.cuts_cap) return fail(JSLP_E_CAPACITY, "slot cut buffer overflow");
                    ns.h_ctl[b] = SlotCtl{SLOT_LOAD, (int)cuts.size(), tot, 0};
                    for (const jslp_cut &c : cuts) {
                        ns.h_cuts[tot].type = c.type; ns.h_cuts[tot].var_index = c.var_index; ns.h_cuts[tot].value = c.value;
                        tot++;
                    }
                    node_of[b] = next++;
                    hist[2 * b] = CycleHist(); hist[2 * b + 1] = CycleHist();
                } else {
                    ns.h_ctl[b].cmd = node_of[b] >= 0 ? SLOT_CONTINUE : SLOT_IDLE;
                }
            }
        }
        int running = 0;
        for (int b = 0; b < B; b++) running += node_of[b] >= 0;
        if (running > 0) {
            const auto t_g = std::chrono::steady_clock::now();
            CK(cudaGraphLaunch(ns.graph, s));
            ctx->launches += 3 + S;
            CK(cudaStreamSynchronize(s));
            const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_g).count();
            t->slot_ms += 1e-3 * us;
            if (dbg) {
                int loads = 0;
                for (int b = 0; b < B; b++) loads += ns.h_ctl[b].cmd == SLOT_LOAD;
                fprintf(stderr, "slot graph: B %d G %d S %d running %d loads %d: %.0f us\n", B, ns.G, S, running, loads, us);
            }
            for (int b = 0; b < B; b++) {
                const int i = node_of[b];
                if (i < 0) continue;
                const Rec &r = ns.h_out[b].rec;
                if (r.status == ST_ERROR) return fail(JSLP_E_CUDA, "slot batch: selector CTA timed out waiting for row CTAs");
                if (r.log_n > S + 2) return fail(JSLP_E_CAPACITY, "slot pivot log overflow inside one batch");
                bool hit = false;
                if (check_cycles) {
                    const int4 *lg = ns.h_logs + (size_t)b * (S + 2);
                    int cs, cl;
                    for (int k = 0; k < r.log_n && !hit; k++) {
                        CycleHist &h = hist[2 * b + (((lg[k].x >> 30) & 1) ? 1 : 0)];
                        hit = h.push_and_check(((long long)lg[k].z << 32) | (unsigned int)lg[k].w, &cs, &cl);
                    }
                }
                if (hit) { account(i, r.p1 + r.p2); redo.push_back(i); node_of[b] = -1; done++; continue; }
                if (r.status == ST_RUNNING) continue;
                jslp_bnb::NodeEval &ev = nodes[i]->ev;
                ev.valid = true;
                ev.pivots = r.p1 + r.p2;
                ev.optimal = r.status == ST_OPTIMAL;
                ev.bounded = r.status != ST_UNBOUNDED;
                ev.feasible = r.status == ST_OPTIMAL || r.status == ST_UNBOUNDED;
                ev.evaluation = r.status == ST_OPTIMAL ? jslp_round_evaluation(r.eval_raw, t->precision)
                                                       : (r.status == ST_UNBOUNDED ? -INFINITY : 0.0);
                const MipOut &mo = ns.h_out[b].mip;
                ev.is_integral = ev.feasible ? mo.is_integral : 0;
                ev.branch_var = ev.feasible ? mo.var_index : -1;
                ev.branch_value = ev.feasible ? mo.value : 0.0;
                if (ev.feasible && ev.is_integral && ev.evaluation < *U) *U = ev.evaluation;
                account(i, ev.pivots);
                if (dbg) fprintf(stderr, "slot node: slot %d cuts %d pivots %d+%d status %d\n", b, (int)nodes[i]->cuts.size(), r.p1, r.p2, r.status);
                node_of[b] = -1;
                done++;
            }
        }
        if (ex) {  // share the bound; stay in the loop until every rank has finished its nodes
            double v[2] = {*U, done < n ? -1.0 : 0.0};
            rc = ex->min(v, 2);
            if (rc) return rc;
            *U = v[0];
            any = v[1] < 0;
        } else {
            any = done < n;
        }
    }
    for (int i : redo) {
        if (nodes[i]->relaxedEvaluation > *U) { (*pruned)++; continue; }
        rc = eval_node_streaming(t, *nodes[i], check_cycles, nodes[i]->ev);
        if (rc) return rc;
    }
    return JSLP_OK;
}

static size_t node_smem_bytes(int Hcap, int W, int *Ws_out) {
    const int Ws = (W + 1) & ~1;  // even row stride: 16-byte aligned rows (TMA restore)
    if (Ws_out) *Ws_out = Ws;
    return sizeof(double) * ((size_t)Hcap * Ws + 2 * (size_t)Ws + Hcap + 1) + sizeof(CutDev) * (size_t)Hcap +
           sizeof(int) * (2 * (size_t)Hcap + W) + 16;
}

static bool resident_fits(const jslp_tab *t, int Hcap) {
    return t->nOpt == 0 && node_smem_bytes(Hcap, t->W, nullptr) <= (size_t)t->ctx->max_smem_optin - 4096;
}

// Evaluates nodes[0..n) with one CTA each (k_node_batch).  Nodes whose log shows a cycle or that
// hit the in-kernel pivot cap are re-evaluated on the HBM path (exact cycle semantics there).
static int eval_nodes_resident(jslp_tab *t, jslp_bnb::Branch *const *nodes, int n, int check_cycles) {
    jslp_ctx *ctx = t->ctx;
    cudaStream_t s = ctx->stream;
    ResidentBufs &rb = t->rbufs;
    int maxc = 0, totc = 0;
    for (int i = 0; i < n; i++) {
        maxc = std::max(maxc, (int)nodes[i]->cuts.size());
        totc += (int)nodes[i]->cuts.size();
    }
    const int Hcap = t->saved.H + maxc;
    int Ws = 0;
    const size_t smem = node_smem_bytes(Hcap, t->W, &Ws);
    const int log_cap = t->node_log_cap;
    int rc = rb.ensure(n, totc, log_cap);
    if (rc) return rc;
    int off = 0;
    for (int i = 0; i < n; i++) {
        rb.h_off[i] = off;
        for (const jslp_cut &c : nodes[i]->cuts) {
            rb.h_cuts[off].type = c.type; rb.h_cuts[off].var_index = c.var_index; rb.h_cuts[off].value = c.value;
            off++;
        }
    }
    rb.h_off[n] = off;
    NodeBatchDev nb;
    memset(&nb, 0, sizeof(nb));
    nb.rootM = t->saved.M; nb.root_vrow = t->saved.vrow; nb.root_vcol = t->saved.vcol;
    nb.cuts = rb.dv_cuts; nb.cut_off = rb.dv_off; nb.out = rb.dv_out; nb.logs = rb.d_logs;
    nb.H0 = t->saved.H; nb.root_stride = t->stride; nb.first_index = t->saved.lastElementIndex;
    nb.Hcap = Hcap; nb.Ws = Ws; nb.log_cap = log_cap; nb.max_pivots = 100000;
    if (n <= NODE_INLINE_OFF) {
        nb.n_inline = n;
        for (int i = 0; i <= n; i++) nb.cut_off_inline[i] = rb.h_off[i];
    }
    if ((int)smem > rb.smem_set) {
        CK(cudaFuncSetAttribute(k_node_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rb.smem_set = (int)smem;
    }
    k_node_batch<<<n, NODE_THREADS, smem, s>>>(t->d_T, nb);
    ctx->launches += 1;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));  // inputs and outputs are mapped host memory: nothing to copy
    long long slowest = 0;
    bool need_full = false;
    static const bool dbg = getenv("JSLP_DEBUG") != nullptr;  // per-node timeline on stderr (scripts/node_timeline.py)
    for (int i = 0; i < n; i++) {
        if (dbg) {
            const NodeResult &q = rb.h_out[i].r;
            fprintf(stderr, "node r%d n%d i%d cuts %d piv %d+%d: off %lld restore %lld cuts %lld pivots %lld mip %lld end %lld cy %lld %lld %lld %lld %lld %lld\n",
                    (int)ctx->launches, n, i, (int)nodes[i]->cuts.size(), q.p1, q.p2, q.tl[0], q.tl[1], q.tl[2], q.tl[3], q.tl[4], q.t_ns, q.cy[0], q.cy[1], q.cy[2], q.cy[3], q.cy[4], q.cy[5]);
        }
        slowest = std::max(slowest, rb.h_out[i].r.t_ns);
        if (check_cycles && !rb.h_out[i].r.overflow && rb.h_out[i].r.log_n > NODE_LOG_HEAD) need_full = true;
    }
    t->node_kernel_ns += slowest;
    if (need_full) {  // rare: a node with a long pivot sequence, fetch the complete logs
        CK(cudaMemcpyAsync(rb.h_logs, rb.d_logs, sizeof(int4) * (size_t)n * log_cap, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    std::vector<int> redo_list;
    for (int i = 0; i < n; i++) {
        const NodeResult &r = rb.h_out[i].r;
        jslp_bnb::NodeEval &ev = nodes[i]->ev;
        bool redo = r.overflow != 0;
        if (!redo && check_cycles) {
            std::vector<long long> h1, h2;
            int cs, cl;
            const int4 *lg = r.log_n > NODE_LOG_HEAD ? rb.h_logs + (size_t)i * log_cap : rb.h_out[i].log_head;
            for (int k = 0; k < r.log_n && k < log_cap && !redo; k++) {
                std::vector<long long> &h = ((lg[k].x >> 30) & 1) ? h2 : h1;
                h.push_back(((long long)lg[k].z << 32) | (unsigned int)lg[k].w);
                if (cycle_hit(h, &cs, &cl)) redo = true;
            }
        }
        if (redo) { redo_list.push_back(i); continue; }  // second pass: the HBM path may reuse rb's buffers
        ev.valid = true;
        ev.pivots = r.p1 + r.p2;
        ev.optimal = r.status == ST_OPTIMAL;
        ev.bounded = r.status != ST_UNBOUNDED;
        ev.feasible = r.status == ST_OPTIMAL || r.status == ST_UNBOUNDED;
        if (r.status == ST_OPTIMAL) {
            ev.evaluation = jslp_round_evaluation(r.eval_raw, t->precision);
        } else if (r.status == ST_UNBOUNDED) {
            ev.evaluation = -INFINITY;
        } else {
            ev.evaluation = 0.0;  // infeasible: never read by the commit loop; rank-independent on the wire
        }
        ev.is_integral = r.is_integral; ev.branch_var = r.branch_var; ev.branch_value = r.branch_value;
    }
    for (int i : redo_list) {
        rc = eval_node_streaming(t, *nodes[i], check_cycles, nodes[i]->ev);
        if (rc) return rc;
    }
    return JSLP_OK;
}

extern "C" int jslp_branch_and_cut(jslp_tab *t, const jslp_bnb_opts *opts, jslp_bnb_status *out,
                                   jslp_cut *best_cuts, int best_cuts_cap) {
    using namespace jslp_bnb;
    if (!t || !opts) return fail(JSLP_E_INVALID, "NULL argument");
    if (t->n_int <= 0) return fail(JSLP_E_INVALID, "branch_and_cut needs integer variables (upload int_var_indices)");
    const int n_ranks = std::max(1, opts->n_ranks), rank = opts->rank;
    if (n_ranks > 1 && !opts->all_gather && !opts->comm) return fail(JSLP_E_INVALID, "n_ranks > 1 needs a communicator or an all_gather hook");
    if (n_ranks > 1 && t->nOpt > 7) return fail(JSLP_E_UNSUPPORTED, "more than 7 optional objectives across ranks");
    if (opts->service == 1 || opts->service == 2) return bnb_enhanced(t, opts, out, best_cuts, best_cuts_cap);  // sequential by construction
    if (opts->service != 0) return fail(JSLP_E_UNSUPPORTED, "unknown branch-and-cut service");
    jslp_ctx *ctx = t->ctx;
    CK(cudaSetDevice(ctx->device));
    const int64_t launches0 = ctx->launches;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    const auto t_start = now();
    const double timeout_ms = opts->timeout_ms > 0 ? opts->timeout_ms : 0;  // branch-and-cut.ts:61-63
    auto time_up = [&] { return timeout_ms > 0 && ms_since(t_start) >= timeout_ms; };

    Frontier branches;
    int iterations = 0, rounds = 0;
    const double tolerance = opts->tolerance;
    const int check_cycles = opts->check_cycles;
    bool toleranceFlag = true, timed_out = false;
    double bestEvaluation = INFINITY;
    std::unique_ptr<Branch> bestBranch, lastCommitted;
    std::vector<double> bestOpt((size_t)t->nOpt, INFINITY);
    int64_t pivots = 0, nodes = 0, pruned = 0;
    t->node_log.clear();
    t->solutions.clear();
    t->saved.valid = false;
    t->isIntegralFlag = 0;
    t->slot_pivots = 0; t->slot_ms = 0; t->slot_bytes = 0;
    bool early_return = false;
    int K = opts->max_spec_batch > 0 ? opts->max_spec_batch : 16;  // any width commits in the same order
    std::vector<double> wire;
    RoundExchange ex{opts, n_ranks, {}, 0};

    // keep_solutions (branch-and-cut.ts:143-153): the incumbent's tableau state, as generateSolutionSet reads it.
    // Node evaluators return summaries only, so the node is solved once more in place (incumbents are rare).
    auto store_solution = [&](const Branch &b) -> int {
        NodeEval ev;
        int rc = eval_node_streaming(t, b, check_cycles, ev);
        if (rc) return rc;
        jslp_tab::StoredSolution sol;
        sol.evaluation = t->evaluation;
        sol.vrow.resize((size_t)t->H);
        sol.rhs.resize((size_t)t->H);
        rc = jslp_download(t, nullptr, sol.rhs.data(), nullptr, sol.vrow.data(), nullptr, nullptr, nullptr, nullptr);
        if (rc) return rc;
        t->solutions.push_back(std::move(sol));
        return JSLP_OK;
    };

    // commits one evaluated node exactly like one iteration of branch-and-cut.ts:89-192
    auto commit = [&](std::unique_ptr<Branch> active) -> int {
        const NodeEval ev = active->ev;
        iterations++;
        t->feasible = ev.feasible; t->bounded = ev.bounded;
        if (ev.feasible) t->evaluation = ev.evaluation;
        if (ev.optimal) {  // setEvaluation + simplexIters (tableau.ts:420-430, simplex.ts:266-267)
            if (t->simplexIters == 0) t->bestPossibleEval = ev.n_optimal >= 0 ? ev.first_eval : ev.evaluation;
            t->simplexIters += ev.n_optimal >= 0 ? ev.n_optimal : 1;
        }
        NodeLogEntry nl;
        nl.v[0] = iterations; nl.v[1] = (double)active->cuts.size(); nl.v[2] = ev.feasible;
        nl.v[3] = t->evaluation; nl.v[4] = -1; nl.v[5] = -1; nl.v[6] = 0; nl.v[7] = ev.pivots;
        auto done = [&](std::unique_ptr<Branch> keep) {
            t->node_log.push_back(nl);
            lastCommitted = std::move(keep);
        };
        if (!ev.feasible) { done(std::move(active)); return 0; }
        const double evaluation = ev.evaluation;
        if (evaluation > bestEvaluation) { done(std::move(active)); return 0; }
        if (evaluation == bestEvaluation) {  // branch-and-cut.ts:107-127
            bool worse = true;
            for (int o = 0; o < t->nOpt; o++) {
                if (ev.opt0[o] > bestOpt[o]) break;
                if (ev.opt0[o] < bestOpt[o]) { worse = false; break; }
            }
            if (worse) { done(std::move(active)); return 0; }
        }
        if (ev.is_integral) {
            nl.v[4] = 1;
            t->isIntegralFlag = 1;
            if (iterations == 1) { t->node_log.push_back(nl); early_return = true; return 1; }
            bestEvaluation = evaluation;
            for (int o = 0; o < t->nOpt; o++) bestOpt[o] = ev.opt0[o];
            t->node_log.push_back(nl);
            bestBranch.reset(new Branch{active->relaxedEvaluation, active->cuts, NodeEval()});
            if (opts->keep_solutions) {
                int rc = store_solution(*active);
                if (rc) return rc;
            }
            lastCommitted = std::move(active);
        } else {
            if (iterations == 1) {  // snapshot = optimal root tableau (branch-and-cut.ts:155-157)
                int rc = jslp_save(t);
                if (rc) return rc;
            }
            const int varIndex = ev.branch_var;
            const double value = ev.branch_value;
            nl.v[4] = 0; nl.v[5] = varIndex; nl.v[6] = value;
            std::unique_ptr<Branch> high(new Branch{evaluation, {}, NodeEval()}), low(new Branch{evaluation, {}, NodeEval()});
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
            done(std::move(active));
        }
        return 0;
    };

    double eval_ms = 0, commit_ms = 0, root_ms = 0, final_ms = 0;
    t->node_kernel_ns = 0;
    branches.push(std::unique_ptr<Branch>(new Branch{-INFINITY, {}, NodeEval()}));
    bool stop = false;
    while (!stop && !branches.empty() && toleranceFlag) {
        // ---- speculate: evaluate the next K un-evaluated open nodes in pop order ---------------
        const int width = iterations == 0 ? 1 : K;  // the root decides everything: alone
        std::vector<Frontier::Entry> taken;
        std::vector<Branch *> todo;
        // U = incumbent bound of the round (see eval_nodes_slots): committed incumbent and cached integral results
        double U = bestEvaluation;
        // the scan is bounded: once the un-evaluated nodes thin out it would otherwise walk the whole heap every
        // round (entries with a cached result are popped and pushed back); candidates that far down the pop order
        // are not popped before the next round anyway
        const size_t scan_cap = (size_t)width * 4 + 32;
        while (!branches.empty() && (int)todo.size() < width && taken.size() < scan_cap) {
            Frontier::Entry e = branches.pop_entry();
            // nodes that will be skipped at pop (branch-and-cut.ts:90-92) stay in the heap -- their pop
            // still consumes a loop iteration of the reference -- but are never evaluated
            if (e.b->ev.valid) {
                if (e.b->ev.feasible && e.b->ev.is_integral && e.b->ev.evaluation < U) U = e.b->ev.evaluation;
            } else if (!(e.b->relaxedEvaluation > U)) {
                todo.push_back(e.b.get());
            } else {
                pruned++;
            }
            taken.push_back(std::move(e));
        }
        const auto t_eval = now();
        if (!todo.empty()) {
            rounds++;
            int maxc_all = 0;
            for (Branch *b : todo) maxc_all = std::max(maxc_all, (int)b->cuts.size());
            // every rank takes the same decisions from the same (rank-independent) quantities
            // useMIRCuts: the MIR loop (fractional volume, cut rows, re-solves) runs on the single-tableau path
            const bool resident = iterations > 0 && t->saved.valid && t->engine != 1 && t->engine != 2 && !t->use_mir &&
                                  resident_fits(t, t->saved.H + maxc_all);
            const int nslots = (!resident && !t->use_mir && iterations > 0 && t->saved.valid && todo.size() >= 2)
                                   ? slots_for(t, t->saved.H + maxc_all, std::max(K, (int)todo.size())) : 0;
            // The root round is NOT sharded: every rank needs the solved root in its own tableau, it is what
            // jslp_save snapshots and what every later node is derived from.  Rounds of nodes that fit shared
            // memory are not sharded either unless asked: a whole round costs less than one collective.
            const bool sharded = n_ranks > 1 && iterations > 0 && !t->use_mir && (opts->shard_policy == 1 || !resident);
            std::vector<Branch *> mine;
            for (size_t i = 0; i < todo.size(); i++)
                if (!sharded || (int)(i % n_ranks) == rank) mine.push_back(todo[i]);
            if (resident) {
                if (!mine.empty()) {
                    int rc = eval_nodes_resident(t, mine.data(), (int)mine.size(), check_cycles);
                    if (rc) return rc;
                }
            } else if (nslots >= 2) {
                int rc = eval_nodes_slots(t, mine.data(), (int)mine.size(), nslots, check_cycles, &U, sharded ? &ex : nullptr, &pruned);
                if (rc) return rc;
            } else {
                for (Branch *b : mine) {
                    if (b->relaxedEvaluation > U) { pruned++; continue; }
                    int rc = eval_node_streaming(t, *b, check_cycles, b->ev);
                    if (rc) return rc;
                    if (b->ev.feasible && b->ev.is_integral && b->ev.evaluation < U) U = b->ev.evaluation;
                }
            }
            nodes += (int64_t)mine.size();
            if (sharded) {  // all-gather the summaries: rank-major blocks of `per` records
                const int per = (int)((todo.size() + n_ranks - 1) / n_ranks);
                wire.assign((size_t)per * n_ranks * WIRE_DOUBLES, 0.0);
                for (size_t j = 0; j < mine.size(); j++)
                    to_wire(mine[j]->ev, &wire[((size_t)rank * per + j) * WIRE_DOUBLES]);
                int rc = ex.gather(wire.data(), (int64_t)per * WIRE_DOUBLES * 8);
                if (rc) return rc;
                for (size_t i = 0; i < todo.size(); i++)
                    from_wire(todo[i]->ev, &wire[((size_t)(i % n_ranks) * per + i / n_ranks) * WIRE_DOUBLES]);
            }
        }
        for (auto &e : taken) branches.push_entry(std::move(e));  // original seq: order unchanged
        if (iterations == 0) root_ms += ms_since(t_eval);
        eval_ms += ms_since(t_eval);
        const auto t_commit = now();

        // Date.now() < terminalTime (branch-and-cut.ts:76).  One rank: tested before every commit, like the
        // reference's loop condition.  Several ranks: tested once per round and agreed on (min-reduce), so that
        // every rank stops at the same commit.
        bool round_time_up = false;
        if (n_ranks > 1 && timeout_ms > 0) {
            double v = time_up() ? -1.0 : 0.0;
            int rc = ex.min(&v, 1);
            if (rc) return rc;
            round_time_up = v < 0;
        }

        // ---- commit sequentially in the reference's exact order ---------------------------------
        while (!branches.empty() && toleranceFlag) {
            if (opts->max_nodes > 0 && iterations >= opts->max_nodes) { stop = true; break; }
            if (n_ranks > 1 ? round_time_up : time_up()) { timed_out = true; stop = true; break; }
            const double acceptableThreshold = opts->is_minimization ? t->bestPossibleEval * (1 + tolerance)
                                                                     : t->bestPossibleEval * (1 - tolerance);
            // the reference checks the tolerance BEFORE popping, and still processes that pop
            const bool tolHit = tolerance > 0 && bestEvaluation < acceptableThreshold;
            if (branches.h[0].b->relaxedEvaluation <= bestEvaluation && !branches.h[0].b->ev.valid) break;  // next round
            if (tolHit) toleranceFlag = false;
            Frontier::Entry e = branches.pop_entry();
            if (e.b->relaxedEvaluation > bestEvaluation) continue;
            pivots += e.b->ev.pivots;
            int rc = commit(std::move(e.b));
            if (rc < 0) return rc;
            if (rc == 1) { stop = true; break; }
        }
        commit_ms += ms_since(t_commit);
    }

    // Final tableau: the reference re-solves the winner (branch-and-cut.ts:195-197); without a
    // winner it is left at the last evaluated node.  In place on the HBM path.
    int n_best = 0;
    if (!early_return) {
        const Branch *fin = bestBranch ? bestBranch.get() : lastCommitted.get();
        if (fin && (bestBranch || iterations > 1)) {
            const auto t_fin = now();
            NodeEval ev;
            int rc = eval_node_streaming(t, *fin, check_cycles, ev);
            if (rc) return rc;
            nodes++;
            if (bestBranch) {
                pivots += ev.pivots;
                if (ev.optimal) t->simplexIters += ev.n_optimal >= 0 ? ev.n_optimal : 1;
            }
            final_ms = ms_since(t_fin);
        }
        if (bestBranch) {
            n_best = (int)bestBranch->cuts.size();
            if (best_cuts)
                for (int i = 0; i < n_best && i < best_cuts_cap; i++) best_cuts[i] = bestBranch->cuts[i];
        }
    }
    t->bncIterations = iterations;

    float ms = 0.f;
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev1));
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    if (out) {
        memset(out, 0, sizeof(*out));
        out->feasible = t->feasible; out->bounded = t->bounded; out->is_integral = t->isIntegralFlag;
        out->iterations = iterations; out->n_best_cuts = n_best; out->rounds = rounds;
        out->nodes_evaluated = nodes; out->pivots = pivots; out->evaluation = t->evaluation;
        out->best_possible_eval = t->bestPossibleEval; out->gpu_ms = ms;
        out->kernel_launches = ctx->launches - launches0;
        out->host_eval_ms = eval_ms; out->host_commit_ms = commit_ms;
        out->host_root_ms = root_ms; out->host_final_ms = final_ms; out->node_kernel_ms = 1e-6 * (double)t->node_kernel_ns;
        out->timed_out = timed_out ? 1 : 0; out->n_solutions = (int)t->solutions.size();
        out->nodes_pruned = pruned; out->collectives = ex.collectives;
        out->slot_pivots = t->slot_pivots; out->slot_ms = t->slot_ms; out->slot_bytes = t->slot_bytes;
    }
    return JSLP_OK;
}

extern "C" int jslp_bnb_solution(jslp_tab *t, int i, double *evaluation, int32_t *height, int32_t *vrow, double *rhs, int cap) {
    if (!t) return fail(JSLP_E_INVALID, "tab is NULL");
    if (i < 0 || i >= (int)t->solutions.size()) return fail(JSLP_E_INVALID, "solution index out of range");
    const jslp_tab::StoredSolution &s = t->solutions[(size_t)i];
    if (evaluation) *evaluation = s.evaluation;
    if (height) *height = (int32_t)s.vrow.size();
    const size_t n = std::min<size_t>((size_t)std::max(0, cap), s.vrow.size());
    if (vrow) memcpy(vrow, s.vrow.data(), sizeof(int32_t) * n);
    if (rhs) memcpy(rhs, s.rhs.data(), sizeof(double) * n);
    return JSLP_OK;
}

extern "C" int jslp_bnb_node_log(jslp_tab *t, double *entries, int64_t cap, int64_t *n) {
    if (!t || !n) return fail(JSLP_E_INVALID, "NULL argument");
    const int64_t m = std::min<int64_t>((int64_t)t->node_log.size(), cap);
    for (int64_t i = 0; i < m && entries; i++) memcpy(entries + 8 * i, t->node_log[i].v, sizeof(double) * 8);
    *n = (int64_t)t->node_log.size();
    return JSLP_OK;
}
