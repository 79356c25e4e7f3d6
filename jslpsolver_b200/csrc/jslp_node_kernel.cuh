// jslpsolver_b200/csrc/jslp_node_kernel.cuh -- shared-memory-resident node LPs.
//
// One CTA solves one branch-and-cut node (or one small LP) start to finish with the whole tableau
// in shared memory: restore from the root snapshot (backup.ts:53-105), append the node's cut rows
// (cutting-strategies.ts:36-71), run phase 1 / phase 2 to completion (simplex.ts:14-325) and
// report the summary the frontier needs (evaluation, isIntegral, most fractional variable;
// mip-utils.ts:43-61,100-126).  A launch evaluates a whole batch of open nodes, one CTA each:
// this is what makes the node frontier GPU-resident.  Selection is done by ONE warp with shuffle
// reductions (no CTA barrier): the same rules as cta_select, restated at warp scope; the parity tests
// run every fixture through both paths.
#pragma once
#include "jslp_kernels.cuh"

namespace jslp {

struct NodeResult {
    int status;       // ST_* at exit
    int p1, p2;       // pivots per phase
    int log_n;        // pivot-log entries written (selections)
    int overflow;     // pivot cap reached or log overflow: host re-evaluates on the HBM path
    int is_integral;
    int branch_var;   // -1 = none
    int unbounded_var;
    double eval_raw;
    double branch_value;
    long long t_ns;   // CTA lifetime by %globaltimer (reporting only)
    long long pad;
    long long cy[6];  // warp-0 cycles: select, barrier A, division passes, row update, look-ahead pricing, barrier B
    long long tl[6];  // ns since CTA start: offsets read, restored, cut rows built, pivots done, mip scan done, (unused)
};

// What a CTA reports: the summary plus the head of its pivot log (enough for the cycle check of
// almost every node), written straight into mapped pinned host memory -- no D2H copy per round.
constexpr int NODE_LOG_HEAD = 64;
struct NodeOut {
    NodeResult r;
    int4 log_head[NODE_LOG_HEAD];
};

constexpr int NODE_INLINE_OFF = 64;
struct NodeBatchDev {
    const double *rootM;   // root snapshot, row stride = root_stride
    const int *root_vrow, *root_vcol;
    const CutDev *cuts;    // all cuts of the batch, node n owns [cut_off[n], cut_off[n+1]); may be mapped host memory
    const int *cut_off;
    NodeOut *out;          // one record per node; may be mapped host memory
    int4 *logs;            // log_cap entries per node (device memory)
    double *wb_M;          // optional write-back of node 0's final tableau (stride = root_stride)
    int *wb_vrow, *wb_vcol;
    int H0, root_stride, first_index, Hcap, Ws, log_cap, max_pivots;
    int n_inline;          // > 0: cut_off_inline holds the n_inline + 1 offsets (saves a PCIe round trip)
    int cut_off_inline[NODE_INLINE_OFF + 1];
};

constexpr int NODE_THREADS = 256;

__device__ __forceinline__ long long globaltimer_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ---- warp-level selection ----------------------------------------------------------------------
// In the resident kernel one warp decides every pivot.  A lone warp issues a dependent instruction
// every ~8 cycles, so the reductions avoid shuffle trees: values are mapped to order-preserving
// integer keys and reduced with redux.sync (three to four REDUX per arg-min instead of five shuffle
// levels of fp64 compares), and the scans issue their loads and divisions four at a time.
// Phase-2 pricing of the cost row (simplex.ts:140-219 without optional objectives): first batch with
// an improving column, arg-max inside it, lowest column on ties.  Returns the column (0 = none).
// One copy each of the bulky helpers: the pivot loop has to stay inside the instruction cache (an
// earlier fully-inlined build had a 64 KB loop body and spent most of its time fetching instructions).
__device__ __noinline__ int warp_price(const TabDev &T, const double *cost, const int *vcol, int W, int lane, int *neg_out) {
    const int bsz = T.use_partial ? T.batch_size : max(1, W - 1);
    const bool has_unres = T.unres != nullptr;
    PriceAcc acc;
    price_init(acc, T.prec);
#pragma unroll 1
    for (int c = 1 + lane; c < W; c += 32) {
        const double nc = cost[c];
        int label = -1;
        if (has_unres && nc < 0) label = vcol[c];
        price_consider(T, acc, c, nc, label, bsz);
    }
    // lexicographic (batch asc, value desc, column asc)
    const int mb = __reduce_min_sync(0xffffffffu, acc.myb);
    if (mb == INT_MAX) { *neg_out = 0; return 0; }
    const bool in = acc.myb == mb;
    const unsigned long long k = dkey(acc.x.v);
    const unsigned int hi = (unsigned int)(k >> 32), lo = (unsigned int)k;
    const unsigned int mhi = __reduce_max_sync(0xffffffffu, in ? hi : 0u);
    const unsigned int mlo = __reduce_max_sync(0xffffffffu, (in && hi == mhi) ? lo : 0u);
    const bool win = in && hi == mhi && lo == mlo;
    const int col = __reduce_min_sync(0xffffffffu, win ? acc.x.i : INT_MAX);
    *neg_out = __reduce_max_sync(0xffffffffu, (win && acc.x.i == col) ? acc.myneg : 0);
    return col;
}

struct NodePivot {   // warp 0 -> CTA
    int go, r, c, flush;
    double q;
};
struct NodeScan {    // warp 0 -> CTA: which division scan to run (1 = phase-1 column scan of row r, 2 = ratio test of column c)
    int mode, r, c, neg;
};
struct NodePart {    // per-warp partial results of a scan
    double v[8];
    int i[8], d[8], n[8];
};

__global__ void __launch_bounds__(NODE_THREADS) k_node_batch(const TabDev *Tp, const __grid_constant__ NodeBatchDev nb) {
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ TabDev T;
    __shared__ SelSmem sel;
    __shared__ MipOut mip;
    __shared__ NodePivot piv;
    __shared__ NodeScan sc;
    __shared__ NodePart part;
    __shared__ int s_spare;
    __shared__ int s_fin[8];   // status, p1, p2, log_n, overflow, unbounded_var
    __shared__ double s_eval;
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x, NT = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, NW = NT >> 5;
    const int node = blockIdx.x;
    long long t_start = 0;
    if (tid == 0) t_start = globaltimer_ns();
    const int H0 = nb.H0, Ws = nb.Ws;

    // rows live in Hcap + 1 slots: the normalised pivot row is written to the spare slot and the slot
    // table is swapped, so nobody waits for the old pivot row to be overwritten in place.  Ws is even:
    // rows are 16-byte aligned, which is what lets TMA restore them.
    double *Ms = reinterpret_cast<double *>(smraw);
    double *frow = Ms + (size_t)(nb.Hcap + 1) * Ws;
    double *rhs = frow + Ws;                         // Hcap entries: pivot-column staging, final RHS column
    CutDev *cutS = reinterpret_cast<CutDev *>(rhs + nb.Hcap + (nb.Hcap & 1));
    int *vrow = reinterpret_cast<int *>(cutS + (nb.Hcap - H0));
    int *slot = vrow + nb.Hcap;
    if (tid == 0) {
        T = *Tp;
        T.vcol = slot + nb.Hcap;
        s_spare = nb.Hcap;
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    // restore(): root snapshot -> shared memory, one TMA bulk copy per row (backup.ts:53-105)
    if (tid == 0) mbar_expect_tx(&s_bar, (uint32_t)(H0 * Ws * sizeof(double)));
    __syncthreads();
    for (int r = tid; r < H0; r += NT)
        tma_bulk_g2s(Ms + (size_t)r * Ws, nb.rootM + (size_t)r * nb.root_stride, (uint32_t)(Ws * sizeof(double)), &s_bar);
    // the cut list may live in mapped host memory: one PCIe round trip, overlapped with the restore
    int c0, nc;
    if (nb.n_inline > 0) { c0 = nb.cut_off_inline[node]; nc = nb.cut_off_inline[node + 1] - c0; }
    else { c0 = nb.cut_off[node]; nc = nb.cut_off[node + 1] - c0; }
    long long tl0 = 0, tl1 = 0, tl2 = 0, tl3 = 0, tl4 = 0;
    const int W = T.W;
    const double prec = T.prec;
    int *vcol = T.vcol;
    for (int h = tid; h < nc; h += NT) cutS[h] = nb.cuts[c0 + h];
    for (int r = tid; r < H0; r += NT) vrow[r] = __ldg(nb.root_vrow + r);
    for (int c = tid; c < W; c += NT) vcol[c] = __ldg(nb.root_vcol + c);
    for (int r = tid; r < nb.Hcap; r += NT) slot[r] = r;
    const int Hn = H0 + nc;
    if (tid == 0) tl0 = globaltimer_ns() - t_start;
    mbar_wait(&s_bar, 0);
    __syncthreads();
    if (tid == 0) tl1 = globaltimer_ns() - t_start;
    // addCutConstraints(): every cut row is expressed in the ROOT basis, so the rows are independent:
    // one warp per cut, no CTA barrier inside
    for (int h = warp; h < nc; h += NW) {
        const CutDev cut = cutS[h];
        int s_row = -1, s_col = -1;
        for (int r = 1 + lane; r < H0; r += 32) if (vrow[r] == cut.var_index) s_row = r;
        for (int c = 1 + lane; c < W; c += 32) if (vcol[c] == cut.var_index) s_col = c;
        s_row = __reduce_max_sync(0xffffffffu, s_row);
        s_col = __reduce_max_sync(0xffffffffu, s_col);
        const double sign = cut.type == 0 ? -1.0 : 1.0;
        double *crow = Ms + (size_t)(H0 + h) * Ws;
        if (s_row < 0) {
            for (int c = lane; c < W; c += 32) crow[c] = c == 0 ? sign * cut.value : (c == s_col ? sign : 0.0);
        } else {
            const double *vr = Ms + (size_t)s_row * Ws;
            for (int c = lane; c < W; c += 32) crow[c] = c == 0 ? sign * (cut.value - vr[0]) : -sign * vr[c];
        }
        if (lane == 0) vrow[H0 + h] = nb.first_index + h;
    }
    __syncthreads();
    if (tid == 0) tl2 = globaltimer_ns() - t_start;

    // warp 0's registers: the solve state
    int phase = 1, p1 = 0, p2 = 0, log_n = 0, status = ST_RUNNING, overflow = 0, unb = -1;
    int next_c = -1, next_neg = 0;      // look-ahead pricing of the updated cost row (-1 = not priced)
    int rstar = -1, cstar = -1, isneg = 0;
    int4 *plog = nb.logs + (size_t)node * nb.log_cap;
    const bool has_unres = T.unres != nullptr;
    long long cy[6] = {0, 0, 0, 0, 0, 0}, cprev = clock64();
// per-phase cycle counters of warp 0 (S1, S2, S3, D, U, barrier): compile with -DJSLP_NODE_CYCLES to fill them
#ifdef JSLP_NODE_CYCLES
#define NODE_CY(k) do { const long long cnow = clock64(); cy[k] += cnow - cprev; cprev = cnow; } while (0)
#else
#define NODE_CY(k) do { (void)cprev; } while (0)
#endif
    double *pcol = rhs;  // old pivot-column entries of the pivot in flight (the RHS copy is only made at the end)

    // One pivot = five CTA barriers.  fp64 division has a latency of several hundred cycles and does not
    // interleave (each one carries a slow-path branch), so every pass that divides is spread over the
    // whole CTA, one element per thread: (S2) the phase-1 column scan / the ratio test, (D) pivot-row
    // normalisation + pivot-column rewrite.  The cheap decisions (S1, S3) stay in warp 0's registers.
    for (;;) {
        // ---- S1 (warp 0): leaving row (phase 1: simplex.ts:38-54) or entering column (phase 2: 129-269)
        if (warp == 0) {
            const double *cost = Ms + (size_t)slot[0] * Ws;
            int go = 1;
            rstar = -1; cstar = -1; isneg = 0;
            if (p1 + p2 >= nb.max_pivots || log_n >= nb.log_cap) { overflow = 1; go = 0; }
            if (go && phase == 1) {
                VI b = {-prec, INT_MAX};
#pragma unroll 2
                for (int r = 1 + lane; r < Hn; r += 32) {
                    const double v = Ms[(size_t)slot[r] * Ws];
                    if (v < b.v) { b.v = v; b.i = r; }
                }
                b = warp_reduce_vi<true>(b);
                if (b.i == INT_MAX) phase = 2;  // feasible (simplex.ts:51-54)
                else rstar = b.i;
            }
            if (go && rstar < 0) {  // phase 2
                if (next_c < 0) next_c = warp_price(T, cost, vcol, W, lane, &next_neg);
                if (next_c == 0) { status = ST_OPTIMAL; go = 0; }  // simplex.ts:265-269
                else { cstar = next_c; isneg = next_neg; }
            }
            if (lane == 0) {
                sc.mode = !go ? 0 : (rstar >= 0 ? 1 : 2);
                sc.r = rstar; sc.c = cstar; sc.neg = isneg;
                if (!go) {
                    s_fin[0] = status; s_fin[1] = p1; s_fin[2] = p2; s_fin[3] = log_n; s_fin[4] = overflow; s_fin[5] = unb;
                    s_eval = cost[0];
                }
            }
        }
        NODE_CY(0);
        __syncthreads();
        const int mode = sc.mode;
        if (mode == 0) break;
        // ---- S2 (CTA): the scan that divides, one element per thread --------------------------------
        if (mode == 1) {  // entering column of a phase-1 pivot (simplex.ts:56-76)
            const double *cost = Ms + (size_t)slot[0] * Ws;
            const double *lrow = Ms + (size_t)slot[sc.r] * Ws;
            VI e = {-INFINITY, INT_MAX};
#pragma unroll 1
            for (int c = 1 + tid; c < W; c += NT) {
                const double coef = lrow[c];
                if ((has_unres && is_unres(T, vcol[c])) || coef < -prec) {
                    const double quo = ddiv_z(-cost[c], coef);
                    if (e.v < quo) { e.v = quo; e.i = c; }
                }
            }
            e = warp_reduce_vi<false>(e);
            if (lane == 0) { part.v[warp] = e.v; part.i[warp] = e.i; }
        } else {          // ratio test (simplex.ts:271-296)
            const int cs = sc.c, neg = sc.neg;
            VI m = {INFINITY, INT_MAX};
            int dmin = INT_MAX, cnt = 0;
#pragma unroll 1
            for (int r = tid; r < Hn; r += NT) {
                const double *row = Ms + (size_t)slot[r] * Ws;
                const double col = row[cs], rv = row[0];
                if (nz16(col)) cnt++;
                if (r == 0) continue;
                if (-prec < col && col < prec) continue;
                if (col > 0 && prec > rv && rv > -prec) { dmin = min(dmin, r); continue; }
                const double quo = ddiv_z(neg ? -rv : rv, col);  // (-rhs) / col == -rhs / col
                if (quo > prec && m.v > quo) { m.v = quo; m.i = r; }
            }
            dmin = __reduce_min_sync(0xffffffffu, dmin);
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            m = warp_reduce_vi<true>(m);
            if (lane == 0) { part.v[warp] = m.v; part.i[warp] = m.i; part.d[warp] = dmin; part.n[warp] = cnt; }
        }
        NODE_CY(1);
        __syncthreads();
        // ---- S3 (warp 0): combine the partials, stage the pivot ------------------------------------
        if (warp == 0) {
            int go = 1, cnt = 0;
            if (mode == 1) {
                VI e = {-INFINITY, INT_MAX};
                if (lane < NW) { e.v = part.v[lane]; e.i = part.i[lane]; }
                e = warp_reduce_vi<false>(e);
                if (e.i == INT_MAX) { status = ST_INFEASIBLE; go = 0; }  // simplex.ts:73-76
                else {
                    cstar = e.i;
#pragma unroll 1
                    for (int r = lane; r < Hn; r += 32)  // non-zero pivot-column entries (lazy-flush flag)
                        if (nz16(Ms[(size_t)slot[r] * Ws + cstar])) cnt++;
                    cnt = __reduce_add_sync(0xffffffffu, cnt);
                }
            } else {
                VI m = {INFINITY, INT_MAX};
                int dmin = INT_MAX;
                if (lane < NW) { m.v = part.v[lane]; m.i = part.i[lane]; dmin = part.d[lane]; cnt = part.n[lane]; }
                dmin = __reduce_min_sync(0xffffffffu, dmin);
                cnt = __reduce_add_sync(0xffffffffu, cnt);
                if (dmin != INT_MAX) rstar = dmin;
                else {
                    m = warp_reduce_vi<true>(m);
                    if (m.i != INT_MAX) rstar = m.i;
                    else { status = ST_UNBOUNDED; unb = vcol[cstar]; go = 0; }  // simplex.ts:298-303
                }
            }
            if (go) {
                const double q = Ms[(size_t)slot[rstar] * Ws + cstar];
                if (lane == 0) {
                    const int leaving = vrow[rstar], entering = vcol[cstar];
                    plog[log_n] = make_int4(rstar | (phase == 2 ? (1 << 30) : 0), cstar, leaving, entering);
                    vrow[rstar] = entering;  // simplex.ts:339-349
                    vcol[cstar] = leaving;
                    piv.go = 1; piv.r = rstar; piv.c = cstar; piv.q = q;
                    piv.flush = (cnt - (nz16(q) ? 1 : 0)) > 0;
                }
                log_n++;
                if (phase == 1) p1++; else p2++;
                next_c = -1;
            } else if (lane == 0) {
                piv.go = 0;
                s_fin[0] = status; s_fin[1] = p1; s_fin[2] = p2; s_fin[3] = log_n; s_fin[4] = overflow; s_fin[5] = unb;
                s_eval = Ms[(size_t)slot[0] * Ws];
            }
        }
        NODE_CY(2);
        __syncthreads();
        if (!piv.go) break;
        const int prs = piv.r, pcs = piv.c, flush = piv.flush, spare = s_spare;
        const double q = piv.q;
        // ---- D (CTA): both division passes at once: the lower half of the CTA normalises the pivot row
        // (simplex.ts:352-364 + lazy flush 380-382), the upper half rewrites the pivot column
        // (simplex.ts:371-374,386) after staging its old entries for the row updates.
        if (tid < NT / 2) {
            const double *praw = Ms + (size_t)slot[prs] * Ws;
#pragma unroll 1
            for (int c = tid; c < W; c += NT / 2) {
                const double v = praw[c];
                double f = (c == pcs || nz16(v)) ? ddiv_z(c == pcs ? 1.0 : v, q) : 0.0;
                if (flush && !nz16(f) && f != 0.0) f = 0.0;
                frow[c] = f;
            }
        } else {
#pragma unroll 1
            for (int r = tid - NT / 2; r < Hn; r += NT / 2) {
                if (r == prs) { pcol[r] = 0.0; continue; }
                double *e = Ms + (size_t)slot[r] * Ws + pcs;
                const double coef = *e;
                pcol[r] = coef;
                if (nz16(coef)) *e = ddiv_z(-coef, q);  // else untouched (simplex.ts:371; :389-391 is dead code)
            }
        }
        NODE_CY(3);
        __syncthreads();
        // ---- U (CTA): simplex.ts:367-391.  Warp 0 owns the cost row and then prices it (look-ahead, off
        // the other warps' critical path); rows 1.. are dealt over the other warps.  A lane's columns are
        // the same for every row, so its share of the normalised pivot row lives in registers.
#pragma unroll 1
        for (int cb = 0; cb < W; cb += 128) {
            double fr[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int c = cb + lane + 32 * k;
                fr[k] = c < W ? frow[c] : 0.0;
            }
            const int rfirst = warp == 0 ? 0 : warp, rstep = warp == 0 ? Hn : NW - 1;
#pragma unroll 1
            for (int r = rfirst; r < Hn; r += rstep) {
                if (r == prs) {
                    double *dst = Ms + (size_t)spare * Ws;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int c = cb + lane + 32 * k;
                        if (c < W) dst[c] = fr[k];
                    }
                    continue;
                }
                const double coef = pcol[r];
                if (!nz16(coef)) continue;
                double *row = Ms + (size_t)slot[r] * Ws;
                double v[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int c = cb + lane + 32 * k;
                    v[k] = c < W ? row[c] : 0.0;
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int c = cb + lane + 32 * k;
                    if (c < W && c != pcs && nz16(fr[k])) row[c] = __dsub_rn(v[k], __dmul_rn(coef, fr[k]));
                }
            }
        }
        if (warp == 0) {
            __syncwarp();
            next_c = warp_price(T, Ms + (size_t)slot[0] * Ws, vcol, W, lane, &next_neg);
        }
        NODE_CY(4);
        __syncthreads();
        NODE_CY(5);
        if (tid == 0) {
            const int old = slot[prs];
            slot[prs] = spare;
            s_spare = old;
        }
        __syncwarp();  // warp 0 reads the slot table next; every other warp waits at the next barrier
    }

    if (tid == 0) tl3 = globaltimer_ns() - t_start;
    // isIntegral / most fractional variable over the final RHS column
    for (int r = tid; r < Hn; r += NT) rhs[r] = Ms[(size_t)slot[r] * Ws];
    __syncthreads();
    if (T.intpos != nullptr) {
        if (tid == 0) { T.M = rhs; T.stride = 1; T.vrow = vrow; T.H = Hn; }
        __syncthreads();
        cta_mip_scan(T, &mip, sel.red);
    } else if (tid == 0) { mip.is_integral = 1; mip.var_index = -1; mip.value = 0.0; }
    __syncthreads();
    if (tid == 0) tl4 = globaltimer_ns() - t_start;
    NodeOut *out = nb.out + node;
    {
        const int nlog = min(min(s_fin[3], nb.log_cap), NODE_LOG_HEAD);
        for (int k = tid; k < nlog; k += NT) out->log_head[k] = plog[k];
    }
    if (nb.wb_M != nullptr && node == 0) {
        for (int r = warp; r < Hn; r += NW) {
            const double *row = Ms + (size_t)slot[r] * Ws;
            for (int c = lane; c < W; c += 32) nb.wb_M[(size_t)r * nb.root_stride + c] = row[c];
        }
        for (int r = tid; r < Hn; r += NT) nb.wb_vrow[r] = vrow[r];
        for (int c = tid; c < W; c += NT) nb.wb_vcol[c] = vcol[c];
    }
    if (tid == 0) {
        NodeResult r;
        r.status = s_fin[0]; r.p1 = s_fin[1]; r.p2 = s_fin[2]; r.log_n = s_fin[3]; r.overflow = s_fin[4];
        r.is_integral = mip.is_integral; r.branch_var = mip.var_index; r.unbounded_var = s_fin[5];
        r.eval_raw = s_eval; r.branch_value = mip.value;
        r.t_ns = globaltimer_ns() - t_start; r.pad = 0;
        r.tl[0] = tl0; r.tl[1] = tl1; r.tl[2] = tl2; r.tl[3] = tl3; r.tl[4] = tl4; r.tl[5] = 0;
        for (int k = 0; k < 6; k++) r.cy[k] = cy[k];
        out->r = r;
    }
}

}  // namespace jslp
