// jslpsolver_b200/csrc/jslp_node_kernel.cuh -- shared-memory-resident node LPs.
//
// One CTA solves one branch-and-cut node (or one small LP) start to finish with the whole tableau
// in shared memory: restore from the root snapshot (backup.ts:53-105), append the node's cut rows
// (cutting-strategies.ts:36-71), run phase 1 / phase 2 to completion (simplex.ts:14-325) and
// report the summary the frontier needs (evaluation, isIntegral, most fractional variable;
// mip-utils.ts:43-61,100-126).  A launch evaluates a whole batch of open nodes, one CTA each:
// this is what makes the node frontier GPU-resident.  Selection code is the same template as the
// HBM path (cta_select<false>), so both paths obey identical tie-break rules.
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
};

// What a CTA reports: the summary plus the head of its pivot log (enough for the cycle check of
// almost every node), written straight into mapped pinned host memory -- no D2H copy per round.
constexpr int NODE_LOG_HEAD = 64;
struct NodeOut {
    NodeResult r;
    int4 log_head[NODE_LOG_HEAD];
};

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
};

constexpr int NODE_THREADS = 256;

__device__ __forceinline__ long long globaltimer_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(NODE_THREADS) k_node_batch(const TabDev *Tp, NodeBatchDev nb) {
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ TabDev T;
    __shared__ Rec rec;
    __shared__ SelSmem sel;
    __shared__ MipOut mip;
    const int tid = threadIdx.x, NT = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, NW = NT >> 5;
    const int node = blockIdx.x;
    long long t_start = 0;
    if (tid == 0) t_start = globaltimer_ns();
    // issued first, used after the restore: these two may cross PCIe
    const int c0 = nb.cut_off[node], nc = nb.cut_off[node + 1] - c0;
    const int H0 = nb.H0, Ws = nb.Ws;

    double *Ms = reinterpret_cast<double *>(smraw);
    double *prow = Ms + (size_t)nb.Hcap * Ws;
    double *frow = prow + Ws;
    double *pcol = frow + Ws;
    CutDev *cutS = reinterpret_cast<CutDev *>(pcol + nb.Hcap);
    int *vrow = reinterpret_cast<int *>(cutS + (nb.Hcap - H0));
    if (tid == 0) {
        T = *Tp;
        T.vcol = vrow + nb.Hcap;
    }
    __syncthreads();
    const int W = T.W;
    int *vcol = T.vcol;

    // restore(): root snapshot -> shared memory, eight loads in flight per thread before their stores
    {
        int r = 0, c = tid;
        while (c >= W) { c -= W; r++; }
        const int total = H0 * W;
        for (int i0 = tid; i0 < total; i0 += 8 * NT) {
            double v[8];
            int rr[8], cc[8];
#pragma unroll
            for (int k = 0; k < 8; k++) {
                rr[k] = r; cc[k] = c;
                v[k] = (i0 + k * NT < total) ? __ldg(nb.rootM + (size_t)r * nb.root_stride + c) : 0.0;
                c += NT;
                while (c >= W) { c -= W; r++; }
            }
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (i0 + k * NT < total) Ms[rr[k] * Ws + cc[k]] = v[k];
        }
    }
    for (int r = tid; r < H0; r += NT) vrow[r] = __ldg(nb.root_vrow + r);
    for (int c = tid; c < W; c += NT) vcol[c] = __ldg(nb.root_vcol + c);
    for (int h = tid; h < nc; h += NT) cutS[h] = nb.cuts[c0 + h];
    const int Hn = H0 + nc;
    __syncthreads();
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
    if (tid == 0) {
        T.M = Ms; T.vrow = vrow; T.prow = prow; T.pcol = pcol;
        T.stride = Ws; T.H = Hn; T.rowcap = nb.Hcap;
        T.plog = nb.logs + (size_t)node * nb.log_cap;
        T.plog_cap = nb.log_cap;
        rec.status = ST_RUNNING; rec.phase = 1; rec.has_pivot = 0; rec.r = rec.c = 0; rec.is_neg = 0;
        rec.flush = 0; rec.done = 0; rec.p1 = rec.p2 = 0; rec.stop_at = -1; rec.log_n = 0;
        rec.unbounded_var = -1; rec.only_phase = 0; rec.ticket = 0; rec.q = 0; rec.eval_raw = 0;
    }
    __syncthreads();

    int overflow = 0;
    for (;;) {
        cta_select<false>(T, &rec, sel);
        __syncthreads();
        if (!rec.has_pivot) break;
        if (rec.done >= nb.max_pivots || rec.log_n > nb.log_cap) { overflow = 1; break; }
        const int rstar = rec.r, cstar = rec.c, flush = rec.flush, phase = rec.phase;
        const double q = rec.q;
        for (int c = tid; c < W; c += NT) {  // simplex.ts:352-364 (+ lazy flush 380-382)
            const double v = prow[c];
            double f = nz16(v) ? v / q : 0.0;
            if (c == cstar) f = 1.0 / q;
            if (flush && !nz16(f) && f != 0.0) f = 0.0;
            frow[c] = f;
        }
        __syncthreads();
        if (tid == 0) {  // every thread has read the record before the barrier above
            rec.done += 1;
            if (phase == 1) rec.p1 += 1; else rec.p2 += 1;
            rec.has_pivot = 0;
        }
        // simplex.ts:367-391, one row per warp: rows with a zero pivot-column entry cost one test
        for (int r = warp; r < Hn; r += NW) {
            double *row = Ms + r * Ws;
            if (r == rstar) {
                for (int c = lane; c < W; c += 32) row[c] = frow[c];
                continue;
            }
            const double coef = pcol[r];
            if (nz16(coef)) {
                for (int c = lane; c < W; c += 32) {
                    if (c == cstar) { row[c] = -coef / q; continue; }
                    const double v0 = frow[c];
                    if (nz16(v0)) row[c] = __dsub_rn(row[c], __dmul_rn(coef, v0));
                }
            } else if (coef != 0.0 && lane == 0) {
                row[cstar] = 0.0;
            }
        }
        __syncthreads();
    }

    if (T.intpos != nullptr) cta_mip_scan(T, &mip, sel.red);
    else if (tid == 0) { mip.is_integral = 1; mip.var_index = -1; mip.value = 0.0; }
    __syncthreads();
    NodeOut *out = nb.out + node;
    {
        const int nlog = min(min(rec.log_n, nb.log_cap), NODE_LOG_HEAD);
        for (int k = tid; k < nlog; k += NT) out->log_head[k] = T.plog[k];
    }
    if (nb.wb_M != nullptr && node == 0) {
        for (int i = tid; i < Hn * W; i += NT) {
            const int r = i / W, c = i - r * W;
            nb.wb_M[(size_t)r * nb.root_stride + c] = Ms[r * Ws + c];
        }
        for (int r = tid; r < Hn; r += NT) nb.wb_vrow[r] = vrow[r];
        for (int c = tid; c < W; c += NT) nb.wb_vcol[c] = vcol[c];
    }
    if (tid == 0) {
        NodeResult r;
        r.status = rec.status; r.p1 = rec.p1; r.p2 = rec.p2; r.log_n = rec.log_n; r.overflow = overflow;
        r.is_integral = mip.is_integral; r.branch_var = mip.var_index; r.unbounded_var = rec.unbounded_var;
        r.eval_raw = rec.eval_raw; r.branch_value = mip.value;
        r.t_ns = globaltimer_ns() - t_start; r.pad = 0;
        out->r = r;
    }
}

}  // namespace jslp
