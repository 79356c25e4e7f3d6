// jslpsolver_b200/csrc/jslp_comm.cuh -- multi-GPU exchange of the branch-and-cut frontier over NCCL
// (included by jslp_api.cu).
//
// The path shards at node granularity only: per speculative round the ranks all-gather their
// node summaries (128 bytes per node) and, while node LPs are running, all-reduce(min) the incumbent bound so
// that every rank can drop speculative nodes the reference would skip at pop (branch-and-cut.ts:90-92).
// Both collectives run on the context's stream out of one small device staging buffer.
//
// NCCL is resolved at RUN time (dlopen of libnccl.so.2): a single-GPU deployment needs no NCCL at all, and
// inside a process that already loaded a NCCL (PyTorch bundles one) the same library instance is reused.
#pragma once
#include <dlfcn.h>
#include <nccl.h>  // types and enums only; every function is looked up with dlsym

struct NcclApi {
    bool tried = false, ok = false;
    std::string why;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi &nccl_api() {
    static NcclApi a;
    if (a.tried) return a;
    a.tried = true;
    void *h = nullptr;
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
        h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) { a.why = std::string("libnccl not found: ") + (dlerror() ? dlerror() : "?"); return a; }
    auto sym = [&](const char *n) { return dlsym(h, n); };
    a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
    a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
    a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.AllReduce && a.GetErrorString;
    if (!a.ok) a.why = "libnccl lacks a required symbol";
    return a;
}

struct jslp_comm {
    jslp_ctx *ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, n_ranks = 1;
    unsigned char *d_buf = nullptr, *h_buf = nullptr;  // device staging / pinned host staging
    size_t cap = 0;
    int64_t collectives = 0;
};

#define NCK(call)                                                                                        \
    do {                                                                                                 \
        ncclResult_t r_ = (call);                                                                        \
        if (r_ != ncclSuccess) return fail(JSLP_E_CUDA, std::string(#call) + ": " + nccl_api().GetErrorString(r_)); \
    } while (0)

extern "C" int jslp_comm_unique_id(uint8_t *id128) {
    if (!id128) return fail(JSLP_E_INVALID, "id is NULL");
    NcclApi &a = nccl_api();
    if (!a.ok) return fail(JSLP_E_UNSUPPORTED, a.why);
    ncclUniqueId id;
    NCK(a.GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    return JSLP_OK;
}

static int comm_reserve(jslp_comm *c, size_t bytes) {
    if (bytes <= c->cap) return JSLP_OK;
    CK(cudaStreamSynchronize(c->ctx->stream));
    cudaFree(c->d_buf); cudaFreeHost(c->h_buf);
    c->d_buf = nullptr; c->h_buf = nullptr;
    const size_t cap = std::max<size_t>(bytes * 2, 1 << 16);
    CK(cudaMalloc(&c->d_buf, cap));
    CK(cudaMallocHost(&c->h_buf, cap));
    c->cap = cap;
    return JSLP_OK;
}

extern "C" int jslp_comm_create(jslp_ctx *ctx, const uint8_t *id128, int rank, int n_ranks, jslp_comm **out) {
    if (!ctx || !id128 || !out) return fail(JSLP_E_INVALID, "NULL argument");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(JSLP_E_INVALID, "bad rank / n_ranks");
    NcclApi &a = nccl_api();
    if (!a.ok) return fail(JSLP_E_UNSUPPORTED, a.why);
    CK(cudaSetDevice(ctx->device));
    jslp_comm *c = new jslp_comm();
    c->ctx = ctx; c->rank = rank; c->n_ranks = n_ranks;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclResult_t r = a.CommInitRank(&c->comm, n_ranks, id, rank);
    if (r != ncclSuccess) {
        delete c;
        return fail(JSLP_E_CUDA, std::string("ncclCommInitRank: ") + a.GetErrorString(r));
    }
    int rc = comm_reserve(c, 1 << 16);
    if (rc) { a.CommDestroy(c->comm); delete c; return rc; }
    *out = c;
    return JSLP_OK;
}

extern "C" void jslp_comm_destroy(jslp_comm *c) {
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    if (c->comm) nccl_api().CommDestroy(c->comm);
    cudaFree(c->d_buf); cudaFreeHost(c->h_buf);
    delete c;
}

// In-place all-gather of `bytes_per_rank` bytes per rank in the rank-major HOST buffer `buf` (the contract of
// the jslp_bnb_opts.all_gather hook it supersedes).
extern "C" int jslp_comm_all_gather(jslp_comm *c, void *buf, int64_t bytes_per_rank) {
    if (!c || !buf || bytes_per_rank < 0) return fail(JSLP_E_INVALID, "bad argument");
    if (bytes_per_rank == 0) return JSLP_OK;
    const size_t per = (size_t)bytes_per_rank, tot = per * c->n_ranks;
    int rc = comm_reserve(c, tot);
    if (rc) return rc;
    cudaStream_t s = c->ctx->stream;
    CK(cudaSetDevice(c->ctx->device));
    memcpy(c->h_buf, (unsigned char *)buf + per * c->rank, per);
    CK(cudaMemcpyAsync(c->d_buf + per * c->rank, c->h_buf, per, cudaMemcpyHostToDevice, s));
    NCK(nccl_api().AllGather(c->d_buf + per * c->rank, c->d_buf, per, ncclChar, c->comm, s));
    CK(cudaMemcpyAsync(c->h_buf, c->d_buf, tot, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    memcpy(buf, c->h_buf, tot);
    c->collectives++;
    return JSLP_OK;
}

// Element-wise all-reduce(min) of n doubles (host buffer, in place): the incumbent bound and the flags that
// keep the ranks' poll loops in lockstep.
extern "C" int jslp_comm_all_reduce_min(jslp_comm *c, double *vals, int n) {
    if (!c || !vals || n < 0) return fail(JSLP_E_INVALID, "bad argument");
    if (n == 0) return JSLP_OK;
    const size_t bytes = sizeof(double) * (size_t)n;
    int rc = comm_reserve(c, bytes);
    if (rc) return rc;
    cudaStream_t s = c->ctx->stream;
    CK(cudaSetDevice(c->ctx->device));
    memcpy(c->h_buf, vals, bytes);
    CK(cudaMemcpyAsync(c->d_buf, c->h_buf, bytes, cudaMemcpyHostToDevice, s));
    NCK(nccl_api().AllReduce(c->d_buf, c->d_buf, (size_t)n, ncclDouble, ncclMin, c->comm, s));
    CK(cudaMemcpyAsync(c->h_buf, c->d_buf, bytes, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    memcpy(vals, c->h_buf, bytes);
    c->collectives++;
    return JSLP_OK;
}
