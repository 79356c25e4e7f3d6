// Launch-cadence probe: N dependent launches of a tiny kernel as (a) a linear CUDA graph of N kernel
// nodes, (b) a graph WHILE node whose body is the kernel (device-side cudaGraphSetConditional).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)
__global__ void body(int *ctr, int limit, cudaGraphConditionalHandle h, int use_h) {
    __shared__ int dummy;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const int v = atomicAdd(ctr, 1) + 1;
        if (use_h) cudaGraphSetConditional(h, v < limit ? 1 : 0);
    }
    dummy = 0;
}
int main() {
    int *ctr; CK(cudaMalloc(&ctr, 4));
    cudaStream_t s; CK(cudaStreamCreate(&s));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int N = 1024;
    for (int grid : {1, 296}) {
        // (a) linear graph
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal));
        for (int i = 0; i < N; i++) body<<<grid, 256, 0, s>>>(ctr, N, 0, 0);
        CK(cudaStreamEndCapture(s, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaMemsetAsync(ctr, 0, 4, s));
            CK(cudaEventRecord(e0, s)); CK(cudaGraphLaunch(ge, s)); CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep == 2) printf("grid %3d linear graph: %.3f us/launch\n", grid, 1e3 * ms / N);
        }
        // (b) while graph
        cudaGraph_t gw; CK(cudaGraphCreate(&gw, 0));
        cudaGraphConditionalHandle h;
        CK(cudaGraphConditionalHandleCreate(&h, gw, 1, cudaGraphCondAssignDefault));
        cudaGraphNodeParams p = {};
        p.type = cudaGraphNodeTypeConditional;
        p.conditional.handle = h;
        p.conditional.type = cudaGraphCondTypeWhile;
        p.conditional.size = 1;
        cudaGraphNode_t wn;
        CK(cudaGraphAddNode(&wn, gw, nullptr, 0, &p));
        cudaGraph_t bodyg = p.conditional.phGraph_out[0];
        cudaKernelNodeParams kp = {};
        int lim = N, useh = 1;
        void *args[] = {&ctr, &lim, &h, &useh};
        kp.func = (void *)body; kp.gridDim = dim3(grid); kp.blockDim = dim3(256); kp.kernelParams = args;
        cudaGraphNode_t kn;
        CK(cudaGraphAddKernelNode(&kn, bodyg, nullptr, 0, &kp));
        cudaGraphExec_t gwe;
        CK(cudaGraphInstantiate(&gwe, gw, 0));
        for (int rep = 0; rep < 3; rep++) {
            CK(cudaMemsetAsync(ctr, 0, 4, s));
            CK(cudaEventRecord(e0, s)); CK(cudaGraphLaunch(gwe, s)); CK(cudaEventRecord(e1, s)); CK(cudaStreamSynchronize(s));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            int hc; cudaMemcpy(&hc, ctr, 4, cudaMemcpyDeviceToHost);
            if (rep == 2) printf("grid %3d while graph:  %.3f us/iteration (%d iterations)\n", grid, 1e3 * ms / hc, hc);
        }
    }
    return 0;
}
