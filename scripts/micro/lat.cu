// Single-warp latency probes (cycles per dependent op) for the ops the resident node kernel chains.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double a, double b, int n, long long *out, double *sink) {
    __shared__ double sm[256];
    __shared__ int si[256];
    const int lane = threadIdx.x;
    sm[lane] = a + lane; si[lane] = (lane * 7) & 31;
    __syncthreads();
    double x = a + lane, y = b;
    long long t0, t1;
    // DFMA chain
    t0 = clock64();
    for (int i = 0; i < n; i++) x = fma(x, y, a);
    t1 = clock64(); if (lane == 0) out[0] = (t1 - t0);
    // DADD chain
    t0 = clock64();
    for (int i = 0; i < n; i++) x = x + y;
    t1 = clock64(); if (lane == 0) out[1] = (t1 - t0);
    // DDIV chain
    t0 = clock64();
    for (int i = 0; i < n; i++) x = x / y + a;
    t1 = clock64(); if (lane == 0) out[2] = (t1 - t0);
    // 4 independent DDIV
    double z0 = x, z1 = x + 1, z2 = x + 2, z3 = x + 3;
    t0 = clock64();
    for (int i = 0; i < n; i++) { z0 = z0 / y + a; z1 = z1 / y + a; z2 = z2 / y + a; z3 = z3 / y + a; }
    t1 = clock64(); if (lane == 0) out[3] = (t1 - t0);
    x = z0 + z1 + z2 + z3;
    // REDUX chain
    unsigned int u = (unsigned int)lane + (unsigned int)x;
    t0 = clock64();
    for (int i = 0; i < n; i++) u = __reduce_min_sync(0xffffffffu, u + lane);
    t1 = clock64(); if (lane == 0) out[4] = (t1 - t0);
    // SHFL chain
    t0 = clock64();
    for (int i = 0; i < n; i++) u = __shfl_xor_sync(0xffffffffu, u, 1) + 1;
    t1 = clock64(); if (lane == 0) out[5] = (t1 - t0);
    // DSETP + select chain
    t0 = clock64();
    for (int i = 0; i < n; i++) x = (x < y) ? x + 1.0 : y;
    t1 = clock64(); if (lane == 0) out[6] = (t1 - t0);
    // dependent LDS chain (pointer chase)
    int idx = lane;
    t0 = clock64();
    for (int i = 0; i < n; i++) idx = si[idx];
    t1 = clock64(); if (lane == 0) out[7] = (t1 - t0);
    // __syncthreads with 8 warps is measured by the caller configuration (blockDim 256)
    t0 = clock64();
    for (int i = 0; i < n; i++) __syncthreads();
    t1 = clock64(); if (lane == 0) out[8] = (t1 - t0);
    // DMUL->DSUB pair (the update)
    t0 = clock64();
    for (int i = 0; i < n; i++) x = __dsub_rn(x, __dmul_rn(y, x));
    t1 = clock64(); if (lane == 0) out[9] = (t1 - t0);
    sink[threadIdx.x] = x + u + idx;
}
int main() {
    long long *out; double *sink;
    cudaMallocManaged(&out, 80); cudaMalloc(&sink, 8 * 256);
    const int n = 256;
    const char *names[] = {"DFMA", "DADD", "DDIV+DADD", "4xDDIV+DADD (per iter)", "REDUX(+IADD)", "SHFL(+IADD)", "DSETP+SEL+DADD", "LDS chase", "BAR.SYNC", "DMUL+DSUB"};
    for (int threads : {32, 256}) {
        k<<<1, threads>>>(1.000001, 0.999999, n, out, sink); cudaDeviceSynchronize();
        k<<<1, threads>>>(1.000001, 0.999999, n, out, sink); cudaDeviceSynchronize();
        printf("threads %d\n", threads);
        for (int i = 0; i < 10; i++) printf("  %-26s %.1f cycles\n", names[i], (double)out[i] / n);
    }
    return 0;
}
