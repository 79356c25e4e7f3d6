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
    // zero dividend: does 0/y leave the division fast path?
    t0 = clock64();
    for (int i = 0; i < n; i++) x = (x - x) / y + a;
    t1 = clock64(); if (lane == 0) out[10] = (t1 - t0);
    // half the lanes divide zero
    t0 = clock64();
    for (int i = 0; i < n; i++) x = ((lane & 1) ? x : 0.0) / y + a;
    t1 = clock64(); if (lane == 0) out[11] = (t1 - t0);
    // tiny quotient (1e-300 scale)
    double w = 1e-300 * (1 + lane);
    t0 = clock64();
    for (int i = 0; i < n; i++) w = w / y;
    t1 = clock64(); if (lane == 0) out[12] = (t1 - t0);
    // huge divisor / small dividend mix, typical tableau magnitudes
    w = 3.5 + lane;
    t0 = clock64();
    for (int i = 0; i < n; i++) w = 1.0 / (w + 1e-3);
    t1 = clock64(); if (lane == 0) out[13] = (t1 - t0);
    sink[threadIdx.x] = x + u + idx + w;
}
int main() {
    long long *out; double *sink;
    cudaMallocManaged(&out, 160); cudaMalloc(&sink, 8 * 256);
    const int n = 256;
    const char *names[] = {"DFMA", "DADD", "DDIV+DADD", "4xDDIV+DADD (per iter)", "REDUX(+IADD)", "SHFL(+IADD)", "DSETP+SEL+DADD", "LDS chase", "BAR.SYNC", "DMUL+DSUB", "0/y + DADD", "half lanes 0/y + DADD", "1e-300/y", "1/(w+eps)"};
    for (int threads : {32, 256}) {
        k<<<1, threads>>>(1.000001, 0.999999, n, out, sink); cudaDeviceSynchronize();
        k<<<1, threads>>>(1.000001, 0.999999, n, out, sink); cudaDeviceSynchronize();
        printf("threads %d\n", threads);
        for (int i = 0; i < 14; i++) printf("  %-26s %.1f cycles\n", names[i], (double)out[i] / n);
    }
    return 0;
}
