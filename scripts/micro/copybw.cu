// scripts/micro/copybw.cu -- ping-pong copy bandwidth by load/store flavour, footprint and launch geometry.
// Which streaming flavour suits the HBM regime (tableau pairs beyond L2)?  nvcc -O3 -arch=sm_100a copybw.cu -o copybw
#include <cstdio>
#include <cuda_runtime.h>
// element = 16 bytes (MODE 0..3: 128-bit accesses) or 32 bytes (MODE 4..5: the 256-bit accesses sm_100 adds)
struct E32 { unsigned long long a, b, c, d; };
template <int MODE> struct Elem { typedef double2 T; static constexpr int BYTES = 16; };
template <> struct Elem<4> { typedef E32 T; static constexpr int BYTES = 32; };
template <> struct Elem<5> { typedef E32 T; static constexpr int BYTES = 32; };
template <> struct Elem<8> { typedef E32 T; static constexpr int BYTES = 32; };
template <> struct Elem<9> { typedef E32 T; static constexpr int BYTES = 32; };
template <int MODE> __device__ __forceinline__ typename Elem<MODE>::T ld(const char *p) {
    typename Elem<MODE>::T v;
    if constexpr (MODE == 0) asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    if constexpr (MODE == 1) asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    if constexpr (MODE == 2) asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    if constexpr (MODE == 3) asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    if constexpr (MODE == 4) asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p));
    if constexpr (MODE == 5 || MODE == 8 || MODE == 9) asm volatile("ld.global.L2::evict_first.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(v.a), "=l"(v.b), "=l"(v.c), "=l"(v.d) : "l"(p));
    if constexpr (MODE == 6) asm volatile("ld.global.cs.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    if constexpr (MODE == 7) asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
template <int MODE> __device__ __forceinline__ void st(char *p, typename Elem<MODE>::T v) {
    if constexpr (MODE == 8) asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(v.a), "l"(v.b), "l"(v.c), "l"(v.d) : "memory");
    if constexpr (MODE == 9) asm volatile("st.global.L2::evict_last.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(v.a), "l"(v.b), "l"(v.c), "l"(v.d) : "memory");
    if constexpr (MODE == 7) asm volatile("st.global.wb.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
    if constexpr (MODE == 0 || MODE == 2 || MODE == 6) asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
    if constexpr (MODE == 1) asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
    if constexpr (MODE == 3) asm volatile("st.global.cg.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
    if constexpr (MODE == 4) asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(v.a), "l"(v.b), "l"(v.c), "l"(v.d) : "memory");
    if constexpr (MODE == 5) asm volatile("st.global.L2::evict_first.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(v.a), "l"(v.b), "l"(v.c), "l"(v.d) : "memory");
}
template <int MODE, int K, int INTERLEAVE>
__global__ void __launch_bounds__(256) copyk(const double *src_, double *dst_, size_t bytes) {
    constexpr int EB = Elem<MODE>::BYTES;
    const char *src = (const char *)src_;
    char *dst = (char *)dst_;
    const size_t n = bytes / EB;
    const int NT = blockDim.x;
    size_t lo, lim, step;
    if (INTERLEAVE) { lo = (size_t)blockIdx.x * NT * K; lim = n; step = (size_t)gridDim.x * NT * K; }
    else { const size_t per = (n + gridDim.x - 1) / gridDim.x; lo = per * blockIdx.x; lim = lo + per < n ? lo + per : n; step = (size_t)NT * K; }
    typename Elem<MODE>::T cur[K], nxt[K];
#pragma unroll
    for (int j = 0; j < K; j++) { const size_t i = lo + threadIdx.x + (size_t)j * NT; if (i < lim) cur[j] = ld<MODE>(src + EB * i); }
    for (size_t i0 = lo + threadIdx.x; i0 < lim; i0 += step) {
#pragma unroll
        for (int j = 0; j < K; j++) { const size_t i = i0 + step + (size_t)j * NT; if (i < lim) nxt[j] = ld<MODE>(src + EB * i); }
#pragma unroll
        for (int j = 0; j < K; j++) { const size_t i = i0 + (size_t)j * NT; if (i < lim) st<MODE>(dst + EB * i, cur[j]); }
#pragma unroll
        for (int j = 0; j < K; j++) cur[j] = nxt[j];
    }
}
template <int MODE, int K, int IL> double run(double *a, double *b, size_t n2, int grid, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) { copyk<MODE, K, IL><<<grid, 256>>>(a, b, n2 * 16); copyk<MODE, K, IL><<<grid, 256>>>(b, a, n2 * 16); }
    cudaEventRecord(e0);
    for (int i = 0; i < iters; i++) { copyk<MODE, K, IL><<<grid, 256>>>(a, b, n2 * 16); copyk<MODE, K, IL><<<grid, 256>>>(b, a, n2 * 16); }
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return 2.0 * iters * 2.0 * n2 * 16 / (ms * 1e-3) / 1e9;
}
int main() {
    const size_t mbs[] = {32, 40, 50, 60, 75, 100};
    for (size_t mb : mbs) {
        const size_t n2 = (mb << 20) / 16;
        double *a, *b; cudaMalloc(&a, n2 * 16); cudaMalloc(&b, n2 * 16); cudaMemset(a, 0, n2 * 16); cudaMemset(b, 0, n2 * 16);
        const int it = 20;
        printf("2 x %zu MB:", mb);
        for (int g : {296, 592}) {
            printf("\n   [grid %4d] plain %.0f | ldNA+st %.0f | ldEF(cs)+st %.0f | ldNA+st.wb %.0f | K8 ldNA %.0f | 256b ldEF+st %.0f | 256b ldEF+st.evict_last %.0f | 256b all EF %.0f", g,
                   run<0, 4, 0>(a, b, n2, g, it), run<2, 4, 0>(a, b, n2, g, it), run<6, 4, 0>(a, b, n2, g, it), run<7, 4, 0>(a, b, n2, g, it),
                   run<2, 8, 0>(a, b, n2, g, it), run<8, 4, 0>(a, b, n2, g, it), run<9, 4, 0>(a, b, n2, g, it), run<5, 4, 0>(a, b, n2, g, it));
        }
        printf("\n");
        cudaFree(a); cudaFree(b);
    }
    return 0;
}
