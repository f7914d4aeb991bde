// Micro-benchmark: issue/throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench_ffma2 bench_ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
template <int MODE> __global__ void k(float* out, int iters, float s) {
    float a[8], b = s;
    unsigned long long p[8], pb = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(s);
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x + i; p[i] = (unsigned long long)(threadIdx.x + i) * 0x100000001ull; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i] = fmaf(a[i], b, b); }                     // 8 FFMA  =  8 flop-pairs
            else if (MODE == 1) { p[i] = fma2(p[i], pb, pb); }              // 8 FFMA2 = 16 flop-pairs
            else { a[i] = fmaf(a[i], b, b); p[i] = fma2(p[i], pb, pb); }    // mixed
        }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> float run(float* d, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(d, 10, 1.0001f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(d, iters, 1.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    int iters = 20000;
    double lanes = 148.0 * 8 * 256 * iters * 8;
    float t0 = run<0>(d, iters), t1 = run<1>(d, iters), t2 = run<2>(d, iters);
    printf("FFMA : %.3f ms  %.2f Tfma/s (lane-fma)\n", t0, lanes / t0 / 1e9);
    printf("FFMA2: %.3f ms  %.2f Tfma/s (2 fma per lane-instr)\n", t1, 2 * lanes / t1 / 1e9);
    printf("mixed: %.3f ms  %.2f Tfma/s\n", t2, 3 * lanes / t2 / 1e9);
    return 0;
}
