// Micro-benchmark: throughput of fma.rn.f32x2 by operand form on sm_100a -- all operands distinct register pairs,
// one operand a scalar broadcast (SASS .F32), operands shared between neighbouring instructions (reuse cache).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench_ffma2_forms bench_ffma2_forms.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ void fma2_acc(u64& acc, u64 a, u64 b) { asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }
__device__ __forceinline__ void fma2_bc(u64& acc, u64 a, float s) {
    asm volatile("{\n.reg .b64 t;\nmov.b64 t, {%2, %2};\nfma.rn.f32x2 %0, %1, t, %0;\n}" : "+l"(acc) : "l"(a), "f"(s));
}
__device__ __forceinline__ u64 mk(float x, float y) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r; }
// MODE 0: acc[i] += x[j] * y[j] (three distinct pairs, x/y change every instruction)
// MODE 1: acc[i] += x[j] * bc(s[j])   (pair, scalar broadcast, pair)
// MODE 2: acc[i] += x[0] * y[0]       (operands shared by all instructions: reuse cache)
// MODE 3: like 1, but the x operand is shared by two neighbouring instructions (the combine loop's pattern)
template <int MODE> __global__ void k(float* out, int iters, const float* __restrict__ in) {
    u64 acc[8], x[4], y[4];
    float s[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = mk(threadIdx.x + i, 1.f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float* q = in + (threadIdx.x & 31) * 20 + i * 5;            // opaque run-time values, different per lane
        x[i] = mk(q[0], q[1]); y[i] = mk(q[2], q[3]); s[i] = q[4];
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int j = (i + rep) & 3;
                if (MODE == 0) fma2_acc(acc[i], x[j], y[(j + rep) & 3]);
                else if (MODE == 1) fma2_bc(acc[i], x[j], s[(j + rep) & 3]);
                else if (MODE == 2) fma2_acc(acc[i], x[0], y[0]);
                else fma2_bc(acc[i], x[(i >> 1) & 3], s[(i + rep) & 3]);
            }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += __uint_as_float((unsigned)acc[i]) + __uint_as_float((unsigned)(acc[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> float run(float* d, int iters, int ctas_per_sm, int threads) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * ctas_per_sm, threads>>>(d, 10, d + 148 * 8 * 1024);
    cudaEventRecord(e0);
    k<MODE><<<148 * ctas_per_sm, threads>>>(d, iters, d + 148 * 8 * 1024);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 1024 * 4 + 4096); cudaMemset(d, 0, 148 * 8 * 1024 * 4 + 4096);
    const int iters = 5000;
    const char* names[4] = {"3 distinct pairs", "scalar broadcast", "shared operands ", "combine pattern "};
    for (int cfg = 0; cfg < 2; ++cfg) {
        const int ctas = cfg ? 1 : 8, threads = cfg ? 640 : 256;
        const double instr = 148.0 * ctas * (threads / 32) * iters * 32.0;        // warp instructions
        float t[4] = {run<0>(d, iters, ctas, threads), run<1>(d, iters, ctas, threads), run<2>(d, iters, ctas, threads),
                      run<3>(d, iters, ctas, threads)};
        for (int m = 0; m < 4; ++m)
            printf("%d warps/SM  %s: %.3f ms  %.2f cycles per FFMA2 per SMSP (at 1.965 GHz)  %.1f TFLOP/s\n", ctas * threads / 32,
                   names[m], t[m], t[m] * 1e-3 * 1.965e9 / (instr / (148.0 * 4)), instr * 128 / t[m] / 1e9);
    }
    return 0;
}
