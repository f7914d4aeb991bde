// Micro-benchmark: throughput of the legacy tensor path (mma.sync -> HMMA) on sm_100a for the shapes the channel
// kernel could use: m16n8k8 TF32 and m16n8k16 FP16 / BF16, fp32 accumulate, four independent accumulator chains per warp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench_hmma bench_hmma.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(float* out, int iters, const unsigned* __restrict__ in) {
    float c[4][4];
    unsigned a[4], b[2];
    for (int i = 0; i < 4; ++i) { a[i] = in[threadIdx.x % 32 + i * 32]; for (int j = 0; j < 4; ++j) c[j][i] = 0.f; }
    b[0] = in[200 + threadIdx.x % 32]; b[1] = in[300 + threadIdx.x % 32];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (MODE == 0)
                    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
                else if (MODE == 1)
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
                else
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            }
    }
    float r = 0;
    for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i) r += c[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE> float run(float* d, int iters, int threads) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148, threads>>>(d, 10, (const unsigned*)(d + 148 * 1024));
    cudaEventRecord(e0);
    k<MODE><<<148, threads>>>(d, iters, (const unsigned*)(d + 148 * 1024));
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 1024 * 4 + 4096); cudaMemset(d, 0, 148 * 1024 * 4 + 4096);
    const int iters = 5000;
    const char* names[3] = {"m16n8k8  tf32", "m16n8k16 f16 ", "m16n8k16 bf16"};
    for (int cfg = 0; cfg < 3; ++cfg) {
        const int threads = cfg == 0 ? 128 : cfg == 1 ? 640 : 1024;
        const double per_smsp = (threads / 32) / 4.0 * iters * 16.0;        // HMMA per scheduler
        float t[3] = {run<0>(d, iters, threads), run<1>(d, iters, threads), run<2>(d, iters, threads)};
        for (int m = 0; m < 3; ++m)
            printf("%2d warps/SM  %s: %.3f ms  %.2f cycles per HMMA per SMSP (at 1.965 GHz)\n", threads / 32, names[m], t[m],
                   t[m] * 1e-3 * 1.965e9 / per_smsp);
    }
    return 0;
}
