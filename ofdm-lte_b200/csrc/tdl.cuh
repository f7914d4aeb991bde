// Pieces shared by the staged TDL kernel (channel.cu) and the fused channel + RX-FFT kernel
// (fused.cu): channel parameters, the Jakes polynomial-coefficient kernel, cp.async and packed
// f32x2 helpers.
#pragma once
#include "common.cuh"

#define TDL_THREADS 128
#define TDL_STAGES 2
#define TDL_MAX_PB 8192

struct TdlParams {
    int num_taps;
    int delay[LTE_MAX_TAPS];
    float gain[LTE_MAX_TAPS];           // includes sqrt(2/16)
    double w_cyc[LTE_JAKES_TONES];      // fD cos(alpha_n) / fs   [cycles per sample]
    int pb;                             // polynomial block length in samples (power of two)
    int pb_log2;
    int nbs;                            // polynomial blocks per stream (every block a tile touches exists)
};

// Taylor coefficients of every (rx, tx, tap) fading process for every polynomial block, laid out
// for the packed-pair TDL kernel:
//   coef[b][blk][t][tap][slot][re|im][R2]   (floats; R2 = antennas rounded up to the group size)
// slot k = 0..K:   sum_n g e^{j theta_n} (j x_n)^k / k!      (value coefficients)
// slot K+k, k>=1:  k times slot k                           (derivative coefficients)
// with theta_n = 2 pi (w_n m_c + u_n) reduced in fp64 at the block centre m_c.
// work item = (b, blk, triple), triple = (r*T + t)*taps + tap as in `phases`; the thread walks the
// 16 tones itself (no shuffles; the 64 bytes of phases are four 128-bit loads).
template <int K>
__global__ void __launch_bounds__(256)
jakes_coef_kernel(const TdlParams C, const float* __restrict__ phases, float* __restrict__ coef, int R, int T,
                  int R2, long long total_items) {
    constexpr int NC = 2 * K + 1;
    const int nlt = R * T * C.num_taps;
    const long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (b*nbs + blk)*nlt + trip
    if (it >= total_items) return;
    const int trip = (int)(it % nlt);
    const long long q = it / nlt;                       // b*nbs + blk
    const int blk = (int)(q % C.nbs);
    const long long b = q / C.nbs;
    const double mc = (double)blk * C.pb + 0.5 * (C.pb - 1);
    const float4* up = (const float4*)(phases + ((size_t)b * nlt + trip) * LTE_JAKES_TONES);
    float u[LTE_JAKES_TONES];
#pragma unroll
    for (int i = 0; i < LTE_JAKES_TONES / 4; ++i) {
        const float4 v = __ldg(&up[i]);
        u[4 * i] = v.x; u[4 * i + 1] = v.y; u[4 * i + 2] = v.z; u[4 * i + 3] = v.w;
    }
    float2 a[K + 1];
#pragma unroll
    for (int k = 0; k <= K; ++k) a[k] = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int tone = 0; tone < LTE_JAKES_TONES; ++tone) {
        double turns = C.w_cyc[tone] * mc + (double)u[tone];
        turns -= floor(turns);
        float sn, cs;
        sincospif(2.0f * (float)turns, &sn, &cs);
        const float x = (float)(6.283185307179586 * C.w_cyc[tone]);   // rad / sample
        float2 term = make_float2(cs, sn);
        if constexpr (K == 1) {
            // economised (Chebyshev) linear fit of e^{j phi} over |phi| <= X = x pb/2:
            // cos phi ~ 1 - X^2/4 (max error X^2/4 instead of Taylor's X^2/2), sin phi ~ phi
            const float X = x * 0.5f * (float)C.pb;
            const float c0 = 1.0f - 0.25f * X * X;
            a[0].x += term.x * c0;
            a[0].y += term.y * c0;
        } else {
            a[0].x += term.x;
            a[0].y += term.y;
        }
#pragma unroll
        for (int k = 1; k <= K; ++k) {
            const float f = x / (float)k;
            term = make_float2(-term.y * f, term.x * f);
            a[k].x += term.x;
            a[k].y += term.y;
        }
    }
    const int tap = trip % C.num_taps, rt = trip / C.num_taps;
    const int t = rt % T, r = rt / T;
    const float g = C.gain[tap];
    float* c = coef + ((((size_t)q * T + t) * C.num_taps + tap) * NC) * 2 * R2 + r;
#pragma unroll
    for (int k = 0; k <= K; ++k) {
        c[(k * 2 + 0) * R2] = a[k].x * g;
        c[(k * 2 + 1) * R2] = a[k].y * g;
    }
#pragma unroll
    for (int k = 1; k <= K; ++k) {
        c[((K + k) * 2 + 0) * R2] = a[k].x * g * (float)k;
        c[((K + k) * 2 + 1) * R2] = a[k].y * g * (float)k;
    }
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async8_zfill(void* smem_dst, const void* gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;     // src-size 0 => zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__host__ __device__ constexpr int tdl_xs_stride(int halo, int tile) {
    return (halo + tile) / 8 + ((2 - ((halo + tile) / 8) % 16) + 16) % 16;   // == 2 (mod 16): conflict-free 64-bit rows
}
__host__ __device__ constexpr int tdl_os_stride(int tile) { return (tile / 8) | 1; }   // odd: conflict-free 128-bit rows

// antennas per thread (RG, a multiple of 2: each FFMA2 serves an antenna pair) and CTAs per SM
__host__ __device__ constexpr int tdl_rg(int R) { return R <= 2 ? 2 : 4; }
__host__ __device__ constexpr int tdl_min_blocks(int R) { return R <= 4 ? 4 : 2; }

// packed f32x2 helpers (the two lanes are two receive antennas)
struct pf2 { unsigned long long v; };
__device__ __forceinline__ pf2 ppk(float lo, float hi) { pf2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void pupk(pf2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ pf2 pfma(pf2 a, pf2 b, pf2 c) { pf2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
// acc += a * b in place: the read-write constraint keeps the accumulator in the same register pair
// across loop iterations (separate output registers cost one MOV pair per accumulator per tap)
__device__ __forceinline__ void pfma_acc(pf2& acc, pf2 a, pf2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc.v) : "l"(a.v), "l"(b.v)); }
__device__ __forceinline__ pf2 pmul(pf2 a, pf2 b) { pf2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ pf2 psub(pf2 a, pf2 b) { pf2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }

