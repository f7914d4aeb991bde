// Packed-pair FFT core for sm_100a: every thread group carries TWO independent transforms whose
// samples travel as f32x2 pairs (transform A in the low lane, transform B in the high lane), so each
// butterfly add / mul / fma is one FADD2 / FMUL2 / FFMA2 instruction for both transforms.  Blackwell
// issues the packed fp32 ops at half rate, so the FMA pipe sees the same work, but the issue slots --
// which bound these kernels (ncu: 74-76 % issue-active at 41-46 % FMA-pipe) -- are halved, and the
// shared-memory exchanges move both transforms with one 128-bit access per element.
//
// Layout and schedule are those of fft.cuh: a pair of N-point transforms is carried by TPF = N/16
// threads; on entry thread j holds v[s] = x[j + s*TPF] (s = 0..15), on exit v[s] = X[j + s*TPF];
// Stockham passes of radix 16/8/4 with padded shared-memory exchanges in between.
#pragma once
#include "fft.cuh"

// per-pass twiddle tables behind the plan's exp(-2 pi i m / N) table (plan.cu): 0 = powers by repeated
// multiplication, 1 = second pass from the table, 2 = second and third pass.  Measured on the fused
// channel + FFT kernel (N = 2048): 1.402 / 1.388 / 1.415 ms -- the 2 KB second-pass table stays in L1,
// the 14 KB third-pass table does not pay for its 14 loads per thread.
#ifndef FFT2_TW_TABLE
#define FFT2_TW_TABLE 1
#endif

struct f2 { unsigned long long v; };
struct c2 { f2 re, im; };   // one complex sample of transform A (low) and of transform B (high)

__device__ __forceinline__ f2 pk(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(f2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }

__device__ __forceinline__ c2 cadd2(c2 a, c2 b) { return {add2(a.re, b.re), add2(a.im, b.im)}; }
__device__ __forceinline__ c2 csub2(c2 a, c2 b) { return {sub2(a.re, b.re), sub2(a.im, b.im)}; }
// a * w (both packed)
__device__ __forceinline__ c2 cmul2(c2 a, c2 w) {
    return {sub2(mul2(a.re, w.re), mul2(a.im, w.im)), fma2(a.re, w.im, mul2(a.im, w.re))};
}
// a * w with the same scalar twiddle for both transforms: the broadcast operands fold into the
// scalar form of FFMA2 / FMUL2, and the twiddle powers themselves stay on scalar (full-rate) ops
__device__ __forceinline__ c2 cmul2s(c2 a, float2 w) {
    const f2 wx = pk(w.x, w.x), wy = pk(w.y, w.y), nwy = pk(-w.y, -w.y);
    return {fma2(a.re, wx, mul2(a.im, nwy)), fma2(a.re, wy, mul2(a.im, wx))};
}
// a + (-j) b forward / a + (+j) b inverse, and the matching differences, without materialising j*b
template <bool INV> __device__ __forceinline__ c2 add_mj(c2 a, c2 b) {
    return INV ? c2{sub2(a.re, b.im), add2(a.im, b.re)} : c2{add2(a.re, b.im), sub2(a.im, b.re)};
}
template <bool INV> __device__ __forceinline__ c2 sub_mj(c2 a, c2 b) {
    return INV ? c2{add2(a.re, b.im), sub2(a.im, b.re)} : c2{sub2(a.re, b.im), add2(a.im, b.re)};
}
// a * (c - j s) forward, a * (c + j s) inverse; c, s compile-time scalars
template <bool INV> __device__ __forceinline__ c2 mul_w2(c2 a, float c, float s) {
    const f2 cc = pk(c, c), ss = pk(s, s), ns = pk(-s, -s);
    return INV ? c2{fma2(a.re, cc, mul2(a.im, ns)), fma2(a.im, cc, mul2(a.re, ss))}
               : c2{fma2(a.re, cc, mul2(a.im, ss)), fma2(a.im, cc, mul2(a.re, ns))};
}

// (e + w o, e - w o) for a compile-time twiddle w = c - j s (forward) / c + j s (inverse) in six FFMA2:
// w o = c [(o.re + t o.im) + j (o.im - t o.re)] with t = s / c (Linzer-Feig form), so the twiddle product
// and the butterfly's add / subtract share their multiplies -- two packed instructions fewer than
// mul_w2 followed by cadd2 / csub2.
template <bool INV> __device__ __forceinline__ void bfly_w2(c2 e, c2 o, float c, float s, c2& x, c2& y) {
    const float t = INV ? -s / c : s / c;
    const f2 tt = pk(t, t), nt = pk(-t, -t), cc = pk(c, c), nc = pk(-c, -c);
    const f2 ur = fma2(tt, o.im, o.re), ui = fma2(nt, o.re, o.im);
    x = {fma2(cc, ur, e.re), fma2(cc, ui, e.im)};
    y = {fma2(nc, ur, e.re), fma2(nc, ui, e.im)};
}

template <bool INV> __device__ __forceinline__ void bfly4p(c2& a0, c2& a1, c2& a2, c2& a3) {
    const c2 b0 = cadd2(a0, a2), b1 = csub2(a0, a2), b2 = cadd2(a1, a3), d = csub2(a1, a3);
    a0 = cadd2(b0, b2);
    a2 = csub2(b0, b2);
    a1 = add_mj<INV>(b1, d);
    a3 = sub_mj<INV>(b1, d);
}

template <bool INV> __device__ __forceinline__ void dft8p(c2 (&a)[8]) {
    const float h = 0.70710678118654752440f;
    bfly4p<INV>(a[0], a[2], a[4], a[6]);
    bfly4p<INV>(a[1], a[3], a[5], a[7]);
    const c2 e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6], o0 = a[1], o1 = a[3], o2 = a[5], o3 = a[7];
    a[0] = cadd2(e0, o0); a[4] = csub2(e0, o0);
    bfly_w2<INV>(e1, o1, h, h, a[1], a[5]);
    a[2] = add_mj<INV>(e2, o2); a[6] = sub_mj<INV>(e2, o2);
    bfly_w2<INV>(e3, o3, -h, h, a[3], a[7]);
}

template <bool INV> __device__ __forceinline__ void dft16p(c2 (&a)[16]) {
    c2 e[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { e[i] = a[2 * i]; o[i] = a[2 * i + 1]; }
    dft8p<INV>(e);
    dft8p<INV>(o);
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
    a[0] = cadd2(e[0], o[0]); a[8] = csub2(e[0], o[0]);
    bfly_w2<INV>(e[1], o[1], c1, s1, a[1], a[9]);
    bfly_w2<INV>(e[2], o[2], h, h, a[2], a[10]);
    bfly_w2<INV>(e[3], o[3], s1, c1, a[3], a[11]);
    a[4] = add_mj<INV>(e[4], o[4]); a[12] = sub_mj<INV>(e[4], o[4]);
    bfly_w2<INV>(e[5], o[5], -s1, c1, a[5], a[13]);
    bfly_w2<INV>(e[6], o[6], -h, h, a[6], a[14]);
    bfly_w2<INV>(e[7], o[7], -c1, s1, a[7], a[15]);
}

template <int R, bool INV> __device__ __forceinline__ void dftRp(c2 (&a)[R]) {
    if constexpr (R == 16) dft16p<INV>(a);
    else if constexpr (R == 8) dft8p<INV>(a);
    else { static_assert(R == 4, "radix"); bfly4p<INV>(a[0], a[1], a[2], a[3]); }
}

template <int N, int R, int NS, bool INV>
__device__ __forceinline__ void fft2_pass(c2 (&v)[FFT_ELEMS], const float2* __restrict__ tw, int j) {
    constexpr int TPF = N / FFT_ELEMS, Q = FFT_ELEMS / R;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        c2 a[R];
#pragma unroll
        for (int t = 0; t < R; ++t) a[t] = v[q + Q * t];
        if constexpr (NS > 1) {
            const int k = (j + q * TPF) & (NS - 1);
            float2 w[R];
            if constexpr (FFT2_TW_TABLE >= (NS <= 16 ? 1 : 2)) {
                // every twiddle power from the plan's per-pass table: consecutive threads read consecutive
                // entries (one L1 line per warp and power) instead of spending 4 (R - 2) scalar FMA-pipe ops
                const float2* tp = tw + N + (NS <= 16 ? 0 : FFT2_TW_PASS3) + k;
#pragma unroll
                for (int t = 1; t < R; ++t) {
                    w[t] = __ldg(tp + (t - 1) * NS);
                    if (INV) w[t].y = -w[t].y;
                }
            } else {
                w[1] = __ldg(&tw[k * (N / (NS * R))]);
                if (INV) w[1].y = -w[1].y;
#pragma unroll
                for (int t = 2; t < R; ++t) w[t] = cmul(w[t >> 1], w[(t + 1) >> 1]);
            }
#pragma unroll
            for (int t = 1; t < R; ++t) a[t] = cmul2s(a[t], w[t]);
        }
        dftRp<R, INV>(a);
#pragma unroll
        for (int u = 0; u < R; ++u) v[q + Q * u] = a[u];
    }
}

// Exchange through one padded buffer of float4 (A.re, B.re, A.im, B.im) per element.
template <int N, int R, int NS>
__device__ __forceinline__ void fft2_exchange(c2 (&v)[FFT_ELEMS], float4* s, int j) {
    constexpr int TPF = N / FFT_ELEMS, Q = FFT_ELEMS / R;
    if constexpr (NS > 1) __syncthreads();   // readers of the previous exchange are done with `s`
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int jb = j + q * TPF;
        const int k = jb & (NS - 1);
        const int base = (jb - k) * R + k;
#pragma unroll
        for (int u = 0; u < R; ++u) {
            const c2 x = v[q + Q * u];
            float a, b, c, d;
            upk(x.re, a, b);
            upk(x.im, c, d);
            s[fft_pad(base + u * NS)] = make_float4(a, b, c, d);
        }
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < FFT_ELEMS; ++t) {
        const float4 x = s[fft_pad(j + t * TPF)];
        v[t] = {pk(x.x, x.y), pk(x.z, x.w)};
    }
}

// s: one padded buffer of fft_smem_elems(N) float4 private to this transform pair.
template <int N, bool INV>
__device__ __forceinline__ void fft2_run(c2 (&v)[FFT_ELEMS], float4* s, const float2* __restrict__ tw, int j) {
    if constexpr (N == 2048) {
        fft2_pass<N, 16, 1, INV>(v, tw, j);   fft2_exchange<N, 16, 1>(v, s, j);
        fft2_pass<N, 16, 16, INV>(v, tw, j);  fft2_exchange<N, 16, 16>(v, s, j);
        fft2_pass<N, 8, 256, INV>(v, tw, j);
    } else if constexpr (N == 1024) {
        fft2_pass<N, 16, 1, INV>(v, tw, j);   fft2_exchange<N, 16, 1>(v, s, j);
        fft2_pass<N, 8, 16, INV>(v, tw, j);   fft2_exchange<N, 8, 16>(v, s, j);
        fft2_pass<N, 8, 128, INV>(v, tw, j);
    } else if constexpr (N == 512) {
        fft2_pass<N, 8, 1, INV>(v, tw, j);    fft2_exchange<N, 8, 1>(v, s, j);
        fft2_pass<N, 8, 8, INV>(v, tw, j);    fft2_exchange<N, 8, 8>(v, s, j);
        fft2_pass<N, 8, 64, INV>(v, tw, j);
    } else if constexpr (N == 256) {
        fft2_pass<N, 16, 1, INV>(v, tw, j);   fft2_exchange<N, 16, 1>(v, s, j);
        fft2_pass<N, 16, 16, INV>(v, tw, j);
    } else if constexpr (N == 128) {
        fft2_pass<N, 16, 1, INV>(v, tw, j);   fft2_exchange<N, 16, 1>(v, s, j);
        fft2_pass<N, 8, 16, INV>(v, tw, j);
    } else {
        static_assert(N == 64, "unsupported FFT size");
        fft2_pass<N, 16, 1, INV>(v, tw, j);   fft2_exchange<N, 16, 1>(v, s, j);
        fft2_pass<N, 4, 16, INV>(v, tw, j);
    }
}

// transform PAIRS per CTA and shared-memory bytes for a CTA of FFT_CTA_THREADS threads
__host__ __device__ constexpr int fft2_pairs_per_cta(int n) { return FFT_CTA_THREADS / (n / FFT_ELEMS); }
__host__ __device__ constexpr int fft2_cta_smem_bytes(int n) {
    return fft2_pairs_per_cta(n) * fft_smem_elems(n) * (int)sizeof(float4);
}
