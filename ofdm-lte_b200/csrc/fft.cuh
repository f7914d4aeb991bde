// Batched power-of-two FFT core: 16 complex values per thread in registers,
// Stockham autosort passes of radix 16/8 with shared-memory exchanges in between.
//
// A transform of size N is carried by TPF = N/16 threads.  On entry thread j holds
//   v[s] = x[j + s*TPF], s = 0..15;   on exit   v[s] = X[j + s*TPF].
// Each pass multiplies by the inter-pass twiddles W^(k t), does 16/R radix-R
// butterflies per thread and (except after the last pass) redistributes through
// shared memory.  Padding one float2 every 16 keeps the transposing stores and
// the unit-stride loads free of bank conflicts for the first two passes.
#pragma once
#include "common.cuh"

#define FFT_ELEMS 16
__host__ __device__ constexpr int fft_pad(int i) { return i + (i >> 4); }
__host__ __device__ constexpr int fft_smem_elems(int n) { return n + (n >> 4); }

template <bool INV> __device__ __forceinline__ float2 mul_mj(float2 a) {   // * (-j) forward, * (+j) inverse
    return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
template <bool INV> __device__ __forceinline__ float2 mul_w(float2 a, float c, float s) {
    // a * (c - j s) forward, a * (c + j s) inverse
    return INV ? make_float2(fmaf(a.x, c, -a.y * s), fmaf(a.y, c, a.x * s))
               : make_float2(fmaf(a.x, c, a.y * s), fmaf(a.y, c, -a.x * s));
}

template <bool INV> __device__ __forceinline__ void bfly4(float2& a0, float2& a1, float2& a2, float2& a3) {
    float2 b0 = cadd(a0, a2), b1 = csub(a0, a2), b2 = cadd(a1, a3), b3 = mul_mj<INV>(csub(a1, a3));
    a0 = cadd(b0, b2); a1 = cadd(b1, b3); a2 = csub(b0, b2); a3 = csub(b1, b3);
}

// natural order in, natural order out
template <bool INV> __device__ __forceinline__ void dft8(float2 (&a)[8]) {
    const float h = 0.70710678118654752440f;
    bfly4<INV>(a[0], a[2], a[4], a[6]);          // even -> E0..E3 in a0,a2,a4,a6
    bfly4<INV>(a[1], a[3], a[5], a[7]);          // odd  -> O0..O3 in a1,a3,a5,a7
    float2 o1 = mul_w<INV>(a[3], h, h);
    float2 o2 = mul_mj<INV>(a[5]);
    float2 o3 = mul_w<INV>(a[7], -h, h);
    float2 e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6], o0 = a[1];
    a[0] = cadd(e0, o0); a[4] = csub(e0, o0);
    a[1] = cadd(e1, o1); a[5] = csub(e1, o1);
    a[2] = cadd(e2, o2); a[6] = csub(e2, o2);
    a[3] = cadd(e3, o3); a[7] = csub(e3, o3);
}

template <bool INV> __device__ __forceinline__ void dft16(float2 (&a)[16]) {
    float2 e[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { e[i] = a[2 * i]; o[i] = a[2 * i + 1]; }
    dft8<INV>(e);
    dft8<INV>(o);
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
    o[1] = mul_w<INV>(o[1], c1, s1);
    o[2] = mul_w<INV>(o[2], h, h);
    o[3] = mul_w<INV>(o[3], s1, c1);
    o[4] = mul_mj<INV>(o[4]);
    o[5] = mul_w<INV>(o[5], -s1, c1);
    o[6] = mul_w<INV>(o[6], -h, h);
    o[7] = mul_w<INV>(o[7], -c1, s1);
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = cadd(e[i], o[i]); a[i + 8] = csub(e[i], o[i]); }
}

template <int R, bool INV> __device__ __forceinline__ void dftR(float2 (&a)[R]) {
    if constexpr (R == 16) dft16<INV>(a);
    else if constexpr (R == 8) dft8<INV>(a);
    else if constexpr (R == 4) bfly4<INV>(a[0], a[1], a[2], a[3]);
    else { float2 t = a[0]; a[0] = cadd(t, a[1]); a[1] = csub(t, a[1]); }
}

// twiddle + butterflies of one pass (radix R, NS = product of the earlier radices)
template <int N, int R, int NS, bool INV>
__device__ __forceinline__ void fft_pass(float2 (&v)[FFT_ELEMS], const float2* __restrict__ tw, int j) {
    constexpr int TPF = N / FFT_ELEMS, Q = FFT_ELEMS / R;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        float2 a[R];
#pragma unroll
        for (int t = 0; t < R; ++t) a[t] = v[q + Q * t];
        if constexpr (NS > 1) {
            const int k = (j + q * TPF) & (NS - 1);
            float2 w[R];
            w[1] = __ldg(&tw[k * (N / (NS * R))]);
            if (INV) w[1].y = -w[1].y;
#pragma unroll
            for (int t = 2; t < R; ++t) w[t] = cmul(w[t >> 1], w[(t + 1) >> 1]);
#pragma unroll
            for (int t = 1; t < R; ++t) a[t] = cmul(a[t], w[t]);
        }
        dftR<R, INV>(a);
#pragma unroll
        for (int u = 0; u < R; ++u) v[q + Q * u] = a[u];
    }
}

// Stockham scatter of a pass's outputs, barrier, unit-stride gather for the next pass.
template <int N, int R, int NS>
__device__ __forceinline__ void fft_exchange(float2 (&v)[FFT_ELEMS], float2* s, int j) {
    constexpr int TPF = N / FFT_ELEMS, Q = FFT_ELEMS / R;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int jb = j + q * TPF;
        const int k = jb & (NS - 1);
        const int base = (jb - k) * R + k;
#pragma unroll
        for (int u = 0; u < R; ++u) s[fft_pad(base + u * NS)] = v[q + Q * u];
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < FFT_ELEMS; ++t) v[t] = s[fft_pad(j + t * TPF)];
}

// Radix schedule per size; every CTA thread must call this (it contains barriers).
// sA / sB: two padded buffers of fft_smem_elems(N) float2 private to this transform.
template <int N, bool INV>
__device__ __forceinline__ void fft_run(float2 (&v)[FFT_ELEMS], float2* sA, float2* sB,
                                        const float2* __restrict__ tw, int j) {
    if constexpr (N == 2048) {
        fft_pass<N, 16, 1, INV>(v, tw, j);   fft_exchange<N, 16, 1>(v, sA, j);
        fft_pass<N, 16, 16, INV>(v, tw, j);  fft_exchange<N, 16, 16>(v, sB, j);
        fft_pass<N, 8, 256, INV>(v, tw, j);
    } else if constexpr (N == 1024) {
        fft_pass<N, 16, 1, INV>(v, tw, j);   fft_exchange<N, 16, 1>(v, sA, j);
        fft_pass<N, 8, 16, INV>(v, tw, j);   fft_exchange<N, 8, 16>(v, sB, j);
        fft_pass<N, 8, 128, INV>(v, tw, j);
    } else if constexpr (N == 512) {
        fft_pass<N, 8, 1, INV>(v, tw, j);    fft_exchange<N, 8, 1>(v, sA, j);
        fft_pass<N, 8, 8, INV>(v, tw, j);    fft_exchange<N, 8, 8>(v, sB, j);
        fft_pass<N, 8, 64, INV>(v, tw, j);
    } else if constexpr (N == 256) {
        fft_pass<N, 16, 1, INV>(v, tw, j);   fft_exchange<N, 16, 1>(v, sA, j);
        fft_pass<N, 16, 16, INV>(v, tw, j);
    } else if constexpr (N == 128) {
        fft_pass<N, 16, 1, INV>(v, tw, j);   fft_exchange<N, 16, 1>(v, sA, j);
        fft_pass<N, 8, 16, INV>(v, tw, j);
    } else {
        static_assert(N == 64, "unsupported FFT size");
        fft_pass<N, 16, 1, INV>(v, tw, j);   fft_exchange<N, 16, 1>(v, sA, j);
        fft_pass<N, 4, 16, INV>(v, tw, j);
    }
}

#define FFT_CTA_THREADS 128
// transforms per CTA and shared-memory bytes for a CTA of FFT_CTA_THREADS threads
__host__ __device__ constexpr int fft_per_cta(int n) { return FFT_CTA_THREADS / (n / FFT_ELEMS); }
__host__ __device__ constexpr int fft_cta_smem_bytes(int n) {
    return 2 * fft_per_cta(n) * fft_smem_elems(n) * (int)sizeof(float2);
}
