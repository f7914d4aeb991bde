// Hard-decision slicer, bit-error counting and the per-CTA error reduction shared by the
// detector kernels (detect.cu) and the beamforming link kernel (beamform.cu).
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------ slicer
// argmin_i |c_i - y| with first-minimum ties (core/modulator.py:103-106) == per-axis nearest
// level; thresholds are rounded towards -inf so `y > thr` equals `y > exact midpoint`, and a
// tie at y == 0 picks the lower level (16/64-QAM) or the positive one (QPSK).
__device__ __forceinline__ int slice_axis(const DevPlan& P, float y) {
    if (P.nlev == 2) return y < 0.f ? 1 : 0;
    // number of thresholds below y by bisection; unused entries of thr[] are +inf (plan.cu).  The
    // threshold of each step is picked with selects between constant-bank operands: thr[l + 1] and
    // thr[l] with a run-time l would be indexed constant loads on the dependent chain
    const bool b4 = y > P.thr[3];
    const bool b2 = y > (b4 ? P.thr[5] : P.thr[1]);
    const float t1 = b4 ? (b2 ? P.thr[6] : P.thr[4]) : (b2 ? P.thr[2] : P.thr[0]);
    return (b4 ? 4 : 0) + (b2 ? 2 : 0) + (y > t1 ? 1 : 0);
}
__device__ __forceinline__ int slice_symbol(const DevPlan& P, float2 y) {
    return (slice_axis(P, y.x) << (P.bps >> 1)) | slice_axis(P, y.y);
}
// number of differing bits among the first `valid` (MSB-first) bits of two b-bit indices
__device__ __forceinline__ int bit_errors(int a, int b, int bps, long long valid) {
    if (valid <= 0) return 0;
    int x = a ^ b;
    if (valid < bps) x &= ~((1 << (bps - (int)valid)) - 1);
    return __popc(x);
}

__device__ __forceinline__ void block_add_errors(unsigned int e, unsigned long long* dst) {
    e = (unsigned int)__reduce_add_sync(0xffffffffu, e);
    __shared__ unsigned int red[32];
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) red[w] = e;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned int t = threadIdx.x < nw ? red[threadIdx.x] : 0u;
        t = (unsigned int)__reduce_add_sync(0xffffffffu, t);
        if (threadIdx.x == 0 && t) atomicAdd(dst, (unsigned long long)t);
    }
}

