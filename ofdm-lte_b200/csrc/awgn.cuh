// Lazy AWGN shared by the kernels that consume a noise-free grid (detect.cu, mimo.cu, sm.cu).
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------ lazy AWGN
// The sweep engine keeps the frequency-domain AWGN of lte_rx_fft(noise_domain = 1) out of HBM:
// the noise of grid element (row, symbol, bin) is a pure function of (seed, row id, symbol, bin), so
// the kernels that consume Y add it on the fly -- bit-identical to adding it in the RX epilogue
// (same Philox counter, same fmaf), and only for the elements that are actually read.
struct AwgnArgs {
    const double* power;        // [rows] sum |y|^2 of the stream (lte_channel_tdl / lte_channel_rx_fft)
    const float* snr_lin;       // [rows]
    uint32_t key;
    PhiloxKeys ks;              // round keys of `key` (constant-bank operands in the consumers' loops)
    unsigned long long row_id0;
    float n_stream;             // samples per stream, S * L
    int combine;                // MRC only: one draw per combiner output (see lte_awgn_desc)
};

__device__ __forceinline__ float2 awgn_at(const AwgnArgs& A, float sigma, long long row, int s, int N, int kb, float2 y) {
    const float2 w = lte_noise_sample(A.ks, (uint32_t)(A.row_id0 + (unsigned long long)row), (uint32_t)(s * N + kb));
    return make_float2(fmaf(sigma, w.x, y.x), fmaf(sigma, w.y, y.y));
}

static inline int make_awgn_args(AwgnArgs& A, const lte_plan* p, const lte_awgn_desc* d, int32_t S, int64_t rows) {
    if (!d->power || !d->snr_lin) return LTE_ERR_INVALID_ARG;
    if (rows > 0 && !lte_ids_fit(d->row_id0, (uint64_t)rows)) return LTE_ERR_UNSUPPORTED;
    A.power = d->power;
    A.snr_lin = d->snr_lin;
    A.key = lte_key(d->seed, LTE_DOMAIN_NOISE);
    A.ks = philox_key_schedule(A.key);
    A.row_id0 = d->row_id0;
    A.n_stream = (float)((size_t)S * p->dev.L);
    A.combine = d->combine;
    return LTE_OK;
}

