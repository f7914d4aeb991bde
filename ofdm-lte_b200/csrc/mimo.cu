// Transmit-diversity (SFBC Alamouti) stage kernels.
#include "slicer.cuh"
#include "awgn.cuh"

// ------------------------------------------------------------------------------ SFBC encode
// core/sfbc_alamouti.py:45-78: pairs (k, k+1) of data symbols -> TX0 [s0, -conj(s1)],
// TX1 [s1, conj(s0)].  Input: symbol indices (QAM map fused, core/modulator.py:80-86) or complex
// symbols, npair pairs per OFDM symbol; output rows of Nd symbols per TX antenna (the bins past
// 2*npair stay zero: with an odd number of data bins the last one is nulled, :196-200).
__global__ void __launch_bounds__(256)
sfbc_encode_kernel(const DevPlan P, const uint8_t* __restrict__ idx, const float2* __restrict__ syms,
                   float2* __restrict__ out, float2* __restrict__ qam_out, int npair, int S, long long total) {
    const int h = P.bps >> 1, mask = (1 << h) - 1;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const int pr = (int)(g % npair);
        const long long bs = g / npair;                 // b*S + s
        const long long b = bs / S;
        const int s = (int)(bs % S);
        const size_t in0 = (size_t)bs * 2 * npair + 2 * pr;
        float2 s0, s1;
        if (syms) {
            s0 = syms[in0];
            s1 = syms[in0 + 1];
        } else {
            const int i0 = idx[in0], i1 = idx[in0 + 1];
            s0 = make_float2(P.lev[(i0 >> h) & mask], P.lev[i0 & mask]);
            s1 = make_float2(P.lev[(i1 >> h) & mask], P.lev[i1 & mask]);
            if (qam_out) { qam_out[in0] = s0; qam_out[in0 + 1] = s1; }
        }
        float2* o0 = out + (((size_t)b * 2 + 0) * S + s) * P.Nd + 2 * pr;
        float2* o1 = out + (((size_t)b * 2 + 1) * S + s) * P.Nd + 2 * pr;
        o0[0] = s0;
        o0[1] = make_float2(-s1.x, s1.y);
        o1[0] = s1;
        o1[1] = make_float2(s0.x, -s0.y);
        if (pr == 0 && 2 * npair < P.Nd) {             // null the unused tail bins
            for (int k = 2 * npair; k < P.Nd; ++k) {
                out[(((size_t)b * 2 + 0) * S + s) * P.Nd + k] = make_float2(0.f, 0.f);
                out[(((size_t)b * 2 + 1) * S + s) * P.Nd + k] = make_float2(0.f, 0.f);
            }
        }
    }
}

extern "C" int lte_sfbc_encode(const lte_plan* p, const uint8_t* idx, const lte_c32* symbols, lte_c32* out,
                               lte_c32* qam_out, int64_t B, int32_t S, void* stream) {
    if (!p || (!idx && !symbols) || !out || B < 0 || S < 1) return LTE_ERR_INVALID_ARG;
    if (B == 0) return LTE_OK;
    const int npair = p->dev.Nd / 2;
    if (npair < 1) return LTE_ERR_INVALID_ARG;
    const long long total = (long long)B * S * npair;
    long long grid = (total + 255) / 256;
    if (grid > 148 * 32) grid = 148 * 32;
    sfbc_encode_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(p->dev, idx, (const float2*)symbols,
                                                                        (float2*)out, (float2*)qam_out, npair, S,
                                                                        total);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

// ------------------------------------------------------------------------------ SFBC decode
// core/sfbc_alamouti.py:132-161 per receive antenna, then the plain average over antennas of
// core/ofdm_core.py:2204.  thread = (stream, pair); channel estimates of the slot stay in registers.
// COUNT: the decoded pair goes straight through the slicer and the bit-error count against idx_tx
// (core/modulator.py:90-112, core/ofdm_core.py:245-268) instead of to memory -- the sweep's form.
// NOISY: Y is noise free (lte_channel_rx_fft_mimo) and the AWGN of lte_rx_fft(noise_domain = 1) is added to the
// two bins while they are read (awgn.cuh: same draws, so the noisy grid never exists).
template <int R, bool COUNT, bool NOISY>
__global__ void __launch_bounds__(128)
sfbc_decode_kernel(const DevPlan P, const float2* __restrict__ Y, const float2* __restrict__ H0,
                   const float2* __restrict__ H1, float2* __restrict__ out, const uint8_t* __restrict__ idx_tx,
                   unsigned long long* __restrict__ errors, long long nbits, int k0, int nk, int S, int nslot,
                   int npair, int gx, const AwgnArgs A) {
    const long long b = blockIdx.x / gx;
    const int pr0 = (blockIdx.x % gx) * blockDim.x + threadIdx.x;
    const bool act = pr0 < npair;
    if (!COUNT && !act) return;
    const int pr = act ? pr0 : npair - 1;            // counting CTAs stay whole for the block reduction
    unsigned int e = 0;
    const int ka = P.data_idx[2 * pr] - k0, kb = P.data_idx[2 * pr + 1] - k0;
    const float invR = 1.0f / (float)R;
    float sigma[R];
    if (NOISY) {
#pragma unroll
        for (int r = 0; r < R; ++r) sigma[r] = lte_sigma(A.power[b * R + r], A.n_stream, A.snr_lin[b * R + r]);
    }
    for (int slot = 0; slot < nslot; ++slot) {
        float2 h0a[R], h0b[R], h1a[R], h1b[R];
        float nrm[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const size_t o = (((size_t)b * R + r) * nslot + slot) * nk;
            h0a[r] = H0[o + ka]; h0b[r] = H0[o + kb];
            h1a[r] = H1[o + ka]; h1b[r] = H1[o + kb];
            const float2 m0 = make_float2(0.5f * (h0a[r].x + h0b[r].x), 0.5f * (h0a[r].y + h0b[r].y));
            const float2 m1 = make_float2(0.5f * (h1a[r].x + h1b[r].x), 0.5f * (h1a[r].y + h1b[r].y));
            nrm[r] = cabs2(m0) + cabs2(m1) + 1e-10f;
        }
        const int s_end = min(S, (slot + 1) * LTE_SLOT_SYMBOLS);
        for (int s = slot * LTE_SLOT_SYMBOLS; s < s_end; ++s) {
            float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const size_t o = (((size_t)b * R + r) * S + s) * nk;
                float2 ra = Y[o + ka], rb = Y[o + kb];
                if (NOISY) {
                    ra = awgn_at(A, sigma[r], b * R + r, s, P.N, ka + k0, ra);
                    rb = awgn_at(A, sigma[r], b * R + r, s, P.N, kb + k0, rb);
                }
                const float2 rbc = make_float2(rb.x, -rb.y);
                // s0 = conj(h0a) ra + h1b conj(rb);  s1 = conj(h1a) ra - h0b conj(rb)
                float2 t0 = cmulc(h0a[r], ra), u0 = cmul(h1b[r], rbc);
                float2 t1 = cmulc(h1a[r], ra), u1 = cmul(h0b[r], rbc);
                a0.x += __fdiv_rn(t0.x + u0.x, nrm[r]); a0.y += __fdiv_rn(t0.y + u0.y, nrm[r]);
                a1.x += __fdiv_rn(t1.x - u1.x, nrm[r]); a1.y += __fdiv_rn(t1.y - u1.y, nrm[r]);
            }
            const size_t q = ((size_t)b * S + s) * 2 * npair + 2 * pr;
            const float2 d0 = make_float2(a0.x * invR, a0.y * invR), d1 = make_float2(a1.x * invR, a1.y * invR);
            if (COUNT) {
                const long long left = nbits - ((long long)s * 2 * npair + 2 * pr) * P.bps;    // bits of the stream from here on
                e += bit_errors(slice_symbol(P, d0), idx_tx[q], P.bps, left);
                e += bit_errors(slice_symbol(P, d1), idx_tx[q + 1], P.bps, left - P.bps);
            } else {
                out[q] = d0;
                out[q + 1] = d1;
            }
        }
    }
    if (COUNT) block_add_errors(act ? e : 0u, &errors[b]);
}

static int launch_sfbc_decode(const lte_plan* p, const lte_c32* Y, const lte_c32* H0, const lte_c32* H1, lte_c32* out,
                              const uint8_t* idx_tx, unsigned long long* errors, int64_t nbits, int window, int64_t B,
                              int32_t R, int32_t S, const lte_awgn_desc* awgn, void* stream) {
    const bool count = idx_tx != nullptr;
    if (!p || !Y || !H0 || !H1 || (!count && !out) || (count && !errors) || B < 0 || S < 1) return LTE_ERR_INVALID_ARG;
    int32_t k0, nk;
    int rc = lte_plan_window(p, window, &k0, &nk);
    if (rc) return rc;
    AwgnArgs A = {};
    if (awgn && (rc = make_awgn_args(A, p, awgn, S, B * R))) return rc;
    if (B == 0) return LTE_OK;
    const int npair = p->dev.Nd / 2;
    if (npair < 1) return LTE_ERR_INVALID_ARG;
    const int nslot = (S + LTE_SLOT_SYMBOLS - 1) / LTE_SLOT_SYMBOLS;
    const int gx = (npair + 127) / 128;
    const unsigned grid = (unsigned)((long long)gx * B);
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_SFBC2(RR, CC, NN)                                                                              \
    sfbc_decode_kernel<RR, CC, NN><<<grid, 128, 0, st>>>(p->dev, (const float2*)Y, (const float2*)H0, (const float2*)H1, \
                                                         (float2*)out, idx_tx, errors, nbits, k0, nk, S, nslot, npair, gx, A)
#define LAUNCH_SFBC(RR)                                                                                       \
    case RR:                                                                                                  \
        if (count) { if (awgn) LAUNCH_SFBC2(RR, true, true); else LAUNCH_SFBC2(RR, true, false); }            \
        else { if (awgn) LAUNCH_SFBC2(RR, false, true); else LAUNCH_SFBC2(RR, false, false); }                \
        break;
    switch (R) {
        LAUNCH_SFBC(1) LAUNCH_SFBC(2) LAUNCH_SFBC(3) LAUNCH_SFBC(4) LAUNCH_SFBC(5) LAUNCH_SFBC(6) LAUNCH_SFBC(7)
        LAUNCH_SFBC(8)
        default: return LTE_ERR_INVALID_ARG;
    }
#undef LAUNCH_SFBC
#undef LAUNCH_SFBC2
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

extern "C" int lte_sfbc_decode(const lte_plan* p, const lte_c32* Y, const lte_c32* H0, const lte_c32* H1,
                               lte_c32* out, int window, int64_t B, int32_t R, int32_t S, void* stream) {
    if (!out) return LTE_ERR_INVALID_ARG;
    return launch_sfbc_decode(p, Y, H0, H1, out, nullptr, nullptr, 0, window, B, R, S, nullptr, stream);
}

extern "C" int lte_sfbc_decode_count(const lte_plan* p, const lte_c32* Y, const lte_c32* H0, const lte_c32* H1,
                                     const uint8_t* idx_tx, unsigned long long* errors, int64_t nbits, int window,
                                     int64_t B, int32_t R, int32_t S, const lte_awgn_desc* awgn, void* stream) {
    if (!idx_tx || nbits < 0) return LTE_ERR_INVALID_ARG;
    return launch_sfbc_decode(p, Y, H0, H1, nullptr, idx_tx, errors, nbits, window, B, R, S, awgn, stream);
}
