// Fused stage 3 + stage "a10": tapped-delay-line Rayleigh channel, stream power, CP strip and
// FFT/sqrt(N) in one kernel (core/rayleighchannel.py:20-58, core/ofdm_core.py:361-412,
// core/lte_receiver.py:444-491).  The faded time-domain streams -- R x S x L complex samples per
// subframe, 8 of the 15 GB the staged pipeline moves per 4096 subframes -- never leave the SM:
// a CTA stages the transmitted samples of its OFDM symbol (plus the delay-spread halo) in shared
// memory, forms the faded samples of every receive antenna directly in the FFT's register layout
// (antenna pairs ride in the two lanes of the packed f32x2 transforms of fft2.cuh), accumulates the
// stream power including the CP samples, transforms, and writes only the kept bins of Y.
//
// The AWGN of the link cannot be added here, because sigma depends on the power of the whole
// stream; the sweep engine adds it lazily in the consumers instead (lte_crs_ls_interp_awgn,
// lte_mrc_demap_count_awgn -- same draws as lte_rx_fft(noise_domain = 1)).
//
// Jakes fading is one Taylor polynomial per (antenna, tap, OFDM symbol) around the symbol centre
// (jakes_coef_kernel with block length L); the host checks the truncation bound and reports
// LTE_ERR_UNSUPPORTED when the Doppler spread is too large for one block per symbol, in which
// case the caller uses the staged lte_channel_tdl + lte_rx_fft pair.
#include <type_traits>

#include "fft2.cuh"
#include "tdl.cuh"

template <int N, int K>
__global__ void __launch_bounds__(FFT_CTA_THREADS, (N == 2048 && K >= 4) ? 3 : 4)
channel_rx_fft_kernel(const DevPlan P, const TdlParams C, const float2* __restrict__ tx,
                      const float* __restrict__ coef_g, float2* __restrict__ Y, double* __restrict__ power, int k0,
                      int nk, int S, int R, int R2, int halo, unsigned total, bool wide) {
    constexpr int TPF = N / FFT_ELEMS, PPC = fft2_pairs_per_cta(N);     // PPC symbols per CTA
    constexpr int NC = 2 * K + 1;
    extern __shared__ float4 smem4[];
    const int p_local = threadIdx.x / TPF, j = threadIdx.x % TPF;
    const int L = P.L, cp = P.cp;
    const int win = halo + L;                                   // staged samples per symbol
    const int ncoef = C.num_taps * NC * 2 * R2;                  // floats per (stream, symbol)
    // shared memory: sbuf [PPC][fft_smem_elems] float4 | xw [PPC][win (even)] float2 | coef [PPC][ncoef] float
    float4* sbuf = smem4 + (size_t)p_local * fft_smem_elems(N);
    float2* xw_all = (float2*)(smem4 + (size_t)PPC * fft_smem_elems(N));
    const int win2 = (win + 1) & ~1;
    float2* xw = xw_all + (size_t)p_local * win2;
    float* sc_all = (float*)(xw_all + (size_t)PPC * win2);
    const float* sc = sc_all + (size_t)p_local * ncoef;

    const unsigned f = blockIdx.x * PPC + p_local;               // = b * S + s
    const bool valid = f < total;
    const unsigned b = valid ? f / (unsigned)S : 0u;
    const unsigned s = valid ? f - b * (unsigned)S : 0u;

    // ---- stage the symbol's transmitted samples [s L - halo, (s+1) L) and its coefficients -------
    if (valid) {
        const long long m_lo = (long long)s * L - halo;
        const float2* src = tx + (size_t)b * S * L + m_lo;      // may point before the stream for s = 0
        if (wide) {
            // halo and L are even and tx is 16-byte aligned: sample pairs move as 16-byte copies
            for (int i = 2 * j; i < win; i += 2 * TPF) {
                const bool ok = m_lo + i >= 0;
                cp_async16_zfill(&xw[i], ok ? &src[i] : tx, ok);
            }
        } else {
            for (int i = j; i < win; i += TPF) {
                const bool ok = m_lo + i >= 0;
                cp_async8_zfill(&xw[i], ok ? &src[i] : tx, ok);
            }
        }
        const float* cg = coef_g + (size_t)f * ncoef;
        float* scw = sc_all + (size_t)p_local * ncoef;
        for (int i = j; i < ncoef; i += TPF) cp_async4(&scw[i], &cg[i]);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    const float tau0 = (float)(cp + j) - 0.5f * (float)(L - 1);  // polynomial argument of useful sample j
    const int RP = R2 / 2;
    for (int p = 0; p < RP; ++p) {
        c2 v[FFT_ELEMS];
        float pwa = 0.f, pwb = 0.f;
        if (valid) {
            // FIRST = the first tap: the accumulators start from its product (no zero fill + fma on zero)
            auto tap_pass = [&](int tap, auto first) {
                constexpr bool FIRST = decltype(first)::value;
                const float* ct = sc + (size_t)tap * NC * 2 * R2 + 2 * p;
                f2 cre[K + 1], cim[K + 1];
#pragma unroll
                for (int k = 0; k <= K; ++k) {
                    const float2 a = *(const float2*)(ct + (k * 2 + 0) * R2), c = *(const float2*)(ct + (k * 2 + 1) * R2);
                    cre[k] = pk(a.x, a.y);
                    cim[k] = pk(c.x, c.y);
                }
                const float2* xt = xw + halo + cp + j - C.delay[tap];
#pragma unroll
                for (int e = 0; e < FFT_ELEMS; ++e) {
                    const float tf = tau0 + (float)(e * TPF);
                    const f2 tau = pk(tf, tf);
                    f2 hre = cre[K], him = cim[K];
#pragma unroll
                    for (int k = K - 1; k >= 0; --k) { hre = fma2(hre, tau, cre[k]); him = fma2(him, tau, cim[k]); }
                    const float2 x = xt[e * TPF];
                    const f2 xre = pk(x.x, x.x), xim = pk(x.y, x.y), nxim = pk(-x.y, -x.y);
                    if (FIRST) {
                        v[e].re = fma2(him, nxim, mul2(hre, xre));
                        v[e].im = fma2(him, xre, mul2(hre, xim));
                    } else {
                        v[e].re = fma2(hre, xre, v[e].re);
                        v[e].re = fma2(him, nxim, v[e].re);
                        v[e].im = fma2(hre, xim, v[e].im);
                        v[e].im = fma2(him, xre, v[e].im);
                    }
                }
            };
            tap_pass(0, std::true_type());
            for (int tap = 1; tap < C.num_taps; ++tap) tap_pass(tap, std::false_type());
            // power of this thread's 16 useful samples ...
            f2 acc = pk(0.f, 0.f);
#pragma unroll
            for (int e = 0; e < FFT_ELEMS; ++e) { acc = fma2(v[e].re, v[e].re, acc); acc = fma2(v[e].im, v[e].im, acc); }
            // ... and of its cyclic-prefix samples, which only enter the stream power
            // (core/channel.py:216-218 measures the whole faded stream)
            for (int i = j; i < cp; i += TPF) {
                const float tf = (float)i - 0.5f * (float)(L - 1);
                const f2 tau = pk(tf, tf);
                f2 yre = pk(0.f, 0.f), yim = pk(0.f, 0.f);
                for (int tap = 0; tap < C.num_taps; ++tap) {
                    const float* ct = sc + (size_t)tap * NC * 2 * R2 + 2 * p;
                    float2 a = *(const float2*)(ct + (K * 2 + 0) * R2), c = *(const float2*)(ct + (K * 2 + 1) * R2);
                    f2 hre = pk(a.x, a.y), him = pk(c.x, c.y);
#pragma unroll
                    for (int k = K - 1; k >= 0; --k) {
                        a = *(const float2*)(ct + (k * 2 + 0) * R2);
                        c = *(const float2*)(ct + (k * 2 + 1) * R2);
                        hre = fma2(hre, tau, pk(a.x, a.y));
                        him = fma2(him, tau, pk(c.x, c.y));
                    }
                    const float2 x = xw[halo + i - C.delay[tap]];
                    const f2 xre = pk(x.x, x.x), xim = pk(x.y, x.y), nxim = pk(-x.y, -x.y);
                    yre = fma2(hre, xre, yre); yre = fma2(him, nxim, yre);
                    yim = fma2(hre, xim, yim); yim = fma2(him, xre, yim);
                }
                acc = fma2(yre, yre, acc);
                acc = fma2(yim, yim, acc);
            }
            upk(acc, pwa, pwb);
        }
        else {
#pragma unroll
            for (int e = 0; e < FFT_ELEMS; ++e) v[e] = {pk(0.f, 0.f), pk(0.f, 0.f)};
        }
        // ---- stream power: one atomic per warp (or per symbol when a warp carries several) ---------
        // the tap gains carry the FFT's 1/sqrt(N) (host side), so the power is scaled back by N here
        bool leader;
        if constexpr (TPF >= 32) {
            pwa = warp_sum(pwa);
            pwb = warp_sum(pwb);
            leader = (threadIdx.x & 31) == 0;
        } else {
#pragma unroll
            for (int ofs = TPF / 2; ofs > 0; ofs >>= 1) {
                pwa += __shfl_xor_sync(0xffffffffu, pwa, ofs);
                pwb += __shfl_xor_sync(0xffffffffu, pwb, ofs);
            }
            leader = j == 0;
        }
        const int r0 = 2 * p;
        const bool two = r0 + 1 < R;
        if (valid && leader) {
            atomicAdd(&power[(size_t)b * R + r0], (double)pwa * (double)N);
            if (two) atomicAdd(&power[(size_t)b * R + r0 + 1], (double)pwb * (double)N);
        }

        // ---- FFT of the useful part of both antennas, kept bins to Y --------------------------------
        fft2_run<N, false>(v, sbuf, P.twiddle, j);
        if (valid) {
            // base pointers are formed once; every store is base + compile-time offset under one
            // unsigned range check (the pointer may lie before the row, it is only dereferenced in range)
            float2* o0 = Y + (((size_t)b * R + r0) * S + s) * nk + (j - k0);
            float2* o1 = o0 + (size_t)S * nk;
            const unsigned kb = (unsigned)(j - k0);
            if (two) {
#pragma unroll
                for (int e = 0; e < FFT_ELEMS; ++e) {
                    if (kb + (unsigned)(e * TPF) < (unsigned)nk) {
                        float a, c, d, g;
                        upk(v[e].re, a, c);
                        upk(v[e].im, d, g);
                        o0[e * TPF] = make_float2(a, d);
                        o1[e * TPF] = make_float2(c, g);
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < FFT_ELEMS; ++e) {
                    if (kb + (unsigned)(e * TPF) < (unsigned)nk) {
                        float a, c, d, g;
                        upk(v[e].re, a, c);
                        upk(v[e].im, d, g);
                        o0[e * TPF] = make_float2(a, d);
                    }
                }
            }
        }
        __syncthreads();                                         // exchange buffer free for the next antenna pair
    }
}

// T > 1 transmit antennas (SFBC, spatial multiplexing: core/channel.py:399-466 sums the R x T independently
// faded links per receive antenna).  Same structure, one more loop: the windows of the T transmit streams of the
// symbol pass through a two-slot shared-memory buffer (the copy of step i + 1 is in flight while step i is
// accumulated), the accumulators of an antenna pair run over all T x taps links before the power is taken and the
// pair is transformed.  The cyclic-prefix samples, which only enter the power, are accumulated per thread over
// the T streams as well (cp <= 4 TPF covers normal and extended prefixes).
// coefficients: jakes_coef_kernel's [b][symbol][t][tap][slot][re|im][R2]; phases [B][R][T][taps][16].
template <int N, int K>
__global__ void __launch_bounds__(FFT_CTA_THREADS, 3)
channel_rx_fft_mt_kernel(const DevPlan P, const TdlParams C, const float2* __restrict__ tx,
                         const float* __restrict__ coef_g, float2* __restrict__ Y, double* __restrict__ power, int k0,
                         int nk, int S, int R, int R2, int T, int halo, unsigned total, bool wide) {
    constexpr int TPF = N / FFT_ELEMS, PPC = fft2_pairs_per_cta(N);
    constexpr int NC = 2 * K + 1, CPC = 4;                       // CP samples per thread: ceil(cp / TPF) <= 4
    extern __shared__ float4 smem4[];
    const int p_local = threadIdx.x / TPF, j = threadIdx.x % TPF;
    const int L = P.L, cp = P.cp;
    const int win = halo + L;
    const int ncoef = T * C.num_taps * NC * 2 * R2;              // floats per (stream, symbol)
    // shared memory: sbuf [PPC][fft_smem_elems] float4 | xw [PPC][2][win2] float2 | coef [PPC][ncoef] float
    float4* sbuf = smem4 + (size_t)p_local * fft_smem_elems(N);
    float2* xw_all = (float2*)(smem4 + (size_t)PPC * fft_smem_elems(N));
    const int win2 = (win + 1) & ~1;
    float2* xw = xw_all + (size_t)p_local * 2 * win2;
    float* sc_all = (float*)(xw_all + (size_t)PPC * 2 * win2);
    const float* sc = sc_all + (size_t)p_local * ncoef;

    const unsigned f = blockIdx.x * PPC + p_local;               // = b * S + s
    const bool valid = f < total;
    const unsigned b = valid ? f / (unsigned)S : 0u;
    const unsigned s = valid ? f - b * (unsigned)S : 0u;
    const long long m_lo = (long long)s * L - halo;

    auto stage = [&](int t, int slot) {                          // window of transmit stream t into buffer `slot`
        if (valid) {
            const float2* base = tx + ((size_t)b * T + t) * (size_t)S * L;
            const float2* src = base + m_lo;                     // may point before the stream for s = 0
            float2* dst = xw + (size_t)slot * win2;
            if (wide) {
                for (int i = 2 * j; i < win; i += 2 * TPF) {
                    const bool ok = m_lo + i >= 0;
                    cp_async16_zfill(&dst[i], ok ? &src[i] : base, ok);
                }
            } else {
                for (int i = j; i < win; i += TPF) {
                    const bool ok = m_lo + i >= 0;
                    cp_async8_zfill(&dst[i], ok ? &src[i] : base, ok);
                }
            }
        }
        cp_async_commit();
    };
    if (valid) {
        const float* cg = coef_g + (size_t)f * ncoef;
        float* scw = sc_all + (size_t)p_local * ncoef;
        for (int i = j; i < ncoef; i += TPF) cp_async4(&scw[i], &cg[i]);
    }
    stage(0, 0);                                                 // the coefficients ride in the first group

    const float tau0 = (float)(cp + j) - 0.5f * (float)(L - 1);  // polynomial argument of useful sample j
    const int RP = R2 / 2;
    const int steps = RP * T;
    int step = 0;
    for (int p = 0; p < RP; ++p) {
        c2 v[FFT_ELEMS];
        f2 cre_[CPC], cim_[CPC];                                 // cyclic-prefix samples j + c TPF, summed over the links
#pragma unroll
        for (int e = 0; e < FFT_ELEMS; ++e) v[e] = {pk(0.f, 0.f), pk(0.f, 0.f)};
#pragma unroll
        for (int c = 0; c < CPC; ++c) { cre_[c] = pk(0.f, 0.f); cim_[c] = pk(0.f, 0.f); }
        for (int t = 0; t < T; ++t, ++step) {
            cp_async_wait<0>();
            __syncthreads();                                     // step's window landed; everyone is done with the other slot
            if (step + 1 < steps) stage((step + 1) % T, (step + 1) & 1);
            if (valid) {
                const float2* xs = xw + (size_t)(step & 1) * win2;
                for (int tap = 0; tap < C.num_taps; ++tap) {
                    const float* ct = sc + ((size_t)t * C.num_taps + tap) * NC * 2 * R2 + 2 * p;
                    f2 cre[K + 1], cim[K + 1];
#pragma unroll
                    for (int k = 0; k <= K; ++k) {
                        const float2 a = *(const float2*)(ct + (k * 2 + 0) * R2), c = *(const float2*)(ct + (k * 2 + 1) * R2);
                        cre[k] = pk(a.x, a.y);
                        cim[k] = pk(c.x, c.y);
                    }
                    const float2* xt = xs + halo + cp + j - C.delay[tap];
#pragma unroll
                    for (int e = 0; e < FFT_ELEMS; ++e) {
                        const float tf = tau0 + (float)(e * TPF);
                        const f2 tau = pk(tf, tf);
                        f2 hre = cre[K], him = cim[K];
#pragma unroll
                        for (int k = K - 1; k >= 0; --k) { hre = fma2(hre, tau, cre[k]); him = fma2(him, tau, cim[k]); }
                        const float2 x = xt[e * TPF];
                        const f2 xre = pk(x.x, x.x), xim = pk(x.y, x.y), nxim = pk(-x.y, -x.y);
                        v[e].re = fma2(hre, xre, v[e].re);
                        v[e].re = fma2(him, nxim, v[e].re);
                        v[e].im = fma2(hre, xim, v[e].im);
                        v[e].im = fma2(him, xre, v[e].im);
                    }
#pragma unroll
                    for (int c = 0; c < CPC; ++c) {
                        const int i = j + c * TPF;
                        if (i < cp) {
                            const float tf = (float)i - 0.5f * (float)(L - 1);
                            const f2 tau = pk(tf, tf);
                            f2 hre = cre[K], him = cim[K];
#pragma unroll
                            for (int k = K - 1; k >= 0; --k) { hre = fma2(hre, tau, cre[k]); him = fma2(him, tau, cim[k]); }
                            const float2 x = xs[halo + i - C.delay[tap]];
                            const f2 xre = pk(x.x, x.x), xim = pk(x.y, x.y), nxim = pk(-x.y, -x.y);
                            cre_[c] = fma2(hre, xre, cre_[c]); cre_[c] = fma2(him, nxim, cre_[c]);
                            cim_[c] = fma2(hre, xim, cim_[c]); cim_[c] = fma2(him, xre, cim_[c]);
                        }
                    }
                }
            }
        }
        // power of this thread's useful and cyclic-prefix samples of the antenna pair (core/channel.py:216-218)
        float pwa = 0.f, pwb = 0.f;
        {
            f2 acc = pk(0.f, 0.f);
#pragma unroll
            for (int e = 0; e < FFT_ELEMS; ++e) { acc = fma2(v[e].re, v[e].re, acc); acc = fma2(v[e].im, v[e].im, acc); }
#pragma unroll
            for (int c = 0; c < CPC; ++c) { acc = fma2(cre_[c], cre_[c], acc); acc = fma2(cim_[c], cim_[c], acc); }
            upk(acc, pwa, pwb);
        }
        bool leader;
        if constexpr (TPF >= 32) {
            pwa = warp_sum(pwa);
            pwb = warp_sum(pwb);
            leader = (threadIdx.x & 31) == 0;
        } else {
#pragma unroll
            for (int ofs = TPF / 2; ofs > 0; ofs >>= 1) {
                pwa += __shfl_xor_sync(0xffffffffu, pwa, ofs);
                pwb += __shfl_xor_sync(0xffffffffu, pwb, ofs);
            }
            leader = j == 0;
        }
        const int r0 = 2 * p;
        const bool two = r0 + 1 < R;
        if (valid && leader) {
            atomicAdd(&power[(size_t)b * R + r0], (double)pwa * (double)N);
            if (two) atomicAdd(&power[(size_t)b * R + r0 + 1], (double)pwb * (double)N);
        }

        fft2_run<N, false>(v, sbuf, P.twiddle, j);
        if (valid) {
            float2* o0 = Y + (((size_t)b * R + r0) * S + s) * nk + (j - k0);
            float2* o1 = o0 + (size_t)S * nk;
            const unsigned kb = (unsigned)(j - k0);
#pragma unroll
            for (int e = 0; e < FFT_ELEMS; ++e) {
                if (kb + (unsigned)(e * TPF) < (unsigned)nk) {
                    float a, c, d, g;
                    upk(v[e].re, a, c);
                    upk(v[e].im, d, g);
                    o0[e * TPF] = make_float2(a, d);
                    if (two) o1[e * TPF] = make_float2(c, g);
                }
            }
        }
        __syncthreads();                                         // exchange buffer free for the next antenna pair
    }
}

template <typename F> static int dispatch_n(int N, F&& f) {
    switch (N) {
        case 64: return f(std::integral_constant<int, 64>());
        case 128: return f(std::integral_constant<int, 128>());
        case 256: return f(std::integral_constant<int, 256>());
        case 512: return f(std::integral_constant<int, 512>());
        case 1024: return f(std::integral_constant<int, 1024>());
        case 2048: return f(std::integral_constant<int, 2048>());
        default: return LTE_ERR_UNSUPPORTED;
    }
}

// Host-side checks and polynomial choice of lte_channel_rx_fft; shared with the workspace-size query.
struct FusedSetup {
    TdlParams C;
    int K, halo, R2;
    size_t ncoef, coef_bytes;
};

static int fused_setup(const lte_plan* p, const lte_channel_desc* ch, int32_t B, int32_t R, int32_t S, FusedSetup& U,
                       int32_t T = 1) {
    if (!p || !ch || B < 0 || R < 1 || R > LTE_MAX_RX || S < 1 || T < 1 || T > LTE_MAX_TX) return LTE_ERR_INVALID_ARG;
    if (ch->num_taps < 0 || ch->num_taps > LTE_MAX_TAPS) return LTE_ERR_INVALID_ARG;
    if (ch->num_taps == 0) return LTE_ERR_UNSUPPORTED;          // identity link: lte_rx_fft with rx_div = R
    const int L = p->dev.L;
    TdlParams& C = U.C;
    memset(&C, 0, sizeof(C));
    C.num_taps = ch->num_taps;
    int dmax = 0;
    for (int i = 0; i < ch->num_taps; ++i) {
        if (ch->delay[i] < 0) return LTE_ERR_INVALID_ARG;
        C.delay[i] = ch->delay[i];
        // sqrt(2/16) of the Jakes sum and the receiver FFT's 1/sqrt(N) ride on the tap gain
        C.gain[i] = (float)((double)ch->gain[i] * sqrt(2.0 / LTE_JAKES_TONES) / sqrt((double)p->dev.N));
        if (ch->delay[i] > dmax) dmax = ch->delay[i];
    }
    if (dmax > 144) return LTE_ERR_UNSUPPORTED;
    if (T > 1 && p->dev.cp > 4 * (p->dev.N / FFT_ELEMS)) return LTE_ERR_UNSUPPORTED;     // CP samples per thread of the T > 1 kernel
    U.halo = (dmax + 1) & ~1;
    double wmax = 0.0;
    for (int nn = 0; nn < LTE_JAKES_TONES; ++nn) {
        C.w_cyc[nn] = ch->doppler_hz * cos(2.0 * M_PI * (double)(nn + 1) / LTE_JAKES_TONES) / p->desc.fs;
        if (fabs(C.w_cyc[nn]) > wmax) wmax = fabs(C.w_cyc[nn]);
    }
    // one polynomial block per OFDM symbol, remainder kept below 5e-7 of |h| (1/20 of the 1e-5 parity budget):
    // K = 1 is the economised linear fit (tdl.cuh), x^2/4; K = 2 / 4 / 6 are Taylor polynomials, x^3/6, x^5/120, x^7/5040
    const double x = M_PI * wmax * L;
    if (x <= 1.41e-3) U.K = 1;
    else if (x <= 1.44e-2) U.K = 2;
    else if (x <= 0.075) U.K = 4;
    else if (x <= 0.25) U.K = 6;                                // x^7 / 5040 <= 1.2e-8
    else return LTE_ERR_UNSUPPORTED;
    C.pb = L;
    C.nbs = S;
    const long long total = (long long)B * S;
    if (total >= (1ll << 31)) return LTE_ERR_UNSUPPORTED;
    U.R2 = (R + 1) & ~1;
    const int NC = 2 * U.K + 1;
    U.ncoef = (size_t)T * C.num_taps * NC * 2 * U.R2;
    U.coef_bytes = sizeof(float) * (size_t)total * U.ncoef;
    return LTE_OK;
}

extern "C" int64_t lte_channel_rx_fft_workspace_bytes(const lte_plan* p, const lte_channel_desc* ch, int32_t B,
                                                      int32_t R, int32_t S) {
    FusedSetup U;
    const int rc = fused_setup(p, ch, B, R, S, U);
    return rc ? (int64_t)rc : (int64_t)U.coef_bytes + 16;
}

extern "C" int lte_channel_rx_fft(const lte_plan* p, const lte_channel_desc* ch, const lte_c32* tx,
                                  const float* phases, lte_c32* Y, double* power, void* workspace, int window,
                                  int32_t B, int32_t R, int32_t S, void* stream) {
    FusedSetup U;
    int rc = fused_setup(p, ch, B, R, S, U);
    if (rc) return rc;
    if (!tx || !phases || !Y || !power || !workspace || ((uintptr_t)workspace & 15)) return LTE_ERR_INVALID_ARG;
    int32_t k0, nk;
    rc = lte_plan_window(p, window, &k0, &nk);
    if (rc) return rc;
    if (B == 0) return LTE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int L = p->dev.L;
    const TdlParams& C = U.C;
    const int K = U.K, halo = U.halo, R2 = U.R2;
    const size_t ncoef = U.ncoef;
    const long long total = (long long)B * S;
    float* coef = (float*)workspace;
    if (R2 != R) LTE_CHECK_CUDA(cudaMemsetAsync(coef, 0, U.coef_bytes, st));
    const long long items = total * R * C.num_taps;
    const unsigned cgrid = (unsigned)((items + 255) / 256);
    if (K == 1) jakes_coef_kernel<1><<<cgrid, 256, 0, st>>>(C, phases, coef, R, 1, R2, items);
    else if (K == 2) jakes_coef_kernel<2><<<cgrid, 256, 0, st>>>(C, phases, coef, R, 1, R2, items);
    else if (K == 4) jakes_coef_kernel<4><<<cgrid, 256, 0, st>>>(C, phases, coef, R, 1, R2, items);
    else jakes_coef_kernel<6><<<cgrid, 256, 0, st>>>(C, phases, coef, R, 1, R2, items);
    LTE_CHECK_CUDA(cudaGetLastError());

    return dispatch_n(p->dev.N, [&](auto nn) -> int {
        constexpr int N = decltype(nn)::value;
        constexpr int PPC = fft2_pairs_per_cta(N);
        const int win2 = (halo + L + 1) & ~1;
        const size_t smem = (size_t)fft2_cta_smem_bytes(N) + (size_t)PPC * win2 * sizeof(float2) +
                            (size_t)PPC * ncoef * sizeof(float);
        if (smem > 200 * 1024) return LTE_ERR_UNSUPPORTED;
        const unsigned grid = (unsigned)((total + PPC - 1) / PPC);
        const bool wide = (L & 1) == 0 && ((uintptr_t)tx & 15) == 0;
        auto launch = [&](auto k) -> int {
            LTE_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k<<<grid, FFT_CTA_THREADS, smem, st>>>(p->dev, C, (const float2*)tx, coef, (float2*)Y, power, k0, nk, S, R,
                                                   R2, halo, (unsigned)total, wide);
            LTE_CHECK_CUDA(cudaGetLastError());
            return LTE_OK;
        };
        if (K == 1) return launch(channel_rx_fft_kernel<N, 1>);
        if (K == 2) return launch(channel_rx_fft_kernel<N, 2>);
        return K == 4 ? launch(channel_rx_fft_kernel<N, 4>) : launch(channel_rx_fft_kernel<N, 6>);
    });
}

// ---- T > 1 ---------------------------------------------------------------------------------------------------
extern "C" int64_t lte_channel_rx_fft_mimo_workspace_bytes(const lte_plan* p, const lte_channel_desc* ch, int32_t B,
                                                           int32_t R, int32_t T, int32_t S) {
    FusedSetup U;
    const int rc = fused_setup(p, ch, B, R, S, U, T);
    return rc ? (int64_t)rc : (int64_t)U.coef_bytes + 16;
}

extern "C" int lte_channel_rx_fft_mimo(const lte_plan* p, const lte_channel_desc* ch, const lte_c32* tx,
                                       const float* phases, lte_c32* Y, double* power, void* workspace, int window,
                                       int32_t B, int32_t R, int32_t T, int32_t S, void* stream) {
    FusedSetup U;
    int rc = fused_setup(p, ch, B, R, S, U, T);
    if (rc) return rc;
    if (!tx || !phases || !Y || !power || !workspace || ((uintptr_t)workspace & 15)) return LTE_ERR_INVALID_ARG;
    int32_t k0, nk;
    rc = lte_plan_window(p, window, &k0, &nk);
    if (rc) return rc;
    if (B == 0) return LTE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int L = p->dev.L;
    const TdlParams& C = U.C;
    const int K = U.K, halo = U.halo, R2 = U.R2;
    const size_t ncoef = U.ncoef;
    const long long total = (long long)B * S;
    float* coef = (float*)workspace;
    if (R2 != R) LTE_CHECK_CUDA(cudaMemsetAsync(coef, 0, U.coef_bytes, st));
    const long long items = total * R * T * C.num_taps;
    const unsigned cgrid = (unsigned)((items + 255) / 256);
    if (K == 1) jakes_coef_kernel<1><<<cgrid, 256, 0, st>>>(C, phases, coef, R, T, R2, items);
    else if (K == 2) jakes_coef_kernel<2><<<cgrid, 256, 0, st>>>(C, phases, coef, R, T, R2, items);
    else if (K == 4) jakes_coef_kernel<4><<<cgrid, 256, 0, st>>>(C, phases, coef, R, T, R2, items);
    else jakes_coef_kernel<6><<<cgrid, 256, 0, st>>>(C, phases, coef, R, T, R2, items);
    LTE_CHECK_CUDA(cudaGetLastError());

    return dispatch_n(p->dev.N, [&](auto nn) -> int {
        constexpr int N = decltype(nn)::value;
        constexpr int PPC = fft2_pairs_per_cta(N);
        const int win2 = (halo + L + 1) & ~1;
        const size_t smem = (size_t)fft2_cta_smem_bytes(N) + (size_t)PPC * 2 * win2 * sizeof(float2) +
                            (size_t)PPC * ncoef * sizeof(float);
        if (smem > 200 * 1024) return LTE_ERR_UNSUPPORTED;
        const unsigned grid = (unsigned)((total + PPC - 1) / PPC);
        // 16-byte copies need every stream start (multiples of S L samples) and the halo to be even
        const bool wide = (L & 1) == 0 && ((uintptr_t)tx & 15) == 0;
        auto launch = [&](auto k) -> int {
            LTE_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k<<<grid, FFT_CTA_THREADS, smem, st>>>(p->dev, C, (const float2*)tx, coef, (float2*)Y, power, k0, nk, S, R,
                                                   R2, T, halo, (unsigned)total, wide);
            LTE_CHECK_CUDA(cudaGetLastError());
            return LTE_OK;
        };
        if (K == 1) return launch(channel_rx_fft_mt_kernel<N, 1>);
        if (K == 2) return launch(channel_rx_fft_mt_kernel<N, 2>);
        return K == 4 ? launch(channel_rx_fft_mt_kernel<N, 4>) : launch(channel_rx_fft_mt_kernel<N, 6>);
    });
}
