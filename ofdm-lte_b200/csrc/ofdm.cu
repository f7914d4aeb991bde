// Stage 1+2 (TX: QAM map + resource grid + IFFT + CP) and stage 2 (RX: noise add +
// CP strip + FFT) kernels on the packed-pair FFT core (fft2.cuh): N/16 threads carry two
// transforms (consecutive OFDM symbols) at once, 128-thread CTAs.
#include "fft2.cuh"

// ------------------------------------------------------------------------------ TX
// Replaces core/modulator.py:61-88 (bits_to_symbols), core/resource_mapper.py:181-223
// (map_symbols) and core/modulator.py:242-248 (ifft * sqrt(N), CP prepend).
// The grid of the symbol pair is built data-centrically in the pair's shared-memory buffer: one
// pass over the Nd data symbols (coalesced index bytes -> level LUT -> bin data_idx[d]) and one
// over the Np pilots; guards and DC stay zero.  Both symbols of the pair share the bin, so each
// item is a single 128-bit store (re_A, re_B, im_A, im_B).
struct PaprOut {                       // per-symbol PAPR outputs of the TX kernel (all optional)
    float* papr_db;                    // [rows*S]
    float2* peak_mean;                 // [rows*S] (peak, mean) power of the N useful samples
    unsigned long long* hist;          // [bins] counts of papr_db
    float lo, inv_step;
    int bins;
};

__device__ __forceinline__ void papr_emit(const PaprOut& po, size_t f, float peak, float sum, int n) {
    const float mean = sum / (float)n;
    // core/ofdm_system.py:212-217: 10 log10(peak / mean), 0 when the symbol is all-zero
    const float db = mean > 0.f ? 10.f * log10f(peak / mean) : 0.f;
    if (po.papr_db) po.papr_db[f] = db;
    if (po.peak_mean) po.peak_mean[f] = make_float2(peak, mean);
    if (po.hist) {
        int bin = (int)floorf((db - po.lo) * po.inv_step);
        bin = min(max(bin, 0), po.bins - 1);
        atomicAdd(&po.hist[bin], 1ull);
    }
}

// (max, sum) over the TPF threads of each transform pair, for both transforms of the pair
template <int TPF>
__device__ __forceinline__ void pair_reduce(float (&pmax)[2], float (&psum)[2], float (*red)[2][FFT_CTA_THREADS / 32]) {
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        float mx = pmax[m], sm = psum[m];
        if constexpr (TPF >= 32) {
            mx = warp_max(mx);
            sm = warp_sum(sm);
            const int w = threadIdx.x >> 5;
            if ((threadIdx.x & 31) == 0) { red[m][0][w] = mx; red[m][1][w] = sm; }
        } else {
#pragma unroll
            for (int ofs = TPF / 2; ofs > 0; ofs >>= 1) {
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, ofs));
                sm += __shfl_xor_sync(0xffffffffu, sm, ofs);
            }
        }
        pmax[m] = mx;
        psum[m] = sm;
    }
    if constexpr (TPF >= 32) {
        __syncthreads();
        const int w = threadIdx.x >> 5;          // only the first warp of a pair ends up with the totals
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int q = 1; q < TPF / 32; ++q) {
                pmax[m] = fmaxf(pmax[m], red[m][0][(w + q) % (FFT_CTA_THREADS / 32)]);
                psum[m] += red[m][1][(w + q) % (FFT_CTA_THREADS / 32)];
            }
    }
}

// SFBC (idx only, T = 2): the Alamouti encoder of core/sfbc_alamouti.py:45-78 rides in the index load -- data
// position c of antenna 0 carries lut[idx[c]] (negated real part on odd c: -conj), of antenna 1 lut[idx[c ^ 1]]
// (negated imaginary part on odd c: conj); idx rows hold 2 * (Nd / 2) symbols, a last odd data bin is nulled.
template <int N, bool STATS, bool SYM, bool PAPR, bool SFBC = false>
__global__ void __launch_bounds__(FFT_CTA_THREADS, 5)
tx_map_ifft_kernel(const DevPlan P, const uint8_t* __restrict__ idx, const float2* __restrict__ symbols,
                   int T, float2* __restrict__ tx, float2* __restrict__ qam_out,
                   double* __restrict__ stats, const PaprOut po, int S, unsigned total) {
    constexpr int TPF = N / FFT_ELEMS, PPC = fft2_pairs_per_cta(N);
    extern __shared__ float4 smem4[];
    __shared__ float s_lev[8];
    if (threadIdx.x < 8) s_lev[threadIdx.x] = P.lev[threadIdx.x];
    const int p_local = threadIdx.x / TPF, j = threadIdx.x % TPF;
    float4* sbuf = smem4 + (size_t)p_local * fft_smem_elems(N);
    const unsigned f0 = (blockIdx.x * PPC + p_local) * 2u;   // OFDM symbol ids f0, f0+1 (= row*S + s)
    const int h = P.bps >> 1, mask = (1 << h) - 1;
    const int Nd = P.Nd;

    // symbol ids f0, f0+1 -> (row, symbol); one division, the second symbol follows the first
    const bool valid[2] = {f0 < total, f0 + 1 < total};
    unsigned row[2], s_sym[2];
    size_t ibase[2];
    row[0] = valid[0] ? f0 / (unsigned)S : 0u;      // row = b*T + t
    s_sym[0] = valid[0] ? f0 - row[0] * (unsigned)S : 0u;
    const bool wrap = s_sym[0] + 1 == (unsigned)S;
    row[1] = valid[1] ? (wrap ? row[0] + 1 : row[0]) : 0u;
    s_sym[1] = valid[1] ? (wrap ? 0u : s_sym[0] + 1) : 0u;
    unsigned tq[2], tr[2];                          // row / T, row % T
    tq[0] = T == 1 ? row[0] : row[0] / (unsigned)T;
    tr[0] = row[0] - tq[0] * (unsigned)T;
    tq[1] = row[1] == row[0] ? tq[0] : (T == 1 ? row[1] : row[1] / (unsigned)T);
    tr[1] = row[1] - tq[1] * (unsigned)T;
#pragma unroll
    for (int m = 0; m < 2; ++m)
        ibase[m] = SYM ? ((size_t)row[m] * S + s_sym[m]) * Nd
                       : ((size_t)tq[m] * S + s_sym[m]) * (SFBC ? 2 * (Nd / 2) : Nd);
    // Bin-centric build straight into the IFFT's register layout (thread j owns bins j + e TPF): the
    // class of every bin comes from the plan's bin_map (data slot / pilot / null), all index or
    // symbol loads are issued first, then the constellation lookups; no shared-memory grid, no
    // barrier before the first butterfly pass.
    __syncthreads();          // s_lev
    const float2* pa = P.pilots + (size_t)tr[0] * P.Np;
    const float2* pc = P.pilots + (size_t)tr[1] * P.Np;
    // element e of every thread of the transform covers bins [e TPF, (e + 1) TPF): outside the occupied
    // window [k0_useful, k0_useful + nk_useful) the whole group is null and skips its loads (a uniform branch)
    const int used_lo = P.k0_useful, used_hi = P.k0_useful + P.nk_useful;
    int code[FFT_ELEMS];
#pragma unroll
    for (int e = 0; e < FFT_ELEMS; ++e) {
        const bool used = (e + 1) * TPF > used_lo && e * TPF < used_hi;
        code[e] = used ? (int)__ldg(&P.bin_map[j + e * TPF]) : BIN_NULL;
    }
    c2 v[FFT_ELEMS];
    if constexpr (SYM) {
        const float2* sp0 = symbols + ibase[0];
        const float2* sp1 = symbols + ibase[1];
#pragma unroll
        for (int e = 0; e < FFT_ELEMS; ++e) {
            float2 a = make_float2(0.f, 0.f), g = make_float2(0.f, 0.f);
            if ((e + 1) * TPF > used_lo && e * TPF < used_hi) {
                const int c = code[e];
                const bool pil = c >= 0 && (c & BIN_PILOT_FLAG), dat = c >= 0 && !pil;
                const int q = c & (BIN_PILOT_FLAG - 1);
                if (dat) {
                    if (valid[0]) a = sp0[q];
                    if (valid[1]) g = sp1[q];
                } else if (pil) {
                    if (valid[0]) a = pa[q];
                    if (valid[1]) g = pc[q];
                }
            }
            v[e] = {pk(a.x, g.x), pk(a.y, g.y)};
        }
    } else {
        const uint8_t* ip[2] = {idx + ibase[0], idx + ibase[1]};
        uint8_t ib[2][FFT_ELEMS];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int e = 0; e < FFT_ELEMS; ++e) {
                ib[m][e] = 0;
                if ((e + 1) * TPF > used_lo && e * TPF < used_hi) {
                    const int c = code[e];
                    if (SFBC) {
                        if (valid[m] && c >= 0 && !(c & BIN_PILOT_FLAG) && c < 2 * (Nd / 2)) ib[m][e] = ip[m][c ^ (int)tr[m]];
                    } else if (valid[m] && c >= 0 && !(c & BIN_PILOT_FLAG)) ib[m][e] = ip[m][c];
                }
            }
#pragma unroll
        for (int e = 0; e < FFT_ELEMS; ++e) {
            float2 a = make_float2(0.f, 0.f), g = make_float2(0.f, 0.f);
            if ((e + 1) * TPF > used_lo && e * TPF < used_hi) {
                const int c = code[e];
                const bool pil = c >= 0 && (c & BIN_PILOT_FLAG), dat = c >= 0 && !pil;
                if (dat) {
                    const int ia = ib[0][e], ic = ib[1][e];
                    if (valid[0]) a = make_float2(s_lev[(ia >> h) & mask], s_lev[ia & mask]);
                    if (valid[1]) g = make_float2(s_lev[(ic >> h) & mask], s_lev[ic & mask]);
                    if (SFBC) {
                        if (c >= 2 * (Nd / 2)) { a = make_float2(0.f, 0.f); g = a; }
                        else if (c & 1) {
                            a = tr[0] ? make_float2(a.x, -a.y) : make_float2(-a.x, a.y);
                            g = tr[1] ? make_float2(g.x, -g.y) : make_float2(-g.x, g.y);
                        }
                    }
                    if (!SFBC && qam_out) {
                        if (valid[0]) qam_out[ibase[0] + c] = a;
                        if (valid[1]) qam_out[ibase[1] + c] = g;
                    }
                } else if (pil) {
                    const int q = c & (BIN_PILOT_FLAG - 1);
                    if (valid[0]) a = pa[q];
                    if (valid[1]) g = pc[q];
                }
            }
            v[e] = {pk(a.x, g.x), pk(a.y, g.y)};
        }
    }

    fft2_run<N, true>(v, sbuf, P.twiddle, j);

    const f2 scale = pk(P.inv_sqrt_n, P.inv_sqrt_n);
    float pmax[2] = {0.f, 0.f}, psum[2] = {0.f, 0.f};       // stream statistics (CP counted)
    float qmax[2] = {0.f, 0.f}, qsum[2] = {0.f, 0.f};       // per-symbol statistics (useful part only)
    const int tail0 = N - P.cp;
    float2* o[2];
#pragma unroll
    for (int m = 0; m < 2; ++m) o[m] = tx + (size_t)row[m] * S * P.L + (size_t)s_sym[m] * P.L + P.cp;
    const bool write = !PAPR || tx != nullptr;
#pragma unroll
    for (int e = 0; e < FFT_ELEMS; ++e) {
        const int n = j + e * TPF;
        float re[2], im[2];
        upk(mul2(v[e].re, scale), re[0], re[1]);
        upk(mul2(v[e].im, scale), im[0], im[1]);
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            if (valid[m]) {
                const float2 xo = make_float2(re[m], im[m]);
                if (write) {
                    o[m][n] = xo;
                    if (n >= tail0) o[m][n - N] = xo;     // cyclic prefix: last cp samples again in front
                }
                if (STATS || PAPR) {
                    const float pw = cabs2(xo);
                    if (STATS) {
                        pmax[m] = fmaxf(pmax[m], pw);
                        psum[m] += (n >= tail0) ? 2.f * pw : pw;
                    }
                    if (PAPR) {
                        qmax[m] = fmaxf(qmax[m], pw);
                        qsum[m] += pw;
                    }
                }
            }
        }
    }
    __shared__ float red[2][2][FFT_CTA_THREADS / 32];
    if (STATS) {
        // per-stream peak and total power including the CP (core/ofdm_core.py:131-133)
        pair_reduce<TPF>(pmax, psum, red);
        if (j == 0) {
#pragma unroll
            for (int m = 0; m < 2; ++m)
                if (valid[m]) {
                    atomicMax((unsigned long long*)&stats[2 * row[m]],
                              (unsigned long long)__double_as_longlong((double)pmax[m]));
                    atomicAdd(&stats[2 * row[m] + 1], (double)psum[m]);
                }
        }
    }
    if (PAPR) {
        // per-symbol PAPR of the useful part (core/ofdm_system.py:173-229)
        if (STATS) __syncthreads();
        pair_reduce<TPF>(qmax, qsum, red);
        if (j == 0) {
#pragma unroll
            for (int m = 0; m < 2; ++m)
                if (valid[m]) papr_emit(po, (size_t)row[m] * S + s_sym[m], qmax[m], qsum[m], N);
        }
    }
}

// Per-symbol PAPR of an existing time-domain stream (core/ofdm_system.py:116-171 with the CP,
// :173-229 without); one warp per OFDM symbol.
__global__ void papr_symbols_kernel(const DevPlan P, const float2* __restrict__ x, int include_cp, const PaprOut po,
                                    long long total) {
    const long long f = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (f >= total) return;
    const int lane = threadIdx.x & 31;
    const int n0 = include_cp ? 0 : P.cp, cnt = include_cp ? P.L : P.N;
    const float2* src = x + (size_t)f * P.L + n0;
    float mx = 0.f, sm = 0.f;
    for (int i = lane; i < cnt; i += 32) {
        const float pw = cabs2(src[i]);
        mx = fmaxf(mx, pw);
        sm += pw;
    }
    mx = warp_max(mx);
    sm = warp_sum(sm);
    if (lane == 0) papr_emit(po, (size_t)f, mx, sm, cnt);
}

__global__ void histogram_kernel(const float* __restrict__ x, long long n, float lo, float inv_step, int bins,
                                 unsigned long long* __restrict__ hist) {
    extern __shared__ unsigned int sh[];
    for (int i = threadIdx.x; i < bins; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int bin = (int)floorf((x[i] - lo) * inv_step);
        bin = min(max(bin, 0), bins - 1);
        atomicAdd(&sh[bin], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

// ------------------------------------------------------------------------------ RX
// Replaces core/lte_receiver.py:444-491 (strip CP, fft / sqrt(N)); optionally adds the
// AWGN of core/channel.py:216-232 (sigma from the measured stream power).
// NOISE: 0 none, 1 per time sample before the FFT (injected normals or Philox), 2 Philox on the
// kept output bins after the FFT.  Loader and epilogue are rolled loops through the pair's
// shared-memory buffer so the Philox/Box-Muller body exists once in the instruction stream.
template <int N, int NOISE>
__global__ void __launch_bounds__(FFT_CTA_THREADS)
rx_fft_kernel(const DevPlan P, const float2* __restrict__ rx, int rx_div, const double* __restrict__ power,
              const float* __restrict__ snr_lin, const float2* __restrict__ z, uint32_t key,
              unsigned long long row_id0, float2* __restrict__ Y, int k0, int nk, int S, unsigned total) {
    constexpr int TPF = N / FFT_ELEMS, PPC = fft2_pairs_per_cta(N);
    extern __shared__ float4 smem4[];
    __shared__ float s_sigma[PPC][2];
    const int p_local = threadIdx.x / TPF, j = threadIdx.x % TPF;
    float4* sbuf = smem4 + (size_t)p_local * fft_smem_elems(N);
    const unsigned f0 = (blockIdx.x * PPC + p_local) * 2u;
    const size_t n_stream = (size_t)S * P.L;

    // symbol ids f0, f0+1 -> (row, symbol); one division, the second symbol follows the first
    bool valid[2] = {f0 < total, f0 + 1 < total};
    unsigned row[2], s_sym[2];
    row[0] = valid[0] ? f0 / (unsigned)S : 0u;
    s_sym[0] = valid[0] ? f0 - row[0] * (unsigned)S : 0u;
    const bool wrap = s_sym[0] + 1 == (unsigned)S;
    row[1] = valid[1] ? (wrap ? row[0] + 1 : row[0]) : 0u;
    s_sym[1] = valid[1] ? (wrap ? 0u : s_sym[0] + 1) : 0u;
    uint32_t rid[2];
    const float2* src[2];
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const unsigned srow = rx_div == 1 ? row[m] : row[m] / (unsigned)rx_div;
        src[m] = rx + (size_t)srow * n_stream + (size_t)s_sym[m] * P.L + P.cp;
        rid[m] = (uint32_t)(row_id0 + (unsigned long long)row[m]);
    }
    if (NOISE != 0 && j < 2)       // one thread per symbol turns the stream power into sigma
        s_sigma[p_local][j] = valid[j] ? lte_sigma(power[row[j]], (float)n_stream, snr_lin[row[j]]) : 0.f;

    c2 v[FFT_ELEMS];
    if constexpr (NOISE == 1) {
        __syncthreads();
        const float sigma[2] = {s_sigma[p_local][0], s_sigma[p_local][1]};
#pragma unroll 2
        for (int e = 0; e < FFT_ELEMS; ++e) {
            const int n = j + e * TPF;
            float2 x[2];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                x[m] = make_float2(0.f, 0.f);
                if (valid[m]) {
                    x[m] = src[m][n];
                    const size_t ms = (size_t)s_sym[m] * P.L + P.cp + n;
                    const float2 w = z ? z[(size_t)row[m] * n_stream + ms] : lte_noise_sample(key, rid[m], (uint32_t)ms);
                    x[m].x = fmaf(sigma[m], w.x, x[m].x);
                    x[m].y = fmaf(sigma[m], w.y, x[m].y);
                }
            }
            sbuf[fft_pad(n)] = make_float4(x[0].x, x[1].x, x[0].y, x[1].y);    // read back by this thread only
        }
#pragma unroll
        for (int e = 0; e < FFT_ELEMS; ++e) {
            const float4 q = sbuf[fft_pad(j + e * TPF)];
            v[e] = {pk(q.x, q.y), pk(q.z, q.w)};
        }
        __syncthreads();
    } else {
#pragma unroll
        for (int e = 0; e < FFT_ELEMS; ++e) {
            const float2 a = valid[0] ? src[0][j + e * TPF] : make_float2(0.f, 0.f);
            const float2 b = valid[1] ? src[1][j + e * TPF] : make_float2(0.f, 0.f);
            v[e] = {pk(a.x, b.x), pk(a.y, b.y)};
        }
    }

    fft2_run<N, false>(v, sbuf, P.twiddle, j);

    __syncthreads();          // all exchange reads are done
    const f2 scale = pk(P.inv_sqrt_n, P.inv_sqrt_n);
#pragma unroll
    for (int e = 0; e < FFT_ELEMS; ++e) {
        float a, b, c, d;
        upk(mul2(v[e].re, scale), a, b);
        upk(mul2(v[e].im, scale), c, d);
        sbuf[fft_pad(j + e * TPF)] = make_float4(a, b, c, d);
    }
    float2* o[2];
#pragma unroll
    for (int m = 0; m < 2; ++m) o[m] = Y + ((size_t)row[m] * S + s_sym[m]) * nk;
    float sigma[2] = {0.f, 0.f};
    if (NOISE == 2) { sigma[0] = s_sigma[p_local][0]; sigma[1] = s_sigma[p_local][1]; }
#pragma unroll 2
    for (int e = 0; e < FFT_ELEMS; ++e) {
        const int kb = j + e * TPF;
        const int k = kb - k0;
        if (k >= 0 && k < nk) {
            const float4 q = sbuf[fft_pad(kb)];
            float2 out[2] = {make_float2(q.x, q.z), make_float2(q.y, q.w)};
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                if (valid[m]) {
                    if constexpr (NOISE == 2) {
                        // The unitary FFT maps white Gaussian noise to white Gaussian noise of the same
                        // variance, so the engine draws it directly on the bins it keeps.
                        const float2 w = lte_noise_sample(key, rid[m], (uint32_t)(s_sym[m] * N + kb));
                        out[m].x = fmaf(sigma[m], w.x, out[m].x);
                        out[m].y = fmaf(sigma[m], w.y, out[m].y);
                    }
                    o[m][k] = out[m];
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------ launchers
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

static int launch_tx(const lte_plan* p, const uint8_t* idx, const lte_c32* symbols, int32_t T, lte_c32* tx,
                     lte_c32* qam_out, double* stats, const PaprOut* po, int32_t B, int32_t S, void* stream) {
    if (B == 0) return LTE_OK;
    const long long total = (long long)B * T * S;
    if (total >= (1ll << 31)) return LTE_ERR_UNSUPPORTED;
    const PaprOut none = {nullptr, nullptr, nullptr, 0.f, 1.f, 1};
    return dispatch_n(p->dev.N, [&](auto n) -> int {
        constexpr int N = decltype(n)::value;
        const int smem = fft2_cta_smem_bytes(N);
        const long long per = 2 * fft2_pairs_per_cta(N);
        const long long grid = (total + per - 1) / per;
        auto launch = [&](auto k) -> int {
            LTE_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            k<<<(unsigned)grid, FFT_CTA_THREADS, smem, (cudaStream_t)stream>>>(
                p->dev, idx, (const float2*)symbols, T, (float2*)tx, (float2*)qam_out, stats, po ? *po : none, S,
                (unsigned)total);
            LTE_CHECK_CUDA(cudaGetLastError());
            return LTE_OK;
        };
        if (po) {
            if (symbols) return stats ? launch(tx_map_ifft_kernel<N, true, true, true>) : launch(tx_map_ifft_kernel<N, false, true, true>);
            return stats ? launch(tx_map_ifft_kernel<N, true, false, true>) : launch(tx_map_ifft_kernel<N, false, false, true>);
        }
        if (symbols) return stats ? launch(tx_map_ifft_kernel<N, true, true, false>) : launch(tx_map_ifft_kernel<N, false, true, false>);
        return stats ? launch(tx_map_ifft_kernel<N, true, false, false>) : launch(tx_map_ifft_kernel<N, false, false, false>);
    });
}

extern "C" int lte_tx_map_ifft(const lte_plan* p, const uint8_t* idx, const lte_c32* symbols, int32_t T,
                               lte_c32* tx, lte_c32* qam_out, double* stats, int32_t B, int32_t S,
                               void* stream) {
    if (!p || (!idx && !symbols) || !tx || B < 0 || S < 1 || T < 1 || T > LTE_MAX_TX) return LTE_ERR_INVALID_ARG;
    return launch_tx(p, idx, symbols, T, tx, qam_out, stats, nullptr, B, S, stream);
}

extern "C" int lte_tx_sfbc_ifft(const lte_plan* p, const uint8_t* idx, lte_c32* tx, int32_t B, int32_t S, void* stream) {
    if (!p || !idx || !tx || B < 0 || S < 1) return LTE_ERR_INVALID_ARG;
    if (p->nsets < 2 || p->dev.Nd < 2) return LTE_ERR_INVALID_ARG;          // the two SFBC pilot sets
    if (B == 0) return LTE_OK;
    const long long total = (long long)B * 2 * S;
    if (total >= (1ll << 31)) return LTE_ERR_UNSUPPORTED;
    const PaprOut none = {nullptr, nullptr, nullptr, 0.f, 1.f, 1};
    return dispatch_n(p->dev.N, [&](auto n) -> int {
        constexpr int N = decltype(n)::value;
        const int smem = fft2_cta_smem_bytes(N);
        const long long per = 2 * fft2_pairs_per_cta(N);
        const long long grid = (total + per - 1) / per;
        auto k = tx_map_ifft_kernel<N, false, false, false, true>;
        LTE_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k<<<(unsigned)grid, FFT_CTA_THREADS, smem, (cudaStream_t)stream>>>(p->dev, idx, nullptr, 2, (float2*)tx, nullptr, nullptr,
                                                                       none, S, (unsigned)total);
        LTE_CHECK_CUDA(cudaGetLastError());
        return LTE_OK;
    });
}

static int make_papr_out(PaprOut& po, float* papr_db, float* peak_mean, unsigned long long* hist, float hist_lo,
                         float hist_step, int32_t hist_bins) {
    if (hist && (hist_bins < 1 || !(hist_step > 0.f))) return LTE_ERR_INVALID_ARG;
    po = {papr_db, (float2*)peak_mean, hist, hist_lo, hist ? 1.f / hist_step : 1.f, hist ? hist_bins : 1};
    return LTE_OK;
}

extern "C" int lte_tx_papr(const lte_plan* p, const uint8_t* idx, const lte_c32* symbols, int32_t T, lte_c32* tx,
                           double* stats, float* papr_db, float* peak_mean, unsigned long long* hist, float hist_lo,
                           float hist_step, int32_t hist_bins, int32_t B, int32_t S, void* stream) {
    if (!p || (!idx && !symbols) || B < 0 || S < 1 || T < 1 || T > LTE_MAX_TX) return LTE_ERR_INVALID_ARG;
    if (!papr_db && !peak_mean && !hist) return LTE_ERR_INVALID_ARG;
    PaprOut po;
    int rc = make_papr_out(po, papr_db, peak_mean, hist, hist_lo, hist_step, hist_bins);
    if (rc) return rc;
    return launch_tx(p, idx, symbols, T, tx, nullptr, stats, &po, B, S, stream);
}

extern "C" int lte_papr_symbols(const lte_plan* p, const lte_c32* x, int32_t include_cp, float* papr_db,
                                float* peak_mean, unsigned long long* hist, float hist_lo, float hist_step,
                                int32_t hist_bins, int64_t rows, int32_t S, void* stream) {
    if (!p || !x || rows < 0 || S < 1) return LTE_ERR_INVALID_ARG;
    if (!papr_db && !peak_mean && !hist) return LTE_ERR_INVALID_ARG;
    PaprOut po;
    int rc = make_papr_out(po, papr_db, peak_mean, hist, hist_lo, hist_step, hist_bins);
    if (rc) return rc;
    const long long total = (long long)rows * S;
    if (total == 0) return LTE_OK;
    const long long grid = (total * 32 + 255) / 256;
    if (grid >= (1ll << 31)) return LTE_ERR_UNSUPPORTED;
    papr_symbols_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(p->dev, (const float2*)x, include_cp, po, total);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

extern "C" int lte_histogram(const float* x, int64_t n, float lo, float step, int32_t bins, unsigned long long* hist,
                             void* stream) {
    if (!x || !hist || n < 0 || bins < 1 || bins > 8192 || !(step > 0.f)) return LTE_ERR_INVALID_ARG;
    if (n == 0) return LTE_OK;
    int sms = 148;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = (n + 255) / 256;
    const unsigned grid = (unsigned)(want < 4ll * sms ? want : 4ll * sms);
    histogram_kernel<<<grid, 256, bins * sizeof(unsigned int), (cudaStream_t)stream>>>(x, n, lo, 1.f / step, bins, hist);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

extern "C" int lte_rx_fft(const lte_plan* p, const lte_c32* rx, int32_t rx_div, const double* power,
                          const float* snr_lin, const lte_c32* z, int32_t noise_domain, uint64_t seed,
                          uint64_t row_id0, lte_c32* Y, int window, int64_t rows, int32_t S, void* stream) {
    if (!p || !rx || !Y || rows < 0 || S < 1 || rx_div < 1) return LTE_ERR_INVALID_ARG;
    if (power && !snr_lin) return LTE_ERR_INVALID_ARG;
    if (noise_domain != 0 && noise_domain != 1) return LTE_ERR_INVALID_ARG;
    if (noise_domain == 1 && z) return LTE_ERR_INVALID_ARG;   // replayed normals are time-domain draws
    if (power && !z && !lte_ids_fit(row_id0, (uint64_t)rows)) return LTE_ERR_UNSUPPORTED;
    int32_t k0, nk;
    int rc = lte_plan_window(p, window, &k0, &nk);
    if (rc) return rc;
    if (rows == 0) return LTE_OK;
    const long long total = (long long)rows * S;
    if (total >= (1ll << 31)) return LTE_ERR_UNSUPPORTED;
    const uint32_t key = lte_key(seed, LTE_DOMAIN_NOISE);
    return dispatch_n(p->dev.N, [&](auto n) -> int {
        constexpr int N = decltype(n)::value;
        const int smem = fft2_cta_smem_bytes(N);
        const long long per = 2 * fft2_pairs_per_cta(N);
        const long long grid = (total + per - 1) / per;
        auto launch = [&](auto k) -> int {
            LTE_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            k<<<(unsigned)grid, FFT_CTA_THREADS, smem, (cudaStream_t)stream>>>(
                p->dev, (const float2*)rx, rx_div, power, snr_lin, (const float2*)z, key, row_id0, (float2*)Y, k0,
                nk, S, (unsigned)total);
            LTE_CHECK_CUDA(cudaGetLastError());
            return LTE_OK;
        };
        if (!power) return launch(rx_fft_kernel<N, 0>);
        if (noise_domain == 0) return launch(rx_fft_kernel<N, 1>);
        return launch(rx_fft_kernel<N, 2>);
    });
}
