// Stage 1+2 (TX: QAM map + resource grid + IFFT + CP) and stage 2 (RX: noise add +
// CP strip + FFT) kernels.  One transform per N/16 threads, 128-thread CTAs.
#include "fft.cuh"

// ------------------------------------------------------------------------------ TX
// Replaces core/modulator.py:61-88 (bits_to_symbols), core/resource_mapper.py:181-223
// (map_symbols) and core/modulator.py:242-248 (ifft * sqrt(N), CP prepend).
template <int N>
__global__ void __launch_bounds__(FFT_CTA_THREADS)
tx_map_ifft_kernel(const DevPlan P, const uint8_t* __restrict__ idx, const float2* __restrict__ symbols,
                   int T, float2* __restrict__ tx, float2* __restrict__ qam_out,
                   double* __restrict__ stats, int S, long long total) {
    constexpr int TPF = N / FFT_ELEMS, FPC = fft_per_cta(N);
    extern __shared__ float2 smem[];
    const int f_local = threadIdx.x / TPF, j = threadIdx.x % TPF;
    const long long f = (long long)blockIdx.x * FPC + f_local;   // OFDM symbol id = row*S + s
    const bool valid = f < total;
    float2* sA = smem + (size_t)f_local * 2 * fft_smem_elems(N);
    float2* sB = sA + fft_smem_elems(N);

    const long long row = valid ? f / S : 0;          // row = b*T + t
    const int t_ant = (int)(row % T);
    const long long b = row / T;
    const int s_sym = valid ? (int)(f % S) : 0;
    const int h = P.bps >> 1, mask = (1 << h) - 1;

    float2 v[FFT_ELEMS];
#pragma unroll
    for (int e = 0; e < FFT_ELEMS; ++e) {
        const int k = j + e * TPF;
        float2 val = make_float2(0.f, 0.f);
        if (valid) {
            const int m = P.bin_map[k];
            if (m >= 0) {
                if (m & BIN_PILOT_FLAG) {
                    val = P.pilots[(size_t)t_ant * P.Np + (m & (BIN_PILOT_FLAG - 1))];
                } else if (symbols) {
                    val = symbols[((size_t)row * S + s_sym) * P.Nd + m];
                } else {
                    const size_t o = ((size_t)b * S + s_sym) * P.Nd + m;
                    const int i = idx[o];
                    val = make_float2(P.lev[(i >> h) & mask], P.lev[i & mask]);
                    if (qam_out) qam_out[o] = val;
                }
            }
        }
        v[e] = val;
    }

    fft_run<N, true>(v, sA, sB, P.twiddle, j);

    float pmax = 0.f, psum = 0.f;
    if (valid) {
        float2* o = tx + (size_t)row * S * P.L + (size_t)s_sym * P.L;
        const int tail0 = N - P.cp;
#pragma unroll
        for (int e = 0; e < FFT_ELEMS; ++e) {
            const int n = j + e * TPF;
            const float2 x = cscale(v[e], P.inv_sqrt_n);
            o[P.cp + n] = x;
            const float pw = cabs2(x);
            pmax = fmaxf(pmax, pw);
            psum += pw;
            if (n >= tail0) { o[n - tail0] = x; psum += pw; }
        }
    }
    if (stats) {
        // per-stream peak and total power including the CP (core/ofdm_core.py:131-133)
        if constexpr (TPF >= 32) {
            pmax = warp_max(pmax);
            psum = warp_sum(psum);
            __shared__ float red[2][FFT_CTA_THREADS / 32];
            const int w = threadIdx.x >> 5;
            if ((threadIdx.x & 31) == 0) { red[0][w] = pmax; red[1][w] = psum; }
            __syncthreads();
#pragma unroll
            for (int q = 1; q < TPF / 32; ++q) { pmax = fmaxf(pmax, red[0][w + q]); psum += red[1][w + q]; }
        } else {
#pragma unroll
            for (int o = TPF / 2; o > 0; o >>= 1) {
                pmax = fmaxf(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
                psum += __shfl_xor_sync(0xffffffffu, psum, o);
            }
        }
        if (j == 0 && valid) {
            atomicMax((unsigned long long*)&stats[2 * row],
                      (unsigned long long)__double_as_longlong((double)pmax));
            atomicAdd(&stats[2 * row + 1], (double)psum);
        }
    }
}

// ------------------------------------------------------------------------------ RX
// Replaces core/lte_receiver.py:444-491 (strip CP, fft / sqrt(N)); optionally adds the
// AWGN of core/channel.py:216-232 while loading (sigma from the measured stream power).
template <int N>
__global__ void __launch_bounds__(FFT_CTA_THREADS)
rx_fft_kernel(const DevPlan P, const float2* __restrict__ rx, int rx_div, const double* __restrict__ power,
              const float* __restrict__ snr_lin, const float2* __restrict__ z, uint32_t key,
              unsigned long long row_id0, float2* __restrict__ Y, int k0, int nk, int S, long long total,
              int noise_freq) {
    constexpr int TPF = N / FFT_ELEMS, FPC = fft_per_cta(N);
    extern __shared__ float2 smem[];
    const int f_local = threadIdx.x / TPF, j = threadIdx.x % TPF;
    const long long f = (long long)blockIdx.x * FPC + f_local;   // row*S + s
    const bool valid = f < total;
    float2* sA = smem + (size_t)f_local * 2 * fft_smem_elems(N);
    float2* sB = sA + fft_smem_elems(N);
    const long long row = valid ? f / S : 0;
    const int s_sym = valid ? (int)(f % S) : 0;
    const size_t n_stream = (size_t)S * P.L;

    float2 v[FFT_ELEMS];
    if (valid) {
        const size_t m0 = (size_t)s_sym * P.L + P.cp;
        const float2* src = rx + (size_t)(row / rx_div) * n_stream + m0;
#pragma unroll
        for (int e = 0; e < FFT_ELEMS; ++e) v[e] = src[j + e * TPF];
        if (power && !noise_freq) {
            const float sigma = lte_sigma(power[row], (float)n_stream, snr_lin[row]);
            if (z) {
                const float2* zs = z + (size_t)row * n_stream + m0;
#pragma unroll
                for (int e = 0; e < FFT_ELEMS; ++e) {
                    const float2 w = zs[j + e * TPF];
                    v[e].x = fmaf(sigma, w.x, v[e].x);
                    v[e].y = fmaf(sigma, w.y, v[e].y);
                }
            } else {
                const uint32_t rid = (uint32_t)(row_id0 + (unsigned long long)row);
#pragma unroll
                for (int e = 0; e < FFT_ELEMS; ++e) {
                    const float2 w = lte_noise_sample(key, rid, (uint32_t)(m0 + j + e * TPF));
                    v[e].x = fmaf(sigma, w.x, v[e].x);
                    v[e].y = fmaf(sigma, w.y, v[e].y);
                }
            }
        }
    } else {
#pragma unroll
        for (int e = 0; e < FFT_ELEMS; ++e) v[e] = make_float2(0.f, 0.f);
    }

    fft_run<N, false>(v, sA, sB, P.twiddle, j);

    if (valid) {
        float2* o = Y + ((size_t)row * S + s_sym) * nk;
        if (power && noise_freq) {
            // The unitary FFT maps white Gaussian noise to white Gaussian noise of the same
            // variance, so the engine draws it directly on the bins it keeps.
            const float sigma = lte_sigma(power[row], (float)n_stream, snr_lin[row]);
            const uint32_t rid = (uint32_t)(row_id0 + (unsigned long long)row);
#pragma unroll
            for (int e = 0; e < FFT_ELEMS; ++e) {
                const int kb = j + e * TPF;
                const int k = kb - k0;
                if (k >= 0 && k < nk) {
                    const float2 w = lte_noise_sample(key, rid, (uint32_t)(s_sym * N + kb));
                    o[k] = make_float2(fmaf(v[e].x, P.inv_sqrt_n, sigma * w.x), fmaf(v[e].y, P.inv_sqrt_n, sigma * w.y));
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < FFT_ELEMS; ++e) {
                const int k = j + e * TPF - k0;
                if (k >= 0 && k < nk) o[k] = cscale(v[e], P.inv_sqrt_n);
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

extern "C" int lte_tx_map_ifft(const lte_plan* p, const uint8_t* idx, const lte_c32* symbols, int32_t T,
                               lte_c32* tx, lte_c32* qam_out, double* stats, int32_t B, int32_t S,
                               void* stream) {
    if (!p || (!idx && !symbols) || !tx || B < 0 || S < 1 || T < 1 || T > LTE_MAX_TX) return LTE_ERR_INVALID_ARG;
    if (B == 0) return LTE_OK;
    const long long total = (long long)B * T * S;
    return dispatch_n(p->dev.N, [&](auto n) -> int {
        constexpr int N = decltype(n)::value;
        auto k = tx_map_ifft_kernel<N>;
        const int smem = fft_cta_smem_bytes(N);
        LTE_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const long long grid = (total + fft_per_cta(N) - 1) / fft_per_cta(N);
        k<<<(unsigned)grid, FFT_CTA_THREADS, smem, (cudaStream_t)stream>>>(
            p->dev, idx, (const float2*)symbols, T, (float2*)tx, (float2*)qam_out, stats, S, total);
        LTE_CHECK_CUDA(cudaGetLastError());
        return LTE_OK;
    });
}

extern "C" int lte_rx_fft(const lte_plan* p, const lte_c32* rx, int32_t rx_div, const double* power,
                          const float* snr_lin, const lte_c32* z, int32_t noise_domain, uint64_t seed,
                          uint64_t row_id0, lte_c32* Y, int window, int64_t rows, int32_t S, void* stream) {
    if (!p || !rx || !Y || rows < 0 || S < 1 || rx_div < 1) return LTE_ERR_INVALID_ARG;
    if (power && !snr_lin) return LTE_ERR_INVALID_ARG;
    if (noise_domain != 0 && noise_domain != 1) return LTE_ERR_INVALID_ARG;
    if (noise_domain == 1 && z) return LTE_ERR_INVALID_ARG;   // replayed normals are time-domain draws
    int32_t k0, nk;
    int rc = lte_plan_window(p, window, &k0, &nk);
    if (rc) return rc;
    if (rows == 0) return LTE_OK;
    const long long total = (long long)rows * S;
    const uint32_t key = lte_key(seed, LTE_DOMAIN_NOISE);
    return dispatch_n(p->dev.N, [&](auto n) -> int {
        constexpr int N = decltype(n)::value;
        auto k = rx_fft_kernel<N>;
        const int smem = fft_cta_smem_bytes(N);
        LTE_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const long long grid = (total + fft_per_cta(N) - 1) / fft_per_cta(N);
        k<<<(unsigned)grid, FFT_CTA_THREADS, smem, (cudaStream_t)stream>>>(
            p->dev, (const float2*)rx, rx_div, power, snr_lin, (const float2*)z, key, row_id0, (float2*)Y, k0, nk,
            S, total, noise_domain);
        LTE_CHECK_CUDA(cudaGetLastError());
        return LTE_OK;
    });
}
