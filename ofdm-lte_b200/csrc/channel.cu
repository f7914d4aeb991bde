// Stage 3: ITU-R M.1225 tapped-delay-line Rayleigh channel in the time domain
// (core/rayleighchannel.py:20-58) plus the AWGN of core/channel.py:46-66 / :216-232,
// and the Philox helpers of the sweep engine.
//
// Jakes fading  h_i[m] = sqrt(2/16) * sum_n exp(j(2 pi fD cos(a_n) m / fs + phi_n))
// is evaluated per polynomial block of PB samples: the 16-tone sum and its first K
// derivatives are formed once at the block centre (phase reduced in fp64), then every
// thread evaluates the degree-K Taylor polynomial at its own samples.  PB is chosen on
// the host so that the truncation error stays below 2e-8 (see lte_channel_tdl).
#include "common.cuh"

#define TDL_THREADS 256

struct TdlParams {
    int num_taps;
    int delay[LTE_MAX_TAPS];
    float gain[LTE_MAX_TAPS];           // includes sqrt(2/16)
    double w_cyc[LTE_JAKES_TONES];      // fD cos(alpha_n) / fs   [cycles per sample]
    int halo;                           // >= max delay, multiple of 8
    int pb;                             // polynomial block length (power of two <= tile)
    int xs_stride;                      // columns per shared-memory row, == 2 (mod 16)
};

// R receive antennas, V consecutive samples per thread, degree-K Taylor polynomial.
// Shared-memory sample layout: sample s of the staged window lives at [s & 7][s >> 3], so a
// warp reading "sample 8*t + c" for consecutive threads t touches consecutive float2 (no
// bank conflicts for any tap delay) while the global loads that fill it stay coalesced.
template <int R, int V, int K>
__global__ void __launch_bounds__(TDL_THREADS, 2)
tdl_kernel(const TdlParams C, const float2* __restrict__ tx, const float* __restrict__ phases,
           float2* __restrict__ faded, double* __restrict__ power, int T, long long n, int tiles) {
    constexpr int TILE = TDL_THREADS * V;
    constexpr int NC = 2 * K + 1;                 // K+1 value coefficients, K derivative coefficients
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* sx = (float2*)smem_raw;               // [T][8][xs_stride]
    float2* coef = sx + (size_t)T * 8 * C.xs_stride;   // [nblk][R*T*taps][NC]
    const int nblk = TILE / C.pb;
    const int nlt = R * T * C.num_taps;

    const long long b = blockIdx.x / tiles;
    const long long tile0 = (long long)(blockIdx.x % tiles) * TILE;
    const int tid = threadIdx.x;

    // ---- stage the TX samples (with the delay halo) ---------------------------------
    const int span = C.halo + TILE;
    for (int t = 0; t < T; ++t) {
        const float2* src = tx + ((size_t)b * T + t) * n;
        float2* dst = sx + (size_t)t * 8 * C.xs_stride;
        for (int i = tid; i < span; i += TDL_THREADS) {
            const long long m = tile0 - C.halo + i;
            dst[(i & 7) * C.xs_stride + (i >> 3)] = (m >= 0 && m < n) ? __ldg(&src[m]) : make_float2(0.f, 0.f);
        }
    }

    // ---- Taylor coefficients of every (rx, tx, tap) fading process per block ----------
    // work item = (block, triple, tone); 16 consecutive lanes reduce one sum
    const int items = nblk * nlt * LTE_JAKES_TONES;
    for (int it0 = 0; it0 < items; it0 += TDL_THREADS) {
        const int it = it0 + tid;
        float2 a[K + 1];
#pragma unroll
        for (int k = 0; k <= K; ++k) a[k] = make_float2(0.f, 0.f);
        int blk = 0, trip = 0;
        if (it < items) {
            const int tone = it & (LTE_JAKES_TONES - 1);
            trip = (it >> 4) % nlt;
            blk = (it >> 4) / nlt;
            const double mc = (double)tile0 + (double)blk * C.pb + 0.5 * (C.pb - 1);
            const float u = phases[((size_t)b * nlt + trip) * LTE_JAKES_TONES + tone];
            double turns = C.w_cyc[tone] * mc + (double)u;
            turns -= floor(turns);
            float sn, cs;
            sincospif(2.0f * (float)turns, &sn, &cs);
            const float x = (float)(6.283185307179586 * C.w_cyc[tone]);   // rad / sample
            // e^{j theta} (j x)^k / k!
            float2 term = make_float2(cs, sn);
            a[0] = term;
#pragma unroll
            for (int k = 1; k <= K; ++k) {
                const float f = x / (float)k;
                term = make_float2(-term.y * f, term.x * f);
                a[k] = term;
            }
        }
#pragma unroll
        for (int k = 0; k <= K; ++k) {
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                a[k].x += __shfl_xor_sync(0xffffffffu, a[k].x, o);
                a[k].y += __shfl_xor_sync(0xffffffffu, a[k].y, o);
            }
        }
        if (it < items && (it & (LTE_JAKES_TONES - 1)) == 0) {
            const float g = C.gain[trip % C.num_taps];
            float2* c = coef + ((size_t)blk * nlt + trip) * NC;
#pragma unroll
            for (int k = 0; k <= K; ++k) c[k] = cscale(a[k], g);
#pragma unroll
            for (int k = 1; k <= K; ++k) c[K + k] = cscale(a[k], g * (float)k);   // derivative: k c_k
        }
    }
    __syncthreads();

    // ---- y_r[m] = sum_t sum_i g_i h_{r,t,i}[m] x_t[m - d_i] ----------------------------
    const int l0 = tid * V;                         // first local sample of this thread
    const long long m0 = tile0 + l0;
    const int blk = l0 / C.pb;
    // polynomial argument at the centre of the thread's V samples; linear stepping inside
    const float tau = (float)(l0 - blk * C.pb) + 0.5f * (V - 1) - 0.5f * (C.pb - 1);
    float2 acc[R][V];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int i = 0; i < V; ++i) acc[r][i] = make_float2(0.f, 0.f);

    for (int t = 0; t < T; ++t) {
        const float2* sxt = sx + (size_t)t * 8 * C.xs_stride;
        for (int tap = 0; tap < C.num_taps; ++tap) {
            const int c0 = C.halo - C.delay[tap] + (V == 8 ? 0 : l0 & 7);
            const int colbase = (V == 8) ? tid : (l0 >> 3);
            float2 xv[V];
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const int cc = c0 + i;
                xv[i] = sxt[(cc & 7) * C.xs_stride + colbase + (cc >> 3)];
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float2* c = coef + ((size_t)blk * nlt + (r * T + t) * C.num_taps + tap) * NC;
                float2 h = c[K], dh = c[2 * K];
#pragma unroll
                for (int k = K - 1; k >= 0; --k) {
                    const float2 ck = c[k];
                    h.x = fmaf(h.x, tau, ck.x);
                    h.y = fmaf(h.y, tau, ck.y);
                }
#pragma unroll
                for (int k = K - 1; k >= 1; --k) {
                    const float2 dk = c[K + k];
                    dh.x = fmaf(dh.x, tau, dk.x);
                    dh.y = fmaf(dh.y, tau, dk.y);
                }
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    const float st = (float)i - 0.5f * (V - 1);
                    const float2 hi = make_float2(fmaf(st, dh.x, h.x), fmaf(st, dh.y, h.y));
                    acc[r][i].x = fmaf(hi.x, xv[i].x, fmaf(-hi.y, xv[i].y, acc[r][i].x));
                    acc[r][i].y = fmaf(hi.x, xv[i].y, fmaf(hi.y, xv[i].x, acc[r][i].y));
                }
            }
        }
    }

    // ---- store + power --------------------------------------------------------------------
    __shared__ float pw_red[TDL_THREADS / 32][LTE_MAX_RX];
    const bool vec = ((n & 1) == 0);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float pw = 0.f;
        float2* dst = faded + ((size_t)b * R + r) * n + m0;
        if (m0 + V <= n && vec) {
            float4* d4 = (float4*)dst;
#pragma unroll
            for (int i = 0; i < V; i += 2) {
                d4[i / 2] = make_float4(acc[r][i].x, acc[r][i].y, acc[r][i + 1].x, acc[r][i + 1].y);
                pw += cabs2(acc[r][i]) + cabs2(acc[r][i + 1]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < V; ++i)
                if (m0 + i < n) { dst[i] = acc[r][i]; pw += cabs2(acc[r][i]); }
        }
        pw = warp_sum(pw);
        if ((tid & 31) == 0) pw_red[tid >> 5][r] = pw;
    }
    __syncthreads();
    if (tid < R) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < TDL_THREADS / 32; ++w) s += pw_red[w][tid];
        atomicAdd(&power[(size_t)b * R + tid], (double)s);
    }
}

// power of an un-faded stream (AWGN channel type): power[row] = sum |x[row / x_div]|^2
__global__ void power_kernel(const float2* __restrict__ x, int x_div, double* __restrict__ power, long long n,
                             int gx) {
    const long long row = blockIdx.x / gx;
    const int bx = blockIdx.x % gx;
    const float2* src = x + (size_t)(row / x_div) * n;
    float s = 0.f;
    for (long long i = (long long)bx * blockDim.x + threadIdx.x; i < n; i += (long long)gx * blockDim.x)
        s += cabs2(src[i]);
    s = warp_sum(s);
    __shared__ float red[32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        t = warp_sum(t);
        if (threadIdx.x == 0) atomicAdd(&power[row], (double)t);
    }
}

extern "C" int lte_channel_tdl(const lte_plan* p, const lte_channel_desc* ch, const lte_c32* tx,
                               const float* phases, lte_c32* faded, double* power, int32_t B, int32_t R,
                               int32_t T, int64_t n, void* stream) {
    if (!p || !ch || !tx || !power || B < 0 || R < 1 || R > LTE_MAX_RX || T < 1 || T > LTE_MAX_TX || n < 1)
        return LTE_ERR_INVALID_ARG;
    if (ch->num_taps < 0 || ch->num_taps > LTE_MAX_TAPS) return LTE_ERR_INVALID_ARG;
    if (B == 0) return LTE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (ch->num_taps == 0) {
        // identity link (channel_type 'awgn'): every RX antenna sees the TX stream, so only
        // the per-(stream, antenna) power is produced; `faded` is not written.
        if (T != 1) return LTE_ERR_UNSUPPORTED;
        int gx = (int)((n + 256 * 8 - 1) / (256 * 8));
        if (gx > 32) gx = 32;
        power_kernel<<<(unsigned)((long long)gx * B * R), 256, 0, st>>>((const float2*)tx, R, power, n, gx);
        LTE_CHECK_CUDA(cudaGetLastError());
        return LTE_OK;
    }
    if (!phases || !faded) return LTE_ERR_INVALID_ARG;

    TdlParams C;
    memset(&C, 0, sizeof(C));
    C.num_taps = ch->num_taps;
    int dmax = 0;
    for (int i = 0; i < ch->num_taps; ++i) {
        if (ch->delay[i] < 0) return LTE_ERR_INVALID_ARG;
        C.delay[i] = ch->delay[i];
        C.gain[i] = (float)((double)ch->gain[i] * sqrt(2.0 / LTE_JAKES_TONES));
        if (ch->delay[i] > dmax) dmax = ch->delay[i];
    }
    if (dmax > 4096) return LTE_ERR_UNSUPPORTED;
    C.halo = (dmax + 7) & ~7;
    const double fs = p->desc.fs;
    double wmax = 0.0;
    for (int nn = 0; nn < LTE_JAKES_TONES; ++nn) {   // alpha_n = 2 pi n / 16, n = 1..16
        C.w_cyc[nn] = ch->doppler_hz * cos(2.0 * M_PI * (double)(nn + 1) / LTE_JAKES_TONES) / fs;
        if (fabs(C.w_cyc[nn]) > wmax) wmax = fabs(C.w_cyc[nn]);
    }
    const int V = (R <= 4) ? 8 : 4;
    const int tile = TDL_THREADS * V;
    // Taylor remainder x^(K+1)/(K+1)! with x = 2 pi w PB/2 kept below 2e-8:
    //   K = 2 needs x < 4.9e-3, K = 4 needs x < 0.075; PB is halved until K = 4 fits.
    C.pb = tile;
    int K = 2;
    if (M_PI * wmax * C.pb > 4.9e-3) {
        K = 4;
        while (C.pb > 32 && M_PI * wmax * C.pb > 0.075) C.pb >>= 1;
        if (M_PI * wmax * C.pb > 0.075) return LTE_ERR_UNSUPPORTED;   // Doppler too high for this fs
    }
    const int nblk = tile / C.pb;
    const int ncols = (C.halo + tile) / 8;
    C.xs_stride = ncols + ((2 - ncols % 16) + 16) % 16;
    const size_t smem = sizeof(float2) * ((size_t)T * 8 * C.xs_stride +
                                          (size_t)nblk * R * T * C.num_taps * (2 * K + 1));
    if (smem > 200 * 1024) return LTE_ERR_UNSUPPORTED;
    const int tiles = (int)((n + tile - 1) / tile);
    const unsigned grid = (unsigned)((long long)tiles * B);
#define LAUNCH_TDL_K(RR, VV, KK)                                                                          \
    {                                                                                                     \
        auto k = tdl_kernel<RR, VV, KK>;                                                                  \
        LTE_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
        k<<<grid, TDL_THREADS, smem, st>>>(C, (const float2*)tx, phases, (float2*)faded, power, T,        \
                                           (long long)n, tiles);                                          \
    }
#define LAUNCH_TDL(RR, VV)                                  \
    case RR:                                                \
        if (K == 2) LAUNCH_TDL_K(RR, VV, 2) else LAUNCH_TDL_K(RR, VV, 4) \
        break;
    switch (R) {
        LAUNCH_TDL(1, 8) LAUNCH_TDL(2, 8) LAUNCH_TDL(3, 8) LAUNCH_TDL(4, 8)
        LAUNCH_TDL(5, 4) LAUNCH_TDL(6, 4) LAUNCH_TDL(7, 4) LAUNCH_TDL(8, 4)
    }
#undef LAUNCH_TDL
#undef LAUNCH_TDL_K
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

// ------------------------------------------------------------------------------ AWGN
__global__ void awgn_kernel(const float2* __restrict__ x, int x_div, const double* __restrict__ power,
                            const float* __restrict__ snr_lin, const float2* __restrict__ z, uint32_t key,
                            unsigned long long row_id0, float2* __restrict__ y, long long n, int gx) {
    const long long row = blockIdx.x / gx;
    const int bx = blockIdx.x % gx;
    const float sigma = lte_sigma(power[row], (float)n, snr_lin[row]);
    const float2* src = x + (size_t)(row / x_div) * n;
    float2* dst = y + (size_t)row * n;
    const uint32_t rid = (uint32_t)(row_id0 + (unsigned long long)row);
    for (long long i = (long long)bx * blockDim.x + threadIdx.x; i < n; i += (long long)gx * blockDim.x) {
        const float2 w = z ? z[(size_t)row * n + i] : lte_noise_sample(key, rid, (uint32_t)i);
        const float2 v = src[i];
        dst[i] = make_float2(fmaf(sigma, w.x, v.x), fmaf(sigma, w.y, v.y));
    }
}

extern "C" int lte_awgn_add(const lte_plan* p, const lte_c32* x, int32_t x_div, const double* power,
                            const float* snr_lin, const lte_c32* z, uint64_t seed, uint64_t row_id0,
                            lte_c32* y, int64_t rows, int64_t n, void* stream) {
    if (!p || !x || !power || !snr_lin || !y || rows < 0 || n < 1 || x_div < 1) return LTE_ERR_INVALID_ARG;
    if (rows == 0) return LTE_OK;
    int gx = (int)((n + 255) / 256);
    if (gx > 64) gx = 64;
    awgn_kernel<<<(unsigned)((long long)gx * rows), 256, 0, (cudaStream_t)stream>>>(
        (const float2*)x, x_div, power, snr_lin, (const float2*)z, lte_key(seed, LTE_DOMAIN_NOISE), row_id0,
        (float2*)y, n, gx);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

// ------------------------------------------------------------------------------ engine RNG helpers
__global__ void random_indices_kernel(uint8_t* __restrict__ idx, long long nsym, uint32_t key,
                                      unsigned long long stream_id0, uint32_t mask, int gx) {
    const long long b = blockIdx.x / gx;
    const int bx = blockIdx.x % gx;
    uint8_t* dst = idx + (size_t)b * nsym;
    const uint32_t sid = (uint32_t)(stream_id0 + (unsigned long long)b);
    const long long nq = (nsym + 7) / 8;
    for (long long q = (long long)bx * blockDim.x + threadIdx.x; q < nq; q += (long long)gx * blockDim.x) {
        uint32_t r0, r1;
        philox2x32_10(key, (uint32_t)q, sid, r0, r1);
        r0 &= mask * 0x01010101u;
        r1 &= mask * 0x01010101u;
        const long long o = q * 8;
        if (o + 8 <= nsym && ((((size_t)b * nsym) & 7) == 0)) {
            *(uint2*)(dst + o) = make_uint2(r0, r1);
        } else {
            for (int i = 0; i < 8 && o + i < nsym; ++i) dst[o + i] = (uint8_t)((i < 4 ? r0 >> (8 * i) : r1 >> (8 * (i - 4))) & 0xff);
        }
    }
}

extern "C" int lte_random_indices(const lte_plan* p, uint8_t* idx, int64_t nsym, int64_t B, uint64_t seed,
                                  uint64_t stream_id0, void* stream) {
    if (!p || !idx || nsym < 1 || B < 0) return LTE_ERR_INVALID_ARG;
    if (B == 0) return LTE_OK;
    int gx = (int)(((nsym + 7) / 8 + 255) / 256);
    if (gx > 16) gx = 16;
    random_indices_kernel<<<(unsigned)((long long)gx * B), 256, 0, (cudaStream_t)stream>>>(
        idx, nsym, lte_key(seed, LTE_DOMAIN_BITS), stream_id0, (1u << p->dev.bps) - 1u, gx);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

__global__ void random_phases_kernel(float* __restrict__ ph, long long per_stream, uint32_t key,
                                     unsigned long long stream_id0, int gx) {
    const long long b = blockIdx.x / gx;
    const int bx = blockIdx.x % gx;
    const uint32_t sid = (uint32_t)(stream_id0 + (unsigned long long)b);
    for (long long q = (long long)bx * blockDim.x + threadIdx.x; q < (per_stream + 1) / 2;
         q += (long long)gx * blockDim.x) {
        uint32_t r0, r1;
        philox2x32_10(key, (uint32_t)q, sid, r0, r1);
        float* dst = ph + (size_t)b * per_stream + 2 * q;
        dst[0] = (float)(r0 >> 8) * (1.0f / 16777216.0f);
        if (2 * q + 1 < per_stream) dst[1] = (float)(r1 >> 8) * (1.0f / 16777216.0f);
    }
}

extern "C" int lte_random_phases(float* phases, int64_t per_stream, int64_t B, uint64_t seed,
                                 uint64_t stream_id0, void* stream) {
    if (!phases || per_stream < 1 || B < 0) return LTE_ERR_INVALID_ARG;
    if (B == 0) return LTE_OK;
    const int gx = (int)(((per_stream + 1) / 2 + 127) / 128);
    random_phases_kernel<<<(unsigned)((long long)gx * B), 128, 0, (cudaStream_t)stream>>>(
        phases, per_stream, lte_key(seed, LTE_DOMAIN_PHASE), stream_id0, gx);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}
