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

#define TDL_THREADS 128
#define TDL_STAGES 2
#define TDL_MAX_PB 8192

struct TdlParams {
    int num_taps;
    int delay[LTE_MAX_TAPS];
    float gain[LTE_MAX_TAPS];           // includes sqrt(2/16)
    double w_cyc[LTE_JAKES_TONES];      // fD cos(alpha_n) / fs   [cycles per sample]
    int pb;                             // polynomial block length in samples (power of two)
    int nbs;                            // polynomial blocks per stream = ceil(n / pb)
};

// Taylor coefficients of every (rx, tx, tap) fading process for every polynomial block:
// coef[b][blk][triple][2K+1] = { sum_n g e^{j theta_n} (j x_n)^k / k!  (k = 0..K),
//                                k-scaled copies for the derivative (k = 1..K) }
// with theta_n = 2 pi (w_n m_c + u_n) reduced in fp64 at the block centre m_c.
// work item = (b, blk, triple, tone); 16 consecutive lanes reduce one sum.
template <int K>
__global__ void __launch_bounds__(256)
jakes_coef_kernel(const TdlParams C, const float* __restrict__ phases, float2* __restrict__ coef, int nlt,
                  long long total_items) {
    constexpr int NC = 2 * K + 1;
    const long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    float2 a[K + 1];
#pragma unroll
    for (int k = 0; k <= K; ++k) a[k] = make_float2(0.f, 0.f);
    long long grp = 0;
    int trip = 0;
    if (it < total_items) {
        const int tone = (int)(it & (LTE_JAKES_TONES - 1));
        grp = it >> 4;                                  // (b*nbs + blk)*nlt + trip
        trip = (int)(grp % nlt);
        const long long q = grp / nlt;
        const int blk = (int)(q % C.nbs);
        const long long b = q / C.nbs;
        const double mc = (double)blk * C.pb + 0.5 * (C.pb - 1);
        const float u = phases[((size_t)b * nlt + trip) * LTE_JAKES_TONES + tone];
        double turns = C.w_cyc[tone] * mc + (double)u;
        turns -= floor(turns);
        float sn, cs;
        sincospif(2.0f * (float)turns, &sn, &cs);
        const float x = (float)(6.283185307179586 * C.w_cyc[tone]);   // rad / sample
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
    if (it < total_items && (it & (LTE_JAKES_TONES - 1)) == 0) {
        const float g = C.gain[trip % C.num_taps];
        float2* c = coef + (size_t)grp * NC;
#pragma unroll
        for (int k = 0; k <= K; ++k) c[k] = cscale(a[k], g);
#pragma unroll
        for (int k = 1; k <= K; ++k) c[K + k] = cscale(a[k], g * (float)k);
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
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__host__ __device__ constexpr int tdl_xs_stride(int halo, int tile) {
    return (halo + tile) / 8 + ((2 - ((halo + tile) / 8) % 16) + 16) % 16;   // == 2 (mod 16)
}

// 8 consecutive staged samples starting at (column base, phase PH): row (PH+i)&7, column +(PH+i)>>3
template <int PH, int XS> __device__ __forceinline__ void tdl_load8(float2 (&xv)[8], const float2* base) {
#pragma unroll
    for (int i = 0; i < 8; ++i) xv[i] = base[((PH + i) & 7) * XS + ((PH + i) >> 3)];
}

// Persistent tapped-delay-line kernel: R receive antennas, V consecutive samples per thread,
// degree-K polynomial, compile-time delay halo.  Each CTA owns a contiguous chunk of tiles
// (TDL_THREADS*V samples each) and walks it with a two-stage cp.async pipeline: the next tile's
// samples and coefficients land while this tile is computed.
// Shared-memory sample layout: sample s of the staged window lives at [s & 7][s >> 3], so a
// warp reading "sample 8*t + c" for consecutive threads t touches consecutive float2 (no bank
// conflicts for any tap delay) while the global reads that fill it stay coalesced.
// antenna groups per CTA: at most 4 antennas (64 accumulator registers) per thread.  Splitting
// R = 4 into two groups was measured slower (1.68 vs 1.55 ms per 4096 subframes): the extra
// sample loads cost more than the added occupancy gains.
__host__ __device__ constexpr int tdl_groups(int R) {
    return R <= 4 ? 1 : (R == 5 ? 5 : (R == 6 ? 2 : (R == 7 ? 7 : 2)));
}
__host__ __device__ constexpr int tdl_min_blocks(int R) {
    return tdl_groups(R) == 1 ? 3 : (tdl_groups(R) == 2 ? 2 : 1);
}

template <int R, int RG, int V, int K, int HALO>
__global__ void __launch_bounds__(TDL_THREADS * (R / RG), tdl_min_blocks(R))
tdl_kernel(const TdlParams C, const float2* __restrict__ tx, const float2* __restrict__ coef_g,
           float2* __restrict__ faded, double* __restrict__ power, int T, int n, int tiles, int total_tiles,
           int chunk) {
    // RG antennas per thread; the R/RG thread groups of TDL_THREADS share the staged samples
    constexpr int NG = R / RG, NT = TDL_THREADS * NG;
    constexpr int TILE = TDL_THREADS * V;
    constexpr int NC = 2 * K + 1;
    constexpr int XS = tdl_xs_stride(HALO, TILE);
    constexpr int SPAN = HALO + TILE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nblk = C.pb >= TILE ? 1 : TILE / C.pb;         // polynomial blocks per tile
    const int nlt = R * T * C.num_taps;
    const int ncoef = nblk * nlt * NC;                       // float2 per tile
    const int xs_elems = T * 8 * XS;
    const int stage_elems = xs_elems + ncoef;
    float2* stage_base = (float2*)smem_raw;                  // [STAGES][xs | coef]
    const int tid_all = threadIdx.x;
    const int tid = tid_all % TDL_THREADS;                   // sample-group index
    const int grp = tid_all / TDL_THREADS;                   // antenna group (warp-uniform)

    auto prefetch = [&](int tile_id, int stage) {
        float2* sx = stage_base + stage * stage_elems;
        float2* sc = sx + xs_elems;
        const int b = tile_id / tiles;
        const int tile0 = (tile_id - b * tiles) * TILE;
        const bool interior = (tile0 >= HALO) && (tile0 + TILE <= n);
        for (int t = 0; t < T; ++t) {
            const float2* src = tx + ((size_t)b * T + t) * n + (tile0 - HALO);
            float2* dst = sx + t * 8 * XS;
            if (interior) {
#pragma unroll
                for (int i0 = 0; i0 < SPAN; i0 += NT) {
                    const int i = i0 + tid_all;
                    if (i0 + NT <= SPAN || i < SPAN) cp_async8(&dst[(i & 7) * XS + (i >> 3)], &src[i]);
                }
            } else {
                for (int i = tid_all; i < SPAN; i += NT) {
                    const int m = tile0 - HALO + i;
                    const bool ok = (m >= 0 && m < n);
                    cp_async8_zfill(&dst[(i & 7) * XS + (i >> 3)], ok ? &src[i] : tx, ok);
                }
            }
        }
        const float2* cg = coef_g + ((size_t)b * C.nbs + tile0 / C.pb) * nlt * NC;
        for (int i = tid_all; i < ncoef; i += NT) cp_async8(&sc[i], &cg[i]);
    };

    int tile_id = blockIdx.x * chunk;
    const int tile_end = min(tile_id + chunk, total_tiles);
    int stage = 0;
    if (tile_id < tile_end) prefetch(tile_id, 0);
    cp_async_commit();

    __shared__ float pw_red[TDL_THREADS / 32][LTE_MAX_RX];
    const int l0 = tid * V;                         // first local sample of this thread
    const bool vec = ((n & 1) == 0);
    float pw[RG];
#pragma unroll
    for (int r = 0; r < RG; ++r) pw[r] = 0.f;

    auto flush_power = [&](int b) {                 // block reduction of the per-thread power sums
#pragma unroll
        for (int r = 0; r < RG; ++r) {
            const float v = warp_sum(pw[r]);
            if ((tid & 31) == 0) pw_red[tid >> 5][grp * RG + r] = v;
            pw[r] = 0.f;
        }
        __syncthreads();
        if (tid_all < R) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < TDL_THREADS / 32; ++w) s += pw_red[w][tid_all];
            atomicAdd(&power[(size_t)b * R + tid_all], (double)s);
        }
        __syncthreads();
    };

    for (; tile_id < tile_end; ++tile_id, stage ^= 1) {
        if (tile_id + 1 < tile_end) prefetch(tile_id + 1, stage ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();

        const int b = tile_id / tiles;
        const int tile0 = (tile_id - b * tiles) * TILE;
        const int m0 = tile0 + l0;
        const int blk = m0 / C.pb;                                      // block index within the stream
        const float2* sx = stage_base + stage * stage_elems;
        const float2* coef = sx + xs_elems + (blk - tile0 / C.pb) * nlt * NC;
        // polynomial argument at the centre of the thread's V samples; linear stepping inside
        const float tau = (float)(m0 - blk * C.pb) + 0.5f * (V - 1) - 0.5f * (C.pb - 1);

        float2 acc[RG][V];
#pragma unroll
        for (int r = 0; r < RG; ++r)
#pragma unroll
            for (int i = 0; i < V; ++i) acc[r][i] = make_float2(0.f, 0.f);

        for (int t = 0; t < T; ++t) {
            const float2* sxt = sx + t * 8 * XS;
            for (int tap = 0; tap < C.num_taps; ++tap) {
                float2 xv[V];
                if constexpr (V == 8) {
                    const int c0 = HALO - C.delay[tap];
                    const float2* base = sxt + tid + (c0 >> 3);
                    switch (c0 & 7) {
                        case 0: tdl_load8<0, XS>(xv, base); break;
                        case 1: tdl_load8<1, XS>(xv, base); break;
                        case 2: tdl_load8<2, XS>(xv, base); break;
                        case 3: tdl_load8<3, XS>(xv, base); break;
                        case 4: tdl_load8<4, XS>(xv, base); break;
                        case 5: tdl_load8<5, XS>(xv, base); break;
                        case 6: tdl_load8<6, XS>(xv, base); break;
                        default: tdl_load8<7, XS>(xv, base); break;
                    }
                } else {
                    const int c0 = HALO - C.delay[tap] + (l0 & 7);
                    const float2* xcol = sxt + (l0 >> 3);
#pragma unroll
                    for (int i = 0; i < V; ++i) {
                        const int cc = c0 + i;
                        xv[i] = xcol[(cc & 7) * XS + (cc >> 3)];
                    }
                }
#pragma unroll
                for (int r = 0; r < RG; ++r) {
                    const float2* c = coef + (((grp * RG + r) * T + t) * C.num_taps + tap) * NC;
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

        // ---- store + power ----------------------------------------------------------------
#pragma unroll
        for (int r = 0; r < RG; ++r) {
            float2* dst = faded + ((size_t)b * R + grp * RG + r) * n + m0;
            if (m0 + V <= n && vec) {
                float4* d4 = (float4*)dst;
#pragma unroll
                for (int i = 0; i < V; i += 2) {
                    d4[i / 2] = make_float4(acc[r][i].x, acc[r][i].y, acc[r][i + 1].x, acc[r][i + 1].y);
                    pw[r] += cabs2(acc[r][i]) + cabs2(acc[r][i + 1]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < V; ++i)
                    if (m0 + i < n) { dst[i] = acc[r][i]; pw[r] += cabs2(acc[r][i]); }
            }
        }
        const bool last_of_stream = (tile_id + 1 == tile_end) || ((tile_id + 1) / tiles != b);
        if (last_of_stream) flush_power(b);         // uniform branch; includes the stage barrier
        else __syncthreads();                       // every thread is done reading this stage
    }
    cp_async_wait<0>();
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
    if (dmax > 144) return LTE_ERR_UNSUPPORTED;     // longest ITU-R M.1225 delay at 30.72 MHz is 139 samples
    const int halo = dmax <= 16 ? 16 : 144;
    if (n > (1LL << 30) || (long long)B * n > (1LL << 40)) return LTE_ERR_UNSUPPORTED;
    const double fs = p->desc.fs;
    double wmax = 0.0;
    for (int nn = 0; nn < LTE_JAKES_TONES; ++nn) {   // alpha_n = 2 pi n / 16, n = 1..16
        C.w_cyc[nn] = ch->doppler_hz * cos(2.0 * M_PI * (double)(nn + 1) / LTE_JAKES_TONES) / fs;
        if (fabs(C.w_cyc[nn]) > wmax) wmax = fabs(C.w_cyc[nn]);
    }
    const int V = 8;
    const int tile = TDL_THREADS * V;
    // Taylor remainder x^(K+1)/(K+1)! with x = 2 pi w PB/2 kept below 2e-8:
    //   K = 2 needs x < 4.9e-3, K = 4 needs x < 0.075.  PB is the largest power of two that fits.
    int K = 2;
    C.pb = TDL_MAX_PB;
    while (C.pb > tile && M_PI * wmax * C.pb > 4.9e-3) C.pb >>= 1;
    if (M_PI * wmax * C.pb > 4.9e-3) {
        K = 4;
        C.pb = TDL_MAX_PB;
        while (C.pb > 32 && M_PI * wmax * C.pb > 0.075) C.pb >>= 1;
        if (M_PI * wmax * C.pb > 0.075) return LTE_ERR_UNSUPPORTED;   // Doppler too high for this fs
    }
    C.nbs = (int)((n + C.pb - 1) / C.pb);
    const int nblk = C.pb >= tile ? 1 : tile / C.pb;
    const int nlt = R * T * C.num_taps;
    const int NC = 2 * K + 1;
    const size_t ncoef = (size_t)nblk * nlt * NC;
    const size_t smem = sizeof(float2) * TDL_STAGES * ((size_t)T * 8 * tdl_xs_stride(halo, tile) + ncoef);
    if (smem > 200 * 1024) return LTE_ERR_UNSUPPORTED;
    const int tiles = (int)((n + tile - 1) / tile);
    const long long total_tiles_ll = (long long)tiles * B;
    if (total_tiles_ll > (1LL << 30)) return LTE_ERR_UNSUPPORTED;
    const int total_tiles = (int)total_tiles_ll;

    // per-block polynomial coefficients (scratch owned by the plan, grown on demand)
    const size_t coef_bytes = sizeof(float2) * (size_t)B * C.nbs * nlt * NC;
    lte_plan* pm = const_cast<lte_plan*>(p);
    if (pm->scratch_bytes < coef_bytes) {
        if (pm->scratch) LTE_CHECK_CUDA(cudaFree(pm->scratch));
        pm->scratch = nullptr;
        pm->scratch_bytes = 0;
        LTE_CHECK_CUDA(cudaMalloc(&pm->scratch, coef_bytes));
        pm->scratch_bytes = coef_bytes;
    }
    float2* coef = (float2*)pm->scratch;
    const long long items = (long long)B * C.nbs * nlt * LTE_JAKES_TONES;
    const unsigned cgrid = (unsigned)((items + 255) / 256);
    if (K == 2) jakes_coef_kernel<2><<<cgrid, 256, 0, st>>>(C, phases, coef, nlt, items);
    else jakes_coef_kernel<4><<<cgrid, 256, 0, st>>>(C, phases, coef, nlt, items);
    LTE_CHECK_CUDA(cudaGetLastError());

    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int ctas = sms * tdl_min_blocks(R);
    if (ctas > total_tiles) ctas = total_tiles;
    const int chunk = (total_tiles + ctas - 1) / ctas;
    const unsigned grid = (unsigned)((total_tiles + chunk - 1) / chunk);
#define LAUNCH_TDL_KH(RR, VV, KK, HH)                                                                     \
    {                                                                                                     \
        constexpr int RG = RR / tdl_groups(RR);                                                           \
        auto k = tdl_kernel<RR, RG, VV, KK, HH>;                                                          \
        LTE_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
        k<<<grid, TDL_THREADS * (RR / RG), smem, st>>>(C, (const float2*)tx, coef, (float2*)faded, power, \
                                                       T, (int)n, tiles, total_tiles, chunk);             \
    }
#define LAUNCH_TDL_K(RR, VV, KK) \
    if (halo == 16) LAUNCH_TDL_KH(RR, VV, KK, 16) else LAUNCH_TDL_KH(RR, VV, KK, 144)
#define LAUNCH_TDL(RR, VV)                                  \
    case RR:                                                \
        if (K == 2) { LAUNCH_TDL_K(RR, VV, 2) } else { LAUNCH_TDL_K(RR, VV, 4) } \
        break;
    switch (R) {
        LAUNCH_TDL(1, 8) LAUNCH_TDL(2, 8) LAUNCH_TDL(3, 8) LAUNCH_TDL(4, 8)
        LAUNCH_TDL(5, 8) LAUNCH_TDL(6, 8) LAUNCH_TDL(7, 8) LAUNCH_TDL(8, 8)
    }
#undef LAUNCH_TDL
#undef LAUNCH_TDL_K
#undef LAUNCH_TDL_KH
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
