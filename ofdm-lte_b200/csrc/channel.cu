// Stage 3: ITU-R M.1225 tapped-delay-line Rayleigh channel in the time domain
// (core/rayleighchannel.py:20-58) plus the AWGN of core/channel.py:46-66 / :216-232,
// and the Philox helpers of the sweep engine.
//
// Jakes fading  h_i[m] = sqrt(2/16) * sum_n exp(j(2 pi fD cos(a_n) m / fs + phi_n))
// is evaluated per polynomial block of PB samples: the 16-tone sum and its first K
// derivatives are formed once at the block centre (phase reduced in fp64), then every
// thread evaluates the degree-K Taylor polynomial (and its derivative) at the centre of its
// own 8 samples and steps linearly inside them.  PB is chosen on the host so that the block
// truncation error stays below 2e-8; the linear stepping adds (2 pi fD/fs 3.5)^2 / 2, see tdl_setup.
#include "tdl.cuh"

// Sample tile of the 8-samples-per-thread layout: sample s lives at [s & 7][s >> 3]; one element is
// float4 (re, re, im, im) so a 128-bit load feeds both lanes of the packed multiply-adds.
template <int PH, int XS> __device__ __forceinline__ float4 tdl_x(const float4* base, int i) {
    return base[((PH + i) & 7) * XS + ((PH + i) >> 3)];
}

// Persistent tapped-delay-line kernel.  Each thread produces 8 consecutive samples for RG = 2*RP
// receive antennas; the two antennas of a pair ride in the two lanes of FFMA2 instructions:
//   h(m0 + i) = h + (i - 3.5) dh           (packed over the antenna pair, degree-K polynomial at the
//   acc_re += h_re x_re - h_im x_im         centre of the thread's 8 samples, linear inside)
//   acc_im += h_re x_im + h_im x_re
// i.e. 7 FFMA2 per (antenna pair, tap, sample) instead of 12 FFMA.  Each CTA owns a contiguous chunk
// of tiles (TDL_THREADS*8 samples) and prefetches the next tile's raw samples and coefficients with
// cp.async while the current one is computed.
template <int R, int RG, int K, int HALO>
__global__ void __launch_bounds__(TDL_THREADS * ((R + RG - 1) / RG), tdl_min_blocks(R))
tdl_kernel(const TdlParams C, const float2* __restrict__ tx, const float* __restrict__ coef_g,
           float2* __restrict__ faded, double* __restrict__ power, int T, int n, int tiles, int total_tiles,
           int chunk) {
    constexpr int V = 8, RP = RG / 2;
    constexpr int NG = (R + RG - 1) / RG, NT = TDL_THREADS * NG, R2 = NG * RG;
    constexpr int TILE = TDL_THREADS * V;
    constexpr int NC = 2 * K + 1;
    constexpr int XS = tdl_xs_stride(HALO, TILE), OS = tdl_os_stride(TILE);
    constexpr int SPAN = HALO + TILE;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nblk = C.pb >= TILE ? 1 : TILE / C.pb;         // polynomial blocks per tile
    const int ntt = T * C.num_taps;
    const int ncoef = nblk * ntt * NC * 2 * R2;              // floats per tile
    // shared memory: outb [NG*RP][8][OS] float4 | plane [2][T][8][XS] float2 | coef [2][ncoef] float
    // plane: sample s of the staged window at [s & 7][s >> 3] (filled directly by cp.async);
    // outb: per antenna pair, output sample (8 t + i) at [i][t] for the coalesced stores.
    float4* outb = (float4*)smem_raw;
    float2* plane = (float2*)(outb + (size_t)NG * RP * 8 * OS);
    float* scoef = (float*)(plane + (size_t)2 * T * 8 * XS);
    const int tid_all = threadIdx.x;
    const int tid = tid_all % TDL_THREADS;                   // sample-group index
    const int grp = tid_all / TDL_THREADS;                   // antenna group (warp-uniform)

    auto prefetch = [&](int b, int tin, int stage) {
        float2* sr = plane + (size_t)stage * T * 8 * XS;
        float* sc = scoef + (size_t)stage * ncoef;
        const int tile0 = tin * TILE;
        const bool interior = (tile0 >= HALO) && (tile0 + TILE <= n);
        for (int t = 0; t < T; ++t) {
            const float2* src = tx + ((size_t)b * T + t) * n + (tile0 - HALO);
            float2* dst = sr + t * 8 * XS;
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
        const float* cg = coef_g + ((size_t)b * C.nbs + (tile0 >> C.pb_log2)) * ntt * NC * 2 * R2;
        for (int i = tid_all; i < ncoef; i += NT) cp_async4(&sc[i], &cg[i]);
    };

    int tile_id = blockIdx.x * chunk;
    const int tile_end = min(tile_id + chunk, total_tiles);
    int b = tile_id / tiles, tin = tile_id - b * tiles;      // stream and tile-in-stream, advanced incrementally
    int stage = 0;
    if (tile_id < tile_end) prefetch(b, tin, 0);
    cp_async_commit();

    __shared__ float pw_red[TDL_THREADS / 32][LTE_MAX_RX];
    __shared__ __align__(16) int xoff[LTE_MAX_TAPS * 8];
    if (tid_all < C.num_taps * 8) {                 // sample 8*t + c0 + i lives at [(c0+i) & 7][t + (c0+i) >> 3]
        const int cc = HALO - C.delay[tid_all >> 3] + (tid_all & 7);
        xoff[tid_all] = (cc & 7) * XS + (cc >> 3);
    }
    const int l0 = tid * V;                         // first local sample of this thread
    const bool vec = ((n & 1) == 0);
    pf2 pw[RP];
#pragma unroll
    for (int p = 0; p < RP; ++p) pw[p] = ppk(0.f, 0.f);

    auto flush_power = [&](int b) {                 // block reduction of the per-thread power sums
#pragma unroll
        for (int p = 0; p < RP; ++p) {
            float a, c;
            pupk(pw[p], a, c);
            a = warp_sum(a);
            c = warp_sum(c);
            if ((tid & 31) == 0) { pw_red[tid >> 5][grp * RG + 2 * p] = a; pw_red[tid >> 5][grp * RG + 2 * p + 1] = c; }
            pw[p] = ppk(0.f, 0.f);
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
        cp_async_wait<0>();
        __syncthreads();                            // plane[stage] / coef[stage] landed; outb and the other stage are free
        const bool last_tile_of_stream = (tin + 1 == tiles);
        const int nb = last_tile_of_stream ? b + 1 : b, ntin = last_tile_of_stream ? 0 : tin + 1;
        if (tile_id + 1 < tile_end) prefetch(nb, ntin, stage ^ 1);
        cp_async_commit();
        __syncthreads();

        const int tile0 = tin * TILE;
        const int m0 = tile0 + l0;
        const int blk = m0 >> C.pb_log2;                                // block index within the stream
        const float* sc = scoef + stage * ncoef + (blk - (tile0 >> C.pb_log2)) * ntt * NC * 2 * R2 + grp * RG;
        // polynomial argument at the centre of the thread's V samples; linear stepping inside
        const float tauf = (float)(m0 - (blk << C.pb_log2)) + 0.5f * (V - 1) - 0.5f * (C.pb - 1);
        const pf2 tau = ppk(tauf, tauf);

        pf2 are[RP][V], aim[RP][V];
#pragma unroll
        for (int p = 0; p < RP; ++p)
#pragma unroll
            for (int i = 0; i < V; ++i) { are[p][i] = ppk(0.f, 0.f); aim[p][i] = ppk(0.f, 0.f); }

        for (int t = 0; t < T; ++t) {
            const float2* plt = plane + ((size_t)stage * T + t) * 8 * XS;
            for (int tap = 0; tap < C.num_taps; ++tap) {
                // per-tap offsets of the thread's 8 delayed samples in the transposed plane (uniform)
                const int4 o0 = *(const int4*)&xoff[tap * 8], o1 = *(const int4*)&xoff[tap * 8 + 4];
                const float2* base = plt + tid;
                float2 xv[V];
                xv[0] = base[o0.x]; xv[1] = base[o0.y]; xv[2] = base[o0.z]; xv[3] = base[o0.w];
                xv[4] = base[o1.x]; xv[5] = base[o1.y]; xv[6] = base[o1.z]; xv[7] = base[o1.w];
                const float* ct = sc + (size_t)(t * C.num_taps + tap) * NC * 2 * R2;
#pragma unroll
                for (int p = 0; p < RP; ++p) {
                    // packed Horner for value and derivative of (h_re, h_im) of the antenna pair
                    auto cf = [&](int slot, int ri) { const float2 c = *(const float2*)(ct + (slot * 2 + ri) * R2 + 2 * p); return ppk(c.x, c.y); };
                    pf2 hre = cf(K, 0), him = cf(K, 1), dre = cf(2 * K, 0), dim = cf(2 * K, 1);
#pragma unroll
                    for (int k = K - 1; k >= 0; --k) { hre = pfma(hre, tau, cf(k, 0)); him = pfma(him, tau, cf(k, 1)); }
#pragma unroll
                    for (int k = K - 1; k >= 1; --k) { dre = pfma(dre, tau, cf(K + k, 0)); dim = pfma(dim, tau, cf(K + k, 1)); }
                    const pf2 zero = ppk(0.f, 0.f);
                    const pf2 nhim = psub(zero, him), ndim = psub(zero, dim);
#pragma unroll
                    for (int i = 0; i < V; ++i) {
                        const float stf = (float)i - 0.5f * (V - 1);
                        const pf2 st = ppk(stf, stf);
                        // scalar sample broadcast to both antennas (FFMA2 takes a scalar .F32 operand)
                        const pf2 xre = ppk(xv[i].x, xv[i].x), xim = ppk(xv[i].y, xv[i].y);
                        const pf2 hr = pfma(st, dre, hre), hi = pfma(st, dim, him), nhi = pfma(st, ndim, nhim);
                        pfma_acc(are[p][i], hr, xre);
                        pfma_acc(are[p][i], nhi, xim);
                        pfma_acc(aim[p][i], hr, xim);
                        pfma_acc(aim[p][i], hi, xre);
                    }
                }
            }
        }

        // ---- power, then coalesced stores through shared memory -------------------------------
        // A thread owns 8 consecutive samples (64 B per antenna), so direct stores would hit every
        // 32-byte sector twice with half-sector writes; staging lets each warp instruction write
        // 512 contiguous bytes instead (measured: stores cost 0.55 ms of 1.57 ms before this).
#pragma unroll
        for (int p = 0; p < RP; ++p) {
            if (m0 + V <= n) {
#pragma unroll
                for (int i = 0; i < V; ++i) { pfma_acc(pw[p], are[p][i], are[p][i]); pfma_acc(pw[p], aim[p][i], aim[p][i]); }
            } else {                                // samples past the end of the stream do not count
#pragma unroll
                for (int i = 0; i < V; ++i)
                    if (m0 + i < n) { pfma_acc(pw[p], are[p][i], are[p][i]); pfma_acc(pw[p], aim[p][i], aim[p][i]); }
            }
        }
#pragma unroll
        for (int p = 0; p < RP; ++p) {
            float4* ob = outb + (size_t)(grp * RP + p) * 8 * OS;    // (re_r0, re_r1, im_r0, im_r1) per sample
#pragma unroll
            for (int i = 0; i < V; ++i) {
                float a, c, d, e;
                pupk(are[p][i], a, c);
                pupk(aim[p][i], d, e);
                ob[i * OS + tid] = make_float4(a, c, d, e);
            }
        }
        __syncthreads();
        {
#pragma unroll
            for (int p = 0; p < RP; ++p) {
                const float4* ob = outb + (size_t)(grp * RP + p) * 8 * OS;
                const int r0 = grp * RG + 2 * p;
                float2* d0 = faded + ((size_t)b * R + r0) * n + tile0;
                float2* d1 = d0 + n;
#pragma unroll
                for (int k = 0; k < V / 2; ++k) {
                    const int s = 2 * (k * TDL_THREADS + tid);       // this lane's sample pair within the tile
                    const float4 u = ob[(s & 7) * OS + (s >> 3)], w = ob[((s + 1) & 7) * OS + ((s + 1) >> 3)];
                    if (tile0 + s + 2 <= n && vec) {
                        *(float4*)(d0 + s) = make_float4(u.x, u.z, w.x, w.z);
                        if (r0 + 1 < R) *(float4*)(d1 + s) = make_float4(u.y, u.w, w.y, w.w);
                    } else {
                        if (tile0 + s < n) { d0[s] = make_float2(u.x, u.z); if (r0 + 1 < R) d1[s] = make_float2(u.y, u.w); }
                        if (tile0 + s + 1 < n) { d0[s + 1] = make_float2(w.x, w.z); if (r0 + 1 < R) d1[s + 1] = make_float2(w.y, w.w); }
                    }
                }
            }
        }
        if (last_tile_of_stream || tile_id + 1 == tile_end) flush_power(b);   // uniform branch
        b = nb;
        tin = ntin;
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

// Host-side choice of the polynomial (degree K, block length pb) and the launch geometry of lte_channel_tdl;
// shared by the workspace-size query and the launcher.
struct TdlSetup {
    TdlParams C;
    int K, halo, R2, tiles, total_tiles;
    size_t smem, coef_bytes;
};

static int tdl_setup(const lte_plan* p, const lte_channel_desc* ch, int32_t B, int32_t R, int32_t T, int64_t n,
                     TdlSetup& U) {
    if (!p || !ch || B < 0 || R < 1 || R > LTE_MAX_RX || T < 1 || T > LTE_MAX_TX || n < 1) return LTE_ERR_INVALID_ARG;
    if (ch->num_taps < 1 || ch->num_taps > LTE_MAX_TAPS) return LTE_ERR_INVALID_ARG;
    TdlParams& C = U.C;
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
    U.halo = dmax <= 16 ? 16 : 144;
    if (n > (1LL << 30) || (long long)B * n > (1LL << 40)) return LTE_ERR_UNSUPPORTED;
    const double fs = p->desc.fs;
    double wmax = 0.0;
    for (int nn = 0; nn < LTE_JAKES_TONES; ++nn) {   // alpha_n = 2 pi n / 16, n = 1..16
        C.w_cyc[nn] = ch->doppler_hz * cos(2.0 * M_PI * (double)(nn + 1) / LTE_JAKES_TONES) / fs;
        if (fabs(C.w_cyc[nn]) > wmax) wmax = fabs(C.w_cyc[nn]);
    }
    const int tile = TDL_THREADS * 8;
    // Two error terms, both relative to |h|.  (1) Taylor remainder of the block polynomial, x^(K+1)/(K+1)! with
    // x = pi w PB: K = 2 needs x < 4.9e-3, K = 4 needs x < 0.075 to stay below 2e-8; PB is the largest power of
    // two that fits.  (2) Inside a thread the polynomial is stepped linearly over its 8 samples, which drops
    // (2 pi w 3.5)^2 / 2: 3e-6 for Vehicular_B 120 km/h at 1.92 MHz, 1e-5 at fD / fs = 2e-4 (390 Hz at 1.92 MHz).
    // Beyond 1e-4 (fD / fs > 6.4e-4) the configuration is refused rather than silently degraded.
    const double lin = 2.0 * M_PI * wmax * 3.5;
    if (0.5 * lin * lin > 1e-4) return LTE_ERR_UNSUPPORTED;
    U.K = 2;
    C.pb = TDL_MAX_PB;
    while (C.pb > tile && M_PI * wmax * C.pb > 4.9e-3) C.pb >>= 1;
    if (M_PI * wmax * C.pb > 4.9e-3) {
        U.K = 4;
        C.pb = TDL_MAX_PB;
        while (C.pb > 32 && M_PI * wmax * C.pb > 0.075) C.pb >>= 1;
        if (M_PI * wmax * C.pb > 0.075) return LTE_ERR_UNSUPPORTED;   // Doppler too high for this fs
    }
    C.pb_log2 = 0;
    while ((1 << C.pb_log2) < C.pb) ++C.pb_log2;
    const int nblk = C.pb >= tile ? 1 : tile / C.pb;
    U.tiles = (int)((n + tile - 1) / tile);
    // the kernel prefetches the coefficients of WHOLE tiles: every block of the last tile must exist even when
    // the stream ends inside it
    C.nbs = C.pb >= tile ? (int)((n + C.pb - 1) / C.pb) : U.tiles * nblk;
    const int RG = tdl_rg(R), NG = (R + RG - 1) / RG;
    U.R2 = NG * RG;
    const int NC = 2 * U.K + 1;
    const size_t ncoef = (size_t)nblk * T * C.num_taps * NC * 2 * U.R2;            // floats per tile
    const int xs = tdl_xs_stride(U.halo, tile);
    U.smem = (size_t)(U.R2 / 2) * 8 * tdl_os_stride(tile) * sizeof(float4) +
             (size_t)2 * T * 8 * xs * sizeof(float2) + 2 * ncoef * sizeof(float);
    if (U.smem > 200 * 1024) return LTE_ERR_UNSUPPORTED;
    const long long total_tiles_ll = (long long)U.tiles * B;
    if (total_tiles_ll > (1LL << 30)) return LTE_ERR_UNSUPPORTED;
    U.total_tiles = (int)total_tiles_ll;
    U.coef_bytes = sizeof(float) * (size_t)B * C.nbs * T * C.num_taps * NC * 2 * U.R2;
    return LTE_OK;
}

extern "C" int64_t lte_channel_tdl_workspace_bytes(const lte_plan* p, const lte_channel_desc* ch, int32_t B, int32_t R,
                                                   int32_t T, int64_t n) {
    if (ch && ch->num_taps == 0) return 0;          // identity link: nothing to stage
    TdlSetup U;
    const int rc = tdl_setup(p, ch, B, R, T, n, U);
    return rc ? (int64_t)rc : (int64_t)U.coef_bytes + 16;
}

extern "C" int lte_channel_tdl(const lte_plan* p, const lte_channel_desc* ch, const lte_c32* tx,
                               const float* phases, lte_c32* faded, double* power, void* workspace, int32_t B,
                               int32_t R, int32_t T, int64_t n, void* stream) {
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
    if (!phases || !faded || !workspace || ((uintptr_t)workspace & 15)) return LTE_ERR_INVALID_ARG;
    TdlSetup U;
    int rc = tdl_setup(p, ch, B, R, T, n, U);
    if (rc) return rc;
    const TdlParams& C = U.C;
    const int K = U.K, halo = U.halo, R2 = U.R2, tiles = U.tiles, total_tiles = U.total_tiles;
    const size_t smem = U.smem;

    // per-block polynomial coefficients in the caller's workspace
    float* coef = (float*)workspace;
    if (R2 != R) LTE_CHECK_CUDA(cudaMemsetAsync(coef, 0, U.coef_bytes, st));   // padding antennas stay zero
    const long long items = (long long)B * C.nbs * R * T * C.num_taps;
    const unsigned cgrid = (unsigned)((items + 255) / 256);
    if (K == 2) jakes_coef_kernel<2><<<cgrid, 256, 0, st>>>(C, phases, coef, R, T, R2, items);
    else jakes_coef_kernel<4><<<cgrid, 256, 0, st>>>(C, phases, coef, R, T, R2, items);
    LTE_CHECK_CUDA(cudaGetLastError());

    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
    int ctas = sms * tdl_min_blocks(R);
    if (ctas > total_tiles) ctas = total_tiles;
    const int chunk = (total_tiles + ctas - 1) / ctas;
    const unsigned grid = (unsigned)((total_tiles + chunk - 1) / chunk);
#define LAUNCH_TDL_KH(RR, KK, HH)                                                                         \
    {                                                                                                     \
        constexpr int RGc = tdl_rg(RR), NGc = (RR + RGc - 1) / RGc;                                       \
        auto k = tdl_kernel<RR, RGc, KK, HH>;                                                             \
        LTE_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
        k<<<grid, TDL_THREADS * NGc, smem, st>>>(C, (const float2*)tx, coef, (float2*)faded, power, T,    \
                                                 (int)n, tiles, total_tiles, chunk);                      \
    }
#define LAUNCH_TDL_K(RR, KK) \
    if (halo == 16) LAUNCH_TDL_KH(RR, KK, 16) else LAUNCH_TDL_KH(RR, KK, 144)
#define LAUNCH_TDL(RR)                                  \
    case RR:                                            \
        if (K == 2) { LAUNCH_TDL_K(RR, 2) } else { LAUNCH_TDL_K(RR, 4) } \
        break;
    switch (R) {
        LAUNCH_TDL(1) LAUNCH_TDL(2) LAUNCH_TDL(3) LAUNCH_TDL(4) LAUNCH_TDL(5) LAUNCH_TDL(6) LAUNCH_TDL(7) LAUNCH_TDL(8)
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
    if (!z && (!lte_ids_fit(row_id0, (uint64_t)rows) || n > (1ll << 32))) return LTE_ERR_UNSUPPORTED;
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
    if (!lte_ids_fit(stream_id0, (uint64_t)B)) return LTE_ERR_UNSUPPORTED;
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
    if (!lte_ids_fit(stream_id0, (uint64_t)B)) return LTE_ERR_UNSUPPORTED;
    if (B == 0) return LTE_OK;
    const int gx = (int)(((per_stream + 1) / 2 + 127) / 128);
    random_phases_kernel<<<(unsigned)((long long)gx * B), 128, 0, (cudaStream_t)stream>>>(
        phases, per_stream, lte_key(seed, LTE_DOMAIN_PHASE), stream_id0, gx);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}
