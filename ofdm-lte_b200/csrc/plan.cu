// Plan creation: host-side table work of the reference's L0/L1 layers
// (config.py:101-130, core/resource_mapper.py:45-111, core/modulator.py:28-59)
// uploaded once per configuration as device constant tables.
#include <math.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

static thread_local cudaError_t g_last_cuda = cudaSuccess;

int lte_set_cuda_error(cudaError_t e) {
    g_last_cuda = e;
    cudaGetLastError();   // clear the sticky-less error state
    return LTE_ERR_CUDA;
}

extern "C" int lte_version(void) { return 200; }   // 2.0: caller-owned workspaces, spectral link, compact sweep layout

extern "C" const char* lte_error_string(int code) {
    switch (code) {
        case LTE_OK: return "ok";
        case LTE_ERR_INVALID_ARG: return "invalid argument";
        case LTE_ERR_UNSUPPORTED: return "unsupported configuration";
        case LTE_ERR_CUDA: return cudaGetErrorString(g_last_cuda);
        case LTE_ERR_NO_DEVICE: return "no CUDA device";
        default: return "unknown error";
    }
}

static float round_down_f32(double x) {
    float f = (float)x;
    if ((double)f > x) f = nextafterf(f, -INFINITY);
    return f;
}

extern "C" int lte_plan_create(const lte_plan_desc* d, const lte_c32* pilots_host, lte_plan** out) {
    if (!d || !out) return LTE_ERR_INVALID_ARG;
    const int N = d->N, Nc = d->Nc;
    if (N < 64 || N > 2048 || (N & (N - 1))) return LTE_ERR_UNSUPPORTED;
    if (Nc < 1 || Nc > N || d->cp < 0 || d->cp > N) return LTE_ERR_INVALID_ARG;
    if (d->bits_per_symbol != 2 && d->bits_per_symbol != 4 && d->bits_per_symbol != 6)
        return LTE_ERR_INVALID_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return LTE_ERR_NO_DEVICE;
    }
    const int nsets = d->mode_simple ? 1 : (d->num_tx_pilot_sets < 1 ? 1 : d->num_tx_pilot_sets);
    if (nsets > LTE_MAX_TX) return LTE_ERR_INVALID_ARG;

    lte_plan* p = new lte_plan();
    p->desc = *d;
    p->nsets = nsets;
    p->blob = nullptr;
    cudaGetDevice(&p->device);

    // --- bin classes (core/resource_mapper.py:57-74) -------------------------------
    std::vector<int16_t> bin_map(N, BIN_NULL);
    int k0, nk;
    if (d->mode_simple) {          // core/modulator.py:192-212: symbols on bins 0..Nc-1
        for (int k = 0; k < Nc; ++k) { bin_map[k] = (int16_t)k; p->data_idx_h.push_back(k); }
        k0 = 0; nk = Nc;
    } else {
        const int gl = (N - Nc) / 2, gr = N - Nc - gl, dc = N / 2;
        for (int k = 0; k < N; ++k) {
            if (k < gl || k >= N - gr || k == dc) continue;
            if ((k - gl) % 6 == 3) {
                bin_map[k] = (int16_t)(BIN_PILOT_FLAG | (int)p->pilot_idx_h.size());
                p->pilot_idx_h.push_back(k);
            } else {
                bin_map[k] = (int16_t)p->data_idx_h.size();
                p->data_idx_h.push_back(k);
            }
        }
        k0 = gl; nk = Nc;
    }
    const int Nd = (int)p->data_idx_h.size(), Np = (int)p->pilot_idx_h.size();
    if (Np > 0 && !pilots_host) { delete p; return LTE_ERR_INVALID_ARG; }

    // --- pilot sets: owned pilots, LS inverses, interpolation segments ---------------
    std::vector<float2> pilots((size_t)nsets * (Np ? Np : 1), make_float2(0.f, 0.f));
    std::vector<int16_t> pset_bin((size_t)nsets * (Np ? Np : 1), 0);
    std::vector<float2> pset_inv((size_t)nsets * (Np ? Np : 1), make_float2(0.f, 0.f));
    std::vector<int16_t> pset_seg((size_t)nsets * N, -1);
    std::vector<int> pset_cnt(nsets, 0);
    for (int s = 0; s < nsets; ++s) {
        int cnt = 0;
        for (int i = 0; i < Np; ++i) {
            lte_c32 v = pilots_host[(size_t)s * Np + i];
            pilots[(size_t)s * Np + i] = make_float2(v.re, v.im);
            if (v.re != 0.f || v.im != 0.f) {
                pset_bin[(size_t)s * Np + cnt] = (int16_t)p->pilot_idx_h[i];
                double den = (double)v.re * v.re + (double)v.im * v.im;
                pset_inv[(size_t)s * Np + cnt] = make_float2((float)(v.re / den), (float)(-v.im / den));
                ++cnt;
            }
        }
        pset_cnt[s] = cnt;
        int lo = -1;
        for (int k = 0; k < N; ++k) {
            while (lo + 1 < cnt && pset_bin[(size_t)s * Np + lo + 1] <= k) ++lo;
            pset_seg[(size_t)s * N + k] = (int16_t)lo;
        }
    }

    // --- twiddles exp(-2 pi i m / N) from double ---------------------------------------
    // followed by the per-pass tables of the packed FFT (fft2.cuh): for a pass of radix R over stride NS,
    // T[(t - 1) NS + k] = exp(-2 pi i k t / (NS R)), t = 1..R-1, k < NS, so that the threads of a warp
    // (consecutive k) read consecutive entries; second pass at N, third pass at N + FFT2_TW_PASS3
    std::vector<float2> tw((size_t)N + FFT2_TW_EXTRA, make_float2(0.f, 0.f));
    for (int m = 0; m < N; ++m) {
        double a = -2.0 * M_PI * (double)m / (double)N;
        tw[m] = make_float2((float)cos(a), (float)sin(a));
    }
    {
        int sched[2][2] = {{0, 0}, {0, 0}};                   // (R, NS) of the twiddled passes
        switch (N) {
            case 2048: sched[0][0] = 16; sched[0][1] = 16; sched[1][0] = 8; sched[1][1] = 256; break;
            case 1024: sched[0][0] = 8; sched[0][1] = 16; sched[1][0] = 8; sched[1][1] = 128; break;
            case 512: sched[0][0] = 8; sched[0][1] = 8; sched[1][0] = 8; sched[1][1] = 64; break;
            case 256: sched[0][0] = 16; sched[0][1] = 16; break;
            case 128: sched[0][0] = 8; sched[0][1] = 16; break;
            case 64: sched[0][0] = 4; sched[0][1] = 16; break;
            default: break;
        }
        for (int ps = 0; ps < 2; ++ps) {
            const int R = sched[ps][0], NS = sched[ps][1];
            float2* T = tw.data() + N + (ps ? FFT2_TW_PASS3 : 0);
            for (int t = 1; t < R; ++t)
                for (int k = 0; k < NS; ++k) {
                    double a = -2.0 * M_PI * (double)k * (double)t / ((double)NS * (double)R);
                    T[(size_t)(t - 1) * NS + k] = make_float2((float)cos(a), (float)sin(a));
                }
        }
    }

    // --- bin pairs of the compact sweep layout --------------------------------------------
    std::vector<int16_t> pair_bin;
    int ndp, npp;
    {
        std::vector<int> nulls;                                   // null bins inside the occupied window (DC)
        for (int k = k0; k < k0 + nk; ++k)
            if (bin_map[k] == BIN_NULL) nulls.push_back(k);
        size_t nu = 0;
        auto pad = [&]() {
            if (pair_bin.size() & 1) pair_bin.push_back(nu < nulls.size() ? (int16_t)nulls[nu++] : (int16_t)-1);
        };
        for (int i = 0; i < Nd; ++i) pair_bin.push_back((int16_t)p->data_idx_h[i]);
        pad();
        ndp = (int)pair_bin.size() / 2;
        for (int i = 0; i < Np; ++i) pair_bin.push_back((int16_t)p->pilot_idx_h[i]);
        pad();
        npp = (int)pair_bin.size() / 2 - ndp;
        while (nu < nulls.size()) pair_bin.push_back((int16_t)nulls[nu++]);
        pad();
    }

    // --- single device blob -------------------------------------------------------------
    auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t off_bin = 0;
    size_t off_didx = align(off_bin + sizeof(int16_t) * N);
    size_t off_pidx = align(off_didx + sizeof(int16_t) * (Nd ? Nd : 1));
    size_t off_pil = align(off_pidx + sizeof(int16_t) * (Np ? Np : 1));
    size_t off_pbin = align(off_pil + sizeof(float2) * pilots.size());
    size_t off_pinv = align(off_pbin + sizeof(int16_t) * pset_bin.size());
    size_t off_pseg = align(off_pinv + sizeof(float2) * pset_inv.size());
    size_t off_pcnt = align(off_pseg + sizeof(int16_t) * pset_seg.size());
    size_t off_tw = align(off_pcnt + sizeof(int) * nsets);
    size_t off_pair = align(off_tw + sizeof(float2) * tw.size());
    size_t total = align(off_pair + sizeof(int16_t) * pair_bin.size());
    std::vector<char> host(total, 0);
    std::vector<int16_t> didx16(Nd ? Nd : 1, 0);
    for (int i = 0; i < Nd; ++i) didx16[i] = (int16_t)p->data_idx_h[i];
    memcpy(&host[off_bin], bin_map.data(), sizeof(int16_t) * N);
    memcpy(&host[off_didx], didx16.data(), sizeof(int16_t) * didx16.size());
    std::vector<int16_t> pidx16(Np ? Np : 1, 0);
    for (int i = 0; i < Np; ++i) pidx16[i] = (int16_t)p->pilot_idx_h[i];
    memcpy(&host[off_pidx], pidx16.data(), sizeof(int16_t) * pidx16.size());
    memcpy(&host[off_pil], pilots.data(), sizeof(float2) * pilots.size());
    memcpy(&host[off_pbin], pset_bin.data(), sizeof(int16_t) * pset_bin.size());
    memcpy(&host[off_pinv], pset_inv.data(), sizeof(float2) * pset_inv.size());
    memcpy(&host[off_pseg], pset_seg.data(), sizeof(int16_t) * pset_seg.size());
    memcpy(&host[off_pcnt], pset_cnt.data(), sizeof(int) * nsets);
    memcpy(&host[off_tw], tw.data(), sizeof(float2) * tw.size());
    memcpy(&host[off_pair], pair_bin.data(), sizeof(int16_t) * pair_bin.size());
    if (cudaMalloc(&p->blob, total) != cudaSuccess ||
        cudaMemcpy(p->blob, host.data(), total, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaError_t e = cudaGetLastError();
        delete p;
        return lte_set_cuda_error(e);
    }
    char* b = (char*)p->blob;
    DevPlan& D = p->dev;
    memset(&D, 0, sizeof(D));
    D.N = N; D.Nc = Nc; D.cp = d->cp; D.L = N + d->cp; D.Nd = Nd; D.Np = Np; D.bps = d->bits_per_symbol;
    D.log2N = 0; while ((1 << D.log2N) < N) ++D.log2N;
    D.k0_useful = k0; D.nk_useful = nk;
    D.bin_map = (const int16_t*)(b + off_bin);
    D.data_idx = (const int16_t*)(b + off_didx);
    D.pilot_idx = (const int16_t*)(b + off_pidx);
    D.pilots = (const float2*)(b + off_pil);
    D.pset_bin = (const int16_t*)(b + off_pbin);
    D.pset_inv = (const float2*)(b + off_pinv);
    D.pset_seg = (const int16_t*)(b + off_pseg);
    D.pset_cnt = (const int*)(b + off_pcnt);
    D.twiddle = (const float2*)(b + off_tw);
    D.pair_bin = (const int16_t*)(b + off_pair);
    D.ndp = ndp; D.npp = npp; D.npairs = (int)pair_bin.size() / 2;
    D.inv_sqrt_n = (float)(1.0 / sqrt((double)N));

    // --- constellation axis levels and slicer thresholds (core/modulator.py:28-59) -----
    for (int i = 0; i < 7; ++i) D.thr[i] = INFINITY;     // unused thresholds never fire (slicer bisection)
    if (D.bps == 2) {          // [1+1j, 1-1j, -1+1j, -1-1j]/sqrt(2): index bit 0 -> +, 1 -> -
        D.nlev = 2;
        D.lev[0] = (float)(1.0 / sqrt(2.0));
        D.lev[1] = (float)(-1.0 / sqrt(2.0));
    } else {
        D.nlev = 1 << (D.bps / 2);
        const double nrm = D.bps == 4 ? sqrt(10.0) : sqrt(42.0);
        for (int i = 0; i < D.nlev; ++i) D.lev[i] = (float)((2 * i - (D.nlev - 1)) / nrm);
        // midpoints; y > thr[i] (thr rounded towards -inf) <=> y > exact midpoint for every
        // fp32 y, and a tie (y == 0) resolves to the lower level like np.argmin's first minimum
        for (int i = 0; i < D.nlev - 1; ++i) D.thr[i] = round_down_f32((2 * i - (D.nlev - 2)) / nrm);
    }
    *out = p;
    return LTE_OK;
}

// Bluestein chirp-z tables of the M-point DFT (kernel in dft.cu): plan state, built once per M on the host in fp64.
// Synchronous (cudaMalloc + cudaMemcpy), like lte_plan_create: call it while setting the plan up, not in a stream.
extern "C" int lte_plan_add_dft(lte_plan* p, int32_t M) {
    if (!p || M < 1) return LTE_ERR_INVALID_ARG;
    if (p->dft.count(M)) return LTE_OK;
    int NB = 64;
    while (NB < 2 * M - 1) NB <<= 1;
    if (NB > 2048) return LTE_ERR_UNSUPPORTED;
    std::vector<double> wr(M), wi(M);
    for (int n = 0; n < M; ++n) {
        const long long q = ((long long)n * n) % (2LL * M);        // exact phase reduction
        const double a = -M_PI * (double)q / (double)M;
        wr[n] = cos(a); wi[n] = sin(a);
    }
    // b[m] = conj(w[|m|]) placed circularly; its NB-point DFT by the definition (fp64)
    std::vector<double> br(NB, 0.0), bi(NB, 0.0);
    for (int m = 0; m < M; ++m) {
        br[m] = wr[m]; bi[m] = -wi[m];
        if (m) { br[NB - m] = wr[m]; bi[NB - m] = -wi[m]; }
    }
    std::vector<double> cr(NB), ci(NB);
    for (int i = 0; i < NB; ++i) { const double a = -2.0 * M_PI * i / NB; cr[i] = cos(a); ci[i] = sin(a); }
    std::vector<float2> w(M), bf(NB), tw(NB);
    for (int i = 0; i < NB; ++i) tw[i] = make_float2((float)cr[i], (float)ci[i]);
    const double scale = 1.0 / ((double)NB * sqrt((double)M));
    for (int k = 0; k < NB; ++k) {
        double sr = 0.0, si = 0.0;
        for (int m = 0; m < NB; ++m) {
            if (br[m] == 0.0 && bi[m] == 0.0) continue;
            const int t = (int)(((long long)k * m) & (NB - 1));
            sr += br[m] * cr[t] - bi[m] * ci[t];
            si += br[m] * ci[t] + bi[m] * cr[t];
        }
        bf[k] = make_float2((float)(sr * scale), (float)(si * scale));
    }
    for (int n = 0; n < M; ++n) w[n] = make_float2((float)wr[n], (float)wi[n]);
    DftTables t;
    t.M = M; t.NB = NB;
    t.w = t.bf = t.tw = nullptr;
    if (cudaMalloc(&t.w, sizeof(float2) * M) != cudaSuccess || cudaMalloc(&t.bf, sizeof(float2) * NB) != cudaSuccess ||
        cudaMalloc(&t.tw, sizeof(float2) * NB) != cudaSuccess ||
        cudaMemcpy(t.w, w.data(), sizeof(float2) * M, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(t.bf, bf.data(), sizeof(float2) * NB, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(t.tw, tw.data(), sizeof(float2) * NB, cudaMemcpyHostToDevice) != cudaSuccess) {
        const cudaError_t e = cudaGetLastError();
        if (t.w) cudaFree(t.w);
        if (t.bf) cudaFree(t.bf);
        if (t.tw) cudaFree(t.tw);
        return lte_set_cuda_error(e);
    }
    p->dft[M] = t;
    return LTE_OK;
}

extern "C" int lte_plan_destroy(lte_plan* p) {
    if (!p) return LTE_OK;
    for (auto& kv : p->dft) { cudaFree(kv.second.w); cudaFree(kv.second.bf); cudaFree(kv.second.tw); }
    if (p->blob) cudaFree(p->blob);
    delete p;
    return LTE_OK;
}

extern "C" int lte_plan_num_data(const lte_plan* p) { return p ? p->dev.Nd : LTE_ERR_INVALID_ARG; }
extern "C" int lte_plan_num_pilots(const lte_plan* p) { return p ? p->dev.Np : LTE_ERR_INVALID_ARG; }

extern "C" int lte_plan_compact_shape(const lte_plan* p, int32_t* ndp, int32_t* npp) {
    if (!p || !ndp || !npp) return LTE_ERR_INVALID_ARG;
    *ndp = p->dev.ndp;
    *npp = p->dev.npp;
    return LTE_OK;
}

extern "C" int lte_plan_indices_host(const lte_plan* p, int32_t* data_idx, int32_t* pilot_idx) {
    if (!p) return LTE_ERR_INVALID_ARG;
    if (data_idx) memcpy(data_idx, p->data_idx_h.data(), sizeof(int32_t) * p->data_idx_h.size());
    if (pilot_idx) memcpy(pilot_idx, p->pilot_idx_h.data(), sizeof(int32_t) * p->pilot_idx_h.size());
    return LTE_OK;
}

extern "C" int lte_plan_window(const lte_plan* p, int window, int32_t* k0, int32_t* nk) {
    if (!p || !k0 || !nk) return LTE_ERR_INVALID_ARG;
    if (window == LTE_WINDOW_FULL) { *k0 = 0; *nk = p->dev.N; }
    else if (window == LTE_WINDOW_USEFUL) { *k0 = p->dev.k0_useful; *nk = p->dev.nk_useful; }
    else return LTE_ERR_INVALID_ARG;
    return LTE_OK;
}

// ------------------------------------------------------------------------------ measurement helper
// Peak packed-fp32 throughput of the device as the kernels of this library see it: independent chains of
// fma.rn.f32x2 (FFMA2), 8 per thread, no memory traffic.  bench.py times one launch with CUDA events and divides:
// flops = grid * 256 threads * iters * 8 instructions * 4 (two lanes x multiply-add).  The compute roofline of the
// FMA-bound kernels is quoted against this number, measured on the box the bench runs on.
__global__ void __launch_bounds__(256)
fp32_peak_kernel(float* __restrict__ sink, int iters, float s) {
    unsigned long long p[8];
    const unsigned long long pb = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(s);
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = (unsigned long long)(threadIdx.x + i + 1) * 0x100000001ull;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pb));
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
    sink[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

extern "C" int64_t lte_fp32_peak_launch(float* sink, int32_t iters, void* stream) {
    if (!sink || iters < 1) return LTE_ERR_INVALID_ARG;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess) return LTE_ERR_CUDA;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * 8;
    fp32_peak_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(sink, iters, 1.0001f);
    if (cudaGetLastError() != cudaSuccess) return LTE_ERR_CUDA;
    return (int64_t)grid * 256 * (int64_t)iters * 8 * 4;        // flops of this launch
}
