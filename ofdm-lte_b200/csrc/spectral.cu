// Spectral form of the fading link for the sweep engine (T = 1, low Doppler): the tapped-delay-line
// Rayleigh channel (core/rayleighchannel.py:20-58), the stream power behind the AWGN
// (core/channel.py:216-218), CP strip and fft / sqrt(N) (core/lte_receiver.py:444-491) without ever
// forming a faded time-domain stream and with ONE forward transform per OFDM symbol instead of one
// per receive antenna.
//
// When the Jakes process of every (antenna, tap) is linear over an OFDM symbol to the engine's
// accuracy -- the economised K = 1 fit of tdl.cuh, h(m) = c0 + c1 tau(m), tau(m) = m - (L-1)/2,
// remainder <= 5e-7 |h| for pi fD L / fs <= 1.41e-3 (3 km/h at 2 GHz) -- the delayed copies of the
// useful part are cyclic shifts (d_t <= cp), and bin by bin
//
//     Y_r[k] = sum_t e_t[k] { (c0 + d_t c1) X[k] + c1 (G[k] - N T_t[k]) },   e_t[k] = exp(-2 pi j k d_t / N)
//
// X    the transmitted grid (rebuilt from the index bytes, never read as samples),
// G    fft((n - n_c) u[n]) / sqrt(N), the spectrum of the ramp-weighted useful samples u: antenna- and
//      tap-independent, produced by the TX kernel below right after its IFFT,
// T_t  the partial DFT of the last d_t samples of u -- the samples whose delayed copy arrives through
//      the cyclic prefix and therefore meets the ramp N samples earlier; one Horner sweep in
//      exp(-2 pi j k / N) over the last max(d_t) samples yields every tap's term.
// The stream power is sum_k |Y_r[k]|^2 (Parseval; the leakage of the c1 term outside the occupied
// window is ~1e-7 of the power) plus the CP samples, which are evaluated in the time domain from the
// symbol tails the TX kernel leaves behind (they also carry the inter-symbol leakage).
// tests/spectral_ref.py is the fp64 restatement; tests/test_spectral_identity.py checks it against the
// oracle's sample-by-sample path, tests/test_gpu_spectral.py checks these kernels against both.
#include <math.h>
#include <string.h>

#include <type_traits>

#include "fft2.cuh"
#include "tdl.cuh"

// ------------------------------------------------------------------------------ TX side
// QAM map + resource grid + IFFT (as tx_map_ifft_kernel, core/modulator.py:61-88,214-302,
// core/resource_mapper.py:181-223), then instead of the time-domain stream:
//   tail [B*S][cp]  the last cp samples of every symbol (= its cyclic prefix),
//   G    [B*S][nk]  fft((n - n_c) u[n]) / sqrt(N) on the occupied window.
template <int N>
__global__ void __launch_bounds__(FFT_CTA_THREADS, 5)
tx_spectral_kernel(const DevPlan P, const uint8_t* __restrict__ idx, float2* __restrict__ G,
                   float2* __restrict__ tail, int k0, int nk, unsigned total) {
    constexpr int TPF = N / FFT_ELEMS, PPC = fft2_pairs_per_cta(N);
    extern __shared__ float4 smem4[];
    __shared__ float s_lev[8];
    if (threadIdx.x < 8) s_lev[threadIdx.x] = P.lev[threadIdx.x];
    const int p_local = threadIdx.x / TPF, j = threadIdx.x % TPF;
    float4* sbuf = smem4 + (size_t)p_local * fft_smem_elems(N);
    const unsigned f0 = (blockIdx.x * PPC + p_local) * 2u;   // OFDM symbol ids f0, f0 + 1 (= b*S + s)
    const int h = P.bps >> 1, mask = (1 << h) - 1;
    const bool valid[2] = {f0 < total, f0 + 1 < total};
    const uint8_t* ip[2] = {idx + (size_t)f0 * P.Nd, idx + (size_t)(f0 + 1) * P.Nd};
    __syncthreads();          // s_lev

    const int used_lo = P.k0_useful, used_hi = P.k0_useful + P.nk_useful;
    int code[FFT_ELEMS];
#pragma unroll
    for (int e = 0; e < FFT_ELEMS; ++e) {
        const bool used = (e + 1) * TPF > used_lo && e * TPF < used_hi;
        code[e] = used ? (int)__ldg(&P.bin_map[j + e * TPF]) : BIN_NULL;
    }
    uint8_t ib[2][FFT_ELEMS];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int e = 0; e < FFT_ELEMS; ++e) {
            ib[m][e] = 0;
            if ((e + 1) * TPF > used_lo && e * TPF < used_hi) {
                const int c = code[e];
                if (valid[m] && c >= 0 && !(c & BIN_PILOT_FLAG)) ib[m][e] = ip[m][c];
            }
        }
    c2 v[FFT_ELEMS];
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
            } else if (pil) {
                const float2 pv = P.pilots[c & (BIN_PILOT_FLAG - 1)];
                if (valid[0]) a = pv;
                if (valid[1]) g = pv;
            }
        }
        v[e] = {pk(a.x, g.x), pk(a.y, g.y)};
    }

    fft2_run<N, true>(v, sbuf, P.twiddle, j);

    // tails (tx samples = raw / sqrt(N)) and the ramp (n - n_c) / N: G = fft(ramp * raw) / ... / sqrt(N)
    const int tail0 = N - P.cp;
    const float nc = 0.5f * (float)(P.L - 1) - (float)P.cp;
    const float inv_n = P.inv_sqrt_n * P.inv_sqrt_n;
#pragma unroll
    for (int e = 0; e < FFT_ELEMS; ++e) {
        const int n = j + e * TPF;
        if ((e + 1) * TPF > tail0 && n >= tail0) {
            const f2 sc = pk(P.inv_sqrt_n, P.inv_sqrt_n);
            float re[2], im[2];
            upk(mul2(v[e].re, sc), re[0], re[1]);
            upk(mul2(v[e].im, sc), im[0], im[1]);
#pragma unroll
            for (int m = 0; m < 2; ++m)
                if (valid[m]) tail[(size_t)(f0 + m) * P.cp + (n - tail0)] = make_float2(re[m], im[m]);
        }
        const float ramp = ((float)n - nc) * inv_n;
        const f2 rr = pk(ramp, ramp);
        v[e].re = mul2(v[e].re, rr);
        v[e].im = mul2(v[e].im, rr);
    }
    __syncthreads();          // every reader of the IFFT's last exchange is done with sbuf

    fft2_run<N, false>(v, sbuf, P.twiddle, j);

    float2* o0 = G + (size_t)f0 * nk + (j - k0);
    float2* o1 = o0 + nk;
    const unsigned kb = (unsigned)(j - k0);
#pragma unroll
    for (int e = 0; e < FFT_ELEMS; ++e) {
        if ((e + 1) * TPF > k0 && e * TPF < k0 + nk && kb + (unsigned)(e * TPF) < (unsigned)nk) {
            float a, c, d, g;
            upk(v[e].re, a, c);
            upk(v[e].im, d, g);
            if (valid[0]) o0[e * TPF] = make_float2(a, d);
            if (valid[1]) o1[e * TPF] = make_float2(c, g);
        }
    }
}

// ------------------------------------------------------------------------------ channel side
struct SpecParams {
    int num_taps;
    int delay[LTE_MAX_TAPS];        // ascending
    int ord[LTE_MAX_TAPS];          // tap index (as in `phases`) of sorted position i
    int pos[LTE_MAX_TAPS];          // sorted position of tap index i
    float gain[LTE_MAX_TAPS];       // by tap index, includes sqrt(2/16)
    double w_cyc[LTE_JAKES_TONES];  // fD cos(alpha_n) / fs   [cycles per sample]
    int dmax;
};

#define SPEC_MAX_THREADS 608
#define SPEC_XMAX 352               // staged time samples per symbol: dmax + cp <= 160 + 192
#define SPEC_GMAX 1216              // window bins staged per symbol (2 per thread)

// Linear Jakes fit per (OFDM symbol, antenna, tap) in the layout the channel kernel stages verbatim:
//   coef[f][sorted tap][slot][re|im][R2],  slot 0: a = c0 + d c1,  1: c1,  2: c0
// c0 / c1 are the economised K = 1 coefficients of jakes_coef_kernel<1> (tdl.cuh): fp64 phase reduction at the
// symbol centre, c0 = g sum_n e^{j theta_n} (1 - X_n^2 / 4), c1 = g sum_n e^{j theta_n} j x_n.
__global__ void __launch_bounds__(256)
spectral_coef_kernel(const SpecParams C, const float* __restrict__ phases, float* __restrict__ coef, int R, int R2,
                     int S, int L, long long total_items) {
    const int nlt = R * C.num_taps;
    const long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // f * nlt + (r * taps + tap)
    if (it >= total_items) return;
    const int trip = (int)(it % nlt);
    const long long f = it / nlt;
    const int s = (int)(f % S);
    const long long b = f / S;
    const double mc = (double)s * L + 0.5 * (L - 1);
    const float4* up = (const float4*)(phases + ((size_t)b * nlt + trip) * LTE_JAKES_TONES);
    float u[LTE_JAKES_TONES];
#pragma unroll
    for (int i = 0; i < LTE_JAKES_TONES / 4; ++i) {
        const float4 v = __ldg(&up[i]);
        u[4 * i] = v.x; u[4 * i + 1] = v.y; u[4 * i + 2] = v.z; u[4 * i + 3] = v.w;
    }
    float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int tone = 0; tone < LTE_JAKES_TONES; ++tone) {
        double turns = C.w_cyc[tone] * mc + (double)u[tone];
        turns -= floor(turns);
        float sn, cs;
        sincospif(2.0f * (float)turns, &sn, &cs);
        const float x = (float)(6.283185307179586 * C.w_cyc[tone]);   // rad / sample
        const float X = x * 0.5f * (float)L;
        const float c0 = 1.0f - 0.25f * X * X;
        a0.x += cs * c0;
        a0.y += sn * c0;
        a1.x += -sn * x;
        a1.y += cs * x;
    }
    const int tap = trip % C.num_taps, r = trip / C.num_taps;
    const float g = C.gain[tap];
    const int ts = C.pos[tap];
    const float d = (float)C.delay[ts];
    a0.x *= g; a0.y *= g; a1.x *= g; a1.y *= g;
    float* c = coef + ((size_t)f * C.num_taps + ts) * 6 * R2 + r;
    c[0 * R2] = fmaf(d, a1.x, a0.x);
    c[1 * R2] = fmaf(d, a1.y, a0.y);
    c[2 * R2] = a1.x;
    c[3 * R2] = a1.y;
    c[4 * R2] = a0.x;
    c[5 * R2] = a0.y;
}

__device__ __forceinline__ void cp_async8_s(unsigned d, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(g));
}
__device__ __forceinline__ void cp_async8_zfill_s(unsigned d, const void* g, bool valid) {
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(g), "r"(sz));
}
__device__ __forceinline__ void cp_async16_s(unsigned d, const void* g) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g));
}
__device__ __forceinline__ f2 neg2(f2 a) { float x, y; upk(a, x, y); return pk(-x, -y); }
// acc += a * b in place: the read-write constraint keeps the accumulator in the same register pair
__device__ __forceinline__ void fma2_acc(f2& acc, f2 a, f2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc.v) : "l"(a.v), "l"(b.v)); }

// Persistent CTAs, each walking whole streams (b = blockIdx.x, += gridDim.x) symbol by symbol.  A thread
// owns the SAME two bins kk = tid and kk + half of the occupied window for its whole life, so everything
// that depends on the bin only -- bin class, pilot value, exp(-2 pi j k / N) and its powers e_t[k] -- sits in
// registers, and the antenna-independent arithmetic (Horner sweep over the symbol tail, V_t = e_t X,
// W_t = e_t G - Q_t) runs packed over the two bins.  The combine runs packed over receive-antenna pairs with
// the bin's V / W as scalar-broadcast operands and the coefficients as 128-bit shared-memory broadcasts.
// Per symbol the CTA stages with cp.async (double buffered, one barrier) the symbol's G window, its Jakes
// coefficients and the dmax + cp time samples around the cyclic prefix: the next symbol's bytes are in
// flight while the current one is computed, and no register holds them.  Stream power stays in registers
// across the stream's symbols and leaves as one atomic per warp and antenna.
// COMPACT: Y holds the data bins only ([B*R][S][Nd]) and the pilot bins of every slot's first symbol go
// to Yp ([B*R][slots][Np]) -- nothing else is ever read downstream.
template <int NT, int RP, bool COMPACT, bool Z0>
__global__ void __launch_bounds__(SPEC_MAX_THREADS, 1)
channel_spectral_kernel(const DevPlan P, const SpecParams C, const uint8_t* __restrict__ idx,
                        const float2* __restrict__ G, const float2* __restrict__ tail,
                        const float* __restrict__ coef_g, float2* __restrict__ Y, float2* __restrict__ Yp,
                        double* __restrict__ power, int k0, int nk, int half, int S, int R, int B) {
    constexpr int R2 = 2 * RP, NCF = NT * 6 * R2;               // coefficient floats per symbol
    __shared__ __align__(16) float s_cf[2][NCF];
    __shared__ __align__(16) float2 s_x[2][SPEC_XMAX];          // x[local index i - dmax], i in [-dmax, cp)
    __shared__ __align__(16) float2 s_g[2][SPEC_GMAX];
    __shared__ float s_lev[8];
    const int tid = threadIdx.x;
    const int cp = P.cp, L = P.L, dmax = C.dmax, nx = dmax + cp;
    const int nslot = (S + LTE_SLOT_SYMBOLS - 1) / LTE_SLOT_SYMBOLS;
    if (tid < 8) s_lev[tid] = P.lev[tid];

    // ---- per-thread constants of its two bins ---------------------------------------------------
    const int kkA = tid, kkB = tid + half;
    const bool okA = tid < half, okB = tid < half && kkB < nk;
    const int kA = (okA ? kkA : 0) + k0, kB = (okB ? kkB : 0) + k0;
    const int codeA = okA ? (int)__ldg(&P.bin_map[kA]) : BIN_NULL, codeB = okB ? (int)__ldg(&P.bin_map[kB]) : BIN_NULL;
    const bool datA = codeA >= 0 && !(codeA & BIN_PILOT_FLAG), datB = codeB >= 0 && !(codeB & BIN_PILOT_FLAG);
    const bool pilA = codeA >= 0 && !datA, pilB = codeB >= 0 && !datB;
    const int slotA = codeA & (BIN_PILOT_FLAG - 1), slotB = codeB & (BIN_PILOT_FLAG - 1);   // data index or pilot index
    const float2 pvA = pilA ? P.pilots[slotA] : make_float2(0.f, 0.f);
    const float2 pvB = pilB ? P.pilots[slotB] : make_float2(0.f, 0.f);
    f2 wre, wim, ere[NT], eim[NT];
    {
        const float2 a = __ldg(&P.twiddle[kA]), c = __ldg(&P.twiddle[kB]);
        wre = pk(a.x, c.x);
        wim = pk(a.y, c.y);
#pragma unroll
        for (int ts = 0; ts < NT; ++ts) {
            const int d = C.delay[ts];
            const float2 ea = __ldg(&P.twiddle[(kA * d) & (P.N - 1)]), ec = __ldg(&P.twiddle[(kB * d) & (P.N - 1)]);
            ere[ts] = pk(ea.x, ec.x);
            eim[ts] = pk(ea.y, ec.y);
        }
    }
    const int hb = P.bps >> 1, mask = (1 << hb) - 1;
    const float sqn = sqrtf((float)P.N);
    const f2 nsq = pk(-sqn, -sqn);
    const float tf = (float)tid - 0.5f * (float)(L - 1);        // polynomial argument of CP sample i = tid
    const f2 tau = pk(tf, tf);
    // where this thread's outputs go inside one (antenna, symbol) row, and whether they go anywhere
    const int strideY = COMPACT ? P.Nd : nk;
    const int offA = COMPACT ? slotA : kkA, offB = COMPACT ? slotB : kkB;
    const bool stA = COMPACT ? datA : okA, stB = COMPACT ? datB : okB;

    f2 pw[RP];
#pragma unroll
    for (int p = 0; p < RP; ++p) pw[p] = pk(0.f, 0.f);

    // ---- staging: cp.async of symbol f into buffer `buf`; index bytes into registers -----------
    // shared addresses and per-thread global offsets are formed once; per symbol only f * row-length is added
    const unsigned sgA = (unsigned)__cvta_generic_to_shared(&s_g[0][kkA]);
    const unsigned sxT = (unsigned)__cvta_generic_to_shared(&s_x[0][tid < SPEC_XMAX ? tid : 0]);
    const unsigned scT = (unsigned)__cvta_generic_to_shared(&s_cf[0][tid < NCF / 4 ? 4 * tid : 0]);
    const float2* gT = G + kkA;
    const float2* xT = tail + (tid - dmax);                     // local sample index tid - dmax; < 0: the previous symbol's tail
    const float* cT = coef_g + 4 * tid;
    const uint8_t* iA = idx + slotA;
    const uint8_t* iB = idx + slotB;
    const bool xrow = tid < nx, xhist = tid < dmax, crow = tid < NCF / 4;
    int ibA = 0, ibB = 0;
    auto prefetch = [&](unsigned f, unsigned s, unsigned buf) {
        const float2* g = gT + (unsigned long long)f * (unsigned)nk;
        const unsigned sg = sgA + buf * (unsigned)(SPEC_GMAX * sizeof(float2));
        if (okA) cp_async8_s(sg, g);
        if (okB) cp_async8_s(sg + (unsigned)half * 8u, g + half);
        if (xrow) {
            const bool have = !xhist || s > 0;
            cp_async8_zfill_s(sxT + buf * (unsigned)(SPEC_XMAX * sizeof(float2)),
                              have ? (const void*)(xT + (unsigned long long)f * (unsigned)cp) : (const void*)tail, have);
        }
        if (crow) cp_async16_s(scT + buf * (unsigned)(NCF * sizeof(float)), cT + (unsigned long long)f * (unsigned)NCF);
        const unsigned long long io = (unsigned long long)f * (unsigned)P.Nd;
        if (datA) ibA = iA[io];
        if (datB) ibB = iB[io];
    };

    unsigned b = blockIdx.x, s = 0;
    unsigned cur = 0;
    if (b < (unsigned)B) prefetch(b * (unsigned)S, 0, 0);
    cp_async_commit();
    while (b < (unsigned)B) {
        cp_async_wait<0>();
        __syncthreads();                                        // buffer `cur` landed; everyone is done with `cur ^ 1`
        // transmitted grid values of the two bins (the index bytes were loaded a symbol ahead)
        float2 xa = pvA, xb = pvB;
        if (datA) xa = make_float2(s_lev[(ibA >> hb) & mask], s_lev[ibA & mask]);
        if (datB) xb = make_float2(s_lev[(ibB >> hb) & mask], s_lev[ibB & mask]);
        unsigned nb = b, ns = s + 1;
        if (ns == (unsigned)S) { ns = 0; nb = b + gridDim.x; }
        if (nb < (unsigned)B) prefetch(nb * (unsigned)S + ns, ns, cur ^ 1u);
        cp_async_commit();
        const float* cf = s_cf[cur];

        // ---- cyclic-prefix samples: time domain, only into the stream power -----------------------
        if (tid < cp) {
            f2 yre[RP], yim[RP];
#pragma unroll
            for (int p = 0; p < RP; ++p) { yre[p] = pk(0.f, 0.f); yim[p] = pk(0.f, 0.f); }
#pragma unroll
            for (int ts = 0; ts < NT; ++ts) {
                const float2 x = s_x[cur][dmax + tid - C.delay[ts]];
                const f2 xre = pk(x.x, x.x), xim = pk(x.y, x.y), nxim = pk(-x.y, -x.y);
                const float* ct = cf + ts * 6 * R2;
#pragma unroll
                for (int p = 0; p < RP; ++p) {
                    const float2 c1r = *(const float2*)(ct + 2 * R2 + 2 * p), c1i = *(const float2*)(ct + 3 * R2 + 2 * p);
                    const float2 c0r = *(const float2*)(ct + 4 * R2 + 2 * p), c0i = *(const float2*)(ct + 5 * R2 + 2 * p);
                    const f2 hre = fma2(pk(c1r.x, c1r.y), tau, pk(c0r.x, c0r.y));
                    const f2 him = fma2(pk(c1i.x, c1i.y), tau, pk(c0i.x, c0i.y));
                    yre[p] = fma2(hre, xre, yre[p]); yre[p] = fma2(him, nxim, yre[p]);
                    yim[p] = fma2(hre, xim, yim[p]); yim[p] = fma2(him, xre, yim[p]);
                }
            }
#pragma unroll
            for (int p = 0; p < RP; ++p) {
                pw[p] = fma2(yre[p], yre[p], pw[p]);
                pw[p] = fma2(yim[p], yim[p], pw[p]);
            }
        }

        // ---- the thread's two bins ------------------------------------------------------------------
        if (okA) {
            const float2 cgA = s_g[cur][kkA], cgB = okB ? s_g[cur][kkB] : make_float2(0.f, 0.f);
            const f2 xre = pk(xa.x, xb.x), xim = pk(xa.y, xb.y);
            const f2 gre_ = pk(cgA.x, cgB.x), gim_ = pk(cgA.y, cgB.y);
            f2 hre = pk(0.f, 0.f), him = pk(0.f, 0.f);          // Horner accumulator, packed over the two bins
            f2 yreA[RP], yimA[RP], yreB[RP], yimB[RP];           // packed over the antenna pair
            const float2* up = &s_x[cur][nx - 1];               // u[N - 1 - p] = up[-p]
            const f2 nwim = neg2(wim);
            int pstep = 0;
#pragma unroll
            for (int ts = 0; ts < NT; ++ts) {
                if (!(Z0 && ts == 0)) {
                    // Horner steps up to this tap's delay, two per trip so the accumulator ping-pongs between
                    // two register sets instead of being copied: acc <- u + w acc
                    int n = C.delay[ts] - pstep;
                    pstep = C.delay[ts];
#pragma unroll 1
                    for (; n >= 2; n -= 2, up -= 2) {
                        const float2 u0 = up[0], u1 = up[-1];
                        f2 gre = fma2(wre, hre, pk(u0.x, u0.x)), gim = fma2(wre, him, pk(u0.y, u0.y));
                        fma2_acc(gre, nwim, him);
                        fma2_acc(gim, wim, hre);
                        hre = fma2(wre, gre, pk(u1.x, u1.x));
                        him = fma2(wre, gim, pk(u1.y, u1.y));
                        fma2_acc(hre, nwim, gim);
                        fma2_acc(him, wim, gre);
                    }
                    if (n) {
                        const float2 u0 = up[0];
                        f2 gre = fma2(wre, hre, pk(u0.x, u0.x)), gim = fma2(wre, him, pk(u0.y, u0.y));
                        fma2_acc(gre, nwim, him);
                        fma2_acc(gim, wim, hre);
                        hre = gre;
                        him = gim;
                        up -= 1;
                    }
                }
                f2 vre, vim, qre, qim;
                if (Z0 && ts == 0) {
                    vre = xre; vim = xim; qre = gre_; qim = gim_;
                } else {
                    vre = fma2(neg2(eim[ts]), xim, mul2(ere[ts], xre));
                    vim = fma2(eim[ts], xre, mul2(ere[ts], xim));
                    qre = fma2(neg2(eim[ts]), gim_, fma2(ere[ts], gre_, mul2(nsq, hre)));
                    qim = fma2(eim[ts], gre_, fma2(ere[ts], gim_, mul2(nsq, him)));
                }
                float vAr, vBr, vAi, vBi, qAr, qBr, qAi, qBi;
                upk(vre, vAr, vBr); upk(vim, vAi, vBi); upk(qre, qAr, qBr); upk(qim, qAi, qBi);
                const float* ct = cf + ts * 6 * R2;
#pragma unroll
                for (int p = 0; p < RP; ++p) {
                    const float2 ar = *(const float2*)(ct + 2 * p), ai = *(const float2*)(ct + R2 + 2 * p);
                    const float2 cr = *(const float2*)(ct + 2 * R2 + 2 * p), ci = *(const float2*)(ct + 3 * R2 + 2 * p);
                    const f2 are = pk(ar.x, ar.y), aim = pk(ai.x, ai.y), cre = pk(cr.x, cr.y), cim = pk(ci.x, ci.y);
                    if (ts == 0) {
                        yreA[p] = mul2(are, pk(vAr, vAr)); yimA[p] = mul2(are, pk(vAi, vAi));
                        yreB[p] = mul2(are, pk(vBr, vBr)); yimB[p] = mul2(are, pk(vBi, vBi));
                    } else {
                        fma2_acc(yreA[p], are, pk(vAr, vAr)); fma2_acc(yimA[p], are, pk(vAi, vAi));
                        fma2_acc(yreB[p], are, pk(vBr, vBr)); fma2_acc(yimB[p], are, pk(vBi, vBi));
                    }
                    fma2_acc(yreA[p], aim, pk(-vAi, -vAi)); fma2_acc(yimA[p], aim, pk(vAr, vAr));
                    fma2_acc(yreB[p], aim, pk(-vBi, -vBi)); fma2_acc(yimB[p], aim, pk(vBr, vBr));
                    fma2_acc(yreA[p], cre, pk(qAr, qAr)); fma2_acc(yimA[p], cre, pk(qAi, qAi));
                    fma2_acc(yreB[p], cre, pk(qBr, qBr)); fma2_acc(yimB[p], cre, pk(qBi, qBi));
                    fma2_acc(yreA[p], cim, pk(-qAi, -qAi)); fma2_acc(yimA[p], cim, pk(qAr, qAr));
                    fma2_acc(yreB[p], cim, pk(-qBi, -qBi)); fma2_acc(yimB[p], cim, pk(qBr, qBr));
                }
            }
            // rows (b R + r) S + s of the output; pilots of a slot's first symbol go to their own tensor
            float2* yrow = Y + (unsigned long long)(b * (unsigned)(R * S) + s) * (unsigned)strideY;
            const unsigned rstride = (unsigned)(S * strideY);
            const bool head = COMPACT && s % LTE_SLOT_SYMBOLS == 0;
            float2* prow = head ? Yp + (unsigned long long)(b * (unsigned)(R * nslot) + s / LTE_SLOT_SYMBOLS) * (unsigned)P.Np : nullptr;
            const unsigned pstride = (unsigned)(nslot * P.Np);
#pragma unroll
            for (int p = 0; p < RP; ++p) {
                const int r0 = 2 * p;
                const bool two = r0 + 1 < R;
                pw[p] = fma2(yreA[p], yreA[p], pw[p]);
                pw[p] = fma2(yimA[p], yimA[p], pw[p]);
                float a, c, d, e;
                upk(yreA[p], a, c);
                upk(yimA[p], d, e);
                if (stA) {
                    yrow[r0 * rstride + offA] = make_float2(a, d);
                    if (two) yrow[(r0 + 1) * rstride + offA] = make_float2(c, e);
                } else if (head && pilA) {
                    prow[r0 * pstride + slotA] = make_float2(a, d);
                    if (two) prow[(r0 + 1) * pstride + slotA] = make_float2(c, e);
                }
                if (okB) {
                    pw[p] = fma2(yreB[p], yreB[p], pw[p]);
                    pw[p] = fma2(yimB[p], yimB[p], pw[p]);
                    upk(yreB[p], a, c);
                    upk(yimB[p], d, e);
                    if (stB) {
                        yrow[r0 * rstride + offB] = make_float2(a, d);
                        if (two) yrow[(r0 + 1) * rstride + offB] = make_float2(c, e);
                    } else if (head && pilB) {
                        prow[r0 * pstride + slotB] = make_float2(a, d);
                        if (two) prow[(r0 + 1) * pstride + slotB] = make_float2(c, e);
                    }
                }
            }
        }

        // ---- end of the stream: its power leaves as one atomic per warp and antenna ----------------
        if (ns == 0) {
#pragma unroll
            for (int p = 0; p < RP; ++p) {
                float a, c;
                upk(pw[p], a, c);
                a = warp_sum(a);
                c = warp_sum(c);
                if ((tid & 31) == 0) {
                    atomicAdd(&power[(size_t)b * R + 2 * p], (double)a);
                    if (2 * p + 1 < R) atomicAdd(&power[(size_t)b * R + 2 * p + 1], (double)c);
                }
                pw[p] = pk(0.f, 0.f);
            }
        }
        b = nb;
        s = ns;
        cur ^= 1;
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

extern "C" int lte_tx_spectral(const lte_plan* p, const uint8_t* idx, lte_c32* G, lte_c32* tail, int32_t B, int32_t S,
                               void* stream) {
    if (!p || !idx || !G || !tail || B < 0 || S < 1) return LTE_ERR_INVALID_ARG;
    if (p->dev.cp < 1) return LTE_ERR_UNSUPPORTED;
    if (B == 0) return LTE_OK;
    const long long total = (long long)B * S;
    if (total >= (1ll << 31) - 1) return LTE_ERR_UNSUPPORTED;
    const int k0 = p->dev.k0_useful, nk = p->dev.nk_useful;
    return dispatch_n(p->dev.N, [&](auto n) -> int {
        constexpr int N = decltype(n)::value;
        const int smem = fft2_cta_smem_bytes(N);
        const long long per = 2 * fft2_pairs_per_cta(N);
        const long long grid = (total + per - 1) / per;
        LTE_CHECK_CUDA(cudaFuncSetAttribute(tx_spectral_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        tx_spectral_kernel<N><<<(unsigned)grid, FFT_CTA_THREADS, smem, (cudaStream_t)stream>>>(
            p->dev, idx, (float2*)G, (float2*)tail, k0, nk, (unsigned)total);
        LTE_CHECK_CUDA(cudaGetLastError());
        return LTE_OK;
    });
}

// Checks shared by the size query and the launcher; fills the sorted tap table.
static int spectral_setup(const lte_plan* p, const lte_channel_desc* ch, int32_t R, SpecParams& C) {
    if (!p || !ch || R < 1 || R > LTE_MAX_RX) return LTE_ERR_INVALID_ARG;
    if (ch->num_taps < 0 || ch->num_taps > LTE_MAX_TAPS) return LTE_ERR_INVALID_ARG;
    if (ch->num_taps == 0 || p->dev.cp < 1) return LTE_ERR_UNSUPPORTED;
    memset(&C, 0, sizeof(C));
    C.num_taps = ch->num_taps;
    for (int i = 0; i < ch->num_taps; ++i) {
        if (ch->delay[i] < 0) return LTE_ERR_INVALID_ARG;
        C.ord[i] = i;
        C.gain[i] = (float)((double)ch->gain[i] * sqrt(2.0 / LTE_JAKES_TONES));
    }
    for (int i = 1; i < ch->num_taps; ++i)                      // insertion sort by delay (stable)
        for (int j = i; j > 0 && ch->delay[C.ord[j - 1]] > ch->delay[C.ord[j]]; --j) {
            const int t = C.ord[j]; C.ord[j] = C.ord[j - 1]; C.ord[j - 1] = t;
        }
    for (int i = 0; i < ch->num_taps; ++i) { C.delay[i] = ch->delay[C.ord[i]]; C.pos[C.ord[i]] = i; }
    C.dmax = C.delay[ch->num_taps - 1];
    // delayed copies must stay inside the symbol's own cyclic prefix; the kernel stages dmax + cp samples
    if (C.dmax > p->dev.cp || C.dmax + p->dev.cp > SPEC_XMAX) return LTE_ERR_UNSUPPORTED;
    if (p->dev.nk_useful > SPEC_GMAX) return LTE_ERR_UNSUPPORTED;
    double wmax = 0.0;
    for (int nn = 0; nn < LTE_JAKES_TONES; ++nn) {
        C.w_cyc[nn] = ch->doppler_hz * cos(2.0 * M_PI * (double)(nn + 1) / LTE_JAKES_TONES) / p->desc.fs;
        if (fabs(C.w_cyc[nn]) > wmax) wmax = fabs(C.w_cyc[nn]);
    }
    // linear (economised) Jakes fit per OFDM symbol: remainder x^2 / 4 <= 5e-7 of |h|
    if (M_PI * wmax * p->dev.L > 1.41e-3) return LTE_ERR_UNSUPPORTED;
    return LTE_OK;
}

extern "C" int64_t lte_channel_spectral_workspace_bytes(const lte_plan* p, const lte_channel_desc* ch, int32_t B,
                                                        int32_t R, int32_t S) {
    SpecParams C;
    const int rc = spectral_setup(p, ch, R, C);
    if (rc) return rc;
    if (B < 0 || S < 1) return LTE_ERR_INVALID_ARG;
    const int R2 = (R + 1) & ~1;
    return (int64_t)sizeof(float) * (int64_t)B * S * ch->num_taps * 6 * R2 + 16;
}

extern "C" int lte_channel_spectral(const lte_plan* p, const lte_channel_desc* ch, const uint8_t* idx, const lte_c32* G,
                                    const lte_c32* tail, const float* phases, lte_c32* Y, lte_c32* Ypilot,
                                    double* power, void* workspace, int32_t B, int32_t R, int32_t S, void* stream) {
    SpecParams C;
    int rc = spectral_setup(p, ch, R, C);
    if (rc) return rc;
    if (((uintptr_t)workspace & 15) || ((uintptr_t)G & 7) || ((uintptr_t)tail & 7)) return LTE_ERR_INVALID_ARG;
    if (!idx || !G || !tail || !phases || !Y || !power || !workspace || B < 0 || S < 1) return LTE_ERR_INVALID_ARG;
    if (Ypilot && p->dev.Np == 0) return LTE_ERR_INVALID_ARG;
    if (B == 0) return LTE_OK;
    const long long total = (long long)B * S;
    if (total >= (1ll << 31) - 1) return LTE_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int R2 = (R + 1) & ~1;
    float* coef = (float*)workspace;
    if (R2 != R)
        LTE_CHECK_CUDA(cudaMemsetAsync(coef, 0, sizeof(float) * (size_t)total * ch->num_taps * 6 * R2, st));
    const long long items = total * R * ch->num_taps;
    spectral_coef_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(C, phases, coef, R, R2, S, p->dev.L, items);
    LTE_CHECK_CUDA(cudaGetLastError());
    const int k0 = p->dev.k0_useful, nk = p->dev.nk_useful;
    const int half = (nk + 1) / 2;
    int need = half > C.dmax + p->dev.cp ? half : C.dmax + p->dev.cp;
    if (need < 64) need = 64;
    const int threads = (need + 31) & ~31;
    if (threads > SPEC_MAX_THREADS || C.dmax + p->dev.cp > SPEC_XMAX) return LTE_ERR_UNSUPPORTED;
    const bool z0 = C.delay[0] == 0;
    auto launch = [&](auto k) -> int {
        int per_sm = 1, dev = p->device, sms = 148;
        LTE_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, 0));
        LTE_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        long long grid = (long long)sms * (per_sm < 1 ? 1 : per_sm);
        if (grid > B) grid = B;
        k<<<(unsigned)grid, threads, 0, st>>>(p->dev, C, idx, (const float2*)G, (const float2*)tail, coef, (float2*)Y,
                                             (float2*)Ypilot, power, k0, nk, half, S, R, B);
        LTE_CHECK_CUDA(cudaGetLastError());
        return LTE_OK;
    };
#define LAUNCH_SPEC_RP(NT, RP)                                                                                   \
    (Ypilot ? (z0 ? launch(channel_spectral_kernel<NT, RP, true, true>) : launch(channel_spectral_kernel<NT, RP, true, false>)) \
            : (z0 ? launch(channel_spectral_kernel<NT, RP, false, true>) : launch(channel_spectral_kernel<NT, RP, false, false>)))
#define LAUNCH_SPEC(NT)                                                     \
    case NT:                                                                \
        return R2 == 2 ? LAUNCH_SPEC_RP(NT, 1) : R2 == 4 ? LAUNCH_SPEC_RP(NT, 2) \
             : R2 == 6 ? LAUNCH_SPEC_RP(NT, 3) : LAUNCH_SPEC_RP(NT, 4);
    switch (ch->num_taps) {
#ifdef SPEC_DEV                     // development builds: the headline shape only (seconds instead of a minute)
        case 4: return LAUNCH_SPEC_RP(4, 2);
#else
        LAUNCH_SPEC(1) LAUNCH_SPEC(2) LAUNCH_SPEC(3) LAUNCH_SPEC(4) LAUNCH_SPEC(5) LAUNCH_SPEC(6) LAUNCH_SPEC(7)
        LAUNCH_SPEC(8)
#endif
        default: return LTE_ERR_INVALID_ARG;
    }
#undef LAUNCH_SPEC
#undef LAUNCH_SPEC_RP
}
