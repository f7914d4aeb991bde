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
// STD: the plan uses the LTE profile's occupied-bin count for this FFT size (config.py:104-107: 76 / 150 / 300 /
// 600 / 1200 of 128 / 256 / 512 / 1024 / 2048), so the occupied window is a compile-time constant and every
// "is this group of 128 bins inside the window" test folds away; other layouts take the run-time tests.
__host__ __device__ constexpr int lte_std_nc(int n) { return n == 2048 ? 1200 : n == 1024 ? 600 : n == 512 ? 300 : n == 256 ? 150 : n == 128 ? 76 : 0; }

template <int N, bool STD>
__global__ void __launch_bounds__(FFT_CTA_THREADS, 5)
tx_spectral_kernel(const DevPlan P, const uint8_t* __restrict__ idx, float2* __restrict__ G,
                   float2* __restrict__ tail, int k0_rt, int nk_rt, unsigned total) {
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

    const int k0 = STD ? (N - lte_std_nc(N)) / 2 : k0_rt, nk = STD ? lte_std_nc(N) : nk_rt;
    const int used_lo = k0, used_hi = k0 + nk;
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

    // tails (tx samples = raw / sqrt(N)) and the ramp (n - n_c) / N: G = fft(ramp * raw) / ... / sqrt(N).
    // (One copy of the transform code looped twice -- forward as conj(IFFT(conj(.))) -- halves the instruction
    // footprint but the loop-carried state spills at 96 registers: 0.58 ms against 0.47 ms, measured.)
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

#define SPEC_MAX_WARPS 19           // warps per CTA: ceil(600 bin pairs / 32) at 20 MHz
#define SPEC_DMAX 160               // longest supported delay spread in samples

// Linear Jakes fit per (OFDM symbol, antenna, tap) in the layout the channel kernel stages verbatim:
//   coef[f] = { [sorted tap][R2] (a.re, a.im, c1.re, c1.im) | [sorted tap][R2] (c0.re, c0.im) },  a = c0 + d c1
// (one 128-bit shared-memory load per (tap, antenna) in the bin loop; c0 only enters the CP samples)
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
    float* cf = coef + (size_t)f * C.num_taps * 6 * R2;
    *(float4*)(cf + ((size_t)ts * R2 + r) * 4) = make_float4(fmaf(d, a1.x, a0.x), fmaf(d, a1.y, a0.y), a1.x, a1.y);
    *(float2*)(cf + (size_t)C.num_taps * R2 * 4 + ((size_t)ts * R2 + r) * 2) = a0;
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
// base + off * scale as ONE instruction (IMAD.WIDE.U32): all per-symbol addresses are a uniform 64-bit
// base plus a per-lane 32-bit element offset
__device__ __forceinline__ unsigned long long gaddr(unsigned long long base, unsigned off, unsigned scale) {
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(off), "r"(scale), "l"(base));
    return r;
}
__device__ __forceinline__ void stg64(unsigned long long a, float x, float y) {
    asm volatile("st.global.v2.f32 [%0], {%1, %2};" ::"l"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ float2 ldg64(unsigned long long a) {
    float2 v;
    asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(a));
    return v;
}
__device__ __forceinline__ unsigned ldgu8(unsigned long long a) {
    unsigned v;
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(a));
    return v;
}
__device__ __forceinline__ float2 lds64(unsigned a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ float4 lds128(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}

__device__ __forceinline__ void stg128(unsigned long long a, float x, float y, float z, float w) {
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ float ldg32(unsigned long long a) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(a));
    return v;
}
// scalar (both lanes) operand of a packed multiply-add: SASS takes it as a .F32 broadcast, no register pair
__device__ __forceinline__ f2 bc(float x) { return pk(x, x); }

// ---- mbarrier / bulk-copy (TMA engine) primitives ------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// one contiguous run of global memory -> shared memory by the bulk-copy engine; completion is signalled on `bar`
__device__ __forceinline__ void bulk_g2s(unsigned dst, unsigned long long src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ float lds32(unsigned a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}

// both halves of a packed register pair straight from two shared-memory words
__device__ __forceinline__ f2 lds32x2(unsigned a, unsigned b) {
    f2 r;
    asm volatile("{\n.reg .f32 lo, hi;\nld.shared.f32 lo, [%1];\nld.shared.f32 hi, [%2];\nmov.b64 %0, {lo, hi};\n}"
                 : "=l"(r.v) : "r"(a), "r"(b));
    return r;
}
// The same with loads ptxas may not merge: two neighbouring bins' (re, im) would otherwise arrive as two 64-bit
// loads and the (re, re) / (im, im) pairs be rebuilt with register moves wherever they are used.
__device__ __forceinline__ f2 lds32x2v(unsigned a, unsigned b) {
    f2 r;
    asm volatile("{\n.reg .f32 lo, hi;\nld.volatile.shared.f32 lo, [%1];\nld.volatile.shared.f32 hi, [%2];\nmov.b64 %0, {lo, hi};\n}"
                 : "=l"(r.v) : "r"(a), "r"(b));
    return r;
}
// Horner steps acc <- u + w acc on the packed accumulator (hre, him), in place: two steps / one step.
// One asm statement each, so the loop-carried accumulator keeps its register pair (no copies at the back edge).
__device__ __forceinline__ void horner2(f2& hre, f2& him, f2 wre, f2 wim, f2 nwim, float2 u0, float2 u1) {
    asm("{\n.reg .b64 tr, ti, a, b, c, d;\n"
        "mov.b64 a, {%5, %5};\nmov.b64 b, {%6, %6};\nmov.b64 c, {%7, %7};\nmov.b64 d, {%8, %8};\n"
        "fma.rn.f32x2 tr, %2, %0, a;\nfma.rn.f32x2 ti, %2, %1, b;\n"
        "fma.rn.f32x2 tr, %4, %1, tr;\nfma.rn.f32x2 ti, %3, %0, ti;\n"
        "fma.rn.f32x2 %0, %2, tr, c;\nfma.rn.f32x2 %1, %2, ti, d;\n"
        "fma.rn.f32x2 %0, %4, ti, %0;\nfma.rn.f32x2 %1, %3, tr, %1;\n}"
        : "+l"(hre.v), "+l"(him.v)
        : "l"(wre.v), "l"(wim.v), "l"(nwim.v), "f"(u0.x), "f"(u0.y), "f"(u1.x), "f"(u1.y));
}
__device__ __forceinline__ void horner1(f2& hre, f2& him, f2 wre, f2 wim, f2 nwim, float2 u0) {
    asm("{\n.reg .b64 tr, a, b;\n"
        "mov.b64 a, {%5, %5};\nmov.b64 b, {%6, %6};\n"
        "fma.rn.f32x2 tr, %2, %0, a;\n"
        "fma.rn.f32x2 tr, %4, %1, tr;\n"
        "fma.rn.f32x2 %1, %2, %1, b;\n"
        "fma.rn.f32x2 %1, %3, %0, %1;\n"
        "mov.b64 %0, tr;\n}"
        : "+l"(hre.v), "+l"(him.v)
        : "l"(wre.v), "l"(wim.v), "l"(nwim.v), "f"(u0.x), "f"(u0.y));
}

#define SPEC_STAGES 3

// Tap delays known at compile time.  The Horner sweep over the symbol tail has one segment per tap whose length is
// the difference of two tap delays; with run-time delays that is a loop per segment (loop control on the uniform
// datapath, a register rotation at every back edge, loads that cannot leave their trip).  For the delay sets of the
// reference's ITU Pedestrian_A profile at the LTE sample rates (config.py:34-60 x config.py:104-107, rounded as
// core/rayleighchannel.py:36 does) the kernel is instantiated with the delays as constants: the sweep is straight-line
// code and every shared-memory offset is an immediate.  DK = 0 keeps the run-time loops for everything else.
// (1.92 MHz has an odd cyclic prefix, which the bulk-copy staging does not take: no entry.)
#define SPEC_NUM_DK 4
__host__ __device__ constexpr int spec_ct_delay(int dk, int ts) {
    return dk == 1 ? (ts == 1 ? 3 : ts == 2 ? 6 : ts == 3 ? 13 : 0)      // 30.72 MHz: 0 / 110 / 190 / 410 ns
         : dk == 2 ? (ts == 1 ? 2 : ts == 2 ? 3 : ts == 3 ? 6 : 0)       // 15.36 MHz
         : dk == 3 ? (ts == 1 ? 1 : ts == 2 ? 1 : ts == 3 ? 3 : 0)       //  7.68 MHz
         : dk == 4 ? (ts == 1 ? 0 : ts == 2 ? 1 : ts == 3 ? 2 : 0)       //  3.84 MHz
         : 0;
}

// Persistent, warp-specialised kernel.  A CTA walks whole streams (b = blockIdx.x, += gridDim.x) symbol by
// symbol.  ONE producer warp feeds a 3-stage shared-memory ring with the bulk-copy (TMA) engine: per OFDM
// symbol three contiguous runs -- the symbol's G window, the dmax + cp time samples around its cyclic prefix,
// its Jakes coefficients -- each one cp.async.bulk that completes on the stage's `full` mbarrier; compute
// warps hand a stage back through its `empty` mbarrier.  No thread ever spends an instruction on staging and
// there is no CTA-wide barrier in the loop.
// Lane l of compute warp w owns the SAME bin pair 32 w + l of the plan's pair table (two consecutive data
// symbols, or two consecutive pilots, plus the window's null bins) for its whole life, so everything that
// depends on the bin only -- class, pilot value, exp(-2 pi j k / N) and its powers e_t[k] -- sits in
// registers.  ALL arithmetic on the bins runs packed over the pair (f32x2 lanes = the two bins): the Horner
// sweep over the symbol tail, V_t = e_t X, W_t = e_t G - Q_t, and the per-antenna combine, whose
// coefficients enter as scalar-broadcast operands straight from 128-bit shared-memory loads -- no packing
// or unpacking instruction on the path.  The cyclic-prefix samples (time domain, power only) are the producer
// warp's second job, one antenna per lane.  Stream power stays in registers across the stream's symbols and
// leaves as one atomic per warp and antenna.
// DK > 0: the tap delays are the compile-time set spec_ct_delay(DK, .), see there.
// PLANAR flag = the sweep's compact layout: a data pair leaves as ONE 128-bit store of two consecutive complex
// values into Y [B*R][S][2 ndp]; pilot pairs only on every slot's first symbol, into Yp [B*R][slots][2 npp];
// nothing else is ever read downstream.  Otherwise Y is the windowed grid [B*R][S][nk].
template <int NT, int R2, bool PLANAR, bool Z0, int DK>
__global__ void __launch_bounds__(32 * (SPEC_MAX_WARPS + 1), 1)
channel_spectral_kernel(const DevPlan P, const SpecParams C, const uint8_t* __restrict__ idx,
                        const float2* __restrict__ G, const float2* __restrict__ tail,
                        const float* __restrict__ coef_g, float* __restrict__ Y, float* __restrict__ Yp,
                        double* __restrict__ power, int k0, int nk, int S, int R, int B, int nwarps, int dmax2) {
    constexpr int NCF = NT * 6 * R2;                            // coefficient floats per symbol
    constexpr bool CT = DK > 0;                                 // tap delays are compile-time constants
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int cp = P.cp, L = P.L;
    const int nslot = (S + LTE_SLOT_SYMBOLS - 1) / LTE_SLOT_SYMBOLS;
    // shared memory: [mbarriers full[], empty[]] [X lookup: re[64], im[64]] stages x {G row | x region | coefficients}
    const unsigned s0 = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned bar_full = s0, bar_empty = s0 + 8u * SPEC_STAGES;
    const unsigned lut = s0 + 64u;
    const unsigned gbytes = (unsigned)nk * 8u, xbytes = (unsigned)(dmax2 + cp) * 8u, cbytes = (unsigned)(NCF * sizeof(float));
    const unsigned stage_bytes = gbytes + xbytes + cbytes;
    const unsigned stage0 = s0 + 64u + 512u;
    {
        const int hb0 = P.bps >> 1, mask0 = (1 << hb0) - 1;
        float* l = (float*)(smem_raw + 64);
        for (int v = threadIdx.x; v < 64; v += blockDim.x) {
            l[v] = P.lev[(v >> hb0) & mask0];
            l[64 + v] = P.lev[v & mask0];
        }
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int st = 0; st < SPEC_STAGES; ++st) {
            mbar_init(bar_full + 8u * st, 1u);
            mbar_init(bar_empty + 8u * st, (unsigned)nwarps + 1u);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const unsigned xo = gbytes + 8u * (unsigned)dmax2;         // local sample 0 (first CP sample) of the symbol
    const unsigned co = gbytes + xbytes;                        // coefficients: [ts][R2] x 16 B, then c0 [ts][R2] x 8 B
    constexpr unsigned C0OFF = (unsigned)(NT * R2 * 16);

    // ================================ producer + cyclic-prefix warp ============================
    // Lane 0 keeps the ring two symbols ahead: at the top of symbol i it waits for the stage of symbol i - 1 to
    // be handed back by every warp and refills it with symbol i + 2.  Then the whole warp evaluates the CP samples
    // of symbol i: they only enter the stream power (core/channel.py:216-218 measures the whole faded stream) but
    // carry the inter-symbol leakage, so they are done in the time domain, y_r[i] = sum_t (c0 + c1 tau_i) x[i - d_t],
    // packed over sample pairs (i, i + 1) with the coefficients as scalar-broadcast operands like in the bin loop.
    // A lane owns ONE antenna (lane % R2), keeps that antenna's coefficients in registers for the symbol and walks
    // the sample pairs lane / R2 + k (32 / R2): short rounds of 26 packed operations without idle lanes, three of
    // them in flight.  This single warp is what the 19 bin warps end up waiting for when its rounds are long and
    // serial (32 sample pairs x all antennas per lane, the third trip with 8 live lanes: 0.12 ms of the kernel's
    // 0.90 ms at the headline shape, measured by leaving the CP samples out; 0.06 ms in this form).
    if (w == nwarps) {
        auto produce = [&](unsigned f, unsigned st) {
            const unsigned dst = stage0 + st * stage_bytes, bar = bar_full + 8u * st;
            // samples [f cp - dmax2, f cp + cp) of the tail array; the very first symbol has nothing before it
            const unsigned skip = f == 0 ? (unsigned)dmax2 * 8u : 0u;
            mbar_expect_tx(bar, stage_bytes - skip);
            bulk_g2s(dst, gaddr((unsigned long long)G, f, gbytes), gbytes, bar);
            bulk_g2s(dst + gbytes + skip,
                     gaddr((unsigned long long)tail, f, (unsigned)cp * 8u) - (unsigned long long)dmax2 * 8ull + skip,
                     xbytes - skip, bar);
            bulk_g2s(dst + gbytes + xbytes, gaddr((unsigned long long)coef_g, f, cbytes), cbytes, bar);
        };
        // the CTA's symbol sequence: (b, s), b = blockIdx.x + k gridDim.x; `pb, ps` run two symbols ahead of `b, s`
        unsigned pb = blockIdx.x, ps = 0, pstage = 0, pphase = 0;
        auto produce_next = [&]() {
            if (pb < (unsigned)B) {
                if (lane == 0) {
                    mbar_wait(bar_empty + 8u * pstage, pphase ^ 1u);
                    produce(pb * (unsigned)S + ps, pstage);
                }
                if (++ps == (unsigned)S) { ps = 0; pb += gridDim.x; }
                if (++pstage == SPEC_STAGES) { pstage = 0; pphase ^= 1u; }
            }
        };
#pragma unroll
        for (int i = 0; i < SPEC_STAGES - 1; ++i) produce_next();
        constexpr int PPR = 32 / R2;                            // sample pairs per round
        constexpr int CPU = 3;                                  // rounds in flight
        const int ra = lane % R2, pl = lane / R2;               // the lane's antenna and its place in a round
        const int npair = pl < PPR ? cp / 2 : 0;                // 32 % R2 lanes stay idle (R2 = 6)
        const float tau0 = -0.5f * (float)(L - 1);
        f2 pwc = pk(0.f, 0.f);
        unsigned stage = 0, phase = 0;
        for (unsigned b = blockIdx.x; b < (unsigned)B; b += gridDim.x) {
            for (unsigned s = 0; s < (unsigned)S; ++s) {
                produce_next();
                __syncwarp();
                mbar_wait(bar_full + 8u * stage, phase);
                const unsigned sb = stage0 + stage * stage_bytes;
                if (s == 0) {
                    // nothing precedes the stream: the dmax2 samples in front of its first symbol (the previous
                    // stream's last ones) count as zeros.  Only this warp reads them.
                    for (int i = lane; i < dmax2; i += 32)
                        asm volatile("st.shared.v2.f32 [%0], {%1, %1};" ::"r"(sb + gbytes + 8u * (unsigned)i), "f"(0.f) : "memory");
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // the stage is refilled by the bulk-copy engine
                    __syncwarp();
                }
                // the lane's antenna: (c0, c1) per tap, once per symbol
                float2 k0[NT], k1[NT];
#pragma unroll
                for (int ts = 0; ts < NT; ++ts) {
                    const float4 ac = lds128(sb + co + 16u * (unsigned)ra + (unsigned)(ts * R2 * 16));   // (a.re, a.im, c1.re, c1.im)
                    k0[ts] = lds64(sb + co + C0OFF + 8u * (unsigned)ra + (unsigned)(ts * R2 * 8));
                    k1[ts] = make_float2(ac.z, ac.w);
                }
                // a round past the end works on the lane's first pair and is left out of the sum
                for (int pp = pl; pp < npair; pp += CPU * PPR) {
                    unsigned xa[CPU];
                    f2 tau[CPU], yre[CPU], yim[CPU];
                    bool ok[CPU];
#pragma unroll
                    for (int u = 0; u < CPU; ++u) {
                        ok[u] = pp + u * PPR < npair;
                        const int i0 = 2 * (ok[u] ? pp + u * PPR : pl);
                        const float tf = (float)i0 + tau0;
                        tau[u] = pk(tf, tf + 1.f);
                        xa[u] = sb + xo + 8u * (unsigned)i0;
                    }
#pragma unroll
                    for (int ts = 0; ts < NT; ++ts) {
                        const int d = CT ? spec_ct_delay(DK, ts) : C.delay[ts];
#pragma unroll
                        for (int u = 0; u < CPU; ++u) {
                            const unsigned x0 = xa[u] - 8u * (unsigned)d;
                            const f2 xr = lds32x2v(x0, x0 + 8u), xi = lds32x2v(x0 + 4u, x0 + 12u);
                            const f2 hr = fma2(tau[u], bc(k1[ts].x), bc(k0[ts].x)), hi = fma2(tau[u], bc(k1[ts].y), bc(k0[ts].y));
                            if (ts == 0) {
                                yre[u] = mul2(hr, xr);
                                yim[u] = mul2(hr, xi);
                            } else {
                                fma2_acc(yre[u], hr, xr);
                                fma2_acc(yim[u], hr, xi);
                            }
                            fma2_acc(yre[u], neg2(hi), xi);
                            fma2_acc(yim[u], hi, xr);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < CPU; ++u) {
                        const f2 m = pk(ok[u] ? 1.f : 0.f, ok[u] ? 1.f : 0.f);
                        yre[u] = mul2(yre[u], m);
                        yim[u] = mul2(yim[u], m);
                        fma2_acc(pwc, yre[u], yre[u]);
                        fma2_acc(pwc, yim[u], yim[u]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_empty + 8u * stage);
                if (++stage == SPEC_STAGES) { stage = 0; phase ^= 1u; }
            }
            // the stream's CP power per antenna: sum over the lanes of the same antenna (ra + k R2)
            float a, c;
            upk(pwc, a, c);
            const float mine = a + c;
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < (32 + R2 - 1) / R2; ++k) {
                const int src = ra + k * R2;
                const float o = __shfl_sync(0xffffffffu, mine, src & 31);
                if (src < 32) t += o;
            }
            if (lane < R) atomicAdd(&power[(size_t)b * R + lane], (double)t);
            pwc = pk(0.f, 0.f);
        }
        return;
    }

    // ================================ bin warps ==================================================
    // ---- per-lane constants of its bin pair -------------------------------------------------------
    const int pi = 32 * w + lane;
    const bool live = pi < P.npairs;
    const int binA = live ? (int)P.pair_bin[2 * pi] : -1, binB = live ? (int)P.pair_bin[2 * pi + 1] : -1;
    const bool okA = binA >= 0, okB = binB >= 0;
    const bool isdat = pi < P.ndp, ispil = !isdat && pi < P.ndp + P.npp;
    const int codeA = okA ? (int)__ldg(&P.bin_map[binA]) : BIN_NULL, codeB = okB ? (int)__ldg(&P.bin_map[binB]) : BIN_NULL;
    const bool datA = isdat && codeA >= 0, datB = isdat && codeB >= 0;      // a data pair may end in a null pad
    const unsigned slotA = (unsigned)(codeA & (BIN_PILOT_FLAG - 1)), slotB = (unsigned)(codeB & (BIN_PILOT_FLAG - 1));
    float2 pvA = make_float2(0.f, 0.f), pvB = pvA;
    if (ispil && codeA >= 0) pvA = P.pilots[slotA];
    if (ispil && codeB >= 0) pvB = P.pilots[slotB];
    const unsigned kkA = (unsigned)((okA ? binA : k0) - k0), kkB = (unsigned)((okB ? binB : k0) - k0);
    f2 wre, wim, nwim, ere[NT], eim[NT];
    {
        const float2 z = make_float2(0.f, 0.f);
        const float2 a = okA ? __ldg(&P.twiddle[binA]) : z, c = okB ? __ldg(&P.twiddle[binB]) : z;
        wre = pk(a.x, c.x);
        wim = pk(a.y, c.y);
        nwim = pk(-a.y, -c.y);
#pragma unroll
        for (int ts = 0; ts < NT; ++ts) {
            const int d = CT ? spec_ct_delay(DK, ts) : C.delay[ts];
            const float2 ea = okA ? __ldg(&P.twiddle[(binA * d) & (P.N - 1)]) : z;
            const float2 ec = okB ? __ldg(&P.twiddle[(binB * d) & (P.N - 1)]) : z;
            ere[ts] = pk(ea.x, ec.x);
            eim[ts] = pk(ea.y, ec.y);
        }
    }
    const f2 vmask = pk(okA ? 1.f : 0.f, okB ? 1.f : 0.f);      // a missing bin contributes nothing
    const bool partial = live && !(okA && okB);
    const float sqn = sqrtf((float)P.N);
    const f2 nsq = pk(-sqn, -sqn);
    const unsigned gA = 8u * kkA, gB = 8u * kkB;                // byte offsets of the pair's G values inside a stage
    const unsigned long long iA = (unsigned long long)idx + slotA, iB = (unsigned long long)idx + slotB;

    f2 pw[R2];                                                  // lanes: the two bins
#pragma unroll
    for (int r = 0; r < R2; ++r) pw[r] = pk(0.f, 0.f);

    unsigned stage = 0, phase = 0;
    unsigned b = blockIdx.x, s = 0;
    unsigned sl = 0, slot = 0;                                  // s % LTE_SLOT_SYMBOLS, s / LTE_SLOT_SYMBOLS
    unsigned ibA = 0, ibB = 0;                                  // index bytes of the NEXT symbol to compute
    if (b < (unsigned)B) {
        const unsigned f = b * (unsigned)S;
        if (datA) ibA = ldgu8(gaddr(iA, f, (unsigned)P.Nd));
        if (datB) ibB = ldgu8(gaddr(iB, f, (unsigned)P.Nd));
    }
    while (b < (unsigned)B) {
        // ---- transmitted grid values of the pair from the symbol's index bytes; next symbol's bytes in flight -----
        f2 xre = pk(pvA.x, pvB.x), xim = pk(pvA.y, pvB.y);
        if (isdat) {
            xre = lds32x2(lut + 4u * ibA, lut + 4u * ibB);
            xim = lds32x2(lut + 256u + 4u * ibA, lut + 256u + 4u * ibB);
            if (!datB) { xre = mul2(xre, pk(1.f, 0.f)); xim = mul2(xim, pk(1.f, 0.f)); }   // null pad of an odd data count
        }
        unsigned nb = b, ns = s + 1;
        if (ns == (unsigned)S) { ns = 0; nb = b + gridDim.x; }
        if (nb < (unsigned)B) {
            const unsigned f = nb * (unsigned)S + ns;
            if (datA) ibA = ldgu8(gaddr(iA, f, (unsigned)P.Nd));
            if (datB) ibB = ldgu8(gaddr(iB, f, (unsigned)P.Nd));
        }
        mbar_wait(bar_full + 8u * stage, phase);                // the symbol's bytes have landed
        const unsigned sb = stage0 + stage * stage_bytes;
        const f2 gre = lds32x2v(sb + gA, sb + gB), gim = lds32x2v(sb + gA + 4u, sb + gB + 4u);

        f2 hre = pk(0.f, 0.f), him = pk(0.f, 0.f);              // Horner accumulator
        f2 yre[R2], yim[R2];
        unsigned up = sb + xo + 8u * (unsigned)(cp - 1);        // u[N - 1 - p] at up - 8 p
        int pstep = 0;
#pragma unroll
        for (int ts = 0; ts < NT; ++ts) {
            if (CT) {
                // Horner steps up to this tap's delay as straight-line code: acc <- u + w acc
                const int dcur = spec_ct_delay(DK, ts);
#pragma unroll
                for (int q = pstep; q < dcur; ++q) {
                    const float2 u = lds64(up - 8u * (unsigned)q);
                    if (q == 0) {
                        hre = bc(u.x);
                        him = bc(u.y);
                    } else {
                        const f2 tr = fma2(nwim, him, fma2(wre, hre, bc(u.x)));
                        him = fma2(wim, hre, fma2(wre, him, bc(u.y)));
                        hre = tr;
                    }
                }
                pstep = dcur;
            } else if (!(Z0 && ts == 0)) {
                // Horner steps up to this tap's delay, two per trip: acc <- u + w acc
                int n = C.delay[ts] - pstep;
                pstep = C.delay[ts];
#pragma unroll 1
                for (; n >= 2; n -= 2, up -= 16u) horner2(hre, him, wre, wim, nwim, lds64(up), lds64(up - 8u));
                if (n) {
                    horner1(hre, him, wre, wim, nwim, lds64(up));
                    up -= 8u;
                }
            }
            f2 vre, vim, qre, qim;
            if (Z0 && ts == 0) {
                vre = xre; vim = xim; qre = gre; qim = gim;
            } else {
                vre = fma2(neg2(eim[ts]), xim, mul2(ere[ts], xre));
                vim = fma2(eim[ts], xre, mul2(ere[ts], xim));
                qre = fma2(neg2(eim[ts]), gim, fma2(ere[ts], gre, mul2(nsq, hre)));
                qim = fma2(eim[ts], gre, fma2(ere[ts], gim, mul2(nsq, him)));
            }
            const unsigned ct = sb + co + (unsigned)(ts * R2 * 16);
#pragma unroll
            for (int r = 0; r < R2; ++r) {
                const float4 ac = lds128(ct + 16u * r);         // (a.re, a.im, c1.re, c1.im) of antenna r
                if (ts == 0) {
                    yre[r] = mul2(vre, bc(ac.x));
                    yim[r] = mul2(vim, bc(ac.x));
                } else {
                    fma2_acc(yre[r], vre, bc(ac.x));
                    fma2_acc(yim[r], vim, bc(ac.x));
                }
                fma2_acc(yre[r], vim, bc(-ac.y)); fma2_acc(yim[r], vre, bc(ac.y));
                fma2_acc(yre[r], qre, bc(ac.z));  fma2_acc(yim[r], qim, bc(ac.z));
                fma2_acc(yre[r], qim, bc(-ac.w)); fma2_acc(yim[r], qre, bc(ac.w));
            }
        }
        // every shared-memory read of this stage is done: hand it back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8u * stage);
        if (++stage == SPEC_STAGES) { stage = 0; phase ^= 1u; }

        if (partial) {
#pragma unroll
            for (int r = 0; r < R2; ++r) { yre[r] = mul2(yre[r], vmask); yim[r] = mul2(yim[r], vmask); }
        }

        // ---- power and stores ---------------------------------------------------------------------
        if (live) {
            const unsigned row0 = b * (unsigned)(R * S) + s;        // row of antenna 0; antenna r is S rows further
            if (PLANAR) {
                // one destination per lane: a data pair goes to its place in row (b, antenna 0, s) of Y, a pilot pair --
                // on a slot's first symbol only -- to row (b, antenna 0, slot) of Yp; antenna r is `rstride` further
                const bool head = ispil && sl == 0;
                const bool dost = isdat || head;
                const unsigned long long dst =
                    isdat ? gaddr((unsigned long long)Y, row0 * (unsigned)P.ndp + (unsigned)pi, 16u)
                          : gaddr((unsigned long long)Yp, (b * (unsigned)(R * nslot) + slot) * (unsigned)P.npp + (unsigned)(pi - P.ndp), 16u);
                const unsigned rstride = isdat ? (unsigned)S * (unsigned)P.ndp : (unsigned)nslot * (unsigned)P.npp;
#pragma unroll
                for (int r = 0; r < R2; ++r) {
                    fma2_acc(pw[r], yre[r], yre[r]);
                    fma2_acc(pw[r], yim[r], yim[r]);
                    if ((CT || r < R) && dost) {
                        float a, c, d, e;
                        upk(yre[r], a, c);
                        upk(yim[r], d, e);
                        stg128(gaddr(dst, (unsigned)r * rstride, 16u), a, d, c, e);
                    }
                }
            } else {
                const unsigned long long yb = gaddr((unsigned long long)Y, row0, (unsigned)nk * 8u);
#pragma unroll
                for (int r = 0; r < R2; ++r) {
                    fma2_acc(pw[r], yre[r], yre[r]);
                    fma2_acc(pw[r], yim[r], yim[r]);
                    if (r < R) {
                        float a, c, d, e;
                        upk(yre[r], a, c);
                        upk(yim[r], d, e);
                        if (okA) stg64(gaddr(yb, kkA + (unsigned)(r * S) * (unsigned)nk, 8u), a, d);
                        if (okB) stg64(gaddr(yb, kkB + (unsigned)(r * S) * (unsigned)nk, 8u), c, e);
                    }
                }
            }
        }

        // ---- end of the stream: its power leaves as one atomic per warp and antenna ----------------
        if (ns == 0) {
#pragma unroll
            for (int r = 0; r < R2; ++r) {
                float a, c;
                upk(pw[r], a, c);
                const float t = warp_sum(a + c);
                if (lane == 0 && r < R) atomicAdd(&power[(size_t)b * R + r], (double)t);
                pw[r] = pk(0.f, 0.f);
            }
        }
        b = nb;
        s = ns;
        if (++sl == LTE_SLOT_SYMBOLS) { sl = 0; ++slot; }
        if (ns == 0) { sl = 0; slot = 0; }
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
        const bool std_layout = !p->desc.mode_simple && nk == lte_std_nc(N) && k0 == (N - nk) / 2;
        auto launch = [&](auto k) -> int {
            LTE_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            k<<<(unsigned)grid, FFT_CTA_THREADS, smem, (cudaStream_t)stream>>>(p->dev, idx, (float2*)G, (float2*)tail, k0, nk,
                                                                           (unsigned)total);
            LTE_CHECK_CUDA(cudaGetLastError());
            return LTE_OK;
        };
        return std_layout ? launch(tx_spectral_kernel<N, true>) : launch(tx_spectral_kernel<N, false>);
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
    // delayed copies must stay inside the symbol's own cyclic prefix
    if (C.dmax > p->dev.cp || C.dmax > SPEC_DMAX) return LTE_ERR_UNSUPPORTED;
    if (p->dev.npairs > 32 * SPEC_MAX_WARPS) return LTE_ERR_UNSUPPORTED;
    // the bulk-copy engine moves 16-byte aligned runs: rows of G (nk) and of the tail array (cp) must be even
    if ((p->dev.cp & 1) || (p->dev.nk_useful & 1) || p->dev.bps > 6) return LTE_ERR_UNSUPPORTED;
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
    if (Ypilot && ((((uintptr_t)Y) | ((uintptr_t)Ypilot)) & 15)) return LTE_ERR_INVALID_ARG;
    if (!idx || !G || !tail || !phases || !Y || !power || !workspace || B < 0 || S < 1) return LTE_ERR_INVALID_ARG;
    if (Ypilot && p->dev.Np == 0) return LTE_ERR_INVALID_ARG;
    if (B == 0) return LTE_OK;
    const long long total = (long long)B * S;
    if (total * R >= (1ll << 31) - 1) return LTE_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int R2 = (R + 1) & ~1;
    float* coef = (float*)workspace;
    if (R2 != R)
        LTE_CHECK_CUDA(cudaMemsetAsync(coef, 0, sizeof(float) * (size_t)total * ch->num_taps * 6 * R2, st));
    const long long items = total * R * ch->num_taps;
    spectral_coef_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(C, phases, coef, R, R2, S, p->dev.L, items);
    LTE_CHECK_CUDA(cudaGetLastError());
    const int k0 = p->dev.k0_useful, nk = p->dev.nk_useful;
    const int nwarps = (p->dev.npairs + 31) / 32;               // bin warps; one more warp feeds the ring and does the CP samples
    if (nwarps > SPEC_MAX_WARPS) return LTE_ERR_UNSUPPORTED;
    const int threads = 32 * (nwarps + 1);
    const int NCF = ch->num_taps * 6 * R2;
    const int dmax2 = (C.dmax + 1) & ~1;
    const size_t stage = (size_t)nk * 8 + (size_t)(dmax2 + p->dev.cp) * 8 + (size_t)NCF * 4;
    const size_t smem = 64 + 512 + SPEC_STAGES * stage;
    if (smem > 200 * 1024) return LTE_ERR_UNSUPPORTED;
    const bool z0 = C.delay[0] == 0;
    auto launch = [&](auto k) -> int {
        int per_sm = 1, sms = 148;
        LTE_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LTE_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, smem));
        LTE_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device));
        long long grid = (long long)sms * (per_sm < 1 ? 1 : per_sm);
        if (grid > B) grid = B;
        k<<<(unsigned)grid, threads, smem, st>>>(p->dev, C, idx, (const float2*)G, (const float2*)tail, coef, (float*)Y,
                                                 (float*)Ypilot, power, k0, nk, S, R, B, nwarps, dmax2);
        LTE_CHECK_CUDA(cudaGetLastError());
        return LTE_OK;
    };
    // compile-time delay sets (spec_ct_delay): four taps, first one at delay 0, compact output, two or four antennas
    int dk = 0;
    if (ch->num_taps == 4 && z0 && Ypilot && R2 <= 4 && R == R2)
        for (int k = 1; k <= SPEC_NUM_DK && !dk; ++k) {
            bool same = true;
            for (int ts = 0; ts < 4; ++ts) same = same && C.delay[ts] == spec_ct_delay(k, ts);
            if (same) dk = k;
        }
#define LAUNCH_SPEC_RP(NT, RP)                                                                                   \
    (Ypilot ? (z0 ? launch(channel_spectral_kernel<NT, RP, true, true, 0>) : launch(channel_spectral_kernel<NT, RP, true, false, 0>)) \
            : (z0 ? launch(channel_spectral_kernel<NT, RP, false, true, 0>) : launch(channel_spectral_kernel<NT, RP, false, false, 0>)))
#define LAUNCH_SPEC(NT)                                                     \
    case NT:                                                                \
        return R2 == 2 ? LAUNCH_SPEC_RP(NT, 2) : R2 == 4 ? LAUNCH_SPEC_RP(NT, 4) \
             : R2 == 6 ? LAUNCH_SPEC_RP(NT, 6) : LAUNCH_SPEC_RP(NT, 8);
#define LAUNCH_SPEC_DK(K)                                                   \
    case K:                                                                 \
        return R2 == 2 ? launch(channel_spectral_kernel<4, 2, true, true, K>) : launch(channel_spectral_kernel<4, 4, true, true, K>);
    switch (dk) {
#ifdef SPEC_DEV
        case 1: return launch(channel_spectral_kernel<4, 4, true, true, 1>);
#else
        LAUNCH_SPEC_DK(1) LAUNCH_SPEC_DK(2) LAUNCH_SPEC_DK(3) LAUNCH_SPEC_DK(4)
#endif
        default: break;
    }
    switch (ch->num_taps) {
#ifdef SPEC_DEV                     // development builds: the headline shape only (seconds instead of a minute)
        case 4: return LAUNCH_SPEC_RP(4, 4);
#else
        LAUNCH_SPEC(1) LAUNCH_SPEC(2) LAUNCH_SPEC(3) LAUNCH_SPEC(4) LAUNCH_SPEC(5) LAUNCH_SPEC(6) LAUNCH_SPEC(7)
        LAUNCH_SPEC(8)
#endif
        default: return LTE_ERR_INVALID_ARG;
    }
#undef LAUNCH_SPEC_DK
#undef LAUNCH_SPEC
#undef LAUNCH_SPEC_RP
}
