// Shared device/host definitions for liblte_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <map>
#include <vector>

#include "../../include/lte_b200.h"

#define LTE_CHECK_CUDA(expr)                          \
    do {                                              \
        cudaError_t _e = (expr);                      \
        if (_e != cudaSuccess) return lte_set_cuda_error(_e); \
    } while (0)

int lte_set_cuda_error(cudaError_t e);

// bin_map encoding (one int16 per FFT bin, TX loader): <0 null, else data/pilot index
#define BIN_NULL (-1)
#define BIN_PILOT_FLAG 0x4000

// Device view of a plan; passed to kernels by value.
// entries behind the N-point twiddle table: per-pass tables of the packed FFT (fft2.cuh, plan.cu)
#define FFT2_TW_PASS3 256
#define FFT2_TW_EXTRA 2048

struct DevPlan {
    int N, log2N, Nc, cp, L, Nd, Np, bps;
    int k0_useful, nk_useful;       // occupied-bin window
    const int16_t* bin_map;         // [N]
    const int16_t* data_idx;        // [Nd]
    const int16_t* pilot_idx;       // [Np]
    const float2* pilots;           // [nsets][Np]  (zero where not owned)
    // per pilot set: owned pilots (bin, 1/p) and the left-neighbour table for interpolation
    const int16_t* pset_bin;        // [nsets][Np]
    const float2* pset_inv;         // [nsets][Np]  1/p in the reference's sense (Y / p)
    const int16_t* pset_seg;        // [nsets][N]   index of the owned pilot at or left of bin k, -1 before the first
    const int* pset_cnt;            // [nsets]
    const float2* twiddle;          // [N] exp(-2 pi i m / N)
    // bin pairs of the spectral link / compact sweep layout (spectral.cu): every bin of the occupied window
    // exactly once, as ndp pairs of consecutive DATA symbols, then npp pairs of consecutive pilots, then the
    // window's remaining null bins; -1 pads an odd count
    const int16_t* pair_bin;        // [npairs][2]
    int ndp, npp, npairs;
    float lev[8];                   // constellation axis levels in index order
    float thr[7];                   // slicer thresholds, rounded towards -inf
    int nlev;                       // 2, 4, 8
    float inv_sqrt_n;
};

// Bluestein tables of the SC-FDM M-point DFT (dft.cu), device resident, owned by the plan
struct DftTables {
    int M, NB;
    float2* w;      // [M]   chirp exp(-j pi n^2 / M)
    float2* bf;     // [NB]  FFT of the circular chirp filter, pre-scaled by 1/(NB sqrt(M))
    float2* tw;     // [NB]  FFT twiddles exp(-2 pi i m / NB)
};

struct lte_plan {
    lte_plan_desc desc;
    DevPlan dev;
    int device;
    int nsets;
    std::vector<int32_t> data_idx_h, pilot_idx_h;
    void* blob;                     // single device allocation holding all tables
    std::map<int, DftTables> dft;   // M -> tables added by lte_plan_add_dft
};

// ------------------------------------------------------------------ complex helpers
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
// conj(a) * b
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.x, b.y, -a.y * b.x));
}
__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
__device__ __forceinline__ float cabs2(float2 a) { return fmaf(a.x, a.x, a.y * a.y); }
// a / b, IEEE division (no fast-math): used by ZF / MRC / LS so slicer inputs stay accurate
__device__ __forceinline__ float2 cdiv(float2 a, float2 b) {
    float d = cabs2(b);
    float2 n = make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
    return make_float2(__fdiv_rn(n.x, d), __fdiv_rn(n.y, d));
}

// ------------------------------------------------------------------ Philox2x32-10
// Counter-based generator (Salmon et al., SC'11).  One call -> 64 random bits.
__host__ __device__ __forceinline__ void philox2x32_10(uint32_t key, uint32_t c0, uint32_t c1,
                                                       uint32_t& r0, uint32_t& r1) {
    const uint32_t M = 0xD256D193u, W = 0x9E3779B9u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
#ifdef __CUDA_ARCH__
        uint32_t hi = __umulhi(M, c0);
#else
        uint32_t hi = (uint32_t)(((uint64_t)M * c0) >> 32);
#endif
        uint32_t lo = M * c0;
        c0 = hi ^ key ^ c1;
        c1 = lo;
        key += W;
    }
    r0 = c0;
    r1 = c1;
}

// The same generator with the ten round keys key + i W handed in (kernel parameters = constant-bank operands of
// the XOR): a hot loop does not re-derive the schedule for every draw.
struct PhiloxKeys { uint32_t k[10]; };
static inline PhiloxKeys philox_key_schedule(uint32_t key) {
    PhiloxKeys s;
    for (int i = 0; i < 10; ++i) s.k[i] = key + (uint32_t)i * 0x9E3779B9u;
    return s;
}
__device__ __forceinline__ void philox2x32_10_ks(const PhiloxKeys& ks, uint32_t c0, uint32_t c1, uint32_t& r0, uint32_t& r1) {
    const uint32_t M = 0xD256D193u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi = __umulhi(M, c0), lo = M * c0;
        c0 = hi ^ ks.k[i] ^ c1;
        c1 = lo;
    }
    r0 = c0;
    r1 = c1;
}

// Key derivation: mixes the user seed with a domain tag so that bits, phases and
// noise never share a stream.
#define LTE_DOMAIN_NOISE 0x6e6f6973u
#define LTE_DOMAIN_BITS 0x62697473u
#define LTE_DOMAIN_PHASE 0x70686173u
__host__ __device__ __forceinline__ uint32_t lte_key(uint64_t seed, uint32_t domain) {
    uint32_t a, b;
    philox2x32_10(domain, (uint32_t)seed, (uint32_t)(seed >> 32), a, b);
    return a ^ b;
}

// Stream / row identifiers are 32-bit Philox counter words: a launch whose ids would run past 2^32 would repeat
// earlier draws (duplicate trials, overstated confidence), so the entry points refuse it instead.
static inline bool lte_ids_fit(uint64_t id0, uint64_t count) { return id0 <= (1ull << 32) && count <= (1ull << 32) - id0; }

// One complex unit normal (re, im ~ N(0,1)) for (row, sample) by Box-Muller; branch-free.
// u1 = (r0 + 0.5) 2^-32 keeps the full 32-bit tail resolution for small r0 (fp32 is exact
// below 2^24), so |z| reaches sqrt(-2 ln 2^-33) = 6.8 sigma.
__device__ __forceinline__ float2 lte_noise_from_bits(uint32_t r0, uint32_t r1);
__device__ __forceinline__ float2 lte_noise_sample(uint32_t key, uint32_t row, uint32_t sample) {
    uint32_t r0, r1;
    philox2x32_10(key, sample, row, r0, r1);
    return lte_noise_from_bits(r0, r1);
}
__device__ __forceinline__ float2 lte_noise_sample(const PhiloxKeys& ks, uint32_t row, uint32_t sample) {
    uint32_t r0, r1;
    philox2x32_10_ks(ks, sample, row, r0, r1);
    return lte_noise_from_bits(r0, r1);
}
__device__ __forceinline__ float2 lte_noise_from_bits(uint32_t r0, uint32_t r1) {
    const float u1 = fmaf((float)r0, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    const float ang = __uint_as_float(0x3f800000u | (r1 >> 9)) - 1.5f;      // [-0.5, 0.5) turns, 23 bits
    float rad;                                                              // sqrt(-2 ln u1)
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(-1.3862943611198906f * __log2f(u1)));
    float s, c;
    __sincosf(6.283185307179586f * ang, &s, &c);
    return make_float2(rad * c, rad * s);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// sigma of the reference's AWGN: sqrt(mean_power / snr_lin / 2)   (core/channel.py:52,58)
__device__ __forceinline__ float lte_sigma(double power_sum, float n, float snr_lin) {
    return sqrtf((float)power_sum / (n * snr_lin * 2.0f));
}
