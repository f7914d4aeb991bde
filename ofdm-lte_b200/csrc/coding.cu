// Coded chain: transport-block CRC + code-block segmentation + turbo encoder + rate
// matching on the TX side, symbol interleaver, max-log soft demapper, rate de-matching, iterative
// max-log BCJR turbo decoder, CRC check and error count on the RX side
// (core/ofdm_core.py:739-1338 and core/channel_coding/{crc,segmentation,turbo_encoder,
// rate_matching,turbo_decoder}.py).
//
// The transport-block structure is fixed by its length, so the host (lte_b200/coding.py) lays it
// out once per length as small int32 tables in HBM:
//   blk[C][LTE_BLK_COLS] = {K, F filler bits, n info bits, src offset into tb+crc, offset into the
//                           code-block row (sum K), offset into the encoded row (sum 3K+12),
//                           has CRC-24B, offset of the block's QPP permutation in pi_tab}
//   rm_table[sum(3K+12)]  : coded bit i  <- encoded bit rm_table[i]   (-1: constant 0)
//   dm_table[sum(3K+12)]  : decoder input j <- LLR dm_table[j]        (-1: never sent, 0.0)
//   pi_tab                : pi(i) = (f1 i + f2 i^2) mod K per distinct K
// One stream = one transport block; every kernel is batched over B streams.
#include "slicer.cuh"

#define BLK_K 0
#define BLK_F 1
#define BLK_N 2
#define BLK_SRC 3
#define BLK_CB 4
#define BLK_ENC 5
#define BLK_CRC 6
#define BLK_PI 7

#define CRC24A 0x864CFBu   // core/channel_coding/crc.py:32 (0x1864CFB without the D^24 term)
#define CRC24B 0x800063u   // :33

__device__ __forceinline__ uint32_t crc24_step(uint32_t reg, uint32_t bit, uint32_t poly) {
    const uint32_t top = ((reg >> 23) & 1u) ^ bit;
    reg = (reg << 1) & 0xFFFFFFu;
    return top ? reg ^ poly : reg;
}

// ------------------------------------------------------------------ TX: CRC-24A of the transport block
// CRCs are linear over GF(2): crc(a ++ b) = crc(a) x^|b| + crc(b) (mod g).  A CTA of 256 threads splits the block
// into 256 chunks, every thread runs the shift register over its chunk, multiplies its remainder by
// x^(bits after the chunk) mod g (square-and-multiply on 24-bit polynomials) and the remainders are XOR-reduced:
// a 1.6 Mbit image payload costs 6 k dependent steps instead of 1.6 M.
__device__ __forceinline__ uint32_t gf2_mulmod24(uint32_t a, uint32_t b, uint32_t poly) {
    uint32_t r = 0;                                   // a(x) b(x) mod g(x), operands < x^24
    for (int i = 23; i >= 0; --i) {
        r = (r & 0x800000u) ? ((r << 1) & 0xFFFFFFu) ^ poly : (r << 1);
        if ((b >> i) & 1u) r ^= a;
    }
    return r;
}
__device__ __forceinline__ uint32_t gf2_xpow24(long long n, uint32_t poly) {      // x^n mod g
    uint32_t result = 1u, base = 2u;
    while (n > 0) {
        if (n & 1) result = gf2_mulmod24(result, base, poly);
        base = gf2_mulmod24(base, base, poly);
        n >>= 1;
    }
    return result;
}

// CRC-24 register of the bits a thread-block sees through `bit(i)`, i in [0, n): every thread returns the result.
template <typename BitFn>
__device__ __forceinline__ uint32_t block_crc24(long long n, uint32_t poly, BitFn bit) {
    __shared__ uint32_t s_part[32];
    const long long per = (n + blockDim.x - 1) / blockDim.x;
    const long long lo = min((long long)threadIdx.x * per, n), hi = min(lo + per, n);
    uint32_t reg = 0;
    for (long long i = lo; i < hi; ++i) reg = crc24_step(reg, bit(i), poly);
    if (reg) reg = gf2_mulmod24(reg, gf2_xpow24(n - hi, poly), poly);
    reg = __reduce_xor_sync(0xffffffffu, reg);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = reg;
    __syncthreads();
    uint32_t total = 0;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) total ^= s_part[w];
    __syncthreads();
    return total;
}

// CTA = stream.
__global__ void __launch_bounds__(256)
tb_crc_kernel(const uint8_t* __restrict__ bits, long long A, uint8_t* __restrict__ crc, long long B) {
    const long long b = blockIdx.x;
    const uint8_t* x = bits + (size_t)b * A;
    const uint32_t reg = block_crc24(A, CRC24A, [&](long long i) { return (uint32_t)(x[i] & 1u); });
    if (threadIdx.x < 24) crc[(size_t)b * 24 + threadIdx.x] = (reg >> (23 - threadIdx.x)) & 1u;
}

// ------------------------------------------------------------------ TX: segmentation (+ CRC-24B)
// thread = (stream, code block): filler zeros, the block's share of tb = bits ++ crc24a, CRC-24B over
// filler + info when the block carries one (core/channel_coding/segmentation.py:66-199).
__global__ void segment_kernel(const uint8_t* __restrict__ bits, const uint8_t* __restrict__ crc, long long A,
                               const int* __restrict__ blk, int C, long long sumK, uint8_t* __restrict__ cb,
                               long long total) {
    const long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= total) return;
    const long long b = it / C;
    const int r = (int)(it - b * C);
    const int* q = blk + r * LTE_BLK_COLS;
    const int K = q[BLK_K], F = q[BLK_F], n = q[BLK_N], src = q[BLK_SRC];
    uint8_t* o = cb + (size_t)b * sumK + q[BLK_CB];
    const uint8_t* x = bits + (size_t)b * A;
    const uint8_t* c = crc + (size_t)b * 24;
    uint32_t reg = 0;
    for (int k = 0; k < F; ++k) { o[k] = 0; reg = crc24_step(reg, 0u, CRC24B); }
    for (int k = 0; k < n; ++k) {
        const long long p = (long long)src + k;
        const uint8_t v = p < A ? (x[p] & 1u) : c[p - A];
        o[F + k] = v;
        reg = crc24_step(reg, v, CRC24B);
    }
    if (q[BLK_CRC]) {
        for (int i = 0; i < 24; ++i) o[F + n + i] = (reg >> (23 - i)) & 1u;
    } else {
        for (int k = F + n; k < K; ++k) o[k] = 0;      // not reached for valid layouts (F + n == K)
    }
}

// ------------------------------------------------------------------ TX: turbo encoder
// thread = (stream, code block, constituent encoder).  The reference's RSC emits the FEEDBACK bit
// a_k = u_k + s1 + s2 as its "systematic" output (core/channel_coding/turbo_encoder.py:140-149),
// parity a_k + s0 + s2, and three termination steps with u = s1 + s2.
__global__ void turbo_encode_kernel(const uint8_t* __restrict__ cb, const int* __restrict__ blk, int C,
                                    long long sumK, long long sumE, const int* __restrict__ pi_tab,
                                    uint8_t* __restrict__ enc, long long total) {
    const long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= total) return;
    const int e = (int)(it & 1);
    const long long br = it >> 1;
    const long long b = br / C;
    const int r = (int)(br - b * C);
    const int* q = blk + r * LTE_BLK_COLS;
    const int K = q[BLK_K];
    const uint8_t* u = cb + (size_t)b * sumK + q[BLK_CB];
    uint8_t* o = enc + (size_t)b * sumE + q[BLK_ENC];
    const int* pi = pi_tab + q[BLK_PI];
    int s0 = 0, s1 = 0, s2 = 0;
    for (int k = 0; k < K + 3; ++k) {
        const int bit = k < K ? (e ? u[pi[k]] : u[k]) : (s1 ^ s2);
        const int fb = bit ^ s1 ^ s2;
        const int par = fb ^ s0 ^ s2;
        if (k < K) {
            if (e == 0) { o[3 * k] = (uint8_t)fb; o[3 * k + 1] = (uint8_t)par; }
            else o[3 * k + 2] = (uint8_t)par;
        } else {
            const int t = k - K;                              // tails: sys1(3) par1(3) sys2(3) par2(3)
            o[3 * K + (e ? 6 : 0) + t] = (uint8_t)fb;
            o[3 * K + (e ? 9 : 3) + t] = (uint8_t)par;
        }
        s2 = s1; s1 = s0; s0 = fb;
    }
}

// ------------------------------------------------------------------ gathers (rate matching / de-matching)
__global__ void gather_u8_kernel(const uint8_t* __restrict__ src, long long n_src, const int* __restrict__ table,
                                 long long n, uint8_t* __restrict__ out, long long total) {
    const long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= total) return;
    const long long b = it / n;
    const int t = table[it - b * n];
    out[it] = t >= 0 ? src[(size_t)b * n_src + t] : (uint8_t)0;
}
__global__ void gather_f32_kernel(const float* __restrict__ src, long long n_src, const int* __restrict__ table,
                                  long long n, float* __restrict__ out, long long total) {
    const long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= total) return;
    const long long b = it / n;
    const int t = table[it - b * n];
    out[it] = t >= 0 ? src[(size_t)b * n_src + t] : 0.f;
}

extern "C" int lte_tb_encode(const uint8_t* bits, int64_t A, const int32_t* blk, int32_t C, int64_t sumK,
                             int64_t sumE, const int32_t* rm_table, const int32_t* pi_tab, uint8_t* crc,
                             uint8_t* cb, uint8_t* enc, uint8_t* coded, int64_t B, void* stream) {
    if (!bits || !blk || !rm_table || !pi_tab || !crc || !cb || !enc || !coded || A < 1 || C < 1 || sumK < 40 ||
        sumE != 3 * sumK + 12 * (int64_t)C || B < 0)
        return LTE_ERR_INVALID_ARG;
    if (B == 0) return LTE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    tb_crc_kernel<<<(unsigned)B, 256, 0, st>>>(bits, A, crc, B);
    const long long t1 = B * C;
    segment_kernel<<<(unsigned)((t1 + 63) / 64), 64, 0, st>>>(bits, crc, A, blk, C, sumK, cb, t1);
    turbo_encode_kernel<<<(unsigned)((2 * t1 + 63) / 64), 64, 0, st>>>(cb, blk, C, sumK, sumE, pi_tab, enc, 2 * t1);
    const long long t2 = B * sumE;
    gather_u8_kernel<<<(unsigned)((t2 + 255) / 256), 256, 0, st>>>(enc, sumE, rm_table, sumE, coded, t2);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

// ------------------------------------------------------------------ TX: QAM map + block interleaver
// core/ofdm_core.py:1037-1060: the nsym QAM symbols are written row-wise into a [rows][Nd] matrix
// (zero padded -- complex zeros, not the symbol of bit pattern 0) and read column-wise; the result is
// cut into rows OFDM symbols of Nd data positions.  out[c*rows + r] = qam(idx[r*Nd + c]).
__global__ void symbol_interleave_kernel(const DevPlan P, const uint8_t* __restrict__ idx, long long nsym, int rows,
                                         float2* __restrict__ out, long long total) {
    const long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= total) return;
    const long long per = (long long)rows * P.Nd;
    const long long b = it / per;
    const long long qpos = it - b * per;
    const int c = (int)(qpos / rows), r = (int)(qpos - (long long)c * rows);
    const long long j = (long long)r * P.Nd + c;
    float2 v = make_float2(0.f, 0.f);
    if (j < nsym) {
        const int h = P.bps >> 1, mask = (1 << h) - 1, x = idx[(size_t)b * nsym + j];
        v = make_float2(P.lev[(x >> h) & mask], P.lev[x & mask]);
    }
    out[it] = v;
}

extern "C" int lte_symbol_interleave(const lte_plan* p, const uint8_t* idx, int64_t nsym, int32_t rows, lte_c32* out,
                                     int64_t B, void* stream) {
    if (!p || !idx || !out || nsym < 1 || rows < 1 || (int64_t)rows * p->dev.Nd < nsym || B < 0)
        return LTE_ERR_INVALID_ARG;
    if (B == 0) return LTE_OK;
    const long long total = (long long)B * rows * p->dev.Nd;
    symbol_interleave_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        p->dev, idx, nsym, rows, (float2*)out, total);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

// ------------------------------------------------------------------ RX: de-interleave + soft demap
// thread = coded symbol j of stream b.  Its received value sits at interleaved position
// q = (j % Nd) * rows + j / Nd of the equalised data stream, i.e. OFDM symbol q / Nd, data bin q % Nd;
// the channel estimate of that bin (held per 14-symbol slot) sets the per-symbol noise variance
//   sigma2 (AWGN)   or   max(sigma2 / clip(|H|^2, 1e-6, 1e6), sigma2 / 4)     (core/ofdm_core.py:1228-1250)
// QPSK: exact LLR 2 sqrt(2) y / nv, unclipped (:791-815).  16/64-QAM: max-log over the natural-binary
// raster constellation, clipped to +-10 (:817-923); the squared distance separates per axis, so the bits of
// the real level index only need the real part (the other axis cancels in min_1 - min_0).
__global__ void __launch_bounds__(256)
soft_demap_kernel(const DevPlan P, const float2* __restrict__ data, const float2* __restrict__ H, int k0, int nk,
                  int nslot, const float* __restrict__ sigma2, int fading, long long nsym, int rows,
                  float* __restrict__ llr, long long total) {
    const long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= total) return;
    const long long b = it / nsym;
    const long long j = it - b * nsym;
    const int r = (int)(j / P.Nd), c = (int)(j - (long long)r * P.Nd);
    const long long q = (long long)c * rows + r;
    const int s = (int)(q / P.Nd), d = (int)(q - (long long)s * P.Nd);
    const float2 y = data[(size_t)b * rows * P.Nd + q];
    const float s2 = sigma2[b];
    float nv = s2;
    if (fading) {
        const float2 h = H[((size_t)b * nslot + s / LTE_SLOT_SYMBOLS) * nk + (P.data_idx[d] - k0)];
        const float hp = fminf(fmaxf(cabs2(h), 1e-6f), 1e6f);
        nv = fmaxf(__fdiv_rn(s2, hp), s2 * 0.25f);
    }
    float* o = llr + ((size_t)b * nsym + j) * P.bps;
    if (P.bps == 2) {
        const float g = __fdiv_rn(2.0f, nv);
        o[0] = g * y.x * 1.41421356237309505f;
        o[1] = g * y.y * 1.41421356237309505f;
        return;
    }
    const int hb = P.bps >> 1, nl = P.nlev;
    const float den = 2.0f * nv;
#pragma unroll
    for (int axis = 0; axis < 2; ++axis) {
        const float v = axis ? y.y : y.x;
        float d0[3] = {INFINITY, INFINITY, INFINITY}, d1[3] = {INFINITY, INFINITY, INFINITY};
        for (int l = 0; l < nl; ++l) {
            const float e = v - P.lev[l];
            const float dist = e * e;
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                if (p < hb) {
                    if ((l >> (hb - 1 - p)) & 1) d1[p] = fminf(d1[p], dist);
                    else d0[p] = fminf(d0[p], dist);
                }
            }
        }
#pragma unroll
        for (int p = 0; p < 3; ++p)
            if (p < hb) o[axis * hb + p] = fminf(fmaxf(__fdiv_rn(d1[p] - d0[p], den), -10.0f), 10.0f);
    }
}

extern "C" int lte_soft_demap(const lte_plan* p, const lte_c32* data, const lte_c32* H, int window,
                              const float* sigma2, int32_t fading, int64_t nsym, int32_t rows, float* llr, int64_t B,
                              void* stream) {
    if (!p || !data || !sigma2 || !llr || nsym < 1 || rows < 1 || (int64_t)rows * p->dev.Nd < nsym || B < 0 ||
        (fading && !H))
        return LTE_ERR_INVALID_ARG;
    int32_t k0, nk;
    int rc = lte_plan_window(p, window, &k0, &nk);
    if (rc) return rc;
    if (B == 0) return LTE_OK;
    const long long total = (long long)B * nsym;
    const int nslot = (rows + LTE_SLOT_SYMBOLS - 1) / LTE_SLOT_SYMBOLS;
    soft_demap_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        p->dev, (const float2*)data, (const float2*)H, k0, nk, nslot, sigma2, fading, nsym, rows, llr, total);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

// ------------------------------------------------------------------ RX: turbo decoder
// Max-log BCJR, the reference's default mode (core/channel_coding/turbo_decoder.py:24, 158-337), eight
// lanes per code block (lane = trellis state), four code blocks per warp.
//   state = (s0 << 2) | (s1 << 1) | s2;  branch u: a = u ^ s1 ^ s2, parity = a ^ s0 ^ s2, next = (a << 2) | (s0 << 1) | s1
//   gamma = (+-Ls +-Lp +-La) / 2 with Ls signed by a (the reference's "systematic" bit), La by u
//   alpha_0 = beta_n = (0, -inf, ...) over n = K + 3 steps (tail included, a-priori 0 there)
// Forward pass stores every 8th alpha in HBM scratch ([n / 8 + 1][32] floats per warp = four blocks, one 128 B line);
// the backward pass fuses the a-posteriori max over the 16 branches (xor-shuffle reductions in the 8-lane
// group).  Extrinsic values cross the QPP interleaver through HBM scratch as well.
struct Bcjr {
    const float* ls; const float* lp; const float* la;   // systematic / parity / a-priori LLR sources
    const int* perm;                                     // optional permutation applied to ls and la (decoder 2)
    const float* ls_tail; const float* lp_tail;          // 3 tail LLRs each
};

// max*(a, b): max-log-MAP (the reference's default, USE_MAX_LOG_MAP = True) or the exact Jacobian logarithm
// log(e^a + e^b) = max + log1p(e^-|a-b|) of its "true Log-MAP" mode (turbo_decoder.py:64-120); -inf is the
// neutral element in both.
template <bool LOGMAP> __device__ __forceinline__ float max_star(float a, float b) {
    const float m = fmaxf(a, b);
    if (!LOGMAP) return m;
    if (m == -INFINITY) return m;
    return m + log1pf(expf(-fabsf(a - b)));
}

// All shuffles use the full warp mask: the four code blocks of a warp step in lockstep over the longest of
// their trellises (nch_w chunks), shorter or absent blocks just predicate their updates off.  (A per-group
// mask makes nvcc wrap every shuffle in a MATCH/VOTE convergence loop -- 3x the instructions.)
#define FULLMASK 0xffffffffu

template <bool EXTRINSIC_OUT, bool LOGMAP>
__device__ __forceinline__ void bcjr_pass(const Bcjr in, int K, int nch_w, float* __restrict__ alpha,
                                          float* __restrict__ ext_out, bool scatter, uint8_t* __restrict__ bits_out,
                                          int st, float* __restrict__ apost_out = nullptr) {
    const int n = K > 0 ? K + 3 : 0;                     // K == 0: padding group of the last warp
    const int s0 = (st >> 2) & 1, s1 = (st >> 1) & 1, s2 = st & 1;
    // outgoing branches of this state
    const int a0 = s1 ^ s2, a1 = 1 ^ s1 ^ s2;
    const int nx0 = (a0 << 2) | (s0 << 1) | s1, nx1 = (a1 << 2) | (s0 << 1) | s1;
    const float ss0 = a0 ? -0.5f : 0.5f, ss1 = a1 ? -0.5f : 0.5f;
    const float sp0 = (a0 ^ s0 ^ s2) ? -0.5f : 0.5f, sp1 = (a1 ^ s0 ^ s2) ? -0.5f : 0.5f;
    // incoming branches: predecessors p with next(p, u) == st.  st = (a << 2) | (p0 << 1) | p1 fixes p0, p1, a;
    // p2 in {0, 1} and u = a ^ p1 ^ p2.
    const int pa = (st >> 2) & 1, pp0 = (st >> 1) & 1, pp1 = st & 1;
    const int pr0 = (pp0 << 2) | (pp1 << 1) | 0, pr1 = (pp0 << 2) | (pp1 << 1) | 1;
    const int pu0 = pa ^ pp1 ^ 0, pu1 = pa ^ pp1 ^ 1;
    const float is0 = pa ? -0.5f : 0.5f;                                        // same a on both incoming branches
    const float ip0 = (pa ^ pp0 ^ 0) ? -0.5f : 0.5f, ip1 = (pa ^ pp0 ^ 1) ? -0.5f : 0.5f;
    const float iu0 = pu0 ? -0.5f : 0.5f, iu1 = pu1 ? -0.5f : 0.5f;

    // The LLRs of a step are the same for the eight lanes, and the recursion consumes them one step at a
    // time, so each lane fetches the inputs of ONE step of an 8-step chunk (a dependent pi[k] -> LLR chain
    // for decoder 2) and the chunk after the current one is already in flight while this one is processed;
    // the values reach the other lanes by shuffle.  One global-memory latency per 8 steps, hidden.
    auto fetch = [&](int k, float& Ls, float& Lp, float& La, int& kk) {
        Ls = Lp = La = 0.f;
        kk = k;
        if (k < K) {
            kk = in.perm ? in.perm[k] : k;
            Ls = in.ls[3 * kk];
            Lp = in.lp[3 * k];
            La = in.la[kk];
        } else if (k < n) {
            Ls = in.ls_tail[k - K];
            Lp = in.lp_tail[k - K];
        }
    };
    // ---- forward
    float a = st == 0 ? 0.f : -INFINITY;
    if (n) alpha[0] = a;
    {
        float cLs, cLp, cLa, nLs, nLp, nLa;
        int ckk, nkk;
        fetch(st, cLs, cLp, cLa, ckk);
        for (int c = 0; c < nch_w; ++c) {
            fetch((c + 1) * 8 + st, nLs, nLp, nLa, nkk);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = c * 8 + j;
                const float Ls = __shfl_sync(FULLMASK, cLs, j, 8), Lp = __shfl_sync(FULLMASK, cLp, j, 8),
                            La = __shfl_sync(FULLMASK, cLa, j, 8);
                const float g0 = (is0 * Ls + ip0 * Lp) + iu0 * La;
                const float g1 = (is0 * Ls + ip1 * Lp) + iu1 * La;
                const float x0 = __shfl_sync(FULLMASK, a, pr0, 8) + g0;
                const float x1 = __shfl_sync(FULLMASK, a, pr1, 8) + g1;
                if (k < n) a = max_star<LOGMAP>(x0, x1);
            }
            if ((c + 1) * 8 < n) alpha[(size_t)(c + 1) * 32] = a;       // checkpoint: alpha before step 8 (c + 1)
            cLs = nLs; cLp = nLp; cLa = nLa; ckk = nkk;
        }
    }
    // ---- backward + a-posteriori.  Only every 8th alpha was stored: the chunk's other seven are recomputed from
    // its checkpoint with the very operations of the forward pass (bit-identical), which trades ~40 % more
    // arithmetic for an 8x smaller scratch stream -- the decoder is memory-latency bound once the batch fills the SMs.
    float bt = st == 0 ? 0.f : -INFINITY;
    {
        float cLs, cLp, cLa, nLs = 0.f, nLp = 0.f, nLa = 0.f, ck, nck = 0.f;
        int ckk, nkk = 0;
        fetch((nch_w - 1) * 8 + st, cLs, cLp, cLa, ckk);
        ck = (nch_w - 1) * 8 < n ? alpha[(size_t)(nch_w - 1) * 32] : 0.f;
        for (int c = nch_w - 1; c >= 0; --c) {
            if (c > 0) {
                fetch((c - 1) * 8 + st, nLs, nLp, nLa, nkk);
                nck = (c - 1) * 8 < n ? alpha[(size_t)(c - 1) * 32] : 0.f;
            }
            float ca[8];
            ca[0] = ck;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const float Ls = __shfl_sync(FULLMASK, cLs, j, 8), Lp = __shfl_sync(FULLMASK, cLp, j, 8),
                            La = __shfl_sync(FULLMASK, cLa, j, 8);
                const float g0 = (is0 * Ls + ip0 * Lp) + iu0 * La;
                const float g1 = (is0 * Ls + ip1 * Lp) + iu1 * La;
                const float x0 = __shfl_sync(FULLMASK, ca[j], pr0, 8) + g0;
                const float x1 = __shfl_sync(FULLMASK, ca[j], pr1, 8) + g1;
                ca[j + 1] = max_star<LOGMAP>(x0, x1);
            }
#pragma unroll
            for (int j = 7; j >= 0; --j) {
                const int k = c * 8 + j;
                const float Ls = __shfl_sync(FULLMASK, cLs, j, 8), Lp = __shfl_sync(FULLMASK, cLp, j, 8),
                            La = __shfl_sync(FULLMASK, cLa, j, 8);
                const int kk = __shfl_sync(FULLMASK, ckk, j, 8);
                const float g0 = (ss0 * Ls + sp0 * Lp) + 0.5f * La;
                const float g1 = (ss1 * Ls + sp1 * Lp) - 0.5f * La;
                const float b0 = __shfl_sync(FULLMASK, bt, nx0, 8), b1 = __shfl_sync(FULLMASK, bt, nx1, 8);
                float v0 = (ca[j] + g0) + b0, v1 = (ca[j] + g1) + b1;
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) {
                    v0 = max_star<LOGMAP>(v0, __shfl_xor_sync(FULLMASK, v0, o, 8));
                    v1 = max_star<LOGMAP>(v1, __shfl_xor_sync(FULLMASK, v1, o, 8));
                }
                if (k < n) {
                    bt = max_star<LOGMAP>(b0 + g0, b1 + g1);
                    if (st == 0) {
                        const float ap = v0 - v1;
                        if (k < K) {
                            if (EXTRINSIC_OUT) ext_out[scatter ? kk : k] = (ap - La) - Ls;
                            else bits_out[k] = ap < 0.f ? 1 : 0;
                        }
                        if (!EXTRINSIC_OUT && apost_out) apost_out[k] = ap;
                    }
                }
            }
            cLs = nLs; cLp = nLp; cLa = nLa; ckk = nkk; ck = nck;
        }
    }
}

#ifndef TURBO_MIN_CTAS
#define TURBO_MIN_CTAS 4
#endif
template <bool LOGMAP>
__global__ void __launch_bounds__(128, TURBO_MIN_CTAS)
turbo_decode_kernel(const float* __restrict__ dl, const int* __restrict__ blk, int C, long long sumK, long long sumE,
                    const int* __restrict__ pi_tab, int iterations, float* __restrict__ work, long long work_per_blk,
                    int Kmax, uint8_t* __restrict__ cbdec, long long total, const float* __restrict__ apriori,
                    float* __restrict__ apost) {
    long long g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;            // (stream, code block)
    if ((g & ~3ll) >= total) return;                       // whole warp beyond the work: warp-uniform exit
    const bool live = g < total;
    const long long gw = g;
    if (!live) g = total - 1;                              // padding group of the last warp: valid pointers, K = 0
    const int lane = threadIdx.x & 31, st = lane & 7;
    const long long b = g / C;
    const int r = (int)(g - b * C);
    const int* q = blk + r * LTE_BLK_COLS;
    const int K = live ? q[BLK_K] : 0;
    const int nch_w = (__reduce_max_sync(FULLMASK, K) + 3 + 7) >> 3;
    const float* x = dl + (size_t)b * sumE + q[BLK_ENC];
    const int* pi = pi_tab + q[BLK_PI];
    // scratch of the four code blocks of a warp: alpha interleaved [step][lane] so that a warp's store of one
    // step is one 128-byte line, then e12 / e21 (extrinsic, natural order) per block
    float* wbase = work + (size_t)(gw & ~3ll) * work_per_blk;
    float* alpha = wbase + lane;
    float* e12 = wbase + (size_t)((Kmax + 10) / 8 + 1) * 32 + (size_t)(gw & 3) * 2 * Kmax;   // decoder 1 -> 2
    float* e21 = e12 + Kmax;                                                     // decoder 2 -> 1 (scattered through pi)
    for (int k = st; k < K; k += 8) e21[k] = apriori ? apriori[(size_t)b * sumK + q[BLK_CB] + k] : 0.f;
    __syncwarp();
    const int Kq = live ? q[BLK_K] : 0;
    const Bcjr d1 = {x, x + 1, e21, nullptr, x + 3 * Kq, x + 3 * Kq + 3};
    const Bcjr d2 = {x, x + 2, e12, pi, x + 3 * Kq + 6, x + 3 * Kq + 9};
    for (int it = 0; it < iterations; ++it) {
        bcjr_pass<true, LOGMAP>(d1, K, nch_w, alpha, e12, false, nullptr, st);
        __syncwarp();
        bcjr_pass<true, LOGMAP>(d2, K, nch_w, alpha, e21, true, nullptr, st);    // e21[pi[k]] = ext2[k]: the de-interleave
        __syncwarp();
    }
    bcjr_pass<false, LOGMAP>(d1, K, nch_w, alpha, nullptr, false, cbdec + (size_t)b * sumK + q[BLK_CB], st,
                     apost ? apost + (size_t)b * (sumK + 3 * C) + q[BLK_CB] + 3 * r : nullptr);
}

// ------------------------------------------------------------------ RX: de-segmentation, CRC check, errors
// CTA = stream.  The A + 24 bits of the decoded transport block (info bits of the code blocks back to back,
// segmentation.py:202-270) are split into 256 chunks; a thread finds the code block its chunk starts in, walks its
// chunk across block boundaries, and accumulates (a) its piece of the CRC-24A register over the first A bits --
// combined by the GF(2) linearity used in tb_crc_kernel -- (b) the received CRC bits, (c) the bit errors against the
// transmitted block (core/ofdm_core.py:1296-1307), and writes bits_rx on the way.
__global__ void __launch_bounds__(256)
tb_check_kernel(const uint8_t* __restrict__ cbdec, const int* __restrict__ blk, int C, long long sumK, long long A,
                const uint8_t* __restrict__ bits_tx, uint8_t* __restrict__ bits_rx, int* __restrict__ crc_ok,
                unsigned long long* __restrict__ errors, long long B) {
    __shared__ uint32_t s_crc[8], s_got;
    __shared__ unsigned int s_err[8];
    const long long b = blockIdx.x;
    const long long n = A + 24;
    const long long per = (n + blockDim.x - 1) / blockDim.x;
    const long long lo = min((long long)threadIdx.x * per, n), hi = min(lo + per, n);
    if (threadIdx.x == 0) s_got = 0;
    __syncthreads();
    uint32_t reg = 0, got = 0;
    unsigned int e = 0;
    if (lo < hi) {
        int r = 0;
        while (r + 1 < C && blk[(r + 1) * LTE_BLK_COLS + BLK_SRC] <= lo) ++r;      // block holding position lo
        const int* q = blk + r * LTE_BLK_COLS;
        long long k = lo - q[BLK_SRC];
        const uint8_t* c = cbdec + (size_t)b * sumK + q[BLK_CB] + q[BLK_F];
        for (long long p = lo; p < hi; ++p) {
            while (k >= q[BLK_N]) {                                                   // next code block
                ++r;
                q = blk + r * LTE_BLK_COLS;
                k = 0;
                c = cbdec + (size_t)b * sumK + q[BLK_CB] + q[BLK_F];
            }
            const uint32_t v = c[k++] & 1u;
            if (p < A) {
                reg = crc24_step(reg, v, CRC24A);
                if (bits_rx) bits_rx[(size_t)b * A + p] = (uint8_t)v;
                if (bits_tx) e += v != (bits_tx[(size_t)b * A + p] & 1u);
            } else {
                got |= v << (23 - (int)(p - A));
            }
        }
        const long long covered = min(hi, A);
        if (reg) reg = gf2_mulmod24(reg, gf2_xpow24(A - covered, CRC24A), CRC24A);
    }
    reg = __reduce_xor_sync(0xffffffffu, reg);
    e = __reduce_add_sync(0xffffffffu, e);
    got = __reduce_or_sync(0xffffffffu, got);
    if ((threadIdx.x & 31) == 0) {
        s_crc[threadIdx.x >> 5] = reg;
        s_err[threadIdx.x >> 5] = e;
        if (got) atomicOr(&s_got, got);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t crc = 0;
        unsigned long long te = 0;
        for (int w = 0; w < 8; ++w) { crc ^= s_crc[w]; te += s_err[w]; }
        if (crc_ok) crc_ok[b] = crc == s_got ? 1 : 0;
        if (errors) errors[b] = te;
    }
}

// per code block: one alpha checkpoint (8 states) per 8 trellis steps + the two extrinsic vectors
extern "C" int64_t lte_tb_decode_work_floats(int32_t Kmax) { return (int64_t)((Kmax + 10) / 8 + 1) * 8 + 2 * (int64_t)Kmax; }

extern "C" int lte_tb_decode(const float* llr, const int32_t* blk, int32_t C, int64_t sumK, int64_t sumE, int32_t Kmax,
                             const int32_t* dm_table, const int32_t* pi_tab, int32_t iterations, int32_t logmap, float* dematched,
                             float* work, uint8_t* cbdec, int64_t A, const uint8_t* bits_tx, uint8_t* bits_rx,
                             int32_t* crc_ok, unsigned long long* errors, int64_t B, void* stream) {
    if (!llr || !blk || !dm_table || !pi_tab || !dematched || !work || !cbdec || C < 1 || sumK < 40 || Kmax < 40 ||
        Kmax > 6144 || sumE != 3 * sumK + 12 * (int64_t)C || iterations < 0 || A < 1 || B < 0)
        return LTE_ERR_INVALID_ARG;
    if (B == 0) return LTE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const long long t2 = B * sumE;
    gather_f32_kernel<<<(unsigned)((t2 + 255) / 256), 256, 0, st>>>(llr, sumE, dm_table, sumE, dematched, t2);
    const long long nblk = B * C;
    if (logmap)
        turbo_decode_kernel<true><<<(unsigned)((nblk * 8 + 127) / 128), 128, 0, st>>>(
            dematched, blk, C, sumK, sumE, pi_tab, iterations, work, lte_tb_decode_work_floats(Kmax), Kmax, cbdec, nblk,
            nullptr, nullptr);
    else
        turbo_decode_kernel<false><<<(unsigned)((nblk * 8 + 127) / 128), 128, 0, st>>>(
            dematched, blk, C, sumK, sumE, pi_tab, iterations, work, lte_tb_decode_work_floats(Kmax), Kmax, cbdec, nblk,
            nullptr, nullptr);
    tb_check_kernel<<<(unsigned)B, 256, 0, st>>>(cbdec, blk, C, sumK, A, bits_tx, bits_rx, crc_ok, errors, B);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

// ------------------------------------------------------------------ stage-level entry points
// The reference exposes its coding stages as module functions (core/channel_coding/__init__.py); these
// launch the same kernels one stage at a time so that API can sit on the GPU as well.

// CRC of B rows of n bits each: generator `poly` of degree `len` (16 or 24) without its leading term,
// zero initial state, MSB first (core/channel_coding/crc.py:89-134).  out [B][len] bits.
__global__ void crc_rows_kernel(const uint8_t* __restrict__ bits, long long n, uint32_t poly, int len,
                                uint8_t* __restrict__ out, long long B) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint8_t* x = bits + (size_t)b * n;
    const uint32_t msk = len == 32 ? 0xFFFFFFFFu : ((1u << len) - 1u);
    uint32_t reg = 0;
    for (long long i = 0; i < n; ++i) {
        const uint32_t top = ((reg >> (len - 1)) & 1u) ^ (x[i] & 1u);
        reg = (reg << 1) & msk;
        if (top) reg ^= poly;
    }
    for (int i = 0; i < len; ++i) out[(size_t)b * len + i] = (reg >> (len - 1 - i)) & 1u;
}

extern "C" int lte_crc_bits(const uint8_t* bits, int64_t n, uint32_t poly, int32_t len, uint8_t* out, int64_t B,
                            void* stream) {
    if (!bits || !out || n < 0 || len < 1 || len > 32 || B < 0) return LTE_ERR_INVALID_ARG;
    if (B == 0) return LTE_OK;
    crc_rows_kernel<<<(unsigned)((B + 63) / 64), 64, 0, (cudaStream_t)stream>>>(bits, n, poly, len, out, B);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

extern "C" int lte_turbo_encode_blocks(const uint8_t* cb, const int32_t* blk, int32_t C, int64_t sumK, int64_t sumE,
                                       const int32_t* pi_tab, uint8_t* enc, int64_t B, void* stream) {
    if (!cb || !blk || !pi_tab || !enc || C < 1 || sumK < 40 || sumE != 3 * sumK + 12 * (int64_t)C || B < 0)
        return LTE_ERR_INVALID_ARG;
    if (B == 0) return LTE_OK;
    const long long t = 2 * B * C;
    turbo_encode_kernel<<<(unsigned)((t + 63) / 64), 64, 0, (cudaStream_t)stream>>>(cb, blk, C, sumK, sumE, pi_tab, enc, t);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

extern "C" int lte_turbo_decode_blocks(const float* dl, const int32_t* blk, int32_t C, int64_t sumK, int64_t sumE,
                                       int32_t Kmax, const int32_t* pi_tab, int32_t iterations, int32_t logmap,
                                       float* work, uint8_t* cbdec, const float* apriori, float* apost, int64_t B,
                                       void* stream) {
    if (!dl || !blk || !pi_tab || !work || !cbdec || C < 1 || sumK < 1 || Kmax < 1 || Kmax > 6144 ||
        sumE != 3 * sumK + 12 * (int64_t)C || iterations < 0 || B < 0)
        return LTE_ERR_INVALID_ARG;
    if (B == 0) return LTE_OK;
    const long long nblk = B * C;
    if (logmap)
        turbo_decode_kernel<true><<<(unsigned)((nblk * 8 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
            dl, blk, C, sumK, sumE, pi_tab, iterations, work, lte_tb_decode_work_floats(Kmax), Kmax, cbdec, nblk, apriori, apost);
    else
        turbo_decode_kernel<false><<<(unsigned)((nblk * 8 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
            dl, blk, C, sumK, sumE, pi_tab, iterations, work, lte_tb_decode_work_floats(Kmax), Kmax, cbdec, nblk, apriori, apost);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

// out[b][i] = table[i] >= 0 ? src[b][table[i]] : 0 -- interleavers, rate matching and their inverses as tables
extern "C" int lte_gather_u8(const uint8_t* src, int64_t n_src, const int32_t* table, int64_t n, uint8_t* out, int64_t B,
                             void* stream) {
    if (!src || !table || !out || n < 1 || n_src < 1 || B < 0) return LTE_ERR_INVALID_ARG;
    if (B == 0) return LTE_OK;
    const long long t = B * n;
    gather_u8_kernel<<<(unsigned)((t + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, n_src, table, n, out, t);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}
extern "C" int lte_gather_f32(const float* src, int64_t n_src, const int32_t* table, int64_t n, float* out, int64_t B,
                              void* stream) {
    if (!src || !table || !out || n < 1 || n_src < 1 || B < 0) return LTE_ERR_INVALID_ARG;
    if (B == 0) return LTE_OK;
    const long long t = B * n;
    gather_f32_kernel<<<(unsigned)((t + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, n_src, table, n, out, t);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}
