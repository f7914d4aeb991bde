// Stage 4 (CRS LS + linear interpolation), stage 5 (ZF / MRC) and stage 6 (hard demap +
// bit-error count), plus the bit <-> symbol-index helpers of the reference-facing API.
#include <algorithm>
#include <vector>

#include "slicer.cuh"
#include "awgn.cuh"

#include "common.cuh"

// ------------------------------------------------------------------------------ bits <-> indices
// core/modulator.py:74-84: zero-pad, read b bits MSB first.
// One thread per symbol; block = (row, chunk of 256 symbols).  Packed rows (np.packbits layout) are
// read through a 16-bit window, so a symbol costs two byte loads instead of `bps` of them.
__global__ void __launch_bounds__(256)
bits_to_indices_kernel(const uint8_t* __restrict__ bits, long long nbits, int packed, uint8_t* __restrict__ idx,
                       unsigned nsym, int bps, unsigned chunks) {
    const unsigned row = blockIdx.x / chunks;
    const unsigned q = (blockIdx.x - row * chunks) * 256u + threadIdx.x;
    if (q >= nsym) return;
    const long long bi0 = (long long)q * bps;            // first bit of the symbol within the row
    int v = 0;
    if (packed) {
        const long long row_bytes = (nbits + 7) >> 3;
        const uint8_t* src = bits + (size_t)row * row_bytes;
        const long long by = bi0 >> 3;
        const int off = (int)(bi0 & 7);
        const unsigned hi = by < row_bytes ? src[by] : 0u, lo = by + 1 < row_bytes ? src[by + 1] : 0u;
        v = (int)((((hi << 8) | lo) >> (16 - off - bps)) & ((1u << bps) - 1u));
        const long long left = nbits - bi0;              // bits of this symbol inside the row
        if (left < bps) v = left <= 0 ? 0 : (v & ~((1 << (bps - (int)left)) - 1));
    } else {
        const uint8_t* src = bits + (size_t)row * nbits + bi0;
        for (int i = 0; i < bps; ++i) v = (v << 1) | ((bi0 + i < nbits) ? (src[i] & 1) : 0);
    }
    idx[(size_t)row * nsym + q] = (uint8_t)v;
}

// Packed rows, one 4-byte-aligned group of idx per thread (the sweep's e2e path: 57 M symbols per step).  The
// group's four symbols span 4 * bps <= 24 bits at an arbitrary bit offset, i.e. at most four bytes of the row:
// they are read once into a big-endian window and leave as one 32-bit store.  Rows need not be a multiple of four
// symbols long: the (at most two) groups a row shares with its neighbours fall back to byte stores of the
// symbols that belong to this row.  Bits past nbits read as 0.
__device__ __forceinline__ unsigned packed_symbol(const uint8_t* src, long long row_bytes, long long nbits, long long q,
                                                  int bps) {
    const long long bi0 = q * bps, by = bi0 >> 3;
    const int off = (int)(bi0 & 7);
    const unsigned hi = by < row_bytes ? src[by] : 0u, lo = by + 1 < row_bytes ? src[by + 1] : 0u;
    unsigned v = (((hi << 8) | lo) >> (16 - off - bps)) & ((1u << bps) - 1u);
    const long long left = nbits - bi0;
    if (left < bps) v = left <= 0 ? 0u : (v & ~((1u << (bps - (int)left)) - 1u));
    return v;
}

__global__ void __launch_bounds__(256)
bits_to_indices_x4_kernel(const uint8_t* __restrict__ bits, long long nbits, uint8_t* __restrict__ idx, unsigned nsym,
                          int bps, unsigned chunks) {
    const unsigned row = blockIdx.x / chunks;
    const unsigned g = (blockIdx.x - row * chunks) * 256u + threadIdx.x;
    const unsigned long long row0 = (unsigned long long)row * nsym;           // address of the row's first index
    const unsigned long long A = ((row0 >> 2) + g) << 2;                      // this thread's aligned group
    const long long q0 = (long long)A - (long long)row0;                      // its first symbol (may be < 0)
    if (q0 >= (long long)nsym) return;
    const long long row_bytes = (nbits + 7) >> 3;
    const uint8_t* src = bits + (size_t)row * row_bytes;
    if (q0 < 0 || q0 + 3 >= (long long)nsym || (q0 + 4) * bps > nbits) {       // shared with a neighbour row / ragged end
        for (int k = 0; k < 4; ++k) {
            const long long q = q0 + k;
            if (q >= 0 && q < (long long)nsym) idx[A + k] = (uint8_t)packed_symbol(src, row_bytes, nbits, q, bps);
        }
        return;
    }
    const long long bi0 = q0 * bps, by = bi0 >> 3;
    const int off = (int)(bi0 & 7);
    unsigned w = 0;                                                           // bytes by .. by+3, big endian
#pragma unroll
    for (int i = 0; i < 4; ++i) w = (w << 8) | (by + i < row_bytes ? (unsigned)src[by + i] : 0u);
    const unsigned m = (1u << bps) - 1u;
    unsigned out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) out |= ((w >> (32 - off - (k + 1) * bps)) & m) << (8 * k);
    *(unsigned*)(idx + A) = out;
}

// Packed rows, one 16-byte-aligned group of idx per thread: the group's 16 symbols are 16 * bps <= 96 bits at an
// arbitrary bit offset of the row, i.e. at most 13 bytes.  They are read as four aligned 32-bit words, turned
// big endian (np.packbits is MSB first), shifted so that the first symbol starts at bit 0, cut into symbols at
// compile-time positions and leave as ONE 128-bit store.  Groups that touch a row boundary, the ragged end of the
// bit budget or the end of the buffer go through packed_symbol() byte by byte.
template <int BPS>
__global__ void __launch_bounds__(256)
bits_to_indices_x16_kernel(const uint8_t* __restrict__ bits, long long nbits, uint8_t* __restrict__ idx, unsigned nsym,
                           unsigned chunks, long long total_bytes) {
    const unsigned row = blockIdx.x / chunks;
    const unsigned g = (blockIdx.x - row * chunks) * 256u + threadIdx.x;
    const unsigned long long row0 = (unsigned long long)row * nsym;           // address of the row's first index
    const unsigned long long A = ((row0 >> 4) + g) << 4;                      // this thread's aligned group
    const long long q0 = (long long)A - (long long)row0;                      // its first symbol (may be < 0)
    if (q0 >= (long long)nsym) return;
    const long long row_bytes = (nbits + 7) >> 3;
    const uint8_t* src = bits + (size_t)row * row_bytes;
    const long long bi0 = q0 * BPS, by = bi0 >> 3;
    const unsigned long long addr = (unsigned long long)(src + by);
    const unsigned long long a0 = addr & ~3ull;
    // whole group inside the row and its bit budget, and the four aligned words inside the buffer
    const bool fast = q0 >= 0 && q0 + 15 < (long long)nsym && (q0 + 16) * BPS <= nbits &&
                      a0 >= (unsigned long long)bits && (long long)(a0 - (unsigned long long)bits) + 16 <= total_bytes;
    if (!fast) {
        for (int k = 0; k < 16; ++k) {
            const long long q = q0 + k;
            if (q >= 0 && q < (long long)nsym) idx[A + k] = (uint8_t)packed_symbol(src, row_bytes, nbits, q, BPS);
        }
        return;
    }
    const unsigned* wp = (const unsigned*)a0;
    unsigned W0 = __byte_perm(__ldg(wp), 0, 0x0123), W1 = __byte_perm(__ldg(wp + 1), 0, 0x0123);
    unsigned W2 = __byte_perm(__ldg(wp + 2), 0, 0x0123), W3 = __byte_perm(__ldg(wp + 3), 0, 0x0123);
    const unsigned off = (unsigned)(addr & 3ull) * 8u + (unsigned)(bi0 & 7);  // < 32
    const unsigned V[3] = {__funnelshift_l(W1, W0, off), __funnelshift_l(W2, W1, off), __funnelshift_l(W3, W2, off)};
    constexpr unsigned m = (1u << BPS) - 1u;
    unsigned out[4];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        unsigned v = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int pos = (4 * o + i) * BPS;                 // bit position of the symbol, compile time
            const int wi = pos >> 5, sh = pos & 31;
            unsigned x;
            if (sh + BPS <= 32) x = V[wi] >> (32 - sh - BPS);
            else x = __funnelshift_l(V[wi + 1 < 3 ? wi + 1 : 2], V[wi], sh) >> (32 - BPS);
            v |= (x & m) << (8 * i);
        }
        out[o] = v;
    }
    *(uint4*)(idx + A) = make_uint4(out[0], out[1], out[2], out[3]);
}

// core/modulator.py:109-110: format(idx, '0{b}b'); output truncated to nbits.
__global__ void indices_to_bits_kernel(const uint8_t* __restrict__ idx, long long nsym, uint8_t* __restrict__ bits,
                                       long long nbits, int bps, long long total) {
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const long long b = g / nbits, bi = g % nbits;
        const long long q = bi / bps;
        const int i = (int)(bi % bps);
        const int v = q < nsym ? idx[(size_t)b * nsym + q] : 0;
        bits[g] = (uint8_t)((v >> (bps - 1 - i)) & 1);
    }
}

static unsigned grid_for(long long total, int threads) {
    long long g = (total + threads - 1) / threads;
    if (g > 148LL * 32) g = 148LL * 32;
    if (g < 1) g = 1;
    return (unsigned)g;
}

extern "C" int lte_bits_to_indices(const lte_plan* p, const uint8_t* bits, int64_t nbits, uint8_t* idx,
                                   int64_t nsym, int32_t B, void* stream) {
    if (!p || !bits || !idx || nsym < 1 || B < 0) return LTE_ERR_INVALID_ARG;
    if (B == 0) return LTE_OK;
    const int packed = nbits < 0;        // negative nbits: rows are np.packbits() bytes
    const long long nb = packed ? -nbits : nbits;
    if (nsym >= (1ll << 31)) return LTE_ERR_UNSUPPORTED;
    if (packed && ((uintptr_t)idx & 15) == 0 && (p->dev.bps == 2 || p->dev.bps == 4 || p->dev.bps == 6)) {
        const unsigned ch16 = (unsigned)(((nsym + 15) / 16 + 1 + 255) / 256);  // + 1: a row may straddle one more group
        if ((long long)ch16 * B >= (1ll << 31)) return LTE_ERR_UNSUPPORTED;
        const long long total_bytes = (long long)B * ((nb + 7) >> 3);
        const unsigned grid16 = (unsigned)((long long)ch16 * B);
        cudaStream_t st = (cudaStream_t)stream;
        if (p->dev.bps == 2) bits_to_indices_x16_kernel<2><<<grid16, 256, 0, st>>>(bits, nb, idx, (unsigned)nsym, ch16, total_bytes);
        else if (p->dev.bps == 4) bits_to_indices_x16_kernel<4><<<grid16, 256, 0, st>>>(bits, nb, idx, (unsigned)nsym, ch16, total_bytes);
        else bits_to_indices_x16_kernel<6><<<grid16, 256, 0, st>>>(bits, nb, idx, (unsigned)nsym, ch16, total_bytes);
        LTE_CHECK_CUDA(cudaGetLastError());
        return LTE_OK;
    }
    if (packed && ((uintptr_t)idx & 3) == 0) {
        const unsigned ch4 = (unsigned)(((nsym + 3) / 4 + 1 + 255) / 256);     // + 1: a row may straddle one more group
        if ((long long)ch4 * B >= (1ll << 31)) return LTE_ERR_UNSUPPORTED;
        bits_to_indices_x4_kernel<<<(unsigned)((long long)ch4 * B), 256, 0, (cudaStream_t)stream>>>(
            bits, nb, idx, (unsigned)nsym, p->dev.bps, ch4);
        LTE_CHECK_CUDA(cudaGetLastError());
        return LTE_OK;
    }
    const unsigned chunks = (unsigned)((nsym + 255) / 256);
    const long long grid = (long long)chunks * B;
    if (grid >= (1ll << 31)) return LTE_ERR_UNSUPPORTED;
    bits_to_indices_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(bits, nb, packed, idx, (unsigned)nsym,
                                                                            p->dev.bps, chunks);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

extern "C" int lte_indices_to_bits(const lte_plan* p, const uint8_t* idx, int64_t nsym, uint8_t* bits,
                                   int64_t nbits, int32_t B, void* stream) {
    if (!p || !bits || !idx || nsym < 1 || nbits < 1 || B < 0) return LTE_ERR_INVALID_ARG;
    if (B == 0) return LTE_OK;
    const long long total = (long long)B * nbits;
    indices_to_bits_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(idx, nsym, bits, nbits,
                                                                                  p->dev.bps, total);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

// core/modulator.py:84-86: constellation[idx], without the grid / IFFT (SC-FDM precoder input)
__global__ void qam_map_kernel(const DevPlan P, const uint8_t* __restrict__ idx, float2* __restrict__ out, long long n) {
    const int h = P.bps >> 1, mask = (1 << h) - 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int v = idx[i];
        out[i] = make_float2(P.lev[(v >> h) & mask], P.lev[v & mask]);
    }
}

extern "C" int lte_qam_map(const lte_plan* p, const uint8_t* idx, lte_c32* symbols, int64_t n, void* stream) {
    if (!p || !idx || !symbols || n < 0) return LTE_ERR_INVALID_ARG;
    if (n == 0) return LTE_OK;
    qam_map_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(p->dev, idx, (float2*)symbols, n);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

// ------------------------------------------------------------------------------ stage 4
// core/lte_receiver.py:62-87 (LS at the pilots) and :98-133 (edge hold + np.linspace).
// One CTA per (row, slot); the estimate comes from the slot's first symbol (:380-406).
template <bool NOISY>
__global__ void __launch_bounds__(256)
crs_ls_interp_kernel(const DevPlan P, const float2* __restrict__ Y, float2* __restrict__ H, int k0, int nk,
                     int set, int S, int nslot, const AwgnArgs A) {
    extern __shared__ float2 hp[];      // LS estimates at the owned pilots
    const long long row = blockIdx.x / nslot;
    const int slot = blockIdx.x % nslot;
    const float2* y = Y + ((size_t)row * S + (size_t)slot * LTE_SLOT_SYMBOLS) * nk;
    const int cnt = P.pset_cnt[set];
    const int16_t* pbin = P.pset_bin + (size_t)set * P.Np;
    const float2* pinv = P.pset_inv + (size_t)set * P.Np;
    float sigma = 0.f;
    if (NOISY) sigma = lte_sigma(A.power[row], A.n_stream, A.snr_lin[row]);
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        float2 yp = y[pbin[i] - k0];
        if (NOISY) yp = awgn_at(A, sigma, row, slot * LTE_SLOT_SYMBOLS, P.N, pbin[i], yp);
        hp[i] = cmul(yp, pinv[i]);
    }
    __syncthreads();
    // np.linspace step of every pilot-to-pilot segment, (b - a) / div, once per segment
    float2* step = hp + P.Np;
    for (int i = threadIdx.x; i < cnt - 1; i += blockDim.x) {
        const float div = (float)(pbin[i + 1] - pbin[i]);
        step[i] = make_float2(__fdiv_rn(hp[i + 1].x - hp[i].x, div), __fdiv_rn(hp[i + 1].y - hp[i].y, div));
    }
    __syncthreads();
    const int16_t* seg = P.pset_seg + (size_t)set * P.N;
    float2* h = H + ((size_t)row * nslot + slot) * nk;
    for (int kk = threadIdx.x; kk < nk; kk += blockDim.x) {
        const int k = kk + k0;
        const int lo = seg[k];
        float2 v;
        if (lo < 0) v = hp[0];
        else if (lo >= cnt - 1) v = hp[cnt - 1];
        else {
            const int i1 = pbin[lo];
            const float2 a = hp[lo], st = step[lo];
            const float t = (float)(k - i1);
            // np.linspace: start + i * (delta / div)
            v = make_float2(fmaf(t, st.x, a.x), fmaf(t, st.y, a.y));
            if (k == i1) v = a;
        }
        h[kk] = v;
    }
}

static int launch_crs(const lte_plan* p, const lte_c32* Y, lte_c32* H, int window, int pilot_set, int64_t rows,
                      int32_t S, const lte_awgn_desc* awgn, void* stream) {
    if (!p || !Y || !H || rows < 0 || S < 1) return LTE_ERR_INVALID_ARG;
    if (p->dev.Np == 0 || pilot_set < 0 || pilot_set >= p->nsets) return LTE_ERR_INVALID_ARG;
    int32_t k0, nk;
    int rc = lte_plan_window(p, window, &k0, &nk);
    if (rc) return rc;
    AwgnArgs A = {};
    if (awgn && (rc = make_awgn_args(A, p, awgn, S, rows))) return rc;
    if (rows == 0) return LTE_OK;
    const int nslot = (S + LTE_SLOT_SYMBOLS - 1) / LTE_SLOT_SYMBOLS;
    const unsigned grid = (unsigned)(rows * nslot);
    const size_t smem = 2 * sizeof(float2) * p->dev.Np;
    if (awgn)
        crs_ls_interp_kernel<true><<<grid, 128, smem, (cudaStream_t)stream>>>(p->dev, (const float2*)Y, (float2*)H, k0,
                                                                             nk, pilot_set, S, nslot, A);
    else
        crs_ls_interp_kernel<false><<<grid, 128, smem, (cudaStream_t)stream>>>(p->dev, (const float2*)Y, (float2*)H, k0,
                                                                              nk, pilot_set, S, nslot, A);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

extern "C" int lte_crs_ls_interp(const lte_plan* p, const lte_c32* Y, lte_c32* H, int window, int pilot_set,
                                 int64_t rows, int32_t S, void* stream) {
    return launch_crs(p, Y, H, window, pilot_set, rows, S, nullptr, stream);
}

extern "C" int lte_crs_ls_interp_awgn(const lte_plan* p, const lte_c32* Y, lte_c32* H, int window, int pilot_set,
                                      int64_t rows, int32_t S, const lte_awgn_desc* awgn, void* stream) {
    if (!awgn) return LTE_ERR_INVALID_ARG;
    return launch_crs(p, Y, H, window, pilot_set, rows, S, awgn, stream);
}

// ------------------------------------------------------------------------------ stage 6
__global__ void __launch_bounds__(256)
demap_count_kernel(const DevPlan P, const float2* __restrict__ syms, const uint8_t* __restrict__ idx_tx,
                   uint8_t* __restrict__ idx_rx, unsigned long long* __restrict__ errors, long long nsym,
                   long long nbits, int gx) {
    const long long b = blockIdx.x / gx;
    const int bx = blockIdx.x % gx;
    unsigned int e = 0;
    for (long long q = (long long)bx * blockDim.x + threadIdx.x; q < nsym; q += (long long)gx * blockDim.x) {
        const size_t o = (size_t)b * nsym + q;
        const int d = slice_symbol(P, syms[o]);
        if (idx_rx) idx_rx[o] = (uint8_t)d;
        if (idx_tx) e += bit_errors(d, idx_tx[o], P.bps, nbits - q * P.bps);
    }
    if (errors) block_add_errors(e, &errors[b]);
}

extern "C" int lte_demap_count(const lte_plan* p, const lte_c32* syms, const uint8_t* idx_tx, uint8_t* idx_rx,
                               unsigned long long* errors, int64_t nsym, int64_t nbits, int64_t B, void* stream) {
    if (!p || !syms || nsym < 1 || B < 0 || (idx_tx && !errors)) return LTE_ERR_INVALID_ARG;
    if (B == 0) return LTE_OK;
    int gx = (int)((nsym + 256 * 4 - 1) / (256 * 4));
    if (gx < 1) gx = 1;
    demap_count_kernel<<<(unsigned)((long long)gx * B), 256, 0, (cudaStream_t)stream>>>(
        p->dev, (const float2*)syms, idx_tx, idx_rx, idx_tx ? errors : nullptr, nsym, nbits, gx);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

// ------------------------------------------------------------------------------ stage 5
// ZF: Y / (H + 1e-6) (core/lte_receiver.py:174), gathered at the data bins (:303-316).
__global__ void __launch_bounds__(256)
zf_kernel(const DevPlan P, const float2* __restrict__ Y, const float2* __restrict__ H, float2* __restrict__ out,
          int k0, int nk, int S, int nslot, long long total, const AwgnArgs A, int noisy) {
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(g % P.Nd);
        const long long bs = g / P.Nd;            // b*S + s
        const int s = (int)(bs % S);
        const long long b = bs / S;
        const int kk = P.data_idx[d] - k0;
        float2 y = Y[(size_t)bs * nk + kk];
        if (noisy) y = awgn_at(A, lte_sigma(A.power[b], A.n_stream, A.snr_lin[b]), b, s, P.N, kk + k0, y);   // lazy AWGN
        if (H) {
            float2 h = H[((size_t)b * nslot + s / LTE_SLOT_SYMBOLS) * nk + kk];
            h.x += 1e-6f;
            y = cdiv(y, h);
        }
        out[g] = y;
    }
}

static int launch_zf(const lte_plan* p, const lte_c32* Y, const lte_c32* H, lte_c32* out, int window, int64_t B,
                     int32_t S, const lte_awgn_desc* awgn, void* stream) {
    if (!p || !Y || !out || B < 0 || S < 1) return LTE_ERR_INVALID_ARG;
    int32_t k0, nk;
    int rc = lte_plan_window(p, window, &k0, &nk);
    if (rc) return rc;
    AwgnArgs A = {};
    if (awgn && (rc = make_awgn_args(A, p, awgn, S, B))) return rc;
    if (B == 0) return LTE_OK;
    const long long total = (long long)B * S * p->dev.Nd;
    const int nslot = (S + LTE_SLOT_SYMBOLS - 1) / LTE_SLOT_SYMBOLS;
    zf_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(p->dev, (const float2*)Y, (const float2*)H,
                                                                     (float2*)out, k0, nk, S, nslot, total, A, awgn ? 1 : 0);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

extern "C" int lte_equalize_zf(const lte_plan* p, const lte_c32* Y, const lte_c32* H, lte_c32* out, int window,
                               int64_t B, int32_t S, void* stream) {
    return launch_zf(p, Y, H, out, window, B, S, nullptr, stream);
}

extern "C" int lte_equalize_zf_awgn(const lte_plan* p, const lte_c32* Y, const lte_c32* H, lte_c32* out, int window,
                                    int64_t B, int32_t S, const lte_awgn_desc* awgn, void* stream) {
    if (!awgn) return LTE_ERR_INVALID_ARG;
    return launch_zf(p, Y, H, out, window, B, S, awgn, stream);
}

// MRC (core/ofdm_core.py:1484-1532): thread = (stream, data bin); H of the slot is held in
// registers while the thread walks the slot's symbols, so H is read once per 14 symbols.
#ifndef MRC_CHUNK
#define MRC_CHUNK 1
#endif
template <int R, bool COUNT, bool NOISY>
__global__ void __launch_bounds__(128, NOISY ? 8 : 1)
mrc_kernel(const DevPlan P, const float2* __restrict__ Y, const float2* __restrict__ H, float2* __restrict__ out,
           const uint8_t* __restrict__ idx_tx, unsigned long long* __restrict__ errors, int k0, int nk, int S,
           int nslot, long long nbits, int gx, int sps, const AwgnArgs A) {
    // block = (stream b, part of a slot, chunk of 128 data bins); a part is `sps` consecutive symbols
    const int pps = LTE_SLOT_SYMBOLS / sps, nparts = nslot * pps;
    const int chunk = blockIdx.x % gx;
    const long long bp = blockIdx.x / gx;
    const long long b = bp / nparts;
    const int part = (int)(bp - b * nparts);
    const int d = chunk * blockDim.x + threadIdx.x;
    unsigned int e = 0;
    if (d < P.Nd) {
        const int kb = P.data_idx[d];
        const int kk = kb - k0;
        float sigma[R];
        if (NOISY) {
#pragma unroll
            for (int r = 0; r < R; ++r) sigma[r] = lte_sigma(A.power[b * R + r], A.n_stream, A.snr_lin[b * R + r]);
        }
        {
            const int slot = part / pps;
            float2 h[R];
            float den = 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                h[r] = H[(((size_t)b * R + r) * nslot + slot) * nk + kk];
                den += cabs2(h[r]);
            }
            den += 1e-10f;
            const float inv_den = __frcp_rn(den);
            // combined mode: sum_r conj(h_r) sigma_r w_r with independent unit normals w_r is one complex
            // normal of standard deviation sqrt(sum_r |h_r|^2 sigma_r^2) per component
            float csig = 0.f;
            const bool comb = NOISY && A.combine;
            if (comb) {
#pragma unroll
                for (int r = 0; r < R; ++r) csig = fmaf(cabs2(h[r]), sigma[r] * sigma[r], csig);
                csig = sqrtf(csig);
            }
            const int s0 = slot * LTE_SLOT_SYMBOLS + (part - slot * pps) * sps;
            const int s_end = min(S, s0 + sps);
            // register double buffering: the loads of the next group of symbols are in flight while this
            // group is combined, sliced and (noisy variants) gets its Philox / Box-Muller samples
            const float2* yb = Y + ((size_t)b * R * S) * nk + kk;
            const int sym_bits = P.Nd * P.bps;
            const long long valid0 = nbits - ((long long)s0 * P.Nd + d) * P.bps;   // bits left from symbol s0 on
            // running pointers (one per antenna) instead of 64-bit index arithmetic per load
            const float2* yp[R];
#pragma unroll
            for (int r = 0; r < R; ++r) yp[r] = yb + ((size_t)r * S + s0) * nk;
            const uint8_t* ip = COUNT ? idx_tx + ((size_t)b * S + s0) * P.Nd + d : nullptr;
            // every bit of every symbol of this part is inside the first nbits: no per-symbol 64-bit bit budget
            const bool all_valid = valid0 - (long long)(s_end - 1 - s0) * sym_bits >= P.bps;
            // the loads of the next symbol are in flight while this one is combined, sliced and (noisy variants) gets
            // its Philox / Box-Muller sample
            auto fetch = [&](float2 (&y)[R], uint8_t& in, int s) {
                in = 0;
                if (s < s_end) {
#pragma unroll
                    for (int r = 0; r < R; ++r) { y[r] = *yp[r]; yp[r] += nk; }
                    if (COUNT) { in = *ip; ip += P.Nd; }
                }
            };
            auto combine = [&](float2 (&y)[R], uint8_t in, int s) {
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (NOISY && !comb) y[r] = awgn_at(A, sigma[r], b * R + r, s, P.N, kb, y[r]);
                    const float2 t = cmulc(h[r], y[r]);
                    acc.x += t.x;
                    acc.y += t.y;
                }
                if (comb) acc = awgn_at(A, csig, b * R, s, P.N, kb, acc);
                const float2 cc = make_float2(acc.x * inv_den, acc.y * inv_den);
                if (COUNT) {
                    const int dec = slice_symbol(P, cc);
                    e += all_valid ? __popc(dec ^ (int)in)
                                   : bit_errors(dec, in, P.bps, valid0 - (long long)(s - s0) * sym_bits);
                } else {
                    out[((size_t)b * S + s) * P.Nd + d] = cc;
                }
            };
            float2 ya[R], yc[R];
            uint8_t ia, ic;
            fetch(ya, ia, s0);
            if constexpr (!NOISY) {
                // plain MRC is bandwidth bound: two register sets in ping-pong, no set is ever copied
                for (int sg = s0; sg < s_end; sg += 2) {
                    fetch(yc, ic, sg + 1);
                    combine(ya, ia, sg);
                    fetch(ya, ia, sg + 2);
                    if (sg + 1 < s_end) combine(yc, ic, sg + 1);
                }
            } else {
                // the lazy-AWGN variants are issue bound and live on occupancy (72 registers, 7 CTAs / SM): one
                // prefetched set, copied into the working set every symbol
                for (int sg = s0; sg < s_end; ++sg) {
#pragma unroll
                    for (int r = 0; r < R; ++r) yc[r] = ya[r];
                    ic = ia;
                    fetch(ya, ia, sg + 1);
                    combine(yc, ic, sg);
                }
            }
        }
    }
    if (COUNT) block_add_errors(e, &errors[b]);
}

template <bool COUNT>
static int launch_mrc(const lte_plan* p, const lte_c32* Y, const lte_c32* H, lte_c32* out, const uint8_t* idx_tx,
                      unsigned long long* errors, int window, int64_t nbits, int64_t B, int32_t R, int32_t S,
                      const lte_awgn_desc* awgn, void* stream) {
    int32_t k0, nk;
    int rc = lte_plan_window(p, window, &k0, &nk);
    if (rc) return rc;
    AwgnArgs A = {};
    if (awgn && (rc = make_awgn_args(A, p, awgn, S, B * R))) return rc;
    if (B == 0) return LTE_OK;
    const int nslot = (S + LTE_SLOT_SYMBOLS - 1) / LTE_SLOT_SYMBOLS;
    const int gx = (p->dev.Nd + 127) / 128;
    // symbols per thread: H of the slot is re-read once per part (from L2); short parts expose more
    // parallelism, which the noisy variants need to hide the generator's dependent chains
    const int sps = LTE_SLOT_SYMBOLS;
    const long long grid_ll = (long long)gx * B * nslot * (LTE_SLOT_SYMBOLS / sps);
    if (grid_ll >= (1ll << 31)) return LTE_ERR_UNSUPPORTED;
    const unsigned grid = (unsigned)grid_ll;
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_MRC(RR)                                                                                          \
    case RR:                                                                                                    \
        if (awgn)                                                                                               \
            mrc_kernel<RR, COUNT, true><<<grid, 128, 0, st>>>(p->dev, (const float2*)Y, (const float2*)H,      \
                                                              (float2*)out, idx_tx, errors, k0, nk, S, nslot,  \
                                                              nbits, gx, sps, A);                                    \
        else                                                                                                    \
            mrc_kernel<RR, COUNT, false><<<grid, 128, 0, st>>>(p->dev, (const float2*)Y, (const float2*)H,     \
                                                               (float2*)out, idx_tx, errors, k0, nk, S, nslot, \
                                                               nbits, gx, sps, A);                                   \
        break;
    switch (R) {
        LAUNCH_MRC(1) LAUNCH_MRC(2) LAUNCH_MRC(3) LAUNCH_MRC(4) LAUNCH_MRC(5) LAUNCH_MRC(6) LAUNCH_MRC(7)
        LAUNCH_MRC(8)
        default: return LTE_ERR_INVALID_ARG;
    }
#undef LAUNCH_MRC
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

extern "C" int lte_equalize_mrc(const lte_plan* p, const lte_c32* Y, const lte_c32* H, lte_c32* out, int window,
                                int64_t B, int32_t R, int32_t S, void* stream) {
    if (!p || !Y || !H || !out || B < 0 || S < 1) return LTE_ERR_INVALID_ARG;
    return launch_mrc<false>(p, Y, H, out, nullptr, nullptr, window, 0, B, R, S, nullptr, stream);
}

extern "C" int lte_mrc_demap_count(const lte_plan* p, const lte_c32* Y, const lte_c32* H, const uint8_t* idx_tx,
                                   unsigned long long* errors, int window, int64_t nbits, int64_t B, int32_t R,
                                   int32_t S, void* stream) {
    if (!p || !Y || !H || !idx_tx || !errors || B < 0 || S < 1) return LTE_ERR_INVALID_ARG;
    return launch_mrc<true>(p, Y, H, nullptr, idx_tx, errors, window, nbits, B, R, S, nullptr, stream);
}

extern "C" int lte_mrc_demap_count_awgn(const lte_plan* p, const lte_c32* Y, const lte_c32* H, const uint8_t* idx_tx,
                                        unsigned long long* errors, int window, int64_t nbits, int64_t B, int32_t R,
                                        int32_t S, const lte_awgn_desc* awgn, void* stream) {
    if (!p || !Y || !H || !idx_tx || !errors || !awgn || B < 0 || S < 1) return LTE_ERR_INVALID_ARG;
    return launch_mrc<true>(p, Y, H, nullptr, idx_tx, errors, window, nbits, B, R, S, awgn, stream);
}

// ------------------------------------------------------------------------------ compact sweep layout
// The spectral link (spectral.cu) leaves the grid as exactly the elements the receiver reads:
//   Yd [rows][S][2 ndp]       data symbol d of the OFDM symbol at element d (ndp = ceil(Nd / 2); with an odd Nd
//                             the last element of a row is padding)
//   Yp [rows][nslot][2 npp]   pilot i of every slot's first symbol at element i (npp = ceil(Np / 2))

// LS estimate at the pilot positions only (core/lte_receiver.py:62-87 without the interpolation, which the
// compact MRC kernel does per data bin): Hp [rows][nslot][Np] = (Yp [+ lazy AWGN]) / pilot.  One thread per pilot.
template <bool NOISY>
__global__ void __launch_bounds__(256)
crs_ls_compact_kernel(const DevPlan P, const float2* __restrict__ Yp, float2* __restrict__ Hp, int nslot, long long total,
                      const AwgnArgs A) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    const int i = (int)(g % P.Np);
    const long long rs = g / P.Np;                  // row * nslot + slot
    const int slot = (int)(rs % nslot);
    const long long row = rs / nslot;
    float2 yp = Yp[(size_t)rs * (2 * P.npp) + i];
    if (NOISY) {
        const float sigma = lte_sigma(A.power[row], A.n_stream, A.snr_lin[row]);
        yp = awgn_at(A, sigma, row, slot * LTE_SLOT_SYMBOLS, P.N, P.pset_bin[i], yp);
    }
    Hp[g] = cmul(yp, P.pset_inv[i]);
}

extern "C" int lte_crs_ls_compact(const lte_plan* p, const lte_c32* Ypilot, lte_c32* Hp, int64_t rows, int32_t S,
                                  const lte_awgn_desc* awgn, void* stream) {
    if (!p || !Ypilot || !Hp || rows < 0 || S < 1) return LTE_ERR_INVALID_ARG;
    if (p->dev.Np == 0 || p->nsets != 1) return LTE_ERR_UNSUPPORTED;
    AwgnArgs A = {};
    int rc;
    if (awgn && (rc = make_awgn_args(A, p, awgn, S, rows))) return rc;
    if (rows == 0) return LTE_OK;
    const int nslot = (S + LTE_SLOT_SYMBOLS - 1) / LTE_SLOT_SYMBOLS;
    const long long total = (long long)rows * nslot * p->dev.Np;
    const long long grid = (total + 255) / 256;
    if (grid >= (1ll << 31)) return LTE_ERR_UNSUPPORTED;
    if (awgn)
        crs_ls_compact_kernel<true><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(p->dev, (const float2*)Ypilot,
                                                                                      (float2*)Hp, nslot, total, A);
    else
        crs_ls_compact_kernel<false><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(p->dev, (const float2*)Ypilot,
                                                                                       (float2*)Hp, nslot, total, A);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

// MRC + slicer + bit-error count on the compact layout (core/ofdm_core.py:1484-1532, core/modulator.py:90-112,
// core/ofdm_core.py:245-268).  thread = (stream, data symbol d): it interpolates H of its bin between the two
// neighbouring pilots with the very operations of crs_ls_interp_kernel (edge hold, start + i * (delta / div)),
// holds it for the slot, and walks the slot's symbols with one 64-bit load per antenna and symbol (a warp reads 256
// contiguous bytes).  The lazy-AWGN variant is bound by the noise generator (Philox + Box-Muller per output) and
// lives on occupancy like mrc_kernel: one bin per thread, 64 registers, 8 CTAs / SM.  Measured alternatives on
// the headline shape: planar pairs (re0, re1, im0, im1) read as two 32-bit loads 0.61 ms, two bins per thread with
// 128-bit loads 0.61 ms, this layout 0.46 ms.  The draws are those of mrc_kernel, so the counts are bit-identical
// to the windowed layout.
// FULL: every bit of every symbol lies inside the first nbits (the sweep's case: nbits = S Nd bps), so the
// per-symbol 64-bit bit budget -- which the compiler otherwise re-derives for every output at 64 registers --
// is not needed at all: errors are a plain XOR + popcount.
// NOISE: 0 = Y already carries its noise, 1 = lazy AWGN per antenna, 2 = one combined draw per output
#define MRCC_PILOTS 40          // pilots a CTA's 128 consecutive data bins may interpolate between (LTE CRS: 27)
template <int R, int NOISE, bool FULL>
__global__ void __launch_bounds__(128, NOISE ? 8 : 1)
mrc_compact_kernel(const DevPlan P, const float2* __restrict__ Yd, const float2* __restrict__ Hp,
                   const float2* __restrict__ Yp, const uint8_t* __restrict__ idx_tx,
                   unsigned long long* __restrict__ errors, int S, int nslot, long long nbits, int gx, const AwgnArgs A) {
    // (estimate, slope to the next pilot) of the pilots this CTA's 128 bins interpolate between, per antenna, and
    // the noise sigma of every antenna: worked out once per CTA instead of once per thread (the divisions are
    // the IEEE ones of crs_ls_interp_kernel, so the interpolated values are unchanged bit for bit)
    __shared__ float4 seg_s[R * MRCC_PILOTS];
    __shared__ float2 ls_s[R * (MRCC_PILOTS + 1)];              // Yp != NULL: the LS estimates formed here (Hp unused)
    __shared__ float sigma_s[R];
    const int chunk = blockIdx.x % gx;
    const long long b = blockIdx.x / gx;
    const int tid = threadIdx.x;
    const int d0 = chunk * blockDim.x;
    const bool act = d0 + tid < P.Nd;
    const int d = act ? d0 + tid : P.Nd - 1;
    unsigned int e = 0;
    constexpr bool NOISY = NOISE != 0, comb = NOISE == 2;
    const int cnt = P.pset_cnt[0];
    auto segment = [&](int bin) { const int lo = P.pset_seg[bin]; return lo < 0 ? 0 : (lo >= cnt - 1 ? cnt - 1 : lo); };
    const int la0 = segment(P.data_idx[d0]);
    const int npil = segment(P.data_idx[min(d0 + (int)blockDim.x, P.Nd) - 1]) - la0 + 1;    // <= MRCC_PILOTS (host check)
    if (NOISY && tid < R) sigma_s[tid] = lte_sigma(A.power[b * R + tid], A.n_stream, A.snr_lin[b * R + tid]);
    const int kb = P.data_idx[d];
    const int lo = P.pset_seg[kb];
    const int la = segment(kb);
    const bool inner = lo >= 0 && lo < cnt - 1;
    const float t = inner ? (float)(kb - P.pset_bin[la]) : 0.f;     // edge bins hold their pilot: a + 0 * slope
    const int sym_bits = P.Nd * P.bps;
    const int ystride = 2 * P.ndp;
    for (int slot = 0; slot < nslot; ++slot) {
        __syncthreads();                                            // the previous slot's table has been consumed
        if (Yp) {
            // the LS step of crs_ls_compact_kernel for the npil + 1 pilots this CTA needs, with its very operations
            // (and lazy-AWGN draws): no Hp tensor, no estimator launch
            for (int q = tid; q < (npil + 1) * R; q += blockDim.x) {
                const int r = q / (npil + 1), i = la0 + q - r * (npil + 1);
                if (i < cnt) {
                    const long long row = b * R + r;
                    float2 yp = Yp[((size_t)row * nslot + slot) * (2 * P.npp) + i];
                    if (NOISY) yp = awgn_at(A, sigma_s[r], row, slot * LTE_SLOT_SYMBOLS, P.N, P.pset_bin[i], yp);
                    ls_s[r * (MRCC_PILOTS + 1) + i - la0] = cmul(yp, P.pset_inv[i]);
                }
            }
            __syncthreads();
        }
        for (int q = tid; q < npil * R; q += blockDim.x) {
            const int r = q / npil, i = la0 + q - r * npil;
            const float2* hp = Yp ? ls_s + r * (MRCC_PILOTS + 1) - la0 : Hp + (((size_t)b * R + r) * nslot + slot) * P.Np;
            const float2 a = hp[i];
            float2 sl = make_float2(0.f, 0.f);
            if (i < cnt - 1) {
                const float2 c = hp[i + 1];
                const float div = (float)(P.pset_bin[i + 1] - P.pset_bin[i]);
                sl = make_float2(__fdiv_rn(c.x - a.x, div), __fdiv_rn(c.y - a.y, div));
            }
            seg_s[r * MRCC_PILOTS + i - la0] = make_float4(a.x, a.y, sl.x, sl.y);
        }
        __syncthreads();
        {
            float sigma[R];
            if (NOISY) {
#pragma unroll
                for (int r = 0; r < R; ++r) sigma[r] = sigma_s[r];
            }
            float2 h[R];
            float den = 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 as = seg_s[r * MRCC_PILOTS + la - la0];
                h[r] = make_float2(fmaf(t, as.z, as.x), fmaf(t, as.w, as.y));
                den += cabs2(h[r]);
            }
            den += 1e-10f;
            const float inv_den = __frcp_rn(den);
            float csig = 0.f;
            if (comb) {
#pragma unroll
                for (int r = 0; r < R; ++r) csig = fmaf(cabs2(h[r]), sigma[r] * sigma[r], csig);
                csig = sqrtf(csig);
            }
            const int s0 = slot * LTE_SLOT_SYMBOLS, s_end = min(S, s0 + LTE_SLOT_SYMBOLS);
            const float2* yp[R];
#pragma unroll
            for (int r = 0; r < R; ++r) yp[r] = Yd + (((size_t)b * R + r) * S + s0) * ystride + d;
            const uint8_t* ip = idx_tx + ((size_t)b * S + s0) * P.Nd + d;
            const long long valid0 = FULL ? 0 : nbits - ((long long)s0 * P.Nd + d) * P.bps;   // bits left from (s0, d) on
            auto fetch = [&](float2 (&y)[R], uint8_t& in, int s) {
                in = 0;
                if (s < s_end) {
#pragma unroll
                    for (int r = 0; r < R; ++r) { y[r] = *yp[r]; yp[r] += ystride; }
                    in = *ip;
                    ip += P.Nd;
                }
            };
            auto combine = [&](float2 (&y)[R], uint8_t in, int s) {
                float2 acc = make_float2(0.f, 0.f);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (NOISY && !comb) y[r] = awgn_at(A, sigma[r], b * R + r, s, P.N, kb, y[r]);
                    const float2 tt = cmulc(h[r], y[r]);
                    acc.x += tt.x;
                    acc.y += tt.y;
                }
                if (comb) acc = awgn_at(A, csig, b * R, s, P.N, kb, acc);
                const int dec = slice_symbol(P, make_float2(acc.x * inv_den, acc.y * inv_den));
                if (FULL) e += __popc(dec ^ (int)in);
                else e += bit_errors(dec, in, P.bps, valid0 - (long long)(s - s0) * sym_bits);
            };
            float2 ya[R], yc[R];
            uint8_t ia, ic;
            fetch(ya, ia, s0);
            if constexpr (!NOISY) {
                // plain MRC is bandwidth bound: two register sets in ping-pong, no set is ever copied
                for (int sg = s0; sg < s_end; sg += 2) {
                    fetch(yc, ic, sg + 1);
                    combine(ya, ia, sg);
                    fetch(ya, ia, sg + 2);
                    if (sg + 1 < s_end) combine(yc, ic, sg + 1);
                }
            } else {
                // the lazy-AWGN variants are issue bound and live on occupancy: one prefetched set, copied into the
                // working set every symbol
                for (int sg = s0; sg < s_end; ++sg) {
#pragma unroll
                    for (int r = 0; r < R; ++r) yc[r] = ya[r];
                    ic = ia;
                    fetch(ya, ia, sg + 1);
                    combine(yc, ic, sg);
                }
            }
        }
    }
    block_add_errors(act ? e : 0u, &errors[b]);
}

static int mrc_compact_launch(const lte_plan* p, const lte_c32* Ydata, const lte_c32* Hp, const lte_c32* Ypilot,
                              const uint8_t* idx_tx, unsigned long long* errors, int64_t nbits, int64_t B,
                              int32_t R, int32_t S, const lte_awgn_desc* awgn, void* stream) {
    if (!p || !Ydata || (!Hp == !Ypilot) || !idx_tx || !errors || B < 0 || S < 1) return LTE_ERR_INVALID_ARG;
    if (p->dev.Np == 0 || p->nsets != 1) return LTE_ERR_UNSUPPORTED;
    AwgnArgs A = {};
    int rc;
    if (awgn && (rc = make_awgn_args(A, p, awgn, S, B * R))) return rc;
    if (B == 0) return LTE_OK;
    const int nslot = (S + LTE_SLOT_SYMBOLS - 1) / LTE_SLOT_SYMBOLS;
    const int gx = (p->dev.Nd + 127) / 128;
    const long long grid_ll = (long long)gx * B;
    if (grid_ll >= (1ll << 31)) return LTE_ERR_UNSUPPORTED;
    {   // the kernel's shared pilot table holds MRCC_PILOTS entries per antenna: pilot grids denser than that
        // per 128 data bins (no LTE numerology) are refused rather than overrun
        const std::vector<int32_t>& pil = p->pilot_idx_h;
        const std::vector<int32_t>& dat = p->data_idx_h;
        auto segment = [&](int bin) {
            int lo = (int)(std::upper_bound(pil.begin(), pil.end(), bin) - pil.begin()) - 1;
            return lo < 0 ? 0 : lo;
        };
        for (int c = 0; c < gx; ++c) {
            const int last = std::min((c + 1) * 128, (int)dat.size()) - 1;
            if (segment(dat[last]) - segment(dat[c * 128]) + 1 > MRCC_PILOTS) return LTE_ERR_UNSUPPORTED;
        }
    }
    const unsigned grid = (unsigned)grid_ll;
    cudaStream_t st = (cudaStream_t)stream;
    const bool full = nbits >= (int64_t)S * p->dev.Nd * p->dev.bps;
#define LAUNCH_MRCC2(RR, NN, FF)                                                                                      \
    mrc_compact_kernel<RR, NN, FF><<<grid, 128, 0, st>>>(p->dev, (const float2*)Ydata, (const float2*)Hp,            \
                                                         (const float2*)Ypilot, idx_tx, errors, S, nslot, nbits, gx, A)
#define LAUNCH_MRCC(RR)                                                                                               \
    case RR:                                                                                                         \
        if (awgn && A.combine) { if (full) LAUNCH_MRCC2(RR, 2, true); else LAUNCH_MRCC2(RR, 2, false); }                \
        else if (awgn) { if (full) LAUNCH_MRCC2(RR, 1, true); else LAUNCH_MRCC2(RR, 1, false); }                     \
        else { if (full) LAUNCH_MRCC2(RR, 0, true); else LAUNCH_MRCC2(RR, 0, false); }                       \
        break;
    switch (R) {
        LAUNCH_MRCC(1) LAUNCH_MRCC(2) LAUNCH_MRCC(3) LAUNCH_MRCC(4) LAUNCH_MRCC(5) LAUNCH_MRCC(6) LAUNCH_MRCC(7)
        LAUNCH_MRCC(8)
        default: return LTE_ERR_INVALID_ARG;
    }
#undef LAUNCH_MRCC
#undef LAUNCH_MRCC2
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

extern "C" int lte_mrc_demap_count_compact(const lte_plan* p, const lte_c32* Ydata, const lte_c32* Hp,
                                           const uint8_t* idx_tx, unsigned long long* errors, int64_t nbits, int64_t B,
                                           int32_t R, int32_t S, const lte_awgn_desc* awgn, void* stream) {
    return mrc_compact_launch(p, Ydata, Hp, nullptr, idx_tx, errors, nbits, B, R, S, awgn, stream);
}

extern "C" int lte_crs_mrc_demap_count_compact(const lte_plan* p, const lte_c32* Ydata, const lte_c32* Ypilot,
                                               const uint8_t* idx_tx, unsigned long long* errors, int64_t nbits,
                                               int64_t B, int32_t R, int32_t S, const lte_awgn_desc* awgn, void* stream) {
    return mrc_compact_launch(p, Ydata, nullptr, Ypilot, idx_tx, errors, nbits, B, R, S, awgn, stream);
}
