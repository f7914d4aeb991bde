// Spatial multiplexing (TM4-like) stage kernels: layer mapping + codebook precoding, flat MIMO
// channel, and the per-subcarrier MIMO detectors (MMSE / ZF / ordered SIC / MRC) on H_eff = H W.
#include "common.cuh"
#include "awgn.cuh"

#define SM_MAX_TX 8
#define SM_MAX_RX 8
#define SM_MAX_LAYERS 4

struct SmPrecoder {
    float2 w[SM_MAX_TX][SM_MAX_LAYERS];
    int T, L;
};

__device__ __forceinline__ int slice_axis_sm(const DevPlan& P, float y) {
    if (P.nlev == 2) return y < 0.f ? 1 : 0;
    // number of thresholds below y by bisection; unused entries of thr[] are +inf (plan.cu)
    int l = (y > P.thr[3]) ? 4 : 0;
    l += (y > P.thr[l + 1]) ? 2 : 0;
    l += (y > P.thr[l]) ? 1 : 0;
    return l;
}

// ------------------------------------------------------------------------------ layer map + precode
// core/layer_mapper.py:35-86 (round robin: symbol q -> layer q % L, position q / L) and
// core/ofdm_core.py:2630-2640 (x_k = W layers[:, k] on the first ceil(Nd / L) data bins; the other
// data bins carry nothing).  out: [B][T][S][Nd] per-antenna data symbols for lte_tx_map_ifft.
__global__ void __launch_bounds__(256)
sm_precode_kernel(const DevPlan P, const SmPrecoder W, const uint8_t* __restrict__ idx,
                  const float2* __restrict__ syms, float2* __restrict__ out, float2* __restrict__ qam_out, int S,
                  long long total) {
    const int h = P.bps >> 1, mask = (1 << h) - 1;
    const int npos = (P.Nd + W.L - 1) / W.L;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(g % P.Nd);
        const long long bs = g / P.Nd;
        const long long b = bs / S;
        const int s = (int)(bs % S);
        float2 x[SM_MAX_TX];
#pragma unroll
        for (int t = 0; t < SM_MAX_TX; ++t) x[t] = make_float2(0.f, 0.f);
        if (p < npos) {
            for (int l = 0; l < W.L; ++l) {
                const int q = p * W.L + l;
                if (q >= P.Nd) break;                       // zero padding of the last position
                const size_t o = (size_t)bs * P.Nd + q;
                float2 v;
                if (syms) v = syms[o];
                else {
                    const int i = idx[o];
                    v = make_float2(P.lev[(i >> h) & mask], P.lev[i & mask]);
                    if (qam_out) qam_out[o] = v;
                }
#pragma unroll
                for (int t = 0; t < SM_MAX_TX; ++t)
                    if (t < W.T) { const float2 m = cmul(W.w[t][l], v); x[t].x += m.x; x[t].y += m.y; }
            }
        }
#pragma unroll
        for (int t = 0; t < SM_MAX_TX; ++t)
            if (t < W.T) out[(((size_t)b * W.T + t) * S + s) * P.Nd + p] = x[t];
    }
}

static int load_precoder(const lte_c32* W_host, int T, int L, SmPrecoder* out) {
    if (!W_host || T < 1 || T > SM_MAX_TX || L < 1 || L > SM_MAX_LAYERS) return LTE_ERR_INVALID_ARG;
    memset(out, 0, sizeof(*out));
    out->T = T; out->L = L;
    for (int t = 0; t < T; ++t)
        for (int l = 0; l < L; ++l) out->w[t][l] = make_float2(W_host[t * L + l].re, W_host[t * L + l].im);
    return LTE_OK;
}

extern "C" int lte_sm_precode(const lte_plan* p, const uint8_t* idx, const lte_c32* symbols,
                              const lte_c32* W_host, int32_t T, int32_t L, lte_c32* out, lte_c32* qam_out,
                              int64_t B, int32_t S, void* stream) {
    if (!p || (!idx && !symbols) || !out || B < 0 || S < 1) return LTE_ERR_INVALID_ARG;
    SmPrecoder W;
    int rc = load_precoder(W_host, T, L, &W);
    if (rc) return rc;
    if (B == 0) return LTE_OK;
    const long long total = (long long)B * S * p->dev.Nd;
    long long grid = (total + 255) / 256;
    if (grid > 148 * 32) grid = 148 * 32;
    sm_precode_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(p->dev, W, idx, (const float2*)symbols,
                                                                       (float2*)out, (float2*)qam_out, S, total);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

// ------------------------------------------------------------------------------ flat MIMO channel
// core/channel.py:467-480: y_r = sum_t h[r][t] x_t with one coefficient per link.
__global__ void __launch_bounds__(256)
flat_mimo_kernel(const float2* __restrict__ tx, const float2* __restrict__ h, float2* __restrict__ out,
                 double* __restrict__ power, int R, int T, long long n, int gx) {
    const long long br = blockIdx.x / gx;          // b*R + r
    const int bx = blockIdx.x % gx;
    const long long b = br / R;
    float2 hh[SM_MAX_TX];
    for (int t = 0; t < T; ++t) hh[t] = h[br * T + t];
    float pw = 0.f;
    for (long long i = (long long)bx * blockDim.x + threadIdx.x; i < n; i += (long long)gx * blockDim.x) {
        float2 acc = make_float2(0.f, 0.f);
        for (int t = 0; t < T; ++t) {
            const float2 m = cmul(hh[t], tx[((size_t)b * T + t) * n + i]);
            acc.x += m.x; acc.y += m.y;
        }
        out[(size_t)br * n + i] = acc;
        pw += cabs2(acc);
    }
    pw = warp_sum(pw);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = pw;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w];
        atomicAdd(&power[br], (double)s);
    }
}

extern "C" int lte_flat_mimo(const lte_plan* p, const lte_c32* tx, const lte_c32* h, lte_c32* out, double* power,
                             int64_t B, int32_t R, int32_t T, int64_t n, void* stream) {
    if (!p || !tx || !h || !out || !power || B < 0 || R < 1 || R > SM_MAX_RX || T < 1 || T > SM_MAX_TX || n < 1)
        return LTE_ERR_INVALID_ARG;
    if (B == 0) return LTE_OK;
    int gx = (int)((n + 256 * 8 - 1) / (256 * 8));
    if (gx > 64) gx = 64;
    flat_mimo_kernel<<<(unsigned)((long long)gx * B * R), 256, 0, (cudaStream_t)stream>>>(
        (const float2*)tx, (const float2*)h, (float2*)out, power, R, T, (long long)n, gx);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

// ------------------------------------------------------------------------------ per-symbol pilot estimates
// LS estimate at every pilot of every TX antenna's pilot set on EVERY OFDM symbol (the SM receiver estimates per
// symbol, core/ofdm_core.py:2700-2760 with core/lte_receiver.py:62-87): Hp [nsets][rows * S][Np], the first
// pset_cnt[t] entries of a row used.  With lazy AWGN each pilot gets its noise sample once here, instead of once
// per data bin that interpolates from it inside the detector.
__global__ void __launch_bounds__(256)
crs_ls_pilots_kernel(const DevPlan P, const float2* __restrict__ Y, float2* __restrict__ Hp, int k0, int nk, int S,
                     int nsets, long long rows, long long total, const AwgnArgs A, int noisy) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    const int i = (int)(g % P.Np);
    const long long q = g / P.Np;                   // t * rows * S + row * S + s
    const long long rs = q % (rows * S);
    const int t = (int)(q / (rows * S));
    if (i >= P.pset_cnt[t]) return;
    const long long row = rs / S;
    const int s = (int)(rs % S);
    const int bin = P.pset_bin[(size_t)t * P.Np + i];
    float2 y = Y[(size_t)rs * nk + (bin - k0)];
    if (noisy) y = awgn_at(A, lte_sigma(A.power[row], A.n_stream, A.snr_lin[row]), row, s, P.N, bin, y);
    Hp[g] = cmul(y, P.pset_inv[(size_t)t * P.Np + i]);
}

extern "C" int lte_crs_ls_pilots(const lte_plan* p, const lte_c32* Y, lte_c32* Hp, int window, int64_t rows, int32_t S,
                                 const lte_awgn_desc* awgn, void* stream) {
    if (!p || !Y || !Hp || rows < 0 || S < 1 || p->dev.Np == 0) return LTE_ERR_INVALID_ARG;
    int32_t k0, nk;
    int rc = lte_plan_window(p, window, &k0, &nk);
    if (rc) return rc;
    AwgnArgs A = {};
    if (awgn && (rc = make_awgn_args(A, p, awgn, S, rows))) return rc;
    if (rows == 0) return LTE_OK;
    const long long total = (long long)p->nsets * rows * S * p->dev.Np;
    const long long grid = (total + 255) / 256;
    if (grid >= (1ll << 31)) return LTE_ERR_UNSUPPORTED;
    crs_ls_pilots_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(p->dev, (const float2*)Y, (float2*)Hp, k0, nk, S,
                                                                          p->nsets, rows, total, A, awgn ? 1 : 0);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

// ------------------------------------------------------------------------------ MIMO detection
// core/mimo_detector.py:99-369 per data position.  The small-matrix algebra runs in fp64 (the
// inputs are fp32): A = H_eff^H H_eff + sigma^2 I is solved by complex Cholesky.
struct cd { double x, y; };
__device__ __forceinline__ cd cdmul(cd a, cd b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cd cdmulc(cd a, cd b) { return {a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x}; }   // conj(a) b
__device__ __forceinline__ cd cdsub(cd a, cd b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cd cdadd(cd a, cd b) { return {a.x + b.x, a.y + b.y}; }

// solve (G^H G-type Hermitian PD) A x = rhs for n <= 4 in place; A is overwritten
__device__ void chol_solve(cd (*A)[SM_MAX_LAYERS], cd* rhs, int n) {
    for (int j = 0; j < n; ++j) {
        double d = A[j][j].x;
        for (int k = 0; k < j; ++k) d -= A[j][k].x * A[j][k].x + A[j][k].y * A[j][k].y;
        d = sqrt(d);
        A[j][j] = {d, 0.0};
        for (int i = j + 1; i < n; ++i) {
            cd s = A[i][j];
            for (int k = 0; k < j; ++k) s = cdsub(s, cdmul(A[i][k], {A[j][k].x, -A[j][k].y}));
            A[i][j] = {s.x / d, s.y / d};
        }
    }
    for (int i = 0; i < n; ++i) {                // forward: G z = rhs
        cd s = rhs[i];
        for (int k = 0; k < i; ++k) s = cdsub(s, cdmul(A[i][k], rhs[k]));
        rhs[i] = {s.x / A[i][i].x, s.y / A[i][i].x};
    }
    for (int i = n - 1; i >= 0; --i) {           // backward: G^H x = z
        cd s = rhs[i];
        for (int k = i + 1; k < n; ++k) s = cdsub(s, cdmulc(A[k][i], rhs[k]));
        rhs[i] = {s.x / A[i][i].x, s.y / A[i][i].x};
    }
}

// (H^H H + sigma2 I)^-1 H^H y over the `cols` listed columns of Heff; result in x[0..nc)
__device__ void mmse_solve(const cd (*Heff)[SM_MAX_LAYERS], const cd* y, int R, const int* cols, int nc,
                           double sigma2, cd* x) {
    cd A[SM_MAX_LAYERS][SM_MAX_LAYERS];
    for (int i = 0; i < nc; ++i) {
        cd r = {0.0, 0.0};
        for (int q = 0; q < R; ++q) r = cdadd(r, cdmulc(Heff[q][cols[i]], y[q]));
        x[i] = r;
        for (int j = 0; j <= i; ++j) {
            cd a = {0.0, 0.0};
            for (int q = 0; q < R; ++q) a = cdadd(a, cdmulc(Heff[q][cols[j]], Heff[q][cols[i]]));
            // A[i][j] = sum_q conj(H[q][i]) H[q][j] = conj of the above
            A[i][j] = {a.x, -a.y};
        }
        A[i][i].x += sigma2;
        A[i][i].y = 0.0;
    }
    chol_solve(A, x, nc);
}

// The same two steps with compile-time sizes: every loop unrolls, nothing is indexed at run time, so H_eff, the
// Gram matrix and the right-hand side live in registers (the generic versions above keep them in local memory
// because `cols` and the loop bounds are run-time values).  Same operations in the same order.
template <int NL>
__device__ __forceinline__ void chol_solve_fixed(cd (&A)[NL][NL], cd (&rhs)[NL]) {
#pragma unroll
    for (int j = 0; j < NL; ++j) {
        double d = A[j][j].x;
#pragma unroll
        for (int k = 0; k < j; ++k) d -= A[j][k].x * A[j][k].x + A[j][k].y * A[j][k].y;
        d = sqrt(d);
        A[j][j] = {d, 0.0};
#pragma unroll
        for (int i = j + 1; i < NL; ++i) {
            cd s = A[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) s = cdsub(s, cdmul(A[i][k], {A[j][k].x, -A[j][k].y}));
            A[i][j] = {s.x / d, s.y / d};
        }
    }
#pragma unroll
    for (int i = 0; i < NL; ++i) {
        cd s = rhs[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s = cdsub(s, cdmul(A[i][k], rhs[k]));
        rhs[i] = {s.x / A[i][i].x, s.y / A[i][i].x};
    }
#pragma unroll
    for (int i = NL - 1; i >= 0; --i) {
        cd s = rhs[i];
#pragma unroll
        for (int k = i + 1; k < NL; ++k) s = cdsub(s, cdmulc(A[k][i], rhs[k]));
        rhs[i] = {s.x / A[i][i].x, s.y / A[i][i].x};
    }
}

template <int NR, int NL>
__device__ __forceinline__ void mmse_solve_fixed(const cd (&Heff)[NR][NL], const cd (&y)[NR], double sigma2, cd (&x)[NL]) {
    cd A[NL][NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) {
        cd r = {0.0, 0.0};
#pragma unroll
        for (int q = 0; q < NR; ++q) r = cdadd(r, cdmulc(Heff[q][i], y[q]));
        x[i] = r;
#pragma unroll
        for (int j = 0; j <= i; ++j) {
            cd a = {0.0, 0.0};
#pragma unroll
            for (int q = 0; q < NR; ++q) a = cdadd(a, cdmulc(Heff[q][j], Heff[q][i]));
            A[i][j] = {a.x, -a.y};
        }
        A[i][i].x += sigma2;
        A[i][i].y = 0.0;
    }
    chol_solve_fixed<NL>(A, x);
}

#define DET_MMSE 0
#define DET_ZF 1
#define DET_SIC 2
#define DET_MRC 3

__global__ void __launch_bounds__(128)
mimo_detect_kernel(const DevPlan P, const SmPrecoder W, const float2* __restrict__ Y, const float2* __restrict__ H,
                   const float2* __restrict__ Hp, float2* __restrict__ out, int k0, int nk, int R, int S, double sigma2_all,
                   const double* __restrict__ sigma2_streams, int detector, long long rows, long long total,
                   const AwgnArgs A, int noisy) {
    const int L = W.L, T = W.T;
    const int npos = (P.Nd + L - 1) / L;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(g % npos);
        const long long bs = g / npos;
        const long long b = bs / S;
        const int s = (int)(bs % S);
        const int kk = P.data_idx[p] - k0;
        const double sigma2 = sigma2_streams ? sigma2_streams[b] : sigma2_all;
        cd Heff[SM_MAX_RX][SM_MAX_LAYERS], y[SM_MAX_RX];
        for (int r = 0; r < R; ++r) {
            const size_t row = (size_t)b * R + r;
            float2 yv = Y[(row * S + s) * nk + kk];
            float sg = 0.f;
            if (noisy) {                // noise-free grid: the AWGN of lte_rx_fft(noise_domain = 1) joins as it is read
                sg = lte_sigma(A.power[row], A.n_stream, A.snr_lin[row]);
                yv = awgn_at(A, sg, (long long)row, s, P.N, kk + k0, yv);
            }
            y[r] = {(double)yv.x, (double)yv.y};
            for (int l = 0; l < L; ++l) Heff[r][l] = {0.0, 0.0};
            for (int t = 0; t < T; ++t) {
                float2 hv;
                if (H) {
                    hv = H[(((size_t)t * rows + row) * S + s) * nk + kk];
                } else {
                    // the per-symbol CRS estimate of TX antenna t at this bin, straight from the symbol's pilot bins:
                    // LS at the two neighbouring pilots of set t, then start + i * (delta / div) -- operation for
                    // operation what crs_ls_interp_kernel writes into H (core/lte_receiver.py:62-133)
                    const int cnt = P.pset_cnt[t];
                    const int16_t* pbin = P.pset_bin + (size_t)t * P.Np;
                    const float2* pinv = P.pset_inv + (size_t)t * P.Np;
                    const float2* yrow = Y + (row * S + s) * nk;
                    const int k = kk + k0;
                    const int lo = P.pset_seg[(size_t)t * P.N + k];
                    const int la = lo < 0 ? 0 : (lo >= cnt - 1 ? cnt - 1 : lo);
                    const int i1 = pbin[la];
                    const float2* hp = Hp ? Hp + (((size_t)t * rows + row) * S + s) * P.Np : nullptr;   // lte_crs_ls_pilots
                    float2 a;
                    if (hp) a = hp[la];
                    else {
                        float2 ya = yrow[i1 - k0];
                        if (noisy) ya = awgn_at(A, sg, (long long)row, s, P.N, i1, ya);
                        a = cmul(ya, pinv[la]);
                    }
                    hv = a;
                    if (lo >= 0 && lo < cnt - 1 && k != i1) {
                        const int i2 = pbin[la + 1];
                        float2 c;
                        if (hp) c = hp[la + 1];
                        else {
                            float2 yc = yrow[i2 - k0];
                            if (noisy) yc = awgn_at(A, sg, (long long)row, s, P.N, i2, yc);
                            c = cmul(yc, pinv[la + 1]);
                        }
                        const float div = (float)(i2 - i1), tt = (float)(k - i1);
                        hv = make_float2(fmaf(tt, __fdiv_rn(c.x - a.x, div), a.x), fmaf(tt, __fdiv_rn(c.y - a.y, div), a.y));
                    }
                }
                const cd hd = {(double)hv.x, (double)hv.y};
                for (int l = 0; l < L; ++l)
                    Heff[r][l] = cdadd(Heff[r][l], cdmul(hd, {(double)W.w[t][l].x, (double)W.w[t][l].y}));
            }
        }
        cd x[SM_MAX_LAYERS];
        int cols[SM_MAX_LAYERS] = {0, 1, 2, 3};
        if (detector == DET_MMSE) {
            mmse_solve(Heff, y, R, cols, L, sigma2, x);
        } else if (detector == DET_ZF) {
            mmse_solve(Heff, y, R, cols, L, 0.0, x);        // pinv of a full-column-rank H_eff
        } else if (detector == DET_MRC) {
            cd num = {0.0, 0.0};
            double den = 0.0;
            for (int q = 0; q < R; ++q) {
                num = cdadd(num, cdmulc(Heff[q][0], y[q]));
                den += Heff[q][0].x * Heff[q][0].x + Heff[q][0].y * Heff[q][0].y;
            }
            x[0] = {num.x / den, num.y / den};
        } else {                                             // ordered SIC
            double nrm[SM_MAX_LAYERS], tot = 0.0;
            for (int l = 0; l < L; ++l) {
                double a = 0.0;
                for (int q = 0; q < R; ++q) a += Heff[q][l].x * Heff[q][l].x + Heff[q][l].y * Heff[q][l].y;
                nrm[l] = a;
                tot += a;
            }
            bool done[SM_MAX_LAYERS] = {false, false, false, false};
            int nrem = L;
            for (int it = 0; it < L; ++it) {
                int best = -1;                               // highest SINR among all layers, fixed order
                double bv = -1.0;
                for (int l = 0; l < L; ++l) {
                    if (done[l]) continue;
                    const double v = nrm[l] / (tot - nrm[l] + sigma2 + 1e-10);
                    if (v > bv) { bv = v; best = l; }
                }
                int nc = 0, rel = 0;
                for (int l = 0; l < L; ++l)
                    if (!done[l]) { if (l == best) rel = nc; cols[nc++] = l; }
                cd sm;
                if (nrem == 1) {
                    cd num = {0.0, 0.0};
                    for (int q = 0; q < R; ++q) num = cdadd(num, cdmulc(Heff[q][best], y[q]));
                    sm = {num.x / (nrm[best] + sigma2), num.y / (nrm[best] + sigma2)};
                } else {
                    cd xs[SM_MAX_LAYERS];
                    mmse_solve(Heff, y, R, cols, nc, sigma2, xs);
                    sm = xs[rel];
                }
                const int ir = slice_axis_sm(P, (float)sm.x), ii = slice_axis_sm(P, (float)sm.y);
                const cd hard = {(double)P.lev[ir], (double)P.lev[ii]};
                x[best] = hard;
                for (int q = 0; q < R; ++q) y[q] = cdsub(y[q], cdmul(Heff[q][best], hard));
                done[best] = true;
                --nrem;
            }
        }
        for (int l = 0; l < L; ++l) {
            const int q = p * L + l;
            if (q < P.Nd) out[(size_t)bs * P.Nd + q] = make_float2((float)x[l].x, (float)x[l].y);
        }
    }
}

// MMSE / ZF with compile-time antenna and layer counts (the sweep shapes): registers instead of local memory.
template <int NR, int NL>
__global__ void __launch_bounds__(128, 4)
mimo_detect_fixed_kernel(const DevPlan P, const SmPrecoder W, const float2* __restrict__ Y, const float2* __restrict__ H,
                   const float2* __restrict__ Hp, float2* __restrict__ out, int k0, int nk, int R, int S, double sigma2_all,
                   const double* __restrict__ sigma2_streams, int detector, long long rows, long long total,
                   const AwgnArgs A, int noisy) {
    constexpr int L = NL;
    const int T = W.T;
    (void)R;
    const int npos = (P.Nd + L - 1) / L;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total;
         g += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(g % npos);
        const long long bs = g / npos;
        const long long b = bs / S;
        const int s = (int)(bs % S);
        const int kk = P.data_idx[p] - k0;
        const double sigma2 = sigma2_streams ? sigma2_streams[b] : sigma2_all;
        cd Heff[NR][NL], y[NR];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const size_t row = (size_t)b * NR + r;
            float2 yv = Y[(row * S + s) * nk + kk];
            float sg = 0.f;
            if (noisy) {                // noise-free grid: the AWGN of lte_rx_fft(noise_domain = 1) joins as it is read
                sg = lte_sigma(A.power[row], A.n_stream, A.snr_lin[row]);
                yv = awgn_at(A, sg, (long long)row, s, P.N, kk + k0, yv);
            }
            y[r] = {(double)yv.x, (double)yv.y};
#pragma unroll
            for (int l = 0; l < L; ++l) Heff[r][l] = {0.0, 0.0};
            for (int t = 0; t < T; ++t) {
                float2 hv;
                if (H) {
                    hv = H[(((size_t)t * rows + row) * S + s) * nk + kk];
                } else {
                    // the per-symbol CRS estimate of TX antenna t at this bin, straight from the symbol's pilot bins:
                    // LS at the two neighbouring pilots of set t, then start + i * (delta / div) -- operation for
                    // operation what crs_ls_interp_kernel writes into H (core/lte_receiver.py:62-133)
                    const int cnt = P.pset_cnt[t];
                    const int16_t* pbin = P.pset_bin + (size_t)t * P.Np;
                    const float2* pinv = P.pset_inv + (size_t)t * P.Np;
                    const float2* yrow = Y + (row * S + s) * nk;
                    const int k = kk + k0;
                    const int lo = P.pset_seg[(size_t)t * P.N + k];
                    const int la = lo < 0 ? 0 : (lo >= cnt - 1 ? cnt - 1 : lo);
                    const int i1 = pbin[la];
                    const float2* hp = Hp ? Hp + (((size_t)t * rows + row) * S + s) * P.Np : nullptr;   // lte_crs_ls_pilots
                    float2 a;
                    if (hp) a = hp[la];
                    else {
                        float2 ya = yrow[i1 - k0];
                        if (noisy) ya = awgn_at(A, sg, (long long)row, s, P.N, i1, ya);
                        a = cmul(ya, pinv[la]);
                    }
                    hv = a;
                    if (lo >= 0 && lo < cnt - 1 && k != i1) {
                        const int i2 = pbin[la + 1];
                        float2 c;
                        if (hp) c = hp[la + 1];
                        else {
                            float2 yc = yrow[i2 - k0];
                            if (noisy) yc = awgn_at(A, sg, (long long)row, s, P.N, i2, yc);
                            c = cmul(yc, pinv[la + 1]);
                        }
                        const float div = (float)(i2 - i1), tt = (float)(k - i1);
                        hv = make_float2(fmaf(tt, __fdiv_rn(c.x - a.x, div), a.x), fmaf(tt, __fdiv_rn(c.y - a.y, div), a.y));
                    }
                }
                const cd hd = {(double)hv.x, (double)hv.y};
#pragma unroll
                for (int l = 0; l < L; ++l)
                    Heff[r][l] = cdadd(Heff[r][l], cdmul(hd, {(double)W.w[t][l].x, (double)W.w[t][l].y}));
            }
        }
        cd x[NL];
        mmse_solve_fixed<NR, NL>(Heff, y, detector == DET_ZF ? 0.0 : sigma2, x);     // ZF: pinv of a full-column-rank H_eff
#pragma unroll
        for (int l = 0; l < L; ++l) {
            const int q = p * L + l;
            if (q < P.Nd) out[(size_t)bs * P.Nd + q] = make_float2((float)x[l].x, (float)x[l].y);
        }
    }
}

extern "C" int lte_mimo_detect(const lte_plan* p, const lte_c32* Y, const lte_c32* H, const lte_c32* Hpilot,
                               const lte_c32* W_host,
                               int32_t T, int32_t L, double sigma2, const double* sigma2_streams, int32_t detector,
                               lte_c32* out, int window, int64_t B, int32_t R, int32_t S, const lte_awgn_desc* awgn,
                               void* stream) {
    if (!p || !Y || !out || B < 0 || S < 1 || R < 1 || R > SM_MAX_RX) return LTE_ERR_INVALID_ARG;
    if (!H && (p->nsets < T || p->dev.Np == 0)) return LTE_ERR_INVALID_ARG;    // estimating needs the T pilot sets
    if (H && Hpilot) return LTE_ERR_INVALID_ARG;
    if (detector < DET_MMSE || detector > DET_MRC) return LTE_ERR_INVALID_ARG;
    if (R < L) return LTE_ERR_INVALID_ARG;                  // core/mimo_detector.py:34-35
    if (detector == DET_MRC && L != 1) return LTE_ERR_INVALID_ARG;
    SmPrecoder W;
    int rc = load_precoder(W_host, T, L, &W);
    if (rc) return rc;
    int32_t k0, nk;
    rc = lte_plan_window(p, window, &k0, &nk);
    if (rc) return rc;
    AwgnArgs A = {};
    if (awgn && (rc = make_awgn_args(A, p, awgn, S, B * R))) return rc;
    if (B == 0) return LTE_OK;
    const int npos = (p->dev.Nd + L - 1) / L;
    const long long total = (long long)B * S * npos;
    long long grid = (total + 127) / 128;
    if (grid > 148 * 16) grid = 148 * 16;
#define LAUNCH_DET_FIXED(RR, LL)                                                                                  \
    if (R == RR && L == LL) {                                                                                    \
        mimo_detect_fixed_kernel<RR, LL><<<(unsigned)grid, 128, 0, (cudaStream_t)stream>>>(                       \
            p->dev, W, (const float2*)Y, (const float2*)H, (const float2*)Hpilot, (float2*)out, k0, nk, R, S, sigma2,   \
            sigma2_streams,                                                                                      \
            detector, (long long)B * R, total, A, awgn ? 1 : 0);                                                 \
        LTE_CHECK_CUDA(cudaGetLastError());                                                                      \
        return LTE_OK;                                                                                           \
    }
    if (detector == DET_MMSE || detector == DET_ZF) {
        LAUNCH_DET_FIXED(2, 1) LAUNCH_DET_FIXED(2, 2) LAUNCH_DET_FIXED(4, 1) LAUNCH_DET_FIXED(4, 2) LAUNCH_DET_FIXED(4, 3)
        LAUNCH_DET_FIXED(4, 4)
    }
#undef LAUNCH_DET_FIXED
    mimo_detect_kernel<<<(unsigned)grid, 128, 0, (cudaStream_t)stream>>>(
        p->dev, W, (const float2*)Y, (const float2*)H, (const float2*)Hpilot, (float2*)out, k0, nk, R, S, sigma2,
        sigma2_streams, detector, (long long)B * R, total, A, awgn ? 1 : 0);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}
