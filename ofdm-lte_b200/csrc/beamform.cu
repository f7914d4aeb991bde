// Rank-1 beamforming over a flat R x T channel: CSI feedback / precoder selection
// for a batch of channel matrices, and the per-subcarrier link  x = W s,  y = H x + n,  MRC with
// H_eff = H W, slicer, bit-error count  (core/ofdm_core.py:2260-2477, core/csi_feedback.py:55-196,
// core/codebook_lte.py:332-373, core/beamforming_precoder.py:41-66, :176-201).
//
// One stream = one channel realisation (B streams per launch).  The reference never leaves the
// frequency domain on this path (no IFFT / CP / FFT), and neither do these kernels.
#include <math.h>

#include "slicer.cuh"

#define BF_MAX_T 8
#define BF_MAX_R 8
#define BF_MAX_CB 16

struct BfCodebook {
    int ncb;
    float2 w[BF_MAX_CB][BF_MAX_T];
};

// ------------------------------------------------------------------ flat channel draws
// h[b][r][t] ~ CN(0, 1): (randn + j randn) / sqrt(2)  (core/ofdm_core.py:2347-2348), Philox keyed
// (seed, global stream id, r*T + t) so a stream's channel does not depend on the batch or the rank.
__global__ void random_channel_kernel(float2* __restrict__ h, int RT, uint32_t key, unsigned long long stream_id0,
                                      long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long b = i / RT;
    const int e = (int)(i - b * RT);
    const float2 z = lte_noise_sample(key, (uint32_t)(stream_id0 + (unsigned long long)b), (uint32_t)e);
    h[i] = make_float2(z.x * 0.70710678118654752f, z.y * 0.70710678118654752f);
}

extern "C" int lte_random_channel(lte_c32* h, int64_t B, int32_t R, int32_t T, uint64_t seed, uint64_t stream_id0,
                                  void* stream) {
    if (B > 0 && !lte_ids_fit(stream_id0, (uint64_t)B)) return LTE_ERR_UNSUPPORTED;
    if (!h || B < 0 || R < 1 || R > BF_MAX_R || T < 1 || T > BF_MAX_T) return LTE_ERR_INVALID_ARG;
    if (B == 0) return LTE_OK;
    const long long total = (long long)B * R * T;
    random_channel_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (float2*)h, R * T, lte_key(seed, LTE_DOMAIN_PHASE) ^ 0x62666368u, stream_id0, total);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

// ------------------------------------------------------------------ precoder selection
// One thread per stream; the matrices are at most 8 x 8, so everything is done in fp64 registers
// (the reference works in complex128; the only rounding left is the fp32 storage of H and W).
//   pmi  : argmax_i sum_r |(H w_i)_r|^2, first maximum wins     (LTECodebook.select_best_pmi)
//   MRT  : w = conj(mean_r H[r, :]) / ||.||                     (calculate_mrt_weights)
//   gain : 10 log10( ||H w||^2 / (||H||_F^2 / T) )              (calculate_beamforming_gain)
__global__ void __launch_bounds__(128)
bf_weights_kernel(const float2* __restrict__ h, const BfCodebook cb, int mode, float2* __restrict__ W,
                  float2* __restrict__ heff, int* __restrict__ pmi_out, float* __restrict__ gain_db, long long B,
                  int R, int T) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float2* hb = h + (size_t)b * R * T;
    int best = 0;
    double best_m = -INFINITY;
    for (int i = 0; i < cb.ncb; ++i) {
        double m = 0.0;
        for (int r = 0; r < R; ++r) {
            double ar = 0.0, ai = 0.0;
            for (int t = 0; t < T; ++t) {
                const float2 a = hb[r * T + t], w = cb.w[i][t];
                ar += (double)a.x * w.x - (double)a.y * w.y;
                ai += (double)a.x * w.y + (double)a.y * w.x;
            }
            m += ar * ar + ai * ai;
        }
        if (m > best_m) { best_m = m; best = i; }
    }
    if (pmi_out) pmi_out[b] = best;
    double wr[BF_MAX_T], wi[BF_MAX_T];
    if (mode == LTE_BF_MRT) {
        double nrm = 0.0;
        for (int t = 0; t < T; ++t) {
            double sr = 0.0, si = 0.0;
            for (int r = 0; r < R; ++r) { sr += hb[r * T + t].x; si += hb[r * T + t].y; }
            sr /= R; si /= R;
            wr[t] = sr; wi[t] = -si;
            nrm += sr * sr + si * si;
        }
        nrm = sqrt(nrm);
        for (int t = 0; t < T; ++t) { wr[t] /= nrm; wi[t] /= nrm; }
    } else {
        for (int t = 0; t < T; ++t) { wr[t] = cb.w[best][t].x; wi[t] = cb.w[best][t].y; }
    }
    double pbf = 0.0, ph = 0.0;
    for (int r = 0; r < R; ++r) {
        double ar = 0.0, ai = 0.0;
        for (int t = 0; t < T; ++t) {
            const float2 a = hb[r * T + t];
            ar += (double)a.x * wr[t] - (double)a.y * wi[t];
            ai += (double)a.x * wi[t] + (double)a.y * wr[t];
            ph += (double)a.x * a.x + (double)a.y * a.y;
        }
        heff[(size_t)b * R + r] = make_float2((float)ar, (float)ai);
        pbf += ar * ar + ai * ai;
    }
    for (int t = 0; t < T; ++t) W[(size_t)b * T + t] = make_float2((float)wr[t], (float)wi[t]);
    if (gain_db) gain_db[b] = (float)(10.0 * log10(pbf / (ph / T)));
}

extern "C" int lte_bf_weights(const lte_c32* h, const lte_c32* codebook_host, int32_t ncb, int32_t mode, lte_c32* W,
                              lte_c32* heff, int32_t* pmi, float* gain_db, int64_t B, int32_t R, int32_t T,
                              void* stream) {
    if (!h || !W || !heff || B < 0 || R < 1 || R > BF_MAX_R || T < 1 || T > BF_MAX_T) return LTE_ERR_INVALID_ARG;
    if (mode != LTE_BF_MRT && mode != LTE_BF_CODEBOOK) return LTE_ERR_INVALID_ARG;
    if (ncb < 0 || ncb > BF_MAX_CB || (ncb > 0 && !codebook_host)) return LTE_ERR_INVALID_ARG;
    if (mode == LTE_BF_CODEBOOK && ncb == 0) return LTE_ERR_INVALID_ARG;
    if (B == 0) return LTE_OK;
    BfCodebook cb;
    memset(&cb, 0, sizeof(cb));
    cb.ncb = ncb;
    for (int i = 0; i < ncb; ++i)
        for (int t = 0; t < T; ++t) cb.w[i][t] = make_float2(codebook_host[i * T + t].re, codebook_host[i * T + t].im);
    bf_weights_kernel<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        (const float2*)h, cb, mode, (float2*)W, (float2*)heff, pmi, gain_db, B, R, T);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

// ------------------------------------------------------------------ rank / precoder feedback
// RankAdaptation.get_feedback with its default methods (core/rank_adaptation.py:41-130 `eigenvalue`,
// :148-211 `capacity`) for n channel matrices at once, one thread per matrix, fp64 like the reference:
//   RI  : eigenvalues of H^H H (cyclic Jacobi on the T x T Hermitian matrix), count those above
//         rank_threshold x the largest, cap at max_rank, then the SNR gates (< 5 dB: 1, < 10 dB: <= 2)
//   PMI : argmax over the rank's codebook of log2 det(I + snr / ri  H_eff H_eff^H), H_eff = H W, first
//         maximum wins; the determinant is taken of the ri x ri matrix I + snr / ri  H_eff^H H_eff
//         (Sylvester's identity: same value, at most 4 x 4) by Cholesky factorisation.
#define RF_MAX_T 4
#define RF_MAX_RANK 4
struct c64 { double x, y; };
__device__ __forceinline__ c64 c64_mulc(c64 a, c64 b) { return {a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y}; }   // a conj(b)

__global__ void __launch_bounds__(64)
rank_feedback_kernel(const float2* __restrict__ H, const double* __restrict__ snr_db, const float2* __restrict__ codebook,
                     int ncb_stride, int4 ncb, double rank_threshold, int max_rank, int* __restrict__ ri_out,
                     int* __restrict__ pmi_out, long long n, int R, int T) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    c64 h[BF_MAX_R][RF_MAX_T];
    for (int r = 0; r < R; ++r)
        for (int t = 0; t < T; ++t) {
            const float2 v = H[((size_t)i * R + r) * T + t];
            h[r][t] = {(double)v.x, (double)v.y};
        }
    // ---- Gram matrix A = H^H H and its eigenvalues
    c64 A[RF_MAX_T][RF_MAX_T];
    for (int a = 0; a < T; ++a)
        for (int b = 0; b < T; ++b) {
            c64 acc = {0.0, 0.0};
            for (int r = 0; r < R; ++r) {                       // conj(h[r][a]) h[r][b]
                const c64 p = c64_mulc(h[r][b], h[r][a]);
                acc.x += p.x; acc.y += p.y;
            }
            A[a][b] = acc;
        }
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = 0.0, dia = 0.0;
        for (int a = 0; a < T; ++a)
            for (int b = 0; b < T; ++b) {
                const double m = A[a][b].x * A[a][b].x + A[a][b].y * A[a][b].y;
                if (a == b) dia += m; else off += m;
            }
        if (off <= 1e-32 * dia) break;
        for (int p = 0; p < T - 1; ++p)
            for (int q = p + 1; q < T; ++q) {
                const double g = sqrt(A[p][q].x * A[p][q].x + A[p][q].y * A[p][q].y);
                if (g == 0.0) continue;
                // unitary rotation in the (p, q) plane that annihilates A[p][q] = g e^{j phi}
                const c64 ph = {A[p][q].x / g, A[p][q].y / g};
                const double tau = (A[q][q].x - A[p][p].x) / (2.0 * g);
                const double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                const double c = 1.0 / sqrt(1.0 + tt * tt), sn = tt * c;
                // columns: A <- A J, J = [[c, s e^{j phi}], [-s e^{-j phi}, c]] on (p, q)
                for (int k = 0; k < T; ++k) {
                    const c64 akp = A[k][p], akq = A[k][q];
                    // akq e^{-j phi}
                    const c64 u = {akq.x * ph.x + akq.y * ph.y, akq.y * ph.x - akq.x * ph.y};
                    // akp e^{j phi}
                    const c64 v = {akp.x * ph.x - akp.y * ph.y, akp.y * ph.x + akp.x * ph.y};
                    A[k][p] = {c * akp.x - sn * u.x, c * akp.y - sn * u.y};
                    A[k][q] = {sn * v.x + c * akq.x, sn * v.y + c * akq.y};
                }
                // rows: A <- J^H A
                for (int k = 0; k < T; ++k) {
                    const c64 apk = A[p][k], aqk = A[q][k];
                    // aqk e^{j phi}
                    const c64 u = {aqk.x * ph.x - aqk.y * ph.y, aqk.y * ph.x + aqk.x * ph.y};
                    // apk e^{-j phi}
                    const c64 v = {apk.x * ph.x + apk.y * ph.y, apk.y * ph.x - apk.x * ph.y};
                    A[p][k] = {c * apk.x - sn * u.x, c * apk.y - sn * u.y};
                    A[q][k] = {sn * v.x + c * aqk.x, sn * v.y + c * aqk.y};
                }
            }
    }
    double lmax = A[0][0].x;
    for (int a = 1; a < T; ++a) lmax = fmax(lmax, A[a][a].x);
    const double sdb = snr_db[i];
    int ri = 1;
    if (lmax >= 1e-10) {
        int sig = 0;
        for (int a = 0; a < T; ++a) sig += (A[a][a].x / lmax > rank_threshold) ? 1 : 0;
        ri = sig < max_rank ? sig : max_rank;
        if (sdb < 5.0) ri = 1;
        else if (sdb < 10.0) ri = ri < 2 ? ri : 2;
        if (ri < 1) ri = 1;
    }
    // ---- precoder of that rank with the largest capacity metric
    const double scale = pow(10.0, sdb / 10.0) / (double)ri;
    const int nc = ri == 1 ? ncb.x : ri == 2 ? ncb.y : ri == 3 ? ncb.z : ncb.w;
    const float2* cbr = codebook + (size_t)(ri - 1) * ncb_stride * T * RF_MAX_RANK;
    int best = 0;
    double best_v = -INFINITY;
    for (int pmi = 0; pmi < nc; ++pmi) {
        const float2* W = cbr + (size_t)pmi * T * RF_MAX_RANK;  // [T][RF_MAX_RANK], columns < ri used
        c64 he[BF_MAX_R][RF_MAX_RANK];
        for (int r = 0; r < R; ++r)
            for (int l = 0; l < ri; ++l) {
                c64 acc = {0.0, 0.0};
                for (int t = 0; t < T; ++t) {
                    const float2 w = W[t * RF_MAX_RANK + l];
                    acc.x += h[r][t].x * (double)w.x - h[r][t].y * (double)w.y;
                    acc.y += h[r][t].x * (double)w.y + h[r][t].y * (double)w.x;
                }
                he[r][l] = acc;
            }
        c64 M[RF_MAX_RANK][RF_MAX_RANK];                        // I + scale H_eff^H H_eff
        for (int a = 0; a < ri; ++a)
            for (int b = 0; b < ri; ++b) {
                c64 acc = {a == b ? 1.0 : 0.0, 0.0};
                for (int r = 0; r < R; ++r) {
                    const c64 p = c64_mulc(he[r][b], he[r][a]);
                    acc.x += scale * p.x; acc.y += scale * p.y;
                }
                M[a][b] = acc;
            }
        double logdet = 0.0;                                    // Cholesky: det = prod d_k^2
        for (int k = 0; k < ri; ++k) {
            double d = M[k][k].x;
            for (int m = 0; m < k; ++m) d -= M[k][m].x * M[k][m].x + M[k][m].y * M[k][m].y;
            logdet += log2(d);
            const double rd = 1.0 / sqrt(d);
            for (int a = k + 1; a < ri; ++a) {
                c64 v = M[a][k];
                for (int m = 0; m < k; ++m) {                   // M[a][m] conj(M[k][m])
                    const c64 p = c64_mulc(M[a][m], M[k][m]);
                    v.x -= p.x; v.y -= p.y;
                }
                M[a][k] = {v.x * rd, v.y * rd};
            }
        }
        if (logdet > best_v) { best_v = logdet; best = pmi; }
    }
    ri_out[i] = ri;
    pmi_out[i] = best;
}

extern "C" int lte_rank_feedback(const lte_c32* H, const double* snr_db, const lte_c32* codebook, int32_t ncb_stride,
                                 const int32_t* ncb_host, double rank_threshold, int32_t max_rank, int32_t* ri,
                                 int32_t* pmi, int64_t n, int32_t R, int32_t T, void* stream) {
    if (!H || !snr_db || !codebook || !ncb_host || !ri || !pmi || n < 0) return LTE_ERR_INVALID_ARG;
    if (R < 1 || R > BF_MAX_R || T < 1 || T > RF_MAX_T || max_rank < 1 || max_rank > RF_MAX_RANK || max_rank > T)
        return LTE_ERR_INVALID_ARG;
    int nc[RF_MAX_RANK] = {0, 0, 0, 0};
    for (int k = 0; k < max_rank; ++k) {
        if (ncb_host[k] < 1 || ncb_host[k] > ncb_stride) return LTE_ERR_INVALID_ARG;
        nc[k] = ncb_host[k];
    }
    if (n == 0) return LTE_OK;
    rank_feedback_kernel<<<(unsigned)((n + 63) / 64), 64, 0, (cudaStream_t)stream>>>(
        (const float2*)H, snr_db, (const float2*)codebook, ncb_stride, make_int4(nc[0], nc[1], nc[2], nc[3]), rank_threshold,
        max_rank, ri, pmi, n, R, T);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

// ------------------------------------------------------------------ the link
// CTA = (stream b, OFDM symbol s); thread = data position d (strided).  Per data symbol
//   x_t = w_t s            (BeamformingPrecoder.apply_precoding, W @ s)
//   y_r = sum_t h_rt x_t   (core/ofdm_core.py:2392-2395, accumulated in TX order)
//   y_r += sigma (z_re + j z_im), sigma = sqrt(10^(-snr/10) / 2)      (:2398-2401)
//   s^  = sum_r conj(heff_r) y_r / sum_r |heff_r|^2                    (:2416-2424)
// then the slicer and XOR/popcount against the transmitted index.  Noise is either replayed
// (z: [B][S][2][R][Nd], the reference's randn(R, Nd) real block then imaginary block per symbol)
// or Philox keyed (seed, row_id0 + b*R + r, s*Nd + d).
template <bool REPLAY>
__global__ void __launch_bounds__(256, 4)
bf_link_kernel(const DevPlan P, const uint8_t* __restrict__ idx, const float2* __restrict__ h,
               const float2* __restrict__ W, const float2* __restrict__ heff, const float* __restrict__ noise_std,
               const float* __restrict__ z, uint32_t key, unsigned long long row_id0, float2* __restrict__ out,
               unsigned long long* __restrict__ errors, long long nbits, int R, int T, int S) {
    __shared__ float2 sh[BF_MAX_R * BF_MAX_T], sw[BF_MAX_T], se[BF_MAX_R];
    __shared__ float s_inv;
    const long long b = blockIdx.x / S;
    const int s = (int)(blockIdx.x - b * S);
    if (threadIdx.x < R * T) sh[threadIdx.x] = h[(size_t)b * R * T + threadIdx.x];
    if (threadIdx.x < T) sw[threadIdx.x] = W[(size_t)b * T + threadIdx.x];
    if (threadIdx.x < R) se[threadIdx.x] = heff[(size_t)b * R + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        float pn = 0.f;
        for (int r = 0; r < R; ++r) pn += cabs2(se[r]);
        s_inv = pn;
    }
    __syncthreads();
    const float pn = s_inv, sigma = noise_std[b];
    const int Nd = P.Nd, hb = P.bps >> 1, mask = (1 << hb) - 1;
    const uint8_t* ib = idx + ((size_t)b * S + s) * Nd;
    const float* zb = REPLAY ? z + ((size_t)b * S + s) * 2 * R * Nd : nullptr;
    unsigned int e = 0;
    for (int d = threadIdx.x; d < Nd; d += blockDim.x) {
        const int v = ib[d];
        const float2 sym = make_float2(P.lev[(v >> hb) & mask], P.lev[v & mask]);
        float2 x[BF_MAX_T];
#pragma unroll
        for (int t = 0; t < BF_MAX_T; ++t)
            if (t < T) x[t] = cmul(sw[t], sym);
        float2 acc = make_float2(0.f, 0.f);
        for (int r = 0; r < R; ++r) {
            float2 y = make_float2(0.f, 0.f);
#pragma unroll
            for (int t = 0; t < BF_MAX_T; ++t) {
                if (t < T) {
                    const float2 p = cmul(sh[r * T + t], x[t]);
                    y.x += p.x;
                    y.y += p.y;
                }
            }
            float2 w;
            if (REPLAY) w = make_float2(zb[(size_t)r * Nd + d], zb[(size_t)(R + r) * Nd + d]);
            else w = lte_noise_sample(key, (uint32_t)(row_id0 + (unsigned long long)(b * R + r)), (uint32_t)(s * Nd + d));
            y.x = fmaf(sigma, w.x, y.x);
            y.y = fmaf(sigma, w.y, y.y);
            const float2 c = cmulc(se[r], y);
            acc.x += c.x;
            acc.y += c.y;
        }
        const float2 eq = make_float2(__fdiv_rn(acc.x, pn), __fdiv_rn(acc.y, pn));
        if (out) out[((size_t)b * S + s) * Nd + d] = eq;
        if (errors) e += bit_errors(slice_symbol(P, eq), v, P.bps, nbits - ((long long)s * Nd + d) * P.bps);
    }
    if (errors) block_add_errors(e, &errors[b]);
}

extern "C" int lte_bf_link(const lte_plan* p, const uint8_t* idx, const lte_c32* h, const lte_c32* W,
                           const lte_c32* heff, const float* noise_std, const float* z, uint64_t seed,
                           uint64_t row_id0, lte_c32* out, unsigned long long* errors, int64_t nbits, int64_t B,
                           int32_t R, int32_t T, int32_t S, void* stream) {
    if (!z && B > 0 && R > 0 && !lte_ids_fit(row_id0, (uint64_t)B * (uint64_t)R)) return LTE_ERR_UNSUPPORTED;
    if (!p || !idx || !h || !W || !heff || !noise_std || B < 0 || S < 1 || R < 1 || R > BF_MAX_R || T < 1 ||
        T > BF_MAX_T || (!out && !errors))
        return LTE_ERR_INVALID_ARG;
    if (B == 0) return LTE_OK;
    if ((long long)B * S >= (1ll << 31)) return LTE_ERR_UNSUPPORTED;
    const unsigned grid = (unsigned)(B * S);
    const uint32_t key = lte_key(seed, LTE_DOMAIN_NOISE);
    if (z)
        bf_link_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(p->dev, idx, (const float2*)h, (const float2*)W,
                                                                      (const float2*)heff, noise_std, z, key, row_id0,
                                                                      (float2*)out, errors, nbits, R, T, S);
    else
        bf_link_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(p->dev, idx, (const float2*)h, (const float2*)W,
                                                                       (const float2*)heff, noise_std, nullptr, key,
                                                                       row_id0, (float2*)out, errors, nbits, R, T, S);
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}
