// SC-FDM M-point unitary DFT / IDFT (reference core/dft_precoding.py:44-54,86-88 and :175,211:
// a dense M x M matrix product) as a Bluestein chirp-z transform on the shared-memory FFT core.
// M = number of data subcarriers (62 / 125 / 249 / 499 / 749 / 999), none of them 2-3-5 smooth.
//
//   X[k] = w[k] * sum_n (x[n] w[n]) conj(w[k-n]) / sqrt(M),   w[n] = exp(-j pi n^2 / M)
// The convolution runs through two NB-point FFTs (NB = pow2 >= 2M-1) in one kernel; the
// spectrum of the chirp filter is a per-(plan, M) table computed once on the host in fp64 by
// lte_plan_add_dft (plan.cu) -- the launcher itself never allocates or copies.
#include "fft.cuh"
#include "common.cuh"

template <int NB>
__global__ void __launch_bounds__(FFT_CTA_THREADS)
dft_m_kernel(const float2* __restrict__ in, float2* __restrict__ out, const float2* __restrict__ w,
             const float2* __restrict__ bf, const float2* __restrict__ tw, int M, int inverse, long long rows) {
    constexpr int TPF = NB / FFT_ELEMS, FPC = fft_per_cta(NB);
    extern __shared__ float2 smem[];
    const int f_local = threadIdx.x / TPF, j = threadIdx.x % TPF;
    const long long row = (long long)blockIdx.x * FPC + f_local;
    const bool valid = row < rows;
    float2* sA = smem + (size_t)f_local * 2 * fft_smem_elems(NB);
    float2* sB = sA + fft_smem_elems(NB);
    const float sgn = inverse ? -1.f : 1.f;       // IDFT(x) = conj(DFT(conj(x)))

    float2 v[FFT_ELEMS];
#pragma unroll
    for (int e = 0; e < FFT_ELEMS; ++e) {
        const int n = j + e * TPF;
        float2 a = make_float2(0.f, 0.f);
        if (valid && n < M) {
            float2 x = in[(size_t)row * M + n];
            x.y *= sgn;
            a = cmul(x, w[n]);
        }
        v[e] = a;
    }
    fft_run<NB, false>(v, sA, sB, tw, j);
    __syncthreads();          // both exchange buffers are reused by the second transform
#pragma unroll
    for (int e = 0; e < FFT_ELEMS; ++e) v[e] = cmul(v[e], bf[j + e * TPF]);
    fft_run<NB, true>(v, sA, sB, tw, j);
    if (valid) {
#pragma unroll
        for (int e = 0; e < FFT_ELEMS; ++e) {
            const int k = j + e * TPF;
            if (k < M) {
                float2 y = cmul(v[e], w[k]);
                y.y *= sgn;
                out[(size_t)row * M + k] = y;
            }
        }
    }
}

extern "C" int lte_dft_m(const lte_plan* p, const lte_c32* in, lte_c32* out, int32_t M, int32_t inverse,
                         int64_t rows, void* stream) {
    if (!p || !in || !out || M < 1 || rows < 0) return LTE_ERR_INVALID_ARG;
    auto it = p->dft.find(M);
    if (it == p->dft.end()) return LTE_ERR_INVALID_ARG;           // tables are plan state: lte_plan_add_dft(plan, M) first
    if (rows == 0) return LTE_OK;
    const DftTables t = it->second;
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_DFT(NBV)                                                                                     \
    case NBV: {                                                                                             \
        auto k = dft_m_kernel<NBV>;                                                                         \
        const int smem = fft_cta_smem_bytes(NBV);                                                           \
        LTE_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));         \
        const long long grid = (rows + fft_per_cta(NBV) - 1) / fft_per_cta(NBV);                            \
        k<<<(unsigned)grid, FFT_CTA_THREADS, smem, st>>>((const float2*)in, (float2*)out, t.w, t.bf,        \
                                                         t.tw, M, inverse, rows);                           \
    } break;
    switch (t.NB) {
        LAUNCH_DFT(64) LAUNCH_DFT(128) LAUNCH_DFT(256) LAUNCH_DFT(512) LAUNCH_DFT(1024) LAUNCH_DFT(2048)
        default: return LTE_ERR_UNSUPPORTED;
    }
#undef LAUNCH_DFT
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}

// The forward transform with the QAM map folded into its load (the batched SC-FDM sweep, core/modulator.py:80-86):
// no symbol tensor between lte_qam_map and the DFT.  Same arithmetic as lte_qam_map + lte_dft_m.
// (Measured and dropped: the inverse transform with the zero-forcing equaliser and the lazy AWGN in its load and the
// slicer + bit-error count in its store -- 1.31 ms per 8192 x 14 symbols of 499 against 0.27 + 0.50 + 0.19 ms for
// the three kernels: the element-wise work is instruction heavy and runs far better in streaming kernels at full
// occupancy than inside a 100-register transform kernel.)
template <int NB>
__global__ void __launch_bounds__(FFT_CTA_THREADS)
dft_qam_kernel(const DevPlan P, const uint8_t* __restrict__ idx, float2* __restrict__ out, const float2* __restrict__ w,
               const float2* __restrict__ bf, const float2* __restrict__ tw, int M, long long rows) {
    constexpr int TPF = NB / FFT_ELEMS, FPC = fft_per_cta(NB);
    extern __shared__ float2 smem[];
    const int f_local = threadIdx.x / TPF, j = threadIdx.x % TPF;
    const long long row = (long long)blockIdx.x * FPC + f_local;
    const bool valid = row < rows;
    float2* sA = smem + (size_t)f_local * 2 * fft_smem_elems(NB);
    float2* sB = sA + fft_smem_elems(NB);
    const int h = P.bps >> 1, mask = (1 << h) - 1;
    float2 v[FFT_ELEMS];
#pragma unroll
    for (int e = 0; e < FFT_ELEMS; ++e) {
        const int n = j + e * TPF;
        float2 a = make_float2(0.f, 0.f);
        if (valid && n < M) {
            const int q = idx[(size_t)row * M + n];
            a = cmul(make_float2(P.lev[(q >> h) & mask], P.lev[q & mask]), w[n]);
        }
        v[e] = a;
    }
    fft_run<NB, false>(v, sA, sB, tw, j);
    __syncthreads();          // both exchange buffers are reused by the second transform
#pragma unroll
    for (int e = 0; e < FFT_ELEMS; ++e) v[e] = cmul(v[e], bf[j + e * TPF]);
    fft_run<NB, true>(v, sA, sB, tw, j);
    if (valid) {
#pragma unroll
        for (int e = 0; e < FFT_ELEMS; ++e) {
            const int k = j + e * TPF;
            if (k < M) out[(size_t)row * M + k] = cmul(v[e], w[k]);
        }
    }
}

extern "C" int lte_dft_qam(const lte_plan* p, const uint8_t* idx, lte_c32* out, int32_t M, int64_t rows, void* stream) {
    if (!p || !idx || !out || M < 1 || rows < 0) return LTE_ERR_INVALID_ARG;
    auto it = p->dft.find(M);
    if (it == p->dft.end()) return LTE_ERR_INVALID_ARG;
    if (rows == 0) return LTE_OK;
    const DftTables t = it->second;
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_DFTQ(NBV)                                                                                    \
    case NBV: {                                                                                             \
        auto k = dft_qam_kernel<NBV>;                                                                       \
        const int smem = fft_cta_smem_bytes(NBV);                                                           \
        LTE_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));         \
        const long long grid = (rows + fft_per_cta(NBV) - 1) / fft_per_cta(NBV);                            \
        k<<<(unsigned)grid, FFT_CTA_THREADS, smem, st>>>(p->dev, idx, (float2*)out, t.w, t.bf, t.tw, M, rows); \
    } break;
    switch (t.NB) {
        LAUNCH_DFTQ(64) LAUNCH_DFTQ(128) LAUNCH_DFTQ(256) LAUNCH_DFTQ(512) LAUNCH_DFTQ(1024) LAUNCH_DFTQ(2048)
        default: return LTE_ERR_UNSUPPORTED;
    }
#undef LAUNCH_DFTQ
    LTE_CHECK_CUDA(cudaGetLastError());
    return LTE_OK;
}
