// SC-FDM M-point unitary DFT / IDFT (reference core/dft_precoding.py:44-54,86-88 and :175,211:
// a dense M x M matrix product) as a Bluestein chirp-z transform on the shared-memory FFT core.
// M = number of data subcarriers (62 / 125 / 249 / 499 / 749 / 999), none of them 2-3-5 smooth.
//
//   X[k] = w[k] * sum_n (x[n] w[n]) conj(w[k-n]) / sqrt(M),   w[n] = exp(-j pi n^2 / M)
// The convolution runs through two NB-point FFTs (NB = pow2 >= 2M-1) in one kernel; the
// spectrum of the chirp filter is a per-(plan, M) table computed once on the host in fp64.
#include <math.h>

#include <map>
#include <mutex>

#include "fft.cuh"

struct DftTables {
    int M, NB;
    float2* w;      // [M]   chirp exp(-j pi n^2 / M)
    float2* bf;     // [NB]  FFT of the circular chirp filter, pre-scaled by 1/(NB sqrt(M))
    float2* tw;     // [NB]  FFT twiddles exp(-2 pi i m / NB)
};

static std::mutex g_dft_mutex;
static std::map<std::pair<int, int>, DftTables> g_dft_tables;   // (device, M) -> tables

static int get_tables(int M, DftTables* out) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_dft_mutex);
    auto it = g_dft_tables.find({dev, M});
    if (it != g_dft_tables.end()) { *out = it->second; return LTE_OK; }
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
    if (cudaMalloc(&t.w, sizeof(float2) * M) != cudaSuccess) return lte_set_cuda_error(cudaGetLastError());
    if (cudaMalloc(&t.bf, sizeof(float2) * NB) != cudaSuccess) return lte_set_cuda_error(cudaGetLastError());
    cudaMemcpy(t.w, w.data(), sizeof(float2) * M, cudaMemcpyHostToDevice);
    if (cudaMalloc(&t.tw, sizeof(float2) * NB) != cudaSuccess) return lte_set_cuda_error(cudaGetLastError());
    cudaMemcpy(t.bf, bf.data(), sizeof(float2) * NB, cudaMemcpyHostToDevice);
    cudaMemcpy(t.tw, tw.data(), sizeof(float2) * NB, cudaMemcpyHostToDevice);
    g_dft_tables[{dev, M}] = t;
    *out = t;
    return LTE_OK;
}

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
    if (rows == 0) return LTE_OK;
    DftTables t;
    int rc = get_tables(M, &t);
    if (rc) return rc;
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
