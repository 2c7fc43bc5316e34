// K^-1 and K^-1 y per posterior sample for the acquisition model
// (src/bark/optimizer/opt_model.py:54-59,83,101: K_inv = np.linalg.inv(scale * K0_no_null + (1e-6 + noise) I)).
//
// The N x N inverse is never factorised: with Z the N x P one-hot leaf indicators and B = c I + Z^T Z (already
// inverted in the leaf-space state, P ~ 2.3 m << N),
//     K^-1 = (I - Z B^-1 Z^T) / sig,        K^-1 y = (y - Z w) / sig,   w = B^-1 Z^T y,   sig = noise + 1e-6.
// Z B^-1 Z^T is two gather-sums over each point's m leaf columns:
//     G[i, :]  = sum_t B^-1[col_t(i), :]            (n x P, kinv_rows_kernel)
//     S[i, j]  = sum_t G[i, col_t(j)]               (n x n, kinv_tile_kernel: 16 rows of G in shared memory)
#include "common.cuh"
#include "mcmc_state.cuh"

namespace bark {

constexpr int KI_THREADS = 256;
constexpr int KI_ROWS = 16;  // rows of G per tile CTA

struct KinvScratch {
    size_t off_cols, off_bfull, off_g, per_sample, total;
};
__host__ __device__ inline KinvScratch kinv_scratch(const WsLayout& lay) {
    KinvScratch k;
    size_t o = 0;
    k.off_cols = o;  o = align256(o + (size_t)lay.m * lay.npad * sizeof(uint16_t));  // [tree][point]
    k.off_bfull = o; o = align256(o + (size_t)lay.P * lay.P * sizeof(double));       // symmetric B^-1
    k.off_g = o;     o = align256(o + (size_t)lay.npad * lay.P * sizeof(double));    // G = Z B^-1
    k.per_sample = o;
    k.total = o * (size_t)lay.chains;
    return k;
}

// leaf-space column of every (tree, point) and the mirrored B^-1
__global__ void kinv_prep_kernel(WsLayout lay, const void* ws, const uint32_t* __restrict__ leaves, unsigned char* scratch) {
    const int64_t sample = blockIdx.y;
    ChainView cv = chain_view(lay, const_cast<void*>(ws), sample);
    const KinvScratch ks = kinv_scratch(lay);
    unsigned char* base = scratch + (size_t)sample * ks.per_sample;
    uint16_t* cols = reinterpret_cast<uint16_t*>(base + ks.off_cols);
    double* bfull = reinterpret_cast<double*>(base + ks.off_bfull);
    const int64_t n = lay.n, m = lay.m, L = lay.L, P = lay.P, npad = lay.npad;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = tid; e < m * npad; e += nth) {
        const int64_t t = e / npad, i = e % npad;
        cols[e] = (i < n) ? cv.colmap[t * L + leaves[(sample * n + i) * m + t]] : NO_COL;
    }
    for (int64_t e = tid; e < P * P; e += nth) {
        const int64_t r = e / P, c = e % P;
        bfull[e] = (c <= r) ? cv.Binv[e] : cv.Binv[c * P + r];
    }
}

// G[i, k] = sum_t Bfull[col_t(i), k]; one CTA per point, threads over k
__global__ void __launch_bounds__(KI_THREADS) kinv_rows_kernel(WsLayout lay, unsigned char* scratch) {
    extern __shared__ uint16_t ki_cols[];  // [m]
    const int64_t sample = blockIdx.y, i = blockIdx.x;
    const KinvScratch ks = kinv_scratch(lay);
    unsigned char* base = scratch + (size_t)sample * ks.per_sample;
    const uint16_t* cols = reinterpret_cast<const uint16_t*>(base + ks.off_cols);
    const double* bfull = reinterpret_cast<const double*>(base + ks.off_bfull);
    double* G = reinterpret_cast<double*>(base + ks.off_g);
    const int m = (int)lay.m, P = (int)lay.P;
    for (int t = threadIdx.x; t < m; t += KI_THREADS) ki_cols[t] = cols[(size_t)t * lay.npad + i];
    __syncthreads();
    for (int k = threadIdx.x; k < P; k += KI_THREADS) {
        double acc = 0.0;
        for (int t = 0; t < m; ++t) {
            const uint16_t c = ki_cols[t];
            if (c != NO_COL) acc += bfull[(size_t)c * P + k];
        }
        G[(size_t)i * P + k] = acc;
    }
}

// kinv[i, j] = (delta_ij - sum_t G[i, col_t(j)]) / sig for a tile of KI_ROWS rows x all j; kinv_y[i] = (y_i - sum_t w[col_t(i)]) / sig
__global__ void __launch_bounds__(KI_THREADS) kinv_tile_kernel(WsLayout lay, const void* ws, const unsigned char* scratch,
                                                                double* __restrict__ kinv, double* __restrict__ kinv_y) {
    extern __shared__ double ki_g[];  // [KI_ROWS][P + 1]
    const int64_t sample = blockIdx.y;
    const int i0 = blockIdx.x * KI_ROWS;
    ChainView cv = chain_view(lay, const_cast<void*>(ws), sample);
    SharedView sv = shared_view(lay, ws);
    const KinvScratch ks = kinv_scratch(lay);
    const unsigned char* base = scratch + (size_t)sample * ks.per_sample;
    const uint16_t* cols = reinterpret_cast<const uint16_t*>(base + ks.off_cols);
    const double* G = reinterpret_cast<const double*>(base + ks.off_g);
    const int n = (int)lay.n, m = (int)lay.m, P = (int)lay.P;
    const int64_t npad = lay.npad;
    const int rows = min(KI_ROWS, n - i0);
    const double inv_sig = 1.0 / cv.sc->sig;
    for (int e = threadIdx.x; e < rows * P; e += KI_THREADS) ki_g[(size_t)(e / P) * (P + 1) + e % P] = G[(size_t)(i0 + e / P) * P + e % P];
    __syncthreads();
    if (kinv) {
        for (int j = threadIdx.x; j < n; j += KI_THREADS) {
            double acc[KI_ROWS];
#pragma unroll
            for (int r = 0; r < KI_ROWS; ++r) acc[r] = 0.0;
            for (int t = 0; t < m; ++t) {
                const uint16_t c = cols[(size_t)t * npad + j];  // coalesced over j
                if (c != NO_COL) {
#pragma unroll
                    for (int r = 0; r < KI_ROWS; ++r) acc[r] += ki_g[(size_t)r * (P + 1) + c];
                }
            }
#pragma unroll
            for (int r = 0; r < KI_ROWS; ++r)
                if (r < rows) kinv[((size_t)sample * n + i0 + r) * n + j] = (((i0 + r) == j ? 1.0 : 0.0) - acc[r]) * inv_sig;
        }
    }
    if (kinv_y) {
        for (int r = threadIdx.x; r < rows; r += KI_THREADS) {
            double acc = 0.0;
            for (int t = 0; t < m; ++t) {
                const uint16_t c = cols[(size_t)t * npad + i0 + r];
                if (c != NO_COL) acc += cv.w[c];
            }
            kinv_y[(size_t)sample * n + i0 + r] = (sv.y[i0 + r] - acc) * inv_sig;
        }
    }
}

}  // namespace bark

using namespace bark;

extern "C" {

size_t bark_kinv_scratch_bytes(const bark_mcmc_dims* dims) {
    if (!dims || dims->chains < 1 || dims->p_cap < 64 || dims->p_cap % 64) return 0;
    return kinv_scratch(make_layout(*dims)).total;
}

int bark_kinv_export(const bark_mcmc_dims* dims, const void* workspace, const uint32_t* leaves, double* kinv,
                     double* kinv_y, void* scratch, void* stream) {
    BARK_CHECK_ARG(dims && workspace && leaves && scratch, "null pointer");
    BARK_CHECK_ARG(kinv || kinv_y, "nothing to export");
    const WsLayout lay = make_layout(*dims);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem_tile = (size_t)KI_ROWS * (lay.P + 1) * sizeof(double);
    BARK_CHECK_ARG(smem_tile <= 227 * 1024, "p_cap too large for the K^-1 tile kernel");
    BARK_CHECK_ARG(dims->chains <= 65535, "too many samples per call");
    kinv_prep_kernel<<<dim3(148, (unsigned)dims->chains), 256, 0, st>>>(lay, workspace, leaves, (unsigned char*)scratch);
    kinv_rows_kernel<<<dim3((unsigned)lay.n, (unsigned)dims->chains), KI_THREADS, (size_t)lay.m * sizeof(uint16_t), st>>>(
        lay, (unsigned char*)scratch);
    BARK_CUDA(cudaFuncSetAttribute(kinv_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tile));
    kinv_tile_kernel<<<dim3((unsigned)ceil_div(lay.n, KI_ROWS), (unsigned)dims->chains), KI_THREADS, smem_tile, st>>>(
        lay, workspace, (const unsigned char*)scratch, kinv, kinv_y);
    BARK_LAUNCH_CHECK();
    return BARK_OK;
}

}  // extern "C"
