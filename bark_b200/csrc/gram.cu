// a4: leaf co-occurrence Gram.
//   bark_gram_counts    : count[i,j] = #{t : leaf_t(x_i) == leaf_t(x'_j)}  -- exact int32
//                         (the np.sum(np.equal(...)) of src/bark/forest.py:85-88)
//   bark_gram_to_kernel : K = scale * ((1/m) * count) + (jitter + noise) I  (bark_sampler.py:153-156)
//
// Leaf slot ids are < node_limit <= 255, so a (point, tree) id is one byte.  A CTA owns a 64x64 output tile,
// stages 64-tree slabs of both operands as packed bytes in shared memory and compares four trees per
// instruction (__vcmpeq4 + popc); every thread accumulates a 4x4 register block.
#include <algorithm>

#include "common.cuh"

namespace bark {

constexpr int GT = 64;         // output tile edge
constexpr int GK_TREES = 64;   // trees per staged slab
constexpr int GK_WORDS = GK_TREES / 4;
constexpr int G_THREADS = 256;

__global__ void __launch_bounds__(G_THREADS)
gram_counts_kernel(const uint32_t* __restrict__ la, const uint32_t* __restrict__ lb, int64_t na, int64_t nb, int64_t m,
                   int32_t* __restrict__ counts) {
    __shared__ uint32_t As[GT][GK_WORDS + 1];
    __shared__ uint32_t Bs[GT][GK_WORDS + 1];
    const int64_t b = blockIdx.z;
    const int64_t i0 = (int64_t)blockIdx.y * GT, j0 = (int64_t)blockIdx.x * GT;
    const uint32_t* A = la + b * na * m;
    const uint32_t* B = lb + b * nb * m;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;  // 16 x 16 threads, 4x4 outputs each
    int acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0;

    for (int64_t t0 = 0; t0 < m; t0 += GK_TREES) {
        __syncthreads();
        // pack 4 tree ids per word; pad with ids that can never match (0xFF vs 0xFE)
        for (int e = threadIdx.x; e < GT * GK_WORDS; e += G_THREADS) {
            const int r = e / GK_WORDS, w = e % GK_WORDS;
            uint32_t pa = 0, pb = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int64_t t = t0 + w * 4 + q;
                uint32_t va = 0xFFu, vb = 0xFEu;
                if (t < m) {
                    if (i0 + r < na) va = A[(i0 + r) * m + t] & 0xFFu;
                    if (j0 + r < nb) vb = B[(j0 + r) * m + t] & 0xFFu;
                }
                pa |= va << (8 * q);
                pb |= vb << (8 * q);
            }
            As[r][w] = pa;
            Bs[r][w] = pb;
        }
        __syncthreads();
#pragma unroll 4
        for (int w = 0; w < GK_WORDS; ++w) {
            uint32_t a[4], bb[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[ty * 4 + i][w];
#pragma unroll
            for (int j = 0; j < 4; ++j) bb[j] = Bs[tx + 16 * j][w];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += __popc(__vcmpeq4(a[i], bb[j])) >> 3;
        }
    }
    int32_t* C = counts + b * na * nb;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = i0 + ty * 4 + i;
        if (r >= na) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t c = j0 + tx + 16 * j;
            if (c < nb) C[r * nb + c] = acc[i][j];
        }
    }
}

__global__ void gram_to_kernel_kernel(const int32_t* __restrict__ counts, int64_t na, int64_t nb, double inv_m,
                                      const double* __restrict__ scale, const double* __restrict__ noise, double jitter,
                                      int add_diag, double* __restrict__ K) {
    const int64_t b = blockIdx.y;
    const int64_t total = na * nb;
    const double s = scale[b];
    // (1e-6 + noise) * eye : computed once, added with an explicit non-fused add
    const double dg = add_diag ? __dadd_rn(jitter, noise[b]) : 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / nb, c = e % nb;
        // scale * ((1/m) * count): two separately rounded multiplies, as numpy evaluates it
        double v = __dmul_rn(s, __dmul_rn(inv_m, (double)counts[b * total + e]));
        if (add_diag && r == c) v = __dadd_rn(v, dg);
        K[b * total + e] = v;
    }
}

}  // namespace bark

using namespace bark;

extern "C" {

int bark_gram_counts(const uint32_t* leaves_a, const uint32_t* leaves_b, int64_t batch, int64_t na, int64_t nb,
                     int64_t m, int32_t* counts, void* stream) {
    BARK_CHECK_ARG(batch >= 0 && na >= 0 && nb >= 0 && m >= 0, "negative size");
    if (batch == 0 || na == 0 || nb == 0) return BARK_OK;
    BARK_CHECK_ARG(leaves_a && leaves_b && counts, "null pointer");
    BARK_CHECK_ARG(batch <= 65535 && ceil_div(na, GT) <= 65535, "grid too large");
    dim3 grid((unsigned)ceil_div(nb, GT), (unsigned)ceil_div(na, GT), (unsigned)batch);
    gram_counts_kernel<<<grid, G_THREADS, 0, (cudaStream_t)stream>>>(leaves_a, leaves_b, na, nb, m, counts);
    BARK_LAUNCH_CHECK();
    return BARK_OK;
}

int bark_gram_to_kernel(const int32_t* counts, int64_t batch, int64_t na, int64_t nb, int64_t m, const double* scale,
                        const double* noise, double jitter, int add_diag, double* K, void* stream) {
    BARK_CHECK_ARG(batch >= 0 && na >= 0 && nb >= 0 && m >= 1, "bad size");
    if (batch == 0 || na == 0 || nb == 0) return BARK_OK;
    BARK_CHECK_ARG(counts && scale && K && (noise || !add_diag), "null pointer");
    BARK_CHECK_ARG(!add_diag || na == nb, "add_diag needs a square matrix");
    BARK_CHECK_ARG(batch <= 65535, "batch too large");
    const int64_t total = na * nb;
    dim3 grid((unsigned)std::min<int64_t>(ceil_div(total, 256), 148 * 8), (unsigned)batch);
    gram_to_kernel_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(counts, na, nb, 1.0 / (double)m, scale, noise, jitter,
                                                                 add_diag, K);
    BARK_LAUNCH_CHECK();
    return BARK_OK;
}

}  // extern "C"
