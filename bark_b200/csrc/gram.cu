// a4 (epilogue only): K = scale * ((1/m) * count) + (jitter + noise) I  from integer counts
// (src/bark/fitting/bark_sampler.py:153-156).  The counts themselves come from the tcgen05 one-hot GEMM in
// gram_umma.cu, which can also fuse this epilogue.
#include <algorithm>

#include "common.cuh"

namespace bark {

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
