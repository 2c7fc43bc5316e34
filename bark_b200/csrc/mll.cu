// a5: batched FP64 log-marginal-likelihood  mll = 0.5 * (-y^T K^-1 y - log|K|)
// (src/bark/fitting/quick_inverse.py:36-38 evaluated from scratch as in bark_sampler.py:160-162,269-272).
// A thread-block CLUSTER per matrix (as many CTAs -- 1, 2, 4 or 8 -- as keep the whole batch resident in one wave)
// runs the forward block sweep of linalg.cuh (square-root-free block Cholesky): the 128 x 128 tiles of the panel and of
// the trailing update are dealt round-robin over the cluster, the pivot block is inverted by rank 0.
#include "common.cuh"
#include "linalg.cuh"

namespace bark {

struct MllScratch {
    double* ck;
    double* gk;
    double* yv;
};

__host__ __device__ inline size_t mll_scratch_doubles(int64_t n) { return (size_t)n * la::NB * 2 + (size_t)la::NB * la::NB + (size_t)((n + 1) & ~1LL); }

__global__ void __launch_bounds__(la::THREADS, 1)
mll_batched_kernel(double* K, int n, const double* __restrict__ y, double* out_mll, double* out_logdet,
                   double* out_quad, uint32_t* status, double* scratch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    la::Smem& s = *reinterpret_cast<la::Smem*>(smem_raw);
    la::ClusterTeam team{cooperative_groups::this_cluster()};
    const int trank = team.rank(), tsize = team.size();
    const int64_t b = blockIdx.x / tsize;
    double* W = K + b * (int64_t)n * n;
    double* base = scratch + b * mll_scratch_doubles(n);
    double* ck = base;
    double* gk = base + (size_t)n * la::NB;
    double* dg = gk + (size_t)n * la::NB;
    double* yv = dg + (size_t)la::NB * la::NB;
    if (trank == 0)
        for (int i = threadIdx.x; i < n; i += la::THREADS) __stcg(yv + i, y[i]);
    team.sync();
    double quad = 0.0;
    const double logdet = la::block_sweep<false>(W, n, n, ck, gk, dg, yv, &quad, s, status ? status + b : nullptr, team);
    if (trank == 0 && threadIdx.x == 0) {
        if (out_logdet) out_logdet[b] = logdet;
        if (out_quad) out_quad[b] = quad;
        if (out_mll) out_mll[b] = 0.5 * (-quad - logdet);
    }
}

}  // namespace bark

using namespace bark;

extern "C" {

size_t bark_mll_workspace_bytes(int64_t batch, int64_t n) {
    if (batch <= 0 || n <= 0) return 0;
    return (size_t)batch * mll_scratch_doubles(n) * sizeof(double);
}

int bark_mll_batched(double* K, int64_t batch, int64_t n, const double* y, double* out_mll, double* out_logdet,
                     double* out_quad, uint32_t* status, void* workspace, void* stream) {
    BARK_CHECK_ARG(batch >= 0 && n >= 0, "negative size");
    if (batch == 0) return BARK_OK;
    BARK_CHECK_ARG(n >= 1 && n <= (1 << 20), "n out of range");
    BARK_CHECK_ARG(K && y && workspace, "null pointer");
    BARK_CUDA(cudaFuncSetAttribute(mll_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)sizeof(la::Smem)));
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int R = 1;
    while (R < 8 && batch * (R * 2) <= sms && n > 2 * la::TILE * R) R *= 2;  // small matrices have no tiles to share out
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(batch * R));
    cfg.blockDim = dim3(la::THREADS);
    cfg.dynamicSmemBytes = sizeof(la::Smem);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)R;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    BARK_CUDA(cudaLaunchKernelEx(&cfg, mll_batched_kernel, K, (int)n, y, out_mll, out_logdet, out_quad, status,
                                 (double*)workspace));
    BARK_LAUNCH_CHECK();
    return BARK_OK;
}

}  // extern "C"
