// a14-a15: posterior predictive mean / variance in leaf space.
//
// For one posterior sample (forest, noise, scale) with B = c I + Z^T Z, b = Z^T y, w = B^-1 b
// (state built by bark_mcmc_init over the samples), a candidate x with leaf columns c_1..c_m has
//     mu(x)  = k^T K^-1 y        = sum_t w[c_t]
//     var(x) = s - k^T K^-1 k    = sig * z^T B^-1 z = sig * sum_{t,t'} Binv[c_t][c_t']
// which equals forest_predict's  K_xX K^-1 y  and  scale - diag(K_xX K^-1 K_Xx)
// (src/bark/tree_kernels/tree_gps.py:97-112) without forming any n_c x n or n_c x n_c matrix.
#include <algorithm>

#include "common.cuh"
#include "forest_device.cuh"
#include "mcmc_state.cuh"

namespace bark {

constexpr int PR_THREADS = 128;  // candidates per CTA

// walk table: leaves carry their leaf-space column in the threshold bits
__global__ void predict_pack_kernel(WsLayout lay, const void* ws, bark_nodes_soa forest, WalkNode* __restrict__ table) {
    const int64_t total = lay.chains * lay.m * lay.L;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t chain = e / (lay.m * lay.L);
        const int64_t rem = e % (lay.m * lay.L);
        ChainView cv = chain_view(lay, const_cast<void*>(ws), chain);
        WalkNode w = make_walk_node(forest.is_leaf[e], forest.feature[e], forest.threshold[e], forest.left[e], forest.right[e]);
        if (forest.is_leaf[e]) w.thr = __int_as_float((int)cv.colmap[rem]);
        table[e] = w;
    }
}

__global__ void __launch_bounds__(PR_THREADS)
predict_sample_kernel(WsLayout lay, const void* ws, const WalkNode* __restrict__ table, const double* __restrict__ cand,
                      int64_t n_c, double* __restrict__ mu, double* __restrict__ var) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int d = (int)lay.d, m = (int)lay.m, L = (int)lay.L, P = (int)lay.P;
    double* xs = reinterpret_cast<double*>(smem_raw);                              // [d][PR_THREADS + 1]
    uint16_t* cols = reinterpret_cast<uint16_t*>(xs + (size_t)d * (PR_THREADS + 1));  // [m][PR_THREADS]
    int* ftc = reinterpret_cast<int*>(cols + (size_t)m * PR_THREADS + (m * PR_THREADS & 1));

    const int64_t sample = blockIdx.y;
    const int64_t p0 = (int64_t)blockIdx.x * PR_THREADS;
    const int np = (int)min((int64_t)PR_THREADS, n_c - p0);
    ChainView cv = chain_view(lay, const_cast<void*>(ws), sample);
    SharedView sv = shared_view(lay, ws);
    const WalkNode* tb = table + sample * (int64_t)m * L;

    for (int e = threadIdx.x; e < np * d; e += PR_THREADS) xs[(size_t)(e % d) * (PR_THREADS + 1) + e / d] = cand[p0 * d + e];
    for (int e = threadIdx.x; e < d; e += PR_THREADS) ftc[e] = sv.ft[e];
    __syncthreads();
    if ((int)threadIdx.x >= np) return;

    const double* xp = xs + threadIdx.x;
    const double* __restrict__ w = cv.w;
    double mean = 0.0;
    for (int t = 0; t < m; ++t) {
        const WalkNode* wn = tb + (size_t)t * L;
        uint32_t at = 0;
        WalkNode nd = wn[0];
        for (int it = 0; it < L && !(nd.feat_leaf & 0x8000u); ++it) {
            const int f = nd.feat_leaf & 0x7fffu;
            at = goes_left(xp[(size_t)f * (PR_THREADS + 1)], nd.thr, ftc[f]) ? nd.left : nd.right;
            nd = wn[at];
        }
        const int col = __float_as_int(nd.thr);
        cols[(size_t)t * PR_THREADS + threadIdx.x] = (uint16_t)col;
        mean += w[col];
    }
    const double* __restrict__ Binv = cv.Binv;
    double dg = 0.0, off = 0.0;
    for (int t = 0; t < m; ++t) {
        // only the lower triangle of Binv is kept current by the sampler: index (max, min)
        const int ci = cols[(size_t)t * PR_THREADS + threadIdx.x];
        dg += Binv[(size_t)ci * P + ci];
        double part = 0.0;
        for (int t2 = 0; t2 < t; ++t2) {
            const int cj = cols[(size_t)t2 * PR_THREADS + threadIdx.x];
            part += Binv[(size_t)max(ci, cj) * P + min(ci, cj)];
        }
        off += part;
    }
    const int64_t o = sample * n_c + p0 + threadIdx.x;
    mu[o] = mean;
    var[o] = cv.sc->sig * (dg + 2.0 * off);
}

// mixture of Gaussians over the samples after un-standardisation (bark.py:83-91, tree_gps.py:116-131)
__global__ void predict_mixture_kernel(WsLayout lay, const void* ws, const double* __restrict__ mu_s,
                                       const double* __restrict__ var_s, int64_t n_c, double y_mean, double y_std,
                                       int add_noise, double* __restrict__ mu, double* __restrict__ var) {
    const int64_t S = lay.chains;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_c; i += (int64_t)gridDim.x * blockDim.x) {
        double sm = 0.0, s2 = 0.0;
        for (int64_t j = 0; j < S; ++j) {
            const double noise = chain_view(lay, const_cast<void*>(ws), j).sc->noise;
            const double mj = mu_s[j * n_c + i] * y_std + y_mean;
            double vj = var_s[j * n_c + i] * (y_std * y_std);
            if (add_noise) vj += noise;
            sm += mj;
            s2 += vj + mj * mj;
        }
        const double e = sm / (double)S;
        mu[i] = e;
        var[i] = s2 / (double)S - e * e;
    }
}

// ---- diag = False (src/bark/tree_kernels/tree_gps.py:107-112): the full n_c x n_c matrix the reference forms,
//     cov[i][j] = scale - (K_xX K^-1 K_Xx)[i][j] = scale - (scale / m) #{t : leaf_t(x_i) = leaf_t(x_j)} + sig z_i^T B^-1 z_j
// (the reference subtracts from the SCALAR scale, also off the diagonal; reproduced as written).  Three small kernels:
// leaf columns + mean per candidate, G = Z_c B^-1 (n_c x P gathers), cov from m gathers of G per pair.
__global__ void __launch_bounds__(PR_THREADS)
predict_cols_kernel(WsLayout lay, const void* ws, const WalkNode* __restrict__ table, const double* __restrict__ cand,
                    int64_t n_c, uint16_t* __restrict__ cols, double* __restrict__ mu) {
    const int d = (int)lay.d, m = (int)lay.m, L = (int)lay.L;
    const int64_t sample = blockIdx.y;
    const int64_t i = (int64_t)blockIdx.x * PR_THREADS + threadIdx.x;
    if (i >= n_c) return;
    ChainView cv = chain_view(lay, const_cast<void*>(ws), sample);
    SharedView sv = shared_view(lay, ws);
    const WalkNode* tb = table + sample * (int64_t)m * L;
    const double* xp = cand + i * d;
    double mean = 0.0;
    for (int t = 0; t < m; ++t) {
        const WalkNode* wn = tb + (size_t)t * L;
        WalkNode nd = wn[0];
        for (int it = 0; it < L && !(nd.feat_leaf & 0x8000u); ++it) {
            const int f = nd.feat_leaf & 0x7fffu;
            nd = wn[goes_left(xp[f], nd.thr, sv.ft[f]) ? nd.left : nd.right];
        }
        const int col = __float_as_int(nd.thr);
        cols[(sample * n_c + i) * m + t] = (uint16_t)col;
        mean += cv.w[col];
    }
    mu[sample * n_c + i] = mean;
}

__global__ void predict_g_kernel(WsLayout lay, const void* ws, const uint16_t* __restrict__ cols, int64_t n_c,
                                 double* __restrict__ G) {
    const int m = (int)lay.m, P = (int)lay.P;
    const int64_t sample = blockIdx.y, i = blockIdx.x;
    const double* Binv = chain_view(lay, const_cast<void*>(ws), sample).Binv;
    const uint16_t* ci = cols + (sample * n_c + i) * m;
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        double acc = 0.0;
        for (int t = 0; t < m; ++t) {
            const int c = ci[t];
            acc += Binv[(size_t)max(c, p) * P + min(c, p)];  // symmetric; the lower triangle is always current
        }
        G[(sample * n_c + i) * P + p] = acc;
    }
}

__global__ void predict_cov_kernel(WsLayout lay, const void* ws, const uint16_t* __restrict__ cols, const double* __restrict__ G,
                                   int64_t n_c, double* __restrict__ cov) {
    const int m = (int)lay.m, P = (int)lay.P;
    const int64_t sample = blockIdx.y, i = blockIdx.x;
    const ChainScalars* sc = chain_view(lay, const_cast<void*>(ws), sample).sc;
    const double scale = sc->scale, sig = sc->sig, s_m = scale / (double)m;
    const uint16_t* ci = cols + (sample * n_c + i) * m;
    const double* gi = G + (sample * n_c + i) * P;
    for (int64_t j = threadIdx.x; j < n_c; j += blockDim.x) {
        const uint16_t* cj = cols + (sample * n_c + j) * m;
        double q = 0.0;
        int same = 0;
        for (int t = 0; t < m; ++t) {
            q += gi[cj[t]];
            same += (ci[t] == cj[t]);
        }
        cov[(sample * n_c + i) * n_c + j] = scale - s_m * (double)same + sig * q;
    }
}

static size_t table_bytes(const bark_mcmc_dims* dm) {
    return align256((size_t)dm->chains * dm->m * dm->node_limit * sizeof(WalkNode));
}

}  // namespace bark

using namespace bark;

extern "C" {

size_t bark_predict_scratch_bytes(const bark_mcmc_dims* dims, int64_t n_c) {
    if (!dims || n_c < 0) return 0;
    return table_bytes(dims) + 2 * align256((size_t)dims->chains * (size_t)n_c * sizeof(double));
}

int bark_predict_mixture(const bark_mcmc_dims* dims, const void* workspace, const double* mu_s, const double* var_s,
                         int64_t n_c, double y_mean, double y_std, int add_noise, double* mu, double* var, void* stream) {
    BARK_CHECK_ARG(dims && workspace, "null pointer");
    if (n_c <= 0) return BARK_OK;
    BARK_CHECK_ARG(mu_s && var_s && mu && var, "null pointer");
    const WsLayout lay = make_layout(*dims);
    predict_mixture_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n_c, 256), 148 * 8), 256, 0, (cudaStream_t)stream>>>(
        lay, workspace, mu_s, var_s, n_c, y_mean, y_std, add_noise, mu, var);
    BARK_LAUNCH_CHECK();
    return BARK_OK;
}

size_t bark_predict_cov_scratch_bytes(const bark_mcmc_dims* dims, int64_t n_c) {
    if (!dims || n_c < 0) return 0;
    return table_bytes(dims) + align256((size_t)dims->chains * (size_t)n_c * dims->m * sizeof(uint16_t)) +
           align256((size_t)dims->chains * (size_t)n_c * dims->p_cap * sizeof(double));
}

int bark_predict_cov(const bark_mcmc_dims* dims, const void* workspace, bark_nodes_soa forest, const double* candidates,
                     int64_t n_c, double* mu, double* cov, void* scratch, void* stream) {
    BARK_CHECK_ARG(dims && workspace && scratch && forest.is_leaf, "null pointer");
    BARK_CHECK_ARG(n_c >= 0 && n_c <= 16384, "n_c out of range for the full covariance (0..16384)");
    if (n_c == 0) return BARK_OK;
    BARK_CHECK_ARG(candidates && mu && cov, "null pointer");
    BARK_CHECK_ARG(dims->chains <= 65535, "too many samples per call");
    const WsLayout lay = make_layout(*dims);
    cudaStream_t st = (cudaStream_t)stream;
    WalkNode* table = (WalkNode*)scratch;
    uint16_t* cols = (uint16_t*)((unsigned char*)scratch + table_bytes(dims));
    double* G = (double*)((unsigned char*)cols + align256((size_t)dims->chains * (size_t)n_c * dims->m * sizeof(uint16_t)));
    predict_pack_kernel<<<148 * 2, 256, 0, st>>>(lay, workspace, forest, table);
    dim3 g1((unsigned)ceil_div(n_c, PR_THREADS), (unsigned)dims->chains);
    predict_cols_kernel<<<g1, PR_THREADS, 0, st>>>(lay, workspace, table, candidates, n_c, cols, mu);
    dim3 g2((unsigned)n_c, (unsigned)dims->chains);
    predict_g_kernel<<<g2, 256, 0, st>>>(lay, workspace, cols, n_c, G);
    predict_cov_kernel<<<g2, 256, 0, st>>>(lay, workspace, cols, G, n_c, cov);
    BARK_LAUNCH_CHECK();
    return BARK_OK;
}

int bark_predict(const bark_mcmc_dims* dims, const void* workspace, bark_nodes_soa forest, const double* candidates,
                 int64_t n_c, int mode, double y_mean, double y_std, int add_noise, double* mu, double* var,
                 void* scratch, void* stream) {
    BARK_CHECK_ARG(dims && workspace && scratch && forest.is_leaf, "null pointer");
    BARK_CHECK_ARG(mode == 0 || mode == 1, "mode must be 0 or 1");
    BARK_CHECK_ARG(n_c >= 0, "n_c < 0");
    if (n_c == 0) return BARK_OK;
    BARK_CHECK_ARG(candidates && mu && var, "null pointer");
    BARK_CHECK_ARG(dims->chains <= 65535, "too many samples per call");
    const WsLayout lay = make_layout(*dims);
    cudaStream_t st = (cudaStream_t)stream;
    WalkNode* table = (WalkNode*)scratch;
    double* mu_s = (double*)((unsigned char*)scratch + table_bytes(dims));
    double* var_s = (double*)((unsigned char*)mu_s + align256((size_t)dims->chains * (size_t)n_c * sizeof(double)));
    predict_pack_kernel<<<148 * 2, 256, 0, st>>>(lay, workspace, forest, table);
    BARK_LAUNCH_CHECK();
    const size_t smem = (size_t)lay.d * (PR_THREADS + 1) * sizeof(double) + ((size_t)lay.m * PR_THREADS + 1) * sizeof(uint16_t) +
                        (size_t)lay.d * sizeof(int) + 16;
    BARK_CHECK_ARG(smem <= 227 * 1024, "d / m too large for the predict kernel's shared memory");
    BARK_CUDA(cudaFuncSetAttribute(predict_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div(n_c, PR_THREADS), (unsigned)dims->chains);
    double* out_mu = mode == 0 ? mu : mu_s;
    double* out_var = mode == 0 ? var : var_s;
    predict_sample_kernel<<<grid, PR_THREADS, smem, st>>>(lay, workspace, table, candidates, n_c, out_mu, out_var);
    BARK_LAUNCH_CHECK();
    if (mode == 1) {
        predict_mixture_kernel<<<(unsigned)std::min<int64_t>(ceil_div(n_c, 256), 148 * 8), 256, 0, st>>>(
            lay, workspace, mu_s, var_s, n_c, y_mean, y_std, add_noise, mu, var);
        BARK_LAUNCH_CHECK();
    }
    return BARK_OK;
}

}  // extern "C"
