// Samples from the BARK tree prior on the device: _sample_single_forest,
// src/bark/fitting/bark_prior_sampler.py:15-62 -- every tree starts as a root leaf; nodes are popped from a LIFO stack and
// split with probability alpha (1 + depth)^-beta by a rule drawn like a grow proposal's (feature uniform; categorical:
// uniform non-trivial subset of the categories still available at the node; integer: uniform in [lo, hi); continuous:
// U(lo, hi), stored as f32), the children take the first two inactive slots and are pushed left then right.
//
// One thread grows one tree directly in the SoA forest (trees are tiny: a handful of nodes), with its own Philox
// stream keyed by (seed, sample, tree).  Reuses the split-rule / box arithmetic of proposal_device.cuh.
#include "common.cuh"
#include "proposal_device.cuh"

namespace bark {

constexpr int PRIOR_MAX_D = 64;     // features held in the per-thread box
constexpr int PRIOR_STACK = 255;    // >= node_limit

struct PriorRng {
    uint64_t seed;
    uint32_t sample, tree, ctr;
    __device__ double next() {
        uint32_t r[4];
        philox4x32_10(sample, tree, ctr++, 0x5052494fu /* 'PRIO' */, (uint32_t)seed, (uint32_t)(seed >> 32), r);
        return u01_from_bits(r[0], r[1]);
    }
};

__global__ void prior_sample_kernel(bark_nodes_soa f, int64_t n_samples, int64_t m, int L, const double* __restrict__ bounds,
                                    const int32_t* __restrict__ ft, int d, double alpha, double beta, uint64_t seed,
                                    uint32_t* __restrict__ status) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_samples * m) return;
    const int64_t base = gid * L;
    PriorRng rng{seed, (uint32_t)(gid / m), (uint32_t)(gid % m), 0u};
    // empty tree: root leaf, everything else inactive (create_empty_forest, src/bark/forest.py:114-117)
    for (int s = 0; s < L; ++s) {
        f.is_leaf[base + s] = 0; f.active[base + s] = 0; f.feature[base + s] = 0; f.threshold[base + s] = 0.f;
        f.left[base + s] = 0; f.right[base + s] = 0; f.parent[base + s] = 0; f.depth[base + s] = 0;
    }
    f.is_leaf[base] = 1; f.active[base] = 1; f.parent[base] = 0xFFFFFFFFu;

    uint8_t stack[PRIOR_STACK];
    int sp = 0;
    stack[sp++] = 0;
    double box[2 * PRIOR_MAX_D];
    while (sp > 0) {
        const int node = stack[--sp];
        const uint32_t depth = f.depth[base + node];
        if (rng.next() > alpha * pow(1.0 + (double)depth, -beta)) continue;
        // feasible box / category mask at `node` (get_node_subspace, src/bark/fitting/tree_traversal.py:49-86)
        for (int e = 0; e < 2 * d; ++e) box[e] = bounds[e];
        int child = node;
        for (int it = 0; it < L && child != 0; ++it) {
            const int up = (int)f.parent[base + child];
            const int fe = (int)f.feature[base + up];
            const float th = f.threshold[base + up];
            const bool from_left = (uint32_t)child == f.left[base + up];
            if (ft[fe] == FEAT_CAT) {
                const long long have = (long long)box[2 * fe + 1];
                if (from_left) {
                    box[2 * fe + 1] = (double)(((long long)th) & have);
                } else {
                    const long long full = next_pow2_ll(have) - 1;
                    box[2 * fe + 1] = (double)(((long long)((double)full - (double)th)) & have);
                }
            } else if (from_left) {
                box[2 * fe + 1] = fmin((double)th, box[2 * fe + 1]);
            } else {
                box[2 * fe] = fmax((double)th + ((ft[fe] == FEAT_INT) ? 1.0 : 0.0), box[2 * fe]);
            }
            child = up;
        }
        // split rule (sample_splitting_rule, src/bark/fitting/tree_proposals.py:78-97)
        const int fe = min((int)(rng.next() * (double)d), d - 1);
        const double lo = box[2 * fe], hi = box[2 * fe + 1];
        const double ur = rng.next();
        double thr;
        if (ft[fe] == FEAT_CAT) {
            const long long avail = (long long)hi;
            const int nb = __popcll((unsigned long long)avail);
            if (nb < 2) {
                thr = 0.0;
            } else {
                const long long top = (1LL << nb) - 1;
                thr = (double)scatter_bits_ll(avail, 1 + min((long long)(ur * (double)(top - 1)), top - 2));
            }
        } else if (ft[fe] == FEAT_INT) {
            if (lo == hi) {
                thr = hi;
            } else {
                const long long li = (long long)lo, hi_i = (long long)hi;
                thr = (double)(li + min((long long)(ur * (double)(hi_i - li)), hi_i - li - 1));
            }
        } else {
            thr = lo + (hi - lo) * ur;
        }
        const float thr32 = (float)thr;
        if (thr32 == 0.f && ft[fe] == FEAT_CAT) continue;         // no non-trivial subset left (:46-51)
        if ((double)thr32 == hi && ft[fe] == FEAT_INT) continue;  // degenerate integer range (:53-58)
        // first two inactive slots, ascending (_get_two_inactive_nodes, tree_proposals.py:45-58)
        int s0 = -1, s1 = -1;
        for (int s = 0; s < L && s1 < 0; ++s)
            if (!f.active[base + s]) { if (s0 < 0) s0 = s; else s1 = s; }
        if (s1 < 0) {
            atomicOr(status, BARK_ST_TREE_OVERFLOW);  // "The tree container is not large enough"
            break;
        }
        for (int k = 0; k < 2; ++k) {  // grow (tree_proposals.py:146-165)
            const int64_t g = base + (k ? s1 : s0);
            f.is_leaf[g] = 1; f.feature[g] = 0; f.threshold[g] = 0.f; f.left[g] = 0; f.right[g] = 0;
            f.parent[g] = (uint32_t)node; f.depth[g] = depth + 1; f.active[g] = 1;
        }
        f.is_leaf[base + node] = 0; f.feature[base + node] = (uint32_t)fe; f.threshold[base + node] = thr32;
        f.left[base + node] = (uint32_t)s0; f.right[base + node] = (uint32_t)s1;
        if (sp + 2 <= PRIOR_STACK) { stack[sp++] = (uint8_t)s0; stack[sp++] = (uint8_t)s1; }
    }
}

}  // namespace bark

using namespace bark;

extern "C" {

int bark_prior_sample(bark_nodes_soa forest, int64_t n_samples, int64_t m, int64_t node_limit, const double* bounds,
                      const int32_t* feat_types, int64_t d, double alpha, double beta, uint64_t seed, uint32_t* status,
                      void* stream) {
    BARK_CHECK_ARG(n_samples >= 0 && m >= 0, "negative size");
    BARK_CHECK_ARG(node_limit >= 3 && node_limit <= 255, "node_limit out of range (3..255)");
    BARK_CHECK_ARG(d >= 1 && d <= PRIOR_MAX_D, "d out of range for the device prior sampler (1..64)");
    if (n_samples == 0 || m == 0) return BARK_OK;
    BARK_CHECK_ARG(forest.is_leaf && bounds && feat_types && status, "null pointer");
    const int64_t n = n_samples * m;
    prior_sample_kernel<<<(unsigned)ceil_div(n, 64), 64, 0, (cudaStream_t)stream>>>(forest, n_samples, m, (int)node_limit, bounds,
                                                                                    feat_types, (int)d, alpha, beta, seed, status);
    BARK_LAUNCH_CHECK();
    return BARK_OK;
}

}  // extern "C"
