// a1/a2: forest container codec and batched forest traversal.
//   bark_nodes_unpack / bark_nodes_pack : NODE_RECORD_DTYPE AoS <-> SoA      (src/bark/forest.py:8-19)
//   bark_traverse                        : pass_through_forest               (src/bark/forest.py:28-67)
#include <stdarg.h>

#include <algorithm>

#include "common.cuh"
#include "forest_device.cuh"

namespace bark {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

// ------------------------------------------------------------------------------------------------
// codec: one thread per record.  The 26-byte records are unaligned, so fields are assembled bytewise
// from a shared-memory staging copy of a coalesced 128-record chunk.
// ------------------------------------------------------------------------------------------------
constexpr int CODEC_THREADS = 128;

__device__ __forceinline__ uint32_t ld_u32_bytes(const uint8_t* p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__device__ __forceinline__ void st_u32_bytes(uint8_t* p, uint32_t v) {
    p[0] = (uint8_t)v;
    p[1] = (uint8_t)(v >> 8);
    p[2] = (uint8_t)(v >> 16);
    p[3] = (uint8_t)(v >> 24);
}

__global__ void __launch_bounds__(CODEC_THREADS) nodes_unpack_kernel(const uint8_t* __restrict__ aos, int64_t n,
                                                                     bark_nodes_soa soa) {
    __shared__ uint8_t stage[CODEC_THREADS * BARK_NODE_BYTES];
    const int64_t base = (int64_t)blockIdx.x * CODEC_THREADS;
    const int cnt = (int)min((int64_t)CODEC_THREADS, n - base);
    for (int b = threadIdx.x; b < cnt * BARK_NODE_BYTES; b += CODEC_THREADS)
        stage[b] = aos[base * BARK_NODE_BYTES + b];
    __syncthreads();
    if ((int)threadIdx.x < cnt) {
        const uint8_t* r = stage + threadIdx.x * BARK_NODE_BYTES;
        const int64_t i = base + threadIdx.x;
        soa.is_leaf[i] = r[0];
        soa.feature[i] = ld_u32_bytes(r + 1);
        soa.threshold[i] = __uint_as_float(ld_u32_bytes(r + 5));
        soa.left[i] = ld_u32_bytes(r + 9);
        soa.right[i] = ld_u32_bytes(r + 13);
        soa.parent[i] = ld_u32_bytes(r + 17);
        soa.depth[i] = ld_u32_bytes(r + 21);
        soa.active[i] = r[25];
    }
}

__global__ void __launch_bounds__(CODEC_THREADS) nodes_pack_kernel(bark_nodes_soa soa, int64_t n,
                                                                   uint8_t* __restrict__ aos) {
    __shared__ uint8_t stage[CODEC_THREADS * BARK_NODE_BYTES];
    const int64_t base = (int64_t)blockIdx.x * CODEC_THREADS;
    const int cnt = (int)min((int64_t)CODEC_THREADS, n - base);
    if ((int)threadIdx.x < cnt) {
        uint8_t* r = stage + threadIdx.x * BARK_NODE_BYTES;
        const int64_t i = base + threadIdx.x;
        r[0] = soa.is_leaf[i];
        st_u32_bytes(r + 1, soa.feature[i]);
        st_u32_bytes(r + 5, __float_as_uint(soa.threshold[i]));
        st_u32_bytes(r + 9, soa.left[i]);
        st_u32_bytes(r + 13, soa.right[i]);
        st_u32_bytes(r + 17, soa.parent[i]);
        st_u32_bytes(r + 21, soa.depth[i]);
        r[25] = soa.active[i];
    }
    __syncthreads();
    for (int b = threadIdx.x; b < cnt * BARK_NODE_BYTES; b += CODEC_THREADS)
        aos[base * BARK_NODE_BYTES + b] = stage[b];
}

// ------------------------------------------------------------------------------------------------
// traversal: grid = (point tiles, tree groups, forests).  A CTA stages TREES_PER_CTA trees (packed
// 8-byte walk records, slots [0, node_limit)) and a POINTS_PER_CTA x d tile of X in shared memory
// (coalesced loads), then each thread walks its point through the staged trees and writes leaf ids
// (n_forests, n_points, m) u32 -- consecutive trees of one point are contiguous, so the CTA transposes
// through shared memory to store 32-tree runs coalesced.
// ------------------------------------------------------------------------------------------------
constexpr int TRAV_THREADS = 256;   // = points per CTA
constexpr int TRAV_TREES = 32;      // trees per CTA
constexpr int TRAV_STAGE = 32;      // node slots staged per tree; a walk that reaches a higher slot reads it from global memory

// Posterior trees use the lowest slots (a grow takes the first two inactive slots, src/bark/fitting/tree_proposals.py:45-58;
// SURVEY: <= 9 of 100 slots active), so only slots [0, TRAV_STAGE) of every tree are staged: 1/3 of the node bytes of a
// full 100-slot stage, which otherwise outweigh the leaf ids the CTA writes.  Slots beyond the stage stay reachable.
// One node of a tree: the staged copy, or (slots beyond the stage) the global record.
__device__ __forceinline__ WalkNode load_walk_node(const WalkNode* __restrict__ wn, int staged, const bark_nodes_soa& nodes, int64_t gbase,
                                                   uint32_t at) {
    if ((int)at < staged) return wn[at];
    const int64_t g = gbase + at;
    return make_walk_node(nodes.is_leaf[g], nodes.feature[g], nodes.threshold[g], nodes.left[g], nodes.right[g]);
}

__global__ void __launch_bounds__(TRAV_THREADS)
traverse_kernel(bark_nodes_soa nodes, int64_t m, int node_limit, const double* __restrict__ X, int64_t n_points,
                int d, const int32_t* __restrict__ feat_types, uint32_t* __restrict__ leaves) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: WalkNode wn[TRAV_TREES][staged] | double xs[d][TRAV_THREADS+1] (feature-major, padded)
    //         | float xf[d][TRAV_THREADS+1] | int ft[d] | uint32 out[TRAV_THREADS][TRAV_TREES+1]
    const int staged = min(node_limit, TRAV_STAGE);
    WalkNode* wn = reinterpret_cast<WalkNode*>(smem_raw);
    double* xs = reinterpret_cast<double*>(wn + (size_t)TRAV_TREES * staged);
    float* xf = reinterpret_cast<float*>(xs + (size_t)d * (TRAV_THREADS + 1));
    int* ft = reinterpret_cast<int*>(xf + (size_t)(((size_t)d * (TRAV_THREADS + 1) + 1) & ~(size_t)1));
    uint32_t* outs = reinterpret_cast<uint32_t*>(ft + ((d + 1) & ~1));  // [TRAV_THREADS][TRAV_TREES + 1]

    const int64_t forest = blockIdx.z;
    const int64_t t0 = (int64_t)blockIdx.y * TRAV_TREES;
    const int nt = (int)min((int64_t)TRAV_TREES, m - t0);
    const int64_t p0 = (int64_t)blockIdx.x * TRAV_THREADS;
    const int np = (int)min((int64_t)TRAV_THREADS, n_points - p0);

    // stage the low slots of the CTA's trees
    const int64_t node_base = (forest * m + t0) * node_limit;
    for (int e = threadIdx.x; e < nt * staged; e += TRAV_THREADS) {
        const int t = e / staged, sl = e - t * staged;
        const int64_t g = node_base + (int64_t)t * node_limit + sl;
        wn[e] = make_walk_node(nodes.is_leaf[g], nodes.feature[g], nodes.threshold[g], nodes.left[g], nodes.right[g]);
    }
    bool any_cat = false;
    for (int e = threadIdx.x; e < d; e += TRAV_THREADS) ft[e] = feat_types[e];
    for (int e = 0; e < d; ++e) any_cat |= (feat_types[e] == FEAT_CAT);  // CTA-uniform
    // stage X tile feature-major.  Numeric splits are
    // decided in FP32 on the candidates rounded UP: x <= (double)thr with an f32 threshold  <=>  ru_f32(x) <= thr -- the
    // reference's comparison (src/bark/forest.py:33-47) bit for bit, off the FP64 pipe.
    // (thread = point, its d features are consecutive in memory: a warp's loads cover one contiguous 32 d-double block, no
    // integer division per element)
    if ((int)threadIdx.x < np) {
        const double* xr = X + (p0 + threadIdx.x) * d;
        for (int f = 0; f < d; ++f) {
            const double x = xr[f];
            xs[(size_t)f * (TRAV_THREADS + 1) + threadIdx.x] = x;
            xf[(size_t)f * (TRAV_THREADS + 1) + threadIdx.x] = __double2float_ru(x);
        }
    }
    __syncthreads();

    if ((int)threadIdx.x < np) {
        const double* xp = xs + threadIdx.x;
        const float* xfp = xf + threadIdx.x;
        // four trees in flight per thread, branch-free (a leaf steps to itself): the four pointer chases overlap
        for (int t = 0; t < nt; t += 4) {
            WalkNode nd[4];
            uint32_t cur[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int tu = min(t + u, nt - 1);
                cur[u] = 0;
                nd[u] = wn[(size_t)tu * staged];
            }
            for (int it = 0; it < node_limit; ++it) {
                if ((nd[0].feat_leaf & nd[1].feat_leaf & nd[2].feat_leaf & nd[3].feat_leaf) & 0x8000u) break;
                float xv[4];
                int fi[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    fi[u] = min((int)(nd[u].feat_leaf & 0x7fffu), d - 1);  // (a leaf's feature bits are ignored below)
                    xv[u] = xfp[(size_t)fi[u] * (TRAV_THREADS + 1)];
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    bool left = xv[u] <= nd[u].thr;
                    if (any_cat && ft[fi[u]] == FEAT_CAT) left = goes_left(xp[(size_t)fi[u] * (TRAV_THREADS + 1)], nd[u].thr, FEAT_CAT);
                    const uint32_t at = min((uint32_t)(left ? nd[u].left : nd[u].right), (uint32_t)(node_limit - 1));
                    cur[u] = (nd[u].feat_leaf & 0x8000u) ? cur[u] : at;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int tu = min(t + u, nt - 1);
                    nd[u] = load_walk_node(wn + (size_t)tu * staged, staged, nodes, node_base + (int64_t)tu * node_limit, cur[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (t + u < nt) outs[threadIdx.x * (TRAV_TREES + 1) + t + u] = cur[u];
        }
    }
    __syncthreads();
    // coalesced store: leaves[forest][p][t0 + t]
    for (int e = threadIdx.x; e < np * nt; e += TRAV_THREADS) {
        const int p = e / nt, t = e % nt;
        leaves[(forest * n_points + p0 + p) * m + t0 + t] = outs[p * (TRAV_TREES + 1) + t];
    }
}

}  // namespace bark

using namespace bark;

extern "C" {

int bark_abi_version(void) { return BARK_ABI_VERSION; }
const char* bark_last_error(void) { return g_last_error; }

int bark_nodes_unpack(const uint8_t* aos, int64_t n_nodes, bark_nodes_soa soa, void* stream) {
    BARK_CHECK_ARG(n_nodes >= 0, "n_nodes < 0");
    if (n_nodes == 0) return BARK_OK;
    BARK_CHECK_ARG(aos && soa.is_leaf && soa.active && soa.feature && soa.threshold && soa.left && soa.right &&
                       soa.parent && soa.depth, "null pointer");
    nodes_unpack_kernel<<<(unsigned)ceil_div(n_nodes, CODEC_THREADS), CODEC_THREADS, 0, (cudaStream_t)stream>>>(
        aos, n_nodes, soa);
    BARK_LAUNCH_CHECK();
    return BARK_OK;
}

int bark_nodes_pack(bark_nodes_soa soa, int64_t n_nodes, uint8_t* aos, void* stream) {
    BARK_CHECK_ARG(n_nodes >= 0, "n_nodes < 0");
    if (n_nodes == 0) return BARK_OK;
    BARK_CHECK_ARG(aos && soa.is_leaf && soa.active && soa.feature && soa.threshold && soa.left && soa.right &&
                       soa.parent && soa.depth, "null pointer");
    nodes_pack_kernel<<<(unsigned)ceil_div(n_nodes, CODEC_THREADS), CODEC_THREADS, 0, (cudaStream_t)stream>>>(
        soa, n_nodes, aos);
    BARK_LAUNCH_CHECK();
    return BARK_OK;
}

int bark_traverse(bark_nodes_soa nodes, int64_t n_forests, int64_t m, int64_t node_limit, const double* X,
                  int64_t n_points, int64_t d, const int32_t* feat_types, uint32_t* leaves, void* stream) {
    BARK_CHECK_ARG(n_forests >= 0 && m >= 0 && n_points >= 0, "negative size");
    BARK_CHECK_ARG(node_limit >= 1 && node_limit <= 255, "node_limit out of range (1..255)");
    BARK_CHECK_ARG(d >= 1 && d <= 32767, "d out of range");
    if (n_forests == 0 || m == 0 || n_points == 0) return BARK_OK;
    BARK_CHECK_ARG(X && feat_types && leaves && nodes.is_leaf, "null pointer");
    BARK_CHECK_ARG(n_forests <= 65535 && ceil_div(m, TRAV_TREES) <= 65535, "grid too large");
    size_t smem = (size_t)TRAV_TREES * std::min<int64_t>(node_limit, TRAV_STAGE) * sizeof(WalkNode) +
                  (size_t)d * (TRAV_THREADS + 1) * sizeof(double) +
                  (((size_t)d * (TRAV_THREADS + 1) + 1) & ~(size_t)1) * sizeof(float) +
                  (size_t)((d + 1) & ~1) * sizeof(int) + (size_t)TRAV_THREADS * (TRAV_TREES + 1) * sizeof(uint32_t);
    BARK_CHECK_ARG(smem <= 220 * 1024, "d * node_limit too large for the shared-memory staging");
    BARK_CUDA(cudaFuncSetAttribute(traverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div(n_points, TRAV_THREADS), (unsigned)ceil_div(m, TRAV_TREES), (unsigned)n_forests);
    traverse_kernel<<<grid, TRAV_THREADS, smem, (cudaStream_t)stream>>>(nodes, m, (int)node_limit, X, n_points, (int)d,
                                                                        feat_types, leaves);
    BARK_LAUNCH_CHECK();
    return BARK_OK;
}

}  // extern "C"
