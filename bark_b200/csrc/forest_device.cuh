// Device-side tree walk shared by the traversal, MCMC-init and predict kernels.
// Semantics: _pass_one_through_tree, src/bark/forest.py:28-47.
#pragma once
#include "common.cuh"

namespace bark {

// 8-byte walk record staged in shared memory (one 64-bit load per visited node).
struct __align__(8) WalkNode {
    float thr;          // split threshold (numeric) / category bitmask (categorical), as stored (f32)
    uint16_t feat_leaf; // bit 15: is_leaf; bits 0..14: feature index
    uint8_t left;       // child slots (node_limit <= 255)
    uint8_t right;
};

__device__ __forceinline__ WalkNode make_walk_node(uint8_t is_leaf, uint32_t feature, float thr, uint32_t left,
                                                   uint32_t right) {
    WalkNode w;
    w.thr = thr;
    w.feat_leaf = (uint16_t)((feature & 0x7fffu) | (is_leaf ? 0x8000u : 0u));
    w.left = (uint8_t)left;
    w.right = (uint8_t)right;
    return w;
}

// Walk one point (features at xp[f * xstride]) down one tree; returns the leaf's slot index.
// The step cap only guards against corrupt (cyclic) input; a valid tree terminates earlier.
__device__ __forceinline__ uint32_t walk_tree(const WalkNode* __restrict__ wn, const double* __restrict__ xp,
                                              int xstride, const int* __restrict__ ft, int node_limit) {
    uint32_t at = 0;
    for (int it = 0; it < node_limit; ++it) {
        const WalkNode nd = wn[at];
        if (nd.feat_leaf & 0x8000u) return at;
        const int f = nd.feat_leaf & 0x7fffu;
        at = goes_left(xp[(size_t)f * xstride], nd.thr, ft[f]) ? nd.left : nd.right;
    }
    return at;
}

}  // namespace bark
