// Device-side proposal generation: RNG, tree-structure queries, split-rule sampling, log q / prior ratios,
// noise/scale random walks.  Mirrors (and is parity-tested against oracle/bark_oracle.py for):
//   get_tree_proposal            src/bark/fitting/tree_proposals.py:186-256
//   terminal/singly_internal     src/bark/fitting/tree_traversal.py:28-46
//   get_node_subspace            src/bark/fitting/tree_traversal.py:49-86
//   sample_binary_mask           src/bark/utils/bit_operations.py:34-58
//   get_noise_scale_proposal     src/bark/fitting/noise_scale_proposals.py:70-156
#pragma once
#include "common.cuh"

namespace bark {

constexpr int MOVE_GROW = 0, MOVE_PRUNE = 1, MOVE_CHANGE = 2;
constexpr int TAPE_PER_TREE = 5;   // u_type, u_node, u_feat, u_rule, u_accept
constexpr int TAPE_PER_HYPER = 3;  // z_noise, z_scale, u_accept

// ---------------------------------------------------------------- Philox4x32-10 (counter-based RNG)
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ double u01_from_bits(uint32_t a, uint32_t b) {
    // 53 random bits -> [0, 1)
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}
// `count` (<= 6) uniforms for (chain, sweep, item); item = tree index, or m for the hyper step.
__device__ __forceinline__ void rng_uniforms(uint64_t seed, uint32_t chain, uint32_t sweep, uint32_t item, int count,
                                             double* u) {
    uint32_t r[4];
    for (int blk = 0; blk * 2 < count; ++blk) {
        philox4x32_10(chain, sweep, item, (uint32_t)blk, (uint32_t)seed, (uint32_t)(seed >> 32), r);
        u[blk * 2] = u01_from_bits(r[0], r[1]);
        if (blk * 2 + 1 < count) u[blk * 2 + 1] = u01_from_bits(r[2], r[3]);
    }
}
__device__ __forceinline__ double std_normal_from(double u1, double u2) {
    // Box-Muller on (0,1] x [0,1)
    return sqrt(-2.0 * log(1.0 - u1)) * cospi(2.0 * u2);
}

// ---------------------------------------------------------------- one tree staged in shared memory (SoA)
struct TreeSmem {
    uint8_t* is_leaf; uint8_t* active;
    uint32_t* feat; uint32_t* left; uint32_t* right; uint32_t* parent; uint32_t* depth;
    float* thr;
};

struct Prop {
    int move;     // MOVE_*
    int valid;    // 0: proposal is (nodes, -inf) and can never be accepted
    int node;     // edited node slot
    int feat;     // new split feature (grow / change)
    float thr;    // new split threshold / mask, already rounded to f32 (NodeProposal.new_threshold)
    int sl, sr;   // child slots: newly allocated (grow) or existing (prune / change)
    int a, b;     // leaf-space columns: Z' = Z + u (e_a - e_b)^T
    double lqp;   // log q-ratio + log prior-ratio
    unsigned depth;  // depth of the edited node (children of a grow get depth + 1)
    int pad;
};

__device__ __forceinline__ long long next_pow2_ll(long long x) {  // bit_operations.py:5-10
    long long p = 1;
    while (x >= p) p <<= 1;
    return p;
}
__device__ __forceinline__ long long scatter_bits_ll(long long available, long long packed) {  // :52-56
    long long out = 0;
    for (int i = 0; i < 63; ++i) {
        if ((1LL << i) > available) break;
        if (available & (1LL << i)) {
            out |= (packed & 1LL) << i;
            packed >>= 1;
        }
    }
    return out;
}

// log prior ratio in the grow direction at depth d (tree_proposals.py:136-140)
__device__ __forceinline__ double log_prior_ratio_at_depth(uint32_t depth, double alpha, double beta) {
    const double t1 = log(alpha);
    const double t2 = 2.0 * log(1.0 - alpha / pow((double)(2 + (long long)depth), beta));
    const double t3 = -log(pow((double)(1 + (long long)depth), beta) - alpha);
    return __dadd_rn(__dadd_rn(t1, t2), t3);
}

// Executed by ONE full warp (all 32 lanes converge here).  The tree-structure scans are ballot-parallel over
// the node slots; the short serial part (parent walk, rule sampling, ratios) runs on lane 0, which returns the
// proposal; the other lanes' return value is unspecified.
//   u[0..3]  : uniforms (move type, node, feature, rule)
//   box      : shared-memory scratch, 2*d doubles, pre-filled with `bounds`
//   cm       : this tree's leaf -> column map;   colused / P: column allocator bitmap / capacity
// lowest free leaf column (-1 if none), found by a whole warp: P/32 <= 256 words
__device__ __forceinline__ int warp_find_free_col(const uint32_t* colused, int P) {
    const int lane = threadIdx.x & 31;
    const int nwords = P / 32;
    int fcol = -1;
    for (int base = 0; base < nwords && fcol < 0; base += 32) {
        const int w = base + lane;
        const uint32_t fr = (w < nwords) ? ~colused[w] : 0u;
        const unsigned has = __ballot_sync(0xffffffffu, fr != 0u);
        if (has) {
            const int src = __ffs(has) - 1;
            const uint32_t word = __shfl_sync(0xffffffffu, fr, src);
            fcol = (base + src) * 32 + __ffs(word) - 1;
        }
    }
    return fcol;
}

//   logtab[k] = log(k), k <= L+1;  priortab[dep] = log_prior_ratio_at_depth(dep): precomputed once per launch
//   defer_col: a grow's free column (p.a) is left at -1 for the caller to fill in later (warp_find_free_col)
__device__ __forceinline__ Prop propose_tree_warp(const TreeSmem& T, int L, double* box, const int32_t* ft, int d,
                                                  const uint16_t* cm, const uint32_t* colused, int P,
                                                  const bark_params& prm, const double* u, unsigned* status,
                                                  const double* logtab, const double* priortab, bool defer_col) {
    const int lane = threadIdx.x & 31;
    const int fcol = defer_col ? -1 : warp_find_free_col(colused, P);
    const int nch = (L + 31) >> 5;
    uint32_t mT[8], mS[8], mI[8];
    int nT = 0, nS = 0;
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
        mT[ch] = mS[ch] = mI[ch] = 0u;
        if (ch < nch) {
            const int slot = ch * 32 + lane;
            bool term = false, sing = false, inact = false;
            if (slot < L) {
                const bool leaf = T.is_leaf[slot] != 0, act = T.active[slot] != 0;
                const uint32_t l = T.left[slot], r = T.right[slot];
                const bool ll = (l < (uint32_t)L) && T.is_leaf[l] != 0;
                const bool rl = (r < (uint32_t)L) && T.is_leaf[r] != 0;
                term = act && leaf;
                sing = act && !leaf && ll && rl;
                inact = !act;
            }
            mT[ch] = __ballot_sync(0xffffffffu, term);
            mS[ch] = __ballot_sync(0xffffffffu, sing);
            mI[ch] = __ballot_sync(0xffffffffu, inact);
            nT += __popc(mT[ch]);
            nS += __popc(mS[ch]);
        }
    }
    Prop p;
    p.move = 0; p.valid = 0; p.node = 0; p.feat = 0; p.thr = 0.f; p.sl = 0; p.sr = 0; p.a = 0; p.b = 0;
    p.lqp = -INFINITY; p.depth = 0; p.pad = 0;
    if (lane != 0) return p;

    // move type ~ Categorical(weights) by inverse CDF (searchsorted(cumsum(w), r), tree_proposals.py:195)
    const double c0 = prm.proposal_weights[0], c1 = __dadd_rn(c0, prm.proposal_weights[1]);
    int move = 0;
    if (c0 < u[0]) move = 1;
    if (move == 1 && c1 < u[0]) move = 2;
    p.move = move;

    const int cnt = (move == MOVE_GROW) ? nT : nS;
    if (cnt == 0) return p;
    int k = min((int)(u[1] * (double)cnt), cnt - 1);
    int node = -1;
    for (int ch = 0; ch < nch; ++ch) {
        const uint32_t msk = (move == MOVE_GROW) ? mT[ch] : mS[ch];
        const int pc = __popc(msk);
        if (k < pc) {
            node = ch * 32 + (int)__fns(msk, 0, k + 1);
            break;
        }
        k -= pc;
    }
    p.node = node;

    if (move != MOVE_PRUNE) {
        // feasible box / category mask at `node`: walk up to the root (tree_traversal.py:49-86)
        int child = node;
        for (int it = 0; it < L && child != 0; ++it) {
            const int up = (int)T.parent[child];
            const int f = (int)T.feat[up];
            const float th = T.thr[up];
            const bool from_left = (uint32_t)child == T.left[up];
            if (ft[f] == FEAT_CAT) {
                const long long have = (long long)box[2 * f + 1];
                if (from_left) {
                    box[2 * f + 1] = (double)(((long long)th) & have);
                } else {
                    const long long full = next_pow2_ll(have) - 1;
                    const double comp = (double)full - (double)th;
                    box[2 * f + 1] = (double)(((long long)comp) & have);
                }
            } else if (from_left) {
                box[2 * f + 1] = fmin((double)th, box[2 * f + 1]);
            } else {
                const double bump = (ft[f] == FEAT_INT) ? 1.0 : 0.0;
                box[2 * f] = fmax((double)th + bump, box[2 * f]);
            }
            child = up;
        }
        // split rule (tree_proposals.py:78-97)
        const int f = min((int)(u[2] * (double)d), d - 1);
        const double lo = box[2 * f], hi = box[2 * f + 1];
        double thr;
        if (ft[f] == FEAT_CAT) {
            const long long avail = (long long)hi;
            const int nb = __popcll((unsigned long long)avail);
            if (nb < 2) {
                thr = 0.0;
            } else {
                const long long top = (1LL << nb) - 1;
                const long long pick = 1 + min((long long)(u[3] * (double)(top - 1)), top - 2);
                thr = (double)scatter_bits_ll(avail, pick);
            }
        } else if (ft[f] == FEAT_INT) {
            if (lo == hi) {
                thr = hi;
            } else {
                const long long li = (long long)lo, hi_i = (long long)hi;
                thr = (double)(li + min((long long)(u[3] * (double)(hi_i - li)), hi_i - li - 1));
            }
        } else {
            thr = __dadd_rn(lo, __dmul_rn(__dsub_rn(hi, lo), u[3]));
        }
        const float thr32 = (float)thr;
        p.feat = f;
        p.thr = thr32;
        if (thr32 == 0.f && ft[f] == FEAT_CAT) return p;
        if ((double)thr32 == hi && ft[f] == FEAT_INT) return p;
    }

    p.depth = T.depth[node];
    const uint32_t depth = min(T.depth[node], (uint32_t)L);  // index into the per-depth prior table
    if (move == MOVE_GROW) {
        // first two inactive slots in ascending order (tree_proposals.py:45-58)
        int s0 = -1, s1 = -1;
        for (int ch = 0; ch < nch && s1 < 0; ++ch) {
            uint32_t msk = mI[ch];
            while (msk && s1 < 0) {
                const int bit = __ffs(msk) - 1;
                msk &= msk - 1;
                if (s0 < 0) s0 = ch * 32 + bit; else s1 = ch * 32 + bit;
            }
        }
        if (s1 < 0) {
            atomicOr(status, BARK_ST_TREE_OVERFLOW);
            return p;
        }
        // free leaf column for the right child (lowest index first; found above by the whole warp)
        if (fcol < 0 && !defer_col) {
            atomicOr(status, BARK_ST_COL_OVERFLOW);
            return p;
        }
        int was_sing = 0;
        if (node != 0) {
            const int up = (int)T.parent[node];
            was_sing = (mS[up >> 5] >> (up & 31)) & 1u;
        }
        const int w1 = nS - was_sing + 1;  // singly-internal nodes after the grow (tree_proposals.py:104-107)
        const double log_q = __dsub_rn(logtab[nT], logtab[w1]);
        p.lqp = __dadd_rn(log_q, priortab[depth]);
        p.sl = s0; p.sr = s1;
        p.a = fcol; p.b = (int)cm[node];
    } else {
        p.sl = (int)T.left[node];
        p.sr = (int)T.right[node];
        const int pL = (int)cm[p.sl], pR = (int)cm[p.sr];
        if (move == MOVE_PRUNE) {
            const double log_q = __dsub_rn(logtab[nS], logtab[nT - 1]);  // :111-114
            p.lqp = __dadd_rn(log_q, -priortab[depth]);
            p.a = pL; p.b = pR;
        } else {
            p.lqp = 0.0;
            p.a = pR; p.b = pL;
        }
    }
    p.valid = 1;
    return p;
}

// ---------------------------------------------------------------- noise / scale random walk
constexpr double STEP_NOISE = 1.0;          // PROPOSAL_STEP_SIZE (second assignment), noise_scale_proposals.py:10-11
constexpr double STEP_SCALE = 0.00000001;

__device__ __forceinline__ double softplus_walk(double cur, double step, double z) {  // :61-67
    const double raw = log(exp(cur) - 1.0);
    return log(exp(raw + step * z) + 1.0);
}
__device__ __forceinline__ double log_walk(double cur, double step, double z) {  // :42-58
    return exp(log(cur + 1e-30) + step * z);
}
__device__ __forceinline__ double softplus_q_term(double cur, double nw, double step_var) {
    const double dr = log(exp(cur) - 1.0) - log(exp(nw) - 1.0);
    return dr * dr / step_var + log(1.0 - exp(-cur)) - log(1.0 - exp(-nw));
}
__device__ __forceinline__ double half_normal_logpdf(double x, double var) {  // :14-18
    return (x >= 0.0) ? (-0.5 * (x * x) / var - 0.5 * log(var)) : -INFINITY;
}
__device__ __forceinline__ double inverse_gamma_logpdf(double x, double shape, double rate) {  // :31-39
    const double sc = 1.0 / rate;
    return -(shape + 1.0) * log(x) - sc / x - lgamma(shape) + shape * log(sc);
}

struct HyperProp {
    double noise, scale, lqp;
    unsigned status;
};
__device__ __forceinline__ HyperProp propose_noise_scale(double noise, double scale, const bark_params& prm, double zn,
                                                         double zs) {
    HyperProp h;
    h.noise = noise; h.scale = scale; h.lqp = -INFINITY; h.status = 0;
    const bool sp = prm.use_softplus_transform != 0, ss = prm.sample_scale != 0;
    if (sp && !ss) {  // get_noise_proposal_softplus (:134-156)
        h.noise = softplus_walk(noise, STEP_NOISE, zn);
        const double log_q = -softplus_q_term(noise, h.noise, STEP_NOISE * STEP_NOISE);
        const double log_prior = inverse_gamma_logpdf(h.noise, prm.gamma_prior_shape, prm.gamma_prior_rate) -
                                 inverse_gamma_logpdf(noise, prm.gamma_prior_shape, prm.gamma_prior_rate);
        h.lqp = log_q + log_prior;
        return h;
    }
    if (!sp && !ss) {  // NotImplementedError branch (:78-81)
        h.status = BARK_ST_HYPER_MODE;
        return h;
    }
    double log_q;
    if (sp) {  // :100-131
        h.noise = softplus_walk(noise, STEP_NOISE, zn);
        h.scale = softplus_walk(scale, STEP_SCALE, zs);
        log_q = softplus_q_term(noise, h.noise, STEP_NOISE * STEP_NOISE) +
                softplus_q_term(scale, h.scale, STEP_SCALE * STEP_SCALE);
    } else {  // :83-97
        h.noise = log_walk(noise, STEP_NOISE, zn);
        h.scale = log_walk(scale, STEP_SCALE, zs);
        log_q = -log(noise) - log(scale) + log(h.noise) + log(h.scale);
    }
    const double log_prior = half_normal_logpdf(h.noise, 1.0) + half_normal_logpdf(h.scale, 5.0) -
                             half_normal_logpdf(noise, 1.0) - half_normal_logpdf(scale, 5.0);
    h.lqp = log_q + log_prior;
    return h;
}

}  // namespace bark
