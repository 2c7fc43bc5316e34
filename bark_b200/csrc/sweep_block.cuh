// Tree sweep in speculative BLOCKS: the m tree MH steps of one sweep of one chain (bark_sampler.py:233-264) in
// leaf space, KB consecutive trees at a time.
//
// Why blocks.  The proposals of trees t .. t+KB-1 edit DIFFERENT trees, so their structure (node, rule, moved-point
// mask u_j, columns a_j / b_j) does not depend on each other's MH decisions; only the linear algebra does, and only
// through low-rank terms.  With B0 the state at the start of the block:
//     v_j  = Z0^T u_j                      one pass over the leaf bitsets for all KB masks
//     Wv_j = B0^-1 v_j                     ONE pass over the lower triangle of B0^-1 as a P x P x KB product on the
//                                          FP64 tensor pipe (mma.sync m8n8k4 / DMMA), instead of KB matvec passes
//     Wd_j = B0^-1 (e_a - e_b)             two symmetric rows per proposal (gather)
// The proposals are then decided IN ORDER.  When proposal i is accepted (Z' = Z + u_i d_i^T, B'^-1 = B^-1 - W M^-1 W^T,
// W = [Wd_i Wv_i]) every still-pending proposal j > i is brought up to date with O(P) work:
//     v_j  += g_ij d_i,                g_ij = u_i^T u_j  (AND + POPC of the masks)
//     Wd_j -= W M^-1 (W^T d_j)
//     Wv_j += g_ij Wd_i - W M^-1 (W^T v_j)
// so every decision uses exactly the quantities the one-at-a-time algorithm would have (same 2 x 2 capacitance
// matrix, same log-MLL, same accept rule).  The accepted W's stay in shared memory and are applied to the lower
// triangle of B^-1 as ONE rank-2*n_acc DMMA update per block.  Passes over B^-1 per proposal: (1 + 1) / KB instead
// of 1 + acceptance; cluster barriers per proposal: 3 / KB instead of ~3.3.
//
// A thread-block CLUSTER of R CTAs (R = 1..16, 512 threads each) owns one chain.  The CTAs split the bitset scan,
// the DMMA units of the product and of the update, and exchange v / Wv through distributed shared memory; the
// sequential part is executed REDUNDANTLY and deterministically by every CTA on identical inputs (same code, same
// reduction trees), so the CTAs of a cluster cannot disagree on control flow and no decision has to be broadcast.
//
// A grow's new column is allocated lazily, at its decision: nothing computed before depends on it (the free column's
// row of B^-1 is e_f / c, v_f = 0).  Columns freed by a prune inside a block are handed out again only from the next
// block on (their row of B^-1 is reset to e_b / c by the block-end update).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "linalg.cuh"
#include "mcmc_state.cuh"
#include "proposal_device.cuh"

namespace bark {
namespace cg = cooperative_groups;

constexpr int SB_THREADS = 512;
constexpr int SB_WARPS = SB_THREADS / 32;
constexpr int SB_MAX_R = 16;  // CTAs per chain (16 is a non-portable cluster size: opt-in at launch)
constexpr int SB_KB = 8;      // proposals per block (the DMMA n dimension); smaller powers of two for very wide forests

#ifdef BARK_PHASE_TIMING
#define SB_MARK(i)                                           \
    do {                                                     \
        if (tid == 0) {                                      \
            const long long now__ = clock64();               \
            ph_acc[i] += (unsigned long long)(now__ - ph_t); \
            ph_t = now__;                                    \
        }                                                    \
    } while (0)
#else
#define SB_MARK(i) do { } while (0)
#endif

struct SbAccepted {  // one accepted proposal of the current block, kept until the block-end update
    double al, be, ga;  // M^-1 = [[al, be], [be, ga]]
    double eta, nu;
    int slot, a, b, move;
};
struct SbCtl {
    Prop prop[SB_KB];
    double eta[SB_KB], nu[SB_KB], uacc[SB_KB];
    double vWv[SB_KB], vw[SB_KB];  // v_j^T Wv_j and v_j^T w of the pending proposals (refreshed after every accept)
    int G[SB_KB][SB_KB];           // g_ij = u_i^T u_j
    SbAccepted acc[SB_KB];
    int n_acc;
    int walk_from;                 // next proposal the decision walk looks at
    int acc_slot;                  // slot accepted by the last walk, -1: block finished
    double cw_d, cw_v;             // w update coefficients of the last accept
    double q, ldt, mll;
    int p_hi;
};

struct SbLayout {
    size_t off_ctl, off_v, off_wd, off_wv, off_w, off_upos, off_uneg, off_leaf, off_u32, off_cm, off_box, off_ft,
        off_logtab, off_priortab, off_used, off_red, total;
    int ks;  // proposals per block = row stride of V / Wd / Wv
};
__host__ __device__ inline SbLayout sb_layout(int L, int d, int P, int wd, int ks) {
    SbLayout s;
    s.ks = ks;
    size_t o = 0;
    s.off_ctl = o;      o += align256(sizeof(SbCtl));
    s.off_v = o;        o += align256((size_t)P * ks * 8);
    s.off_wd = o;       o += align256((size_t)P * ks * 8);
    s.off_wv = o;       o += align256((size_t)P * ks * 8);
    s.off_w = o;        o += align256((size_t)P * 8);
    s.off_upos = o;     o += align256((size_t)ks * wd * 4);
    s.off_uneg = o;     o += align256((size_t)ks * wd * 4);
    s.off_leaf = o;     o += align256((size_t)ks * L * 2);      // is_leaf, active
    s.off_u32 = o;      o += align256((size_t)ks * L * 4 * 6);  // feat, left, right, parent, depth, thr
    s.off_cm = o;       o += align256((size_t)ks * L * 2);      // leaf -> column maps
    s.off_box = o;      o += align256((size_t)ks * d * 2 * 8);
    s.off_ft = o;       o += align256((size_t)d * 4);
    s.off_logtab = o;   o += align256((size_t)(L + 2) * 8);
    s.off_priortab = o; o += align256((size_t)(L + 1) * 8);
    s.off_used = o;     o += align256((size_t)(P / 32) * 4);
    s.off_red = o;      o += align256((size_t)2 * 2 * SB_WARPS * SB_KB * 8);  // two buffers x two values
    s.total = o;
    return s;
}
// Largest block size (8, 4, 2, 1) whose working set fits the shared-memory budget; 0 if not even KB = 1 does.
__host__ __device__ inline int sb_pick_ks(int L, int d, int P, int wd, size_t budget) {
    for (int ks = SB_KB; ks >= 1; ks >>= 1)
        if (sb_layout(L, d, P, wd, ks).total <= budget) return ks;
    return 0;
}

__device__ __forceinline__ double2 sb_ldcg2(const double* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }

// Sum of (x, y) over all threads whose slot (tid % ks) is the same; every thread receives the totals of ITS slot.
// Deterministic (fixed shuffle tree, fixed order over the warps).  `red`: 2 * SB_WARPS * SB_KB doubles, must not be in
// use by a reduction that other threads may still be reading (callers alternate two buffers).  One __syncthreads.
__device__ __forceinline__ void sb_slot_sum2(double& x, double& y, int ks, double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int o = 16; o >= ks; o >>= 1) {
        const double tx = __shfl_xor_sync(0xffffffffu, x, o);
        const double ty = __shfl_xor_sync(0xffffffffu, y, o);
        x += tx;
        y += ty;
    }
    if (lane < ks) {
        red[wid * SB_KB + lane] = x;
        red[SB_WARPS * SB_KB + wid * SB_KB + lane] = y;
    }
    __syncthreads();
    const int j = lane & (ks - 1);
    double sx = 0.0, sy = 0.0;
#pragma unroll
    for (int w = 0; w < SB_WARPS; ++w) {
        sx += red[w * SB_KB + j];
        sy += red[SB_WARPS * SB_KB + w * SB_KB + j];
    }
    x = sx;
    y = sy;
}

__global__ void __launch_bounds__(SB_THREADS, 1)
sweep_block_kernel(WsLayout lay, void* ws, bark_nodes_soa forest, bark_params prm, int64_t sweep_in_call,
                   int64_t n_sweeps_call, uint64_t seed, int64_t chain_offset, int64_t sweep_offset,
                   const double* __restrict__ tape, double* __restrict__ trace, int ks) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int R = (int)cluster.num_blocks(), cr = (int)cluster.block_rank();
    auto csync = [&]() {
        if (R > 1) cluster.sync(); else __syncthreads();
    };

    const int P = (int)lay.P, L = (int)lay.L, m = (int)lay.m, n = (int)lay.n, wd = (int)lay.wd, npad = (int)lay.npad;
    const int d = (int)lay.d;
    const SbLayout sl = sb_layout(L, d, P, wd, ks);
    SbCtl* ctl = (SbCtl*)(smem_raw + sl.off_ctl);
    double* V = (double*)(smem_raw + sl.off_v);
    double* Wd = (double*)(smem_raw + sl.off_wd);
    double* Wv = (double*)(smem_raw + sl.off_wv);
    double* w_s = (double*)(smem_raw + sl.off_w);
    uint32_t* upos = (uint32_t*)(smem_raw + sl.off_upos);
    uint32_t* uneg = (uint32_t*)(smem_raw + sl.off_uneg);
    uint8_t* leaf_s = smem_raw + sl.off_leaf;
    uint32_t* u32_s = (uint32_t*)(smem_raw + sl.off_u32);
    uint16_t* cm_all = (uint16_t*)(smem_raw + sl.off_cm);
    double* box_all = (double*)(smem_raw + sl.off_box);
    int32_t* ftc = (int32_t*)(smem_raw + sl.off_ft);
    double* logtab = (double*)(smem_raw + sl.off_logtab);
    double* priortab = (double*)(smem_raw + sl.off_priortab);
    uint32_t* used_s = (uint32_t*)(smem_raw + sl.off_used);
    double* red = (double*)(smem_raw + sl.off_red);
    double* red2 = red + 2 * SB_WARPS * SB_KB;

    const int64_t chain = blockIdx.x / R;
    ChainView cv = chain_view(lay, ws, chain);
    SharedView sv = shared_view(lay, ws);
    ChainScalars* sc = cv.sc;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int lq = lane >> 2, lk = lane & 3;  // DMMA fragment coordinates
    const int gw = cr * SB_WARPS + wid, ngw = R * SB_WARPS;  // this warp among the cluster's warps

    // A dead chain (status set by an earlier launch) is skipped by the whole cluster.  The status word is read BEFORE
    // a cluster barrier and only written after it, so every CTA of the cluster takes the same exit.
    const unsigned status0 = __ldcg(&sc->status);
    csync();
    if (status0 & (BARK_ST_COL_OVERFLOW | BARK_ST_TREE_OVERFLOW | BARK_ST_HYPER_MODE)) return;

    const double sig = sc->sig, c = sc->c, yy = sc->yy;
    const double inv_c = 1.0 / c;
    const double nlogsig = (double)n * log(sig);
    if (tid == 0) {
        ctl->q = sc->q; ctl->ldt = sc->ldt; ctl->mll = sc->mll; ctl->p_hi = sc->p_hi;
    }
    for (int e = tid; e < d; e += SB_THREADS) ftc[e] = sv.ft[e];
    for (int e = tid; e < L + 2; e += SB_THREADS) logtab[e] = log((double)e);
    for (int e = tid; e < L + 1; e += SB_THREADS) priortab[e] = log_prior_ratio_at_depth((uint32_t)e, prm.alpha, prm.beta);
    for (int e = tid; e < P / 32; e += SB_THREADS) used_s[e] = __ldcg(cv.colused + e);
    for (int e = tid; e < P; e += SB_THREADS) w_s[e] = __ldcg(cv.w + e);

    unsigned long long n_valid = 0, n_acc_tot = 0, n_acc_move[3] = {0, 0, 0}, n_valid_move[3] = {0, 0, 0};  // thread 0
    unsigned long long blk_eval = 0, blk_upd = 0, cols_scanned = 0;

    const uint32_t g_chain = (uint32_t)(chain_offset + chain), g_sweep = (uint32_t)(sweep_offset + sweep_in_call);
    const size_t tape_base =
        tape ? ((size_t)(chain * n_sweeps_call + sweep_in_call)) * (size_t)(m * TAPE_PER_TREE + TAPE_PER_HYPER) : 0;
    double* trace_base = trace ? trace + ((size_t)(chain * n_sweeps_call + sweep_in_call)) * (size_t)(m + 1) * 3 : nullptr;
#ifdef BARK_PHASE_TIMING
    unsigned long long ph_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long ph_t = clock64();
#endif
    __syncthreads();

    for (int t0 = 0; t0 < m; t0 += ks) {
        const int nb = min(ks, m - t0);  // proposals in this block
        SB_MARK(0);
        // ------------------------------------------------------------------ phase 0: stage the trees, propose
        for (int e = tid; e < nb * L; e += SB_THREADS) {
            const int j = e / L, s = e - j * L;
            const int64_t g = (chain * (int64_t)m + (t0 + j)) * L + s;
            uint8_t* lf = leaf_s + (size_t)j * 2 * L;
            uint32_t* uu = u32_s + (size_t)j * 6 * L;
            lf[s] = __ldcg(forest.is_leaf + g);
            lf[L + s] = __ldcg(forest.active + g);
            uu[s] = __ldcg(forest.feature + g);
            uu[L + s] = __ldcg(forest.left + g);
            uu[2 * L + s] = __ldcg(forest.right + g);
            uu[3 * L + s] = __ldcg(forest.parent + g);
            uu[4 * L + s] = __ldcg(forest.depth + g);
            uu[5 * L + s] = __float_as_uint(__ldcg(forest.threshold + g));
            cm_all[(size_t)j * L + s] = __ldcg(cv.colmap + (size_t)(t0 + j) * L + s);
        }
        for (int e = tid; e < nb * 2 * d; e += SB_THREADS) box_all[e] = sv.bounds[e % (2 * d)];
        __syncthreads();
        if (wid < ks) {
            const int j = wid;
            if (j < nb) {
                double un[6];
                if (tape) {
                    for (int k = 0; k < TAPE_PER_TREE; ++k) un[k] = tape[tape_base + (size_t)(t0 + j) * TAPE_PER_TREE + k];
                } else {
                    rng_uniforms(seed, g_chain, g_sweep, (uint32_t)(t0 + j), TAPE_PER_TREE, un);
                }
                TreeSmem T;
                T.is_leaf = leaf_s + (size_t)j * 2 * L; T.active = T.is_leaf + L;
                T.feat = u32_s + (size_t)j * 6 * L; T.left = T.feat + L; T.right = T.left + L; T.parent = T.right + L;
                T.depth = T.parent + L; T.thr = (float*)(T.depth + L);
                Prop pp = propose_tree_warp(T, L, box_all + (size_t)j * 2 * d, ftc, d, cm_all + (size_t)j * L, used_s, P, prm,
                                            un, &sc->status, logtab, priortab, true);
                if (lane == 0) {
                    if (pp.valid && pp.move == MOVE_GROW) pp.a = -1;  // allocated at the decision
                    ctl->prop[j] = pp;
                    ctl->uacc[j] = un[4];
                }
            } else if (lane == 0) {
                Prop pp;
                pp.move = 0; pp.valid = 0; pp.node = 0; pp.feat = 0; pp.thr = 0.f; pp.sl = 0; pp.sr = 0; pp.a = 0; pp.b = 0;
                pp.lqp = -INFINITY; pp.depth = 0; pp.pad = 0;
                ctl->prop[j] = pp;
                ctl->uacc[j] = 1.0;
            }
        }
        if (tid == 0) { ctl->n_acc = 0; ctl->walk_from = 0; }
        __syncthreads();
        const int p_hi = ctl->p_hi;
        const int E = min(P, (p_hi + nb + 15) & ~15);  // extent of every vector of this block (multiple of 16)
        const int nb8 = E >> 3;
        SB_MARK(1);

        // ------------------------------------------------------------------ Wd = Binv (e_a - e_b): symmetric rows
        // (issued before the masks so that the gathers overlap them; a grow's e_a / c term is added at its decision)
        for (int idx = tid; idx < E * ks; idx += SB_THREADS) {
            const int k = idx / ks, j = idx - k * ks;
            double val = 0.0;
            if (j < nb && ctl->prop[j].valid) {
                const int a = ctl->prop[j].a, b = ctl->prop[j].b;
                const double sb = __ldcg(cv.Binv + ((k <= b) ? ((size_t)b * P + k) : ((size_t)k * P + b)));
                const double sa = (a >= 0) ? __ldcg(cv.Binv + ((k <= a) ? ((size_t)a * P + k) : ((size_t)k * P + a))) : 0.0;
                val = sa - sb;
            }
            Wd[idx] = val;
        }

        // ------------------------------------------------------------------ phase 1: moved-point masks, eta, n_u
        {
            double eta_p[SB_KB];
            int cnt_p[SB_KB];
#pragma unroll
            for (int j = 0; j < SB_KB; ++j) { eta_p[j] = 0.0; cnt_p[j] = 0; }
            for (int w = wid; w < wd; w += SB_WARPS) {
                const int i = w * 32 + lane;
                const double yv = (i < n) ? sv.y[i] : 0.0;
                uint32_t wa[SB_KB], wb[SB_KB];
                double xv[SB_KB];
#pragma unroll
                for (int j = 0; j < SB_KB; ++j) {
                    wa[j] = wb[j] = 0u;
                    xv[j] = 0.0;
                    if (j < nb && ctl->prop[j].valid) {
                        const int mv = ctl->prop[j].move;
                        wb[j] = __ldcg(cv.bits + (size_t)ctl->prop[j].b * wd + w);
                        if (mv == MOVE_CHANGE) wa[j] = __ldcg(cv.bits + (size_t)ctl->prop[j].a * wd + w);
                        if (mv != MOVE_PRUNE && i < n) xv[j] = sv.Xt[(size_t)ctl->prop[j].feat * npad + i];
                    }
                }
#pragma unroll
                for (int j = 0; j < SB_KB; ++j) {
                    if (j >= ks) break;
                    bool pos = false, neg = false;
                    if (j < nb && ctl->prop[j].valid && i < n) {
                        const int mv = ctl->prop[j].move;
                        const bool in_b = (wb[j] >> lane) & 1u;
                        if (mv == MOVE_GROW) {
                            if (in_b) pos = !goes_left(xv[j], ctl->prop[j].thr, ftc[ctl->prop[j].feat]);
                        } else if (mv == MOVE_PRUNE) {
                            pos = in_b;
                        } else {  // change: b = left child's column, a = right child's column
                            const bool in_a = (wa[j] >> lane) & 1u;
                            if (in_a || in_b) {
                                const bool gl = goes_left(xv[j], ctl->prop[j].thr, ftc[ctl->prop[j].feat]);
                                pos = in_b && !gl;
                                neg = in_a && gl;
                            }
                        }
                    }
                    const unsigned bp = __ballot_sync(0xffffffffu, pos), bn = __ballot_sync(0xffffffffu, neg);
                    if (lane == 0) { upos[j * wd + w] = bp; uneg[j * wd + w] = bn; }
                    if (pos) eta_p[j] += yv;
                    if (neg) eta_p[j] -= yv;
                    cnt_p[j] += __popc(bp) + __popc(bn);
                }
            }
            // per-slot totals: eta (fixed tree over lanes, fixed order over warps), n_u (exact integers)
#pragma unroll
            for (int j = 0; j < SB_KB; ++j) {
                if (j >= ks) break;
                const double e = warp_sum(eta_p[j]);
                if (lane == 0) {
                    red[wid * SB_KB + j] = e;
                    red[SB_WARPS * SB_KB + wid * SB_KB + j] = (double)cnt_p[j];
                }
            }
            __syncthreads();
            if (tid < ks) {
                double e = 0.0, cn = 0.0;
                for (int w = 0; w < SB_WARPS; ++w) { e += red[w * SB_KB + tid]; cn += red[SB_WARPS * SB_KB + w * SB_KB + tid]; }
                ctl->eta[tid] = e;
                ctl->nu[tid] = cn;
            }
        }
        SB_MARK(2);

        // ------------------------------------------------------------------ phase 2: V = Z^T U (AND + POPC), G = U^T U
        {
            const int share = E / R;  // E is a multiple of 16 >= R
            const int r0 = cr * share, r1 = r0 + share;
            unsigned negmask = 0, valmask = 0;
            for (int j = 0; j < nb; ++j) {
                if (ctl->prop[j].valid) {
                    valmask |= 1u << j;
                    if (ctl->prop[j].move == MOVE_CHANGE) negmask |= 1u << j;
                }
            }
            for (int base = r0; base < r1; base += SB_THREADS / 4) {
                const int q = base + (tid >> 2), part = tid & 3;
                int cnt[SB_KB];
#pragma unroll
                for (int j = 0; j < SB_KB; ++j) cnt[j] = 0;
                if (q < r1 && q < p_hi) {
                    const uint32_t* bq = cv.bits + (size_t)q * wd;
                    // wd is a multiple of 4: 16-byte loads, four in flight per thread
                    for (int w0 = part * 4; w0 < wd; w0 += 64) {
                        uint4 x[4];
#pragma unroll
                        for (int g = 0; g < 4; ++g)
                            x[g] = (w0 + 16 * g < wd) ? __ldcg(reinterpret_cast<const uint4*>(bq + w0 + 16 * g)) : make_uint4(0, 0, 0, 0);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (w0 + 16 * g >= wd) break;
                            if ((x[g].x | x[g].y | x[g].z | x[g].w) == 0u) continue;
#pragma unroll
                            for (int j = 0; j < SB_KB; ++j) {
                                if (!((valmask >> j) & 1u)) continue;
                                const uint4 up = *reinterpret_cast<const uint4*>(upos + j * wd + w0 + 16 * g);
                                cnt[j] += __popc(x[g].x & up.x) + __popc(x[g].y & up.y) + __popc(x[g].z & up.z) + __popc(x[g].w & up.w);
                                if ((negmask >> j) & 1u) {
                                    const uint4 un = *reinterpret_cast<const uint4*>(uneg + j * wd + w0 + 16 * g);
                                    cnt[j] -= __popc(x[g].x & un.x) + __popc(x[g].y & un.y) + __popc(x[g].z & un.z) + __popc(x[g].w & un.w);
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < SB_KB; ++j) {
                    cnt[j] += __shfl_xor_sync(0xffffffffu, cnt[j], 1);
                    cnt[j] += __shfl_xor_sync(0xffffffffu, cnt[j], 2);
                }
                if (q < r1) {
                    // the four threads of a column write its ks values (two each at ks = 8) to every CTA of the cluster
#pragma unroll
                    for (int j = 0; j < SB_KB; ++j) {
                        if (j < ks && (j & 3) == part) {
                            const double val = (double)cnt[j];
                            for (int r = 0; r < R; ++r) cluster.map_shared_rank(V, r)[(size_t)q * ks + j] = val;
                        }
                    }
                }
            }
            // g_ij = u_i^T u_j, i < j: one warp per pair
            const int npair = ks * (ks - 1) / 2;
            for (int pr = wid; pr < npair; pr += SB_WARPS) {
                int i = 0, rem = pr;
                while (rem >= ks - 1 - i) { rem -= ks - 1 - i; ++i; }
                const int j = i + 1 + rem;
                int g = 0;
                if (((valmask >> i) & 1u) && ((valmask >> j) & 1u)) {
                    for (int w = lane; w < wd; w += 32) {
                        const uint32_t pi = upos[i * wd + w], ni = uneg[i * wd + w], pj = upos[j * wd + w], nj = uneg[j * wd + w];
                        g += __popc(pi & pj) + __popc(ni & nj) - __popc(pi & nj) - __popc(ni & pj);
                    }
                }
                g = warp_sum_int(g);
                if (lane == 0) { ctl->G[i][j] = g; ctl->G[j][i] = g; }
            }
        }
        SB_MARK(3);
        csync();  // (B1) every column of V present on every CTA
        SB_MARK(4);

        // ------------------------------------------------------------------ phase 3: Wv = Binv V on the FP64 tensor pipe
        // Unit I = the 8 rows [8I, 8I+8) of the result: row part (row block I of the lower triangle times V) plus
        // column part (column block I below the diagonal, transposed, times V) -- (nb8 + 1) 8x8 blocks whatever I is,
        // so every unit costs the same.  One warp per unit; C fragment = Y[8I + lq][2 lk, 2 lk + 1].
        for (int I = gw; I < nb8; I += ngw) {
            double c0 = 0.0, c1 = 0.0;
            const int row = 8 * I + lq;
            const bool vec_on = lq < ks;
            const double* rowp = cv.Binv + (size_t)row * P + 2 * lk;
            // row part: k index of step s in block kb is column 8 kb + 2 lk + s (any bijection of the 8 columns works)
#pragma unroll 4
            for (int kb = 0; kb <= I; ++kb) {
                double2 x = sb_ldcg2(rowp + 8 * kb);
                const int col = 8 * kb + 2 * lk;
                if (kb == I) {  // diagonal block: only columns <= row are kept current
                    if (col > row) x.x = 0.0;
                    if (col + 1 > row) x.y = 0.0;
                }
                const double b0 = vec_on ? V[(size_t)col * ks + lq] : 0.0;
                const double b1 = vec_on ? V[(size_t)(col + 1) * ks + lq] : 0.0;
                la::dmma_m8n8k4(c0, c1, x.x, b0);
                la::dmma_m8n8k4(c0, c1, x.y, b1);
            }
            // column part: A[lq][k] = Binv[8 ib + k (+4)][8I + lq], rows strictly below the diagonal
            const double* colp = cv.Binv + (size_t)lk * P + 8 * I + lq;
#pragma unroll 4
            for (int ib = I; ib < nb8; ++ib) {
                double x0 = __ldcg(colp + (size_t)(8 * ib) * P);
                double x1 = __ldcg(colp + (size_t)(8 * ib + 4) * P);
                if (ib == I) {
                    if (lk <= lq) x0 = 0.0;
                    if (lk + 4 <= lq) x1 = 0.0;
                }
                const double b0 = vec_on ? V[(size_t)(8 * ib + lk) * ks + lq] : 0.0;
                const double b1 = vec_on ? V[(size_t)(8 * ib + 4 + lk) * ks + lq] : 0.0;
                la::dmma_m8n8k4(c0, c1, x0, b0);
                la::dmma_m8n8k4(c0, c1, x1, b1);
            }
            if (2 * lk < ks) {
                if (ks >= 2) {
                    const double2 out = make_double2(c0, c1);
                    for (int r = 0; r < R; ++r)
                        *reinterpret_cast<double2*>(cluster.map_shared_rank(Wv, r) + (size_t)row * ks + 2 * lk) = out;
                } else {
                    for (int r = 0; r < R; ++r) cluster.map_shared_rank(Wv, r)[row] = c0;
                }
            }
        }
        SB_MARK(5);
        csync();  // (B2) Wv complete on every CTA
        SB_MARK(6);

        // ------------------------------------------------------------------ phase 4: decide the proposals in order
        // (identical on every CTA of the cluster).  Thread mapping of the O(P) vector work: slot j = tid % ks,
        // rows k = tid / ks + (512 / ks) * it.
        const int my_j = tid & (ks - 1);
        const int k_first = tid / ks, k_step = SB_THREADS / ks;
        auto refresh_scalars = [&](int from, double* buf) {  // vWv_j, vw_j for the pending slots j >= from
            double x = 0.0, y = 0.0;
            if (my_j >= from && my_j < nb) {
                for (int k = k_first; k < E; k += k_step) {
                    const double vv = V[(size_t)k * ks + my_j];
                    x = fma(vv, Wv[(size_t)k * ks + my_j], x);
                    y = fma(vv, w_s[k], y);
                }
            }
            sb_slot_sum2(x, y, ks, buf);
            if (tid < ks) { ctl->vWv[tid] = x; ctl->vw[tid] = y; }
        };
        refresh_scalars(0, red2);
        __syncthreads();
        while (true) {
            if (tid == 0) {
                // ---- the walk: scalar work per proposal (bark_sampler.py:257-264), until one is accepted
                int j = ctl->walk_from;
                int found = -1;
                for (; j < nb; ++j) {
                    Prop p = ctl->prop[j];
                    const double cur_mll = ctl->mll;
                    if (p.valid && p.move == MOVE_GROW) {
                        int f = -1;
                        for (int wq = 0; wq < P / 32 && f < 0; ++wq) {
                            const uint32_t fr = ~used_s[wq];
                            if (fr) f = wq * 32 + __ffs(fr) - 1;
                        }
                        if (f < 0 || f >= E) {
                            // (f >= E cannot happen: E >= p_hi + nb and at most nb columns are taken per block)
                            atomicOr(&sc->status, BARK_ST_COL_OVERFLOW);
                            p.valid = 0;
                            p.lqp = -INFINITY;
                            ctl->prop[j].valid = 0;
                        } else {
                            p.a = f;
                            ctl->prop[j].a = f;
                            Wd[(size_t)f * ks + j] += inv_c;  // Binv e_f = e_f / c for a free column
                        }
                    }
                    if (!p.valid) {
                        if (trace_base && cr == 0) {
                            trace_base[(t0 + j) * 3 + 0] = -INFINITY;
                            trace_base[(t0 + j) * 3 + 1] = cur_mll;
                            trace_base[(t0 + j) * 3 + 2] = 0.0;
                        }
                        continue;
                    }
                    const int a = p.a, b = p.b;
                    const double eta = ctl->eta[j], n_u = ctl->nu[j];
                    const double dWd = Wd[(size_t)a * ks + j] - Wd[(size_t)b * ks + j];
                    const double dWv = Wv[(size_t)a * ks + j] - Wv[(size_t)b * ks + j];
                    const double dw = w_s[a] - w_s[b];
                    const double vWv = ctl->vWv[j], vw = ctl->vw[j];
                    const double M00 = dWd, M01 = 1.0 + dWv, M11 = -n_u + vWv;
                    const double det = M00 * M11 - M01 * M01;  // < 0 for an SPD B'
                    const double new_ldt = ctl->ldt + log(-det);
                    const double bq = ctl->q + 2.0 * eta * dw + eta * eta * dWd;  // b'^T Binv b'
                    const double Ur0 = dw + eta * dWd;                            // d^T r,  r = Binv b'
                    const double Ur1 = vw + eta * dWv;                            // v^T r
                    const double new_q = bq - (M11 * Ur0 * Ur0 - 2.0 * M01 * Ur0 * Ur1 + M00 * Ur1 * Ur1) / det;
                    const double new_mll = 0.5 * (-(yy - new_q) / sig - nlogsig - new_ldt);
                    const double log_alpha = p.lqp + (new_mll - cur_mll);
                    const bool accept = log(ctl->uacc[j]) <= fmin(log_alpha, 0.0);
                    if (trace_base && cr == 0) {
                        trace_base[(t0 + j) * 3 + 0] = p.lqp;
                        trace_base[(t0 + j) * 3 + 1] = new_mll;
                        trace_base[(t0 + j) * 3 + 2] = accept ? 1.0 : 0.0;
                    }
                    ++n_valid;
                    ++n_valid_move[p.move];
                    if (accept) {
                        SbAccepted& A = ctl->acc[ctl->n_acc];
                        A.al = M11 / det; A.be = -M01 / det; A.ga = M00 / det;
                        A.eta = eta; A.nu = n_u; A.slot = j; A.a = a; A.b = b; A.move = p.move;
                        ctl->cw_d = A.al * Ur0 + A.be * Ur1;
                        ctl->cw_v = A.be * Ur0 + A.ga * Ur1;
                        ctl->n_acc += 1;
                        ctl->q = new_q; ctl->ldt = new_ldt; ctl->mll = new_mll;
                        if (p.move == MOVE_GROW) {
                            used_s[a >> 5] |= 1u << (a & 31);
                            if (a + 1 > ctl->p_hi) ctl->p_hi = a + 1;
                        }
                        ++n_acc_tot;
                        ++n_acc_move[p.move];
                        found = j;
                        ++j;
                        break;
                    }
                }
                ctl->walk_from = j;
                ctl->acc_slot = found;
            }
            __syncthreads();
            const int i = ctl->acc_slot;
            if (i < 0) break;
            // ---- accepted slot i: bring the pending proposals j > i, and w, up to date
            const SbAccepted A = ctl->acc[ctl->n_acc - 1];
            const int ai = A.a, bi = A.b;
            const bool pend = my_j > i && my_j < nb && ctl->prop[my_j].valid;
            const double g = pend ? (double)ctl->G[i][my_j] : 0.0;
            if (tid < ks && pend && g != 0.0) {  // v_j += g d_i
                V[(size_t)ai * ks + tid] += g;
                V[(size_t)bi * ks + tid] -= g;
            }
            __syncthreads();
            double t0s = 0.0, t1s = 0.0;
            if (pend) {
                for (int k = k_first; k < E; k += k_step) {
                    const double vv = V[(size_t)k * ks + my_j];
                    t0s = fma(Wd[(size_t)k * ks + i], vv, t0s);
                    t1s = fma(Wv[(size_t)k * ks + i], vv, t1s);
                }
            }
            sb_slot_sum2(t0s, t1s, ks, red);
            {
                double ddj = 0.0, dvj = 0.0;
                if (pend) {
                    const int aj = ctl->prop[my_j].a, bj = ctl->prop[my_j].b;  // aj < 0: a grow's column, not yet allocated
                    ddj = ((aj >= 0) ? Wd[(size_t)aj * ks + i] : 0.0) - Wd[(size_t)bj * ks + i];
                    dvj = ((aj >= 0) ? Wv[(size_t)aj * ks + i] : 0.0) - Wv[(size_t)bj * ks + i];
                }
                const double cd_d = A.al * ddj + A.be * dvj, cd_v = A.be * ddj + A.ga * dvj;    // M^-1 W^T d_j
                const double cv_d = A.al * t0s + A.be * t1s, cv_v = A.be * t0s + A.ga * t1s;    // M^-1 W^T v_j
                const double cw_d = ctl->cw_d, cw_v = ctl->cw_v;
                for (int k = k_first; k < E; k += k_step) {
                    const double wd_i = Wd[(size_t)k * ks + i], wv_i = Wv[(size_t)k * ks + i];
                    if (pend) {
                        Wd[(size_t)k * ks + my_j] -= wd_i * cd_d + wv_i * cd_v;
                        Wv[(size_t)k * ks + my_j] += g * wd_i - wd_i * cv_d - wv_i * cv_v;
                    }
                    if (my_j == 0) w_s[k] = w_s[k] + A.eta * wd_i - wd_i * cw_d - wv_i * cw_v;  // w' = Binv' b'
                }
            }
            __syncthreads();
            if (tid == 0 && A.move == MOVE_PRUNE) w_s[bi] = 0.0;  // column b is an empty leaf from now on
            __syncthreads();
            refresh_scalars(i + 1, red2);
            __syncthreads();
        }
        SB_MARK(7);

        // ------------------------------------------------------------------ phase 5: block-end updates of the global state
        const int na = ctl->n_acc;
        if (tid == 0) {
            blk_eval += (unsigned long long)E * E;
            cols_scanned += (unsigned long long)E;
            if (na) blk_upd += (unsigned long long)E * E;
        }
        if (na > 0) {
            // Binv -= sum_s W_s M_s^-1 W_s^T on the lower triangle: C (8x8) += A (8 x 4) B (4 x 8), two accepted
            // proposals per DMMA step.  Row blocks are paired (I, nb8-1-I) so that every unit costs nb8 + 1 blocks.
            const int nsteps = (na + 1) >> 1;
            const int s_of = lk >> 1, comp = lk & 1;
            int n_prune = 0;
            for (int s = 0; s < na; ++s) n_prune += (ctl->acc[s].move == MOVE_PRUNE);
            for (int pu = gw; pu < (nb8 + 1) / 2; pu += ngw) {
                for (int half = 0; half < 2; ++half) {
                    const int I = half ? nb8 - 1 - pu : pu;
                    if (half && I == pu) break;
                    const int row = 8 * I + lq;
                    double afr[SB_KB / 2];
#pragma unroll
                    for (int st = 0; st < SB_KB / 2; ++st) {
                        const int s = 2 * st + s_of;
                        afr[st] = 0.0;
                        if (st < nsteps && s < na) {
                            const int slot = ctl->acc[s].slot;
                            afr[st] = -(comp ? Wv[(size_t)row * ks + slot] : Wd[(size_t)row * ks + slot]);
                        }
                    }
                    double* rowp = cv.Binv + (size_t)row * P + 2 * lk;
#pragma unroll 2
                    for (int J = 0; J <= I; ++J) {
                        double2 cc = sb_ldcg2(rowp + 8 * J);
                        const int jc = 8 * J + lq;  // B fragment column
#pragma unroll
                        for (int st = 0; st < SB_KB / 2; ++st) {
                            if (st >= nsteps) break;
                            const int s = 2 * st + s_of;
                            double bfr = 0.0;
                            if (s < na) {
                                const SbAccepted& A = ctl->acc[s];
                                const double wdj = Wd[(size_t)jc * ks + A.slot], wvj = Wv[(size_t)jc * ks + A.slot];
                                bfr = comp ? (A.be * wdj + A.ga * wvj) : (A.al * wdj + A.be * wvj);
                            }
                            la::dmma_m8n8k4(cc.x, cc.y, afr[st], bfr);
                        }
                        if (n_prune) {  // a pruned column becomes an empty leaf: its row / column is exactly e_b / c
                            const int col = 8 * J + 2 * lk;
                            for (int s = 0; s < na; ++s) {
                                if (ctl->acc[s].move != MOVE_PRUNE) continue;
                                const int pb = ctl->acc[s].b;
                                if (row == pb || col == pb) cc.x = (row == col) ? inv_c : 0.0;
                                if (row == pb || col + 1 == pb) cc.y = (row == col + 1) ? inv_c : 0.0;
                            }
                        }
                        __stcg(reinterpret_cast<double2*>(rowp + 8 * J), cc);
                    }
                }
            }
            // A' = A + v d^T + d v^T + n_u d d^T for every accepted proposal (exact integers; atomics make the order
            // irrelevant).  v is the slot's V column as it stood at the acceptance.
            {
                const int share = E / R, r0 = cr * share, r1 = r0 + share;
                for (int s = 0; s < na; ++s) {
                    const int slot = ctl->acc[s].slot, a = ctl->acc[s].a, b = ctl->acc[s].b;
                    for (int k = r0 + tid; k < r1; k += SB_THREADS) {
                        const int vk = (int)V[(size_t)k * ks + slot];
                        if (vk != 0) {
                            atomicAdd(cv.A + (size_t)k * P + a, vk);
                            atomicAdd(cv.A + (size_t)k * P + b, -vk);
                            atomicAdd(cv.A + (size_t)a * P + k, vk);
                            atomicAdd(cv.A + (size_t)b * P + k, -vk);
                        }
                    }
                }
            }
            if (cr == 0) {
                for (int s = 0; s < na; ++s) {
                    const SbAccepted A = ctl->acc[s];
                    const int a = A.a, b = A.b, slot = A.slot;
                    // leaf bitsets (the columns of different trees are disjoint, so the order over s is free)
                    for (int w = tid; w < wd; w += SB_THREADS) {
                        const uint32_t ba = __ldcg(cv.bits + (size_t)a * wd + w), bb = __ldcg(cv.bits + (size_t)b * wd + w);
                        const uint32_t up = upos[slot * wd + w], un = uneg[slot * wd + w];
                        __stcg(cv.bits + (size_t)a * wd + w, (ba | up) & ~un);
                        __stcg(cv.bits + (size_t)b * wd + w, (bb & ~up) | un);
                    }
                    if (tid == 0) {
                        const Prop p = ctl->prop[slot];
                        const int nuu = (int)A.nu;  // corner term n_u d d^T
                        atomicAdd(cv.A + (size_t)a * P + a, nuu);
                        atomicAdd(cv.A + (size_t)b * P + b, nuu);
                        atomicAdd(cv.A + (size_t)a * P + b, -nuu);
                        atomicAdd(cv.A + (size_t)b * P + a, -nuu);
                        cv.b[a] += A.eta;
                        cv.b[b] -= A.eta;
                        // forest edit (tree_proposals.py:146-183) + column bookkeeping in global memory
                        const int64_t g0 = (chain * (int64_t)m + (t0 + slot)) * L;
                        uint16_t* cm = cv.colmap + (size_t)(t0 + slot) * L;
                        if (p.move == MOVE_GROW) {
                            const uint32_t dep = p.depth;
                            for (int s2 = 0; s2 < 2; ++s2) {
                                const int64_t g = g0 + (s2 ? p.sr : p.sl);
                                forest.is_leaf[g] = 1; forest.feature[g] = 0; forest.threshold[g] = 0.f; forest.left[g] = 0;
                                forest.right[g] = 0; forest.parent[g] = (uint32_t)p.node; forest.depth[g] = dep + 1;
                                forest.active[g] = 1;
                            }
                            const int64_t g = g0 + p.node;
                            forest.is_leaf[g] = 0; forest.feature[g] = (uint32_t)p.feat; forest.threshold[g] = p.thr;
                            forest.left[g] = (uint32_t)p.sl; forest.right[g] = (uint32_t)p.sr; forest.active[g] = 1;
                            cm[p.sl] = (uint16_t)b;   // left child keeps the old leaf's column
                            cm[p.sr] = (uint16_t)a;   // right child takes the new column
                            cm[p.node] = NO_COL;
                        } else if (p.move == MOVE_PRUNE) {
                            forest.active[g0 + p.sl] = 0;
                            forest.active[g0 + p.sr] = 0;
                            forest.is_leaf[g0 + p.node] = 1;
                            cm[p.node] = (uint16_t)a;  // merged leaf keeps the left child's column
                            cm[p.sl] = NO_COL;
                            cm[p.sr] = NO_COL;
                            cv.b[b] = 0.0;
                        } else {
                            forest.feature[g0 + p.node] = (uint32_t)p.feat;
                            forest.threshold[g0 + p.node] = p.thr;
                        }
                    }
                }
            }
            __syncthreads();
            // columns freed by the prunes of this block can be handed out again from the next block on
            if (tid == 0) {
                for (int s = 0; s < na; ++s)
                    if (ctl->acc[s].move == MOVE_PRUNE) used_s[ctl->acc[s].b >> 5] &= ~(1u << (ctl->acc[s].b & 31));
            }
        }
        SB_MARK(8);
        // (B3) global-memory edits visible to the whole cluster; the exchanged vectors are free for the next block
        csync();
        SB_MARK(9);
    }
    __syncthreads();
#ifdef BARK_PHASE_TIMING
    if (tid == 0 && cr == 0)
        for (int i = 0; i < 12; ++i) sc->phase_cycles[i] += ph_acc[i];
#endif
    if (cr == 0) {
        for (int e = tid; e < P; e += SB_THREADS) cv.w[e] = w_s[e];
        for (int e = tid; e < P / 32; e += SB_THREADS) cv.colused[e] = used_s[e];
        if (tid == 0) {
            sc->q = ctl->q; sc->ldt = ctl->ldt; sc->mll = ctl->mll; sc->p_hi = ctl->p_hi;
            sc->counters[0] += (unsigned long long)m;
            sc->counters[1] += n_valid;
            sc->counters[2] += n_acc_tot;
            sc->counters[5] += n_acc_move[0];
            sc->counters[6] += n_acc_move[1];
            sc->counters[7] += n_acc_move[2];
            sc->counters[8] += n_valid_move[0];
            sc->counters[9] += n_valid_move[1];
            sc->counters[10] += n_valid_move[2];
            sc->counters[11] += blk_eval;      // sum over blocks of extent^2 (one product pass per block)
            sc->counters[12] += blk_upd;       // sum over blocks with an accepted proposal of extent^2 (one update pass)
            sc->counters[13] += cols_scanned;  // leaf-bitset columns scanned for V = Z^T U, per block
        }
    }
}

}  // namespace bark
