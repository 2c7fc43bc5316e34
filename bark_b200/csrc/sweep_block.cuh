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
// per-warp probe of one block of chain 0 (debug build): cycles from the start of the phase to the end of the warp's own work
#define SB_PROBE_BEGIN() long long pr_t0__ = clock64()
#define SB_PROBE_END(ph)                                                                                          \
    do {                                                                                                          \
        if (lane == 0 && chain == 0 && t0 == 80 && sweep_in_call == 0 && sweep_offset >= 100)                     \
            printf("probe ph=%d R=%d cr=%d wid=%d own=%lld\n", ph, R, cr, wid, (long long)(clock64() - pr_t0__)); \
    } while (0)
#else
#define SB_MARK(i) do { } while (0)
#define SB_PROBE_BEGIN() do { } while (0)
#define SB_PROBE_END(ph) do { } while (0)
#endif

struct SbAccepted {  // one accepted proposal of the current block, kept until the block-end update
    double al, be, ga;  // M^-1 = [[al, be], [be, ga]]
    double eta, nu;
    int slot, a, b, move;
};
struct SbCtl {
    Prop prop[SB_KB];
    double eta[SB_KB], nu[SB_KB], luacc[SB_KB];  // luacc = log(u_accept)
    double vWv[SB_KB], vw[SB_KB];  // v_j^T Wv_j and v_j^T w of the pending proposals (refreshed after every accept)
    int G[SB_KB][SB_KB];           // g_ij = u_i^T u_j
    SbAccepted acc[SB_KB];
    int n_acc;
    int walk_from;                 // next proposal the decision walk looks at
    int acc_slot;                  // slot accepted by the last walk; -1: block finished; -2: walk again
    double cw_d, cw_v;             // w update coefficients of the last accept
    double res, ldt, mll;  // res = y^T y - b^T Binv b (tracked directly: no cancellation against y^T y per proposal)
    int p_hi;
};

// Shared-memory plan.  V / Wd / Wv are VECTOR-major, [ks][P + 8] doubles: consecutive rows of one vector are
// consecutive, which makes the O(P) vector passes and the DMMA fragment loads (one 16-byte load = the B operands of
// two k-steps) bank-conflict free; the stride P + 8 (= 8 mod 16 doubles) keeps the eight vectors of a fragment load
// on different banks.
struct SbLayout {
    unsigned off_ctl, off_v, off_wd, off_wv, off_w, off_upos, off_uneg, off_leaf, off_u32, off_cm, off_box, off_ft,
        off_logtab, off_priortab, off_used, off_red, total;
    int ks, ps;  // proposals per block; vector stride in doubles
};
__host__ __device__ inline SbLayout sb_layout(int L, int d, int P, int wd, int ks) {
    SbLayout s;
    s.ks = ks;
    s.ps = P + 8;
    size_t o = 0;
    s.off_ctl = (unsigned)o;      o += align256(sizeof(SbCtl));
    s.off_v = (unsigned)o;        o += align256((size_t)s.ps * ks * 8);
    s.off_wd = (unsigned)o;       o += align256((size_t)s.ps * ks * 8);
    s.off_wv = (unsigned)o;       o += align256((size_t)s.ps * ks * 8);
    s.off_w = (unsigned)o;        o += align256((size_t)P * 8);
    s.off_upos = (unsigned)o;     o += align256((size_t)ks * wd * 4);
    s.off_uneg = (unsigned)o;     o += align256((size_t)ks * wd * 4);
    s.off_leaf = (unsigned)o;     o += align256((size_t)ks * L * 2);      // is_leaf, active
    s.off_u32 = (unsigned)o;      o += align256((size_t)ks * L * 4 * 6);  // feat, left, right, parent, depth, thr
    s.off_cm = (unsigned)o;       o += align256((size_t)ks * L * 2);      // leaf -> column maps
    s.off_box = (unsigned)o;      o += align256((size_t)ks * d * 2 * 8);
    s.off_ft = (unsigned)o;       o += align256((size_t)d * 4);
    s.off_logtab = (unsigned)o;   o += align256((size_t)(L + 2) * 8);
    s.off_priortab = (unsigned)o; o += align256((size_t)(L + 1) * 8);
    s.off_used = (unsigned)o;     o += align256((size_t)(P / 32) * 4);
    s.off_red = (unsigned)o;      o += align256((size_t)2 * 2 * SB_WARPS * 8);  // two buffers x two values per warp
    s.total = (o > 0xFFFFFFFFull) ? 0xFFFFFFFFu : (unsigned)o;
    return s;
}
// Largest block size (8, 4, 2, 1) whose working set fits the shared-memory budget; 0 if not even KB = 1 does.
__host__ __device__ inline int sb_pick_ks(int L, int d, int P, int wd, size_t budget) {
    for (int ks = SB_KB; ks >= 1; ks >>= 1)
        if (sb_layout(L, d, P, wd, ks).total <= budget) return ks;
    return 0;
}

__device__ __forceinline__ double2 sb_ldcg2(const double* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }

// Sum of (x, y) over the threads of one proposal slot (SB_THREADS / KS consecutive threads = SB_WARPS / KS whole
// warps); every thread receives the totals of ITS slot.  Deterministic (fixed shuffle tree, fixed order over the
// slot's warps).  `red`: 2 * SB_WARPS doubles, not in use by a reduction that other threads may still be reading
// (callers alternate two buffers).  One __syncthreads.
template <int KS>
__device__ __forceinline__ void sb_slot_sum2(double& x, double& y, double* red) {
    constexpr int WPS = SB_WARPS / KS;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    warp_sum2(x, y);
    if (lane == 0) {
        red[wid] = x;
        red[SB_WARPS + wid] = y;
    }
    __syncthreads();
    const int w0 = (wid / WPS) * WPS;
    double sx = 0.0, sy = 0.0;
#pragma unroll
    for (int w = 0; w < WPS; ++w) {
        sx += red[w0 + w];
        sy += red[SB_WARPS + w0 + w];
    }
    x = sx;
    y = sy;
}

template <int KS>
__global__ void __launch_bounds__(SB_THREADS, 1)
sweep_block_kernel(WsLayout lay, SbLayout sl, void* ws, bark_nodes_soa forest, bark_params prm, int64_t sweep_in_call,
                   int64_t n_sweeps_call, uint64_t seed, int64_t chain_offset, int64_t sweep_offset,
                   const double* __restrict__ tape, double* __restrict__ trace) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int R = (int)cluster.num_blocks(), cr = (int)cluster.block_rank();
    auto csync = [&]() {
        if (R > 1) cluster.sync(); else __syncthreads();
    };
    constexpr int TPS = SB_THREADS / KS;  // threads per proposal slot in the O(P) vector passes

    const int P = (int)lay.P, L = (int)lay.L, m = (int)lay.m, n = (int)lay.n, wd = (int)lay.wd, npad = (int)lay.npad;
    const int d = (int)lay.d;
    const int PS = sl.ps;
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
    double* red2 = red + 2 * SB_WARPS;

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

    const double sig = sc->sig, c = sc->c;
    const double inv_c = 1.0 / c;
    const double nlogsig = (double)n * log(sig);
    if (tid == 0) {
        ctl->res = sc->res; ctl->ldt = sc->ldt; ctl->mll = sc->mll; ctl->p_hi = sc->p_hi;
    }
    for (int e = tid; e < d; e += SB_THREADS) ftc[e] = sv.ft[e];
    for (int e = tid; e < L + 2; e += SB_THREADS) logtab[e] = log((double)e);
    for (int e = tid; e < L + 1; e += SB_THREADS) priortab[e] = log_prior_ratio_at_depth((uint32_t)e, prm.alpha, prm.beta);
    for (int e = tid; e < P / 32; e += SB_THREADS) used_s[e] = __ldcg(cv.colused + e);
    for (int e = tid; e < P; e += SB_THREADS) w_s[e] = __ldcg(cv.w + e);

    unsigned long long n_valid = 0, n_acc_tot = 0, n_acc_move[3] = {0, 0, 0}, n_valid_move[3] = {0, 0, 0};  // thread 0
    unsigned long long blk_eval = 0, blk_upd = 0, blk_upd_rank = 0, cols_scanned = 0;

    const uint32_t g_chain = (uint32_t)(chain_offset + chain), g_sweep = (uint32_t)(sweep_offset + sweep_in_call);
    const size_t tape_base =
        tape ? ((size_t)(chain * n_sweeps_call + sweep_in_call)) * (size_t)(m * TAPE_PER_TREE + TAPE_PER_HYPER) : 0;
    double* trace_base = trace ? trace + ((size_t)(chain * n_sweeps_call + sweep_in_call)) * (size_t)(m + 1) * 3 : nullptr;
#ifdef BARK_PHASE_TIMING
    unsigned long long ph_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long ph_t = clock64();
#endif
    __syncthreads();

    for (int t0 = 0; t0 < m; t0 += KS) {
        const int nb = min(KS, m - t0);  // proposals in this block
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
        if (wid < KS) {
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
                    ctl->luacc[j] = log(un[4]);
                }
            } else if (lane == 0) {
                Prop pp;
                pp.move = 0; pp.valid = 0; pp.node = 0; pp.feat = 0; pp.thr = 0.f; pp.sl = 0; pp.sr = 0; pp.a = 0; pp.b = 0;
                pp.lqp = -INFINITY; pp.depth = 0; pp.pad = 0;
                ctl->prop[j] = pp;
                ctl->luacc[j] = 0.0;
            }
        }
        if (tid == 0) { ctl->n_acc = 0; ctl->walk_from = 0; }
        __syncthreads();
        const int p_hi = ctl->p_hi;
        const int E = min(P, (p_hi + nb + 15) & ~15);  // extent of every vector of this block (multiple of 16)
        const int nb8 = E >> 3;
        SB_MARK(1);
        SB_PROBE_BEGIN();

        // ------------------------------------------------------------------ Wd = Binv (e_a - e_b): symmetric rows
        // (issued before the masks so that the gathers overlap them; a grow's e_a / c term is added at its decision)
        {
            int ga[KS], gb[KS];
#pragma unroll
            for (int j = 0; j < KS; ++j) {
                const bool ok = j < nb && ctl->prop[j].valid;
                ga[j] = ok ? ctl->prop[j].a : -1;
                gb[j] = ok ? ctl->prop[j].b : -1;
            }
            for (int k = tid; k < E; k += SB_THREADS) {
                double sa[KS], sb[KS];
#pragma unroll
                for (int j = 0; j < KS; ++j) {  // all gathers in flight before the first shared-memory store
                    const int a = ga[j], b = gb[j];
                    sa[j] = sb[j] = 0.0;
                    if (b >= 0) sb[j] = __ldcg(cv.Binv + ((k <= b) ? ((size_t)b * P + k) : ((size_t)k * P + b)));
                    if (a >= 0) sa[j] = __ldcg(cv.Binv + ((k <= a) ? ((size_t)a * P + k) : ((size_t)k * P + a)));
                }
#pragma unroll
                for (int j = 0; j < KS; ++j) Wd[(size_t)j * PS + k] = sa[j] - sb[j];
            }
        }

        SB_PROBE_END(0);  // the Wd gathers
        // ------------------------------------------------------------------ phase 1: moved-point masks, eta, n_u
        // SB_WARPS / KS warps per proposal; a warp takes every (SB_WARPS / KS)-th 32-point word of its proposal.  The
        // leaf bitsets come in coalesced (one word per lane per 1024 points) and are handed out by shuffles.
        {
            constexpr int nparts = SB_WARPS / KS;
            const int j = wid & (KS - 1), part = wid / KS;
            const bool live = j < nb && ctl->prop[j].valid;
            const int mv = ctl->prop[j].move, pa = ctl->prop[j].a, pb = ctl->prop[j].b;
            const float pthr = ctl->prop[j].thr;
            const int pf = live ? ctl->prop[j].feat : 0;
            const int ftype = ftc[pf];
            const double* xf = sv.Xt + (size_t)pf * npad;
            double eta_p = 0.0;
            int cnt_p = 0;
            const bool numeric = ftype != FEAT_CAT;
            const double thr_d = (double)pthr;
            for (int ch = 0; ch * 32 < wd; ++ch) {
                const int wl = ch * 32 + lane;
                uint32_t wb_l = 0u, wa_l = 0u;
                if (live && wl < wd) {
                    wb_l = __ldcg(cv.bits + (size_t)pb * wd + wl);
                    if (mv == MOVE_CHANGE) wa_l = __ldcg(cv.bits + (size_t)pa * wd + wl);
                }
                if (!live || mv == MOVE_PRUNE) {
                    // prune: u is the whole right leaf, so the mask IS its bitset (and eta its entry of b = Z^T y, below);
                    // a dead slot gets empty masks.  One coalesced pass, no point data needed.
                    if (wl < wd && (wl % nparts) == part) {
                        upos[j * wd + wl] = wb_l;
                        uneg[j * wd + wl] = 0u;
                        cnt_p += __popc(wb_l);
                    }
                    continue;
                }
                for (int ww0 = part; ww0 < 32; ww0 += 8 * nparts) {
                    double xv[8], yv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int ww = ww0 + u * nparts, i = (ch * 32 + ww) * 32 + lane;
                        xv[u] = yv[u] = 0.0;
                        if (ww < 32 && i < n) {
                            yv[u] = sv.y[i];
                            xv[u] = xf[i];
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int ww = ww0 + u * nparts, w = ch * 32 + ww, i = w * 32 + lane;
                        if (ww >= 32 || w >= wd) break;  // warp-uniform
                        const uint32_t wbw = __shfl_sync(0xffffffffu, wb_l, ww);
                        const bool gl = numeric ? (xv[u] <= thr_d) : goes_left(xv[u], pthr, ftype);
                        const bool in_b = ((wbw >> lane) & 1u) && i < n;
                        const bool pos = in_b && !gl;  // grow: the leaf's points that go right; change: left child's that go right
                        bool neg = false;
                        if (mv == MOVE_CHANGE) {       // ... and the right child's (column a) that now go left
                            const uint32_t waw = __shfl_sync(0xffffffffu, wa_l, ww);
                            neg = ((waw >> lane) & 1u) && gl && i < n;
                        }
                        const unsigned bp = __ballot_sync(0xffffffffu, pos), bn = __ballot_sync(0xffffffffu, neg);
                        if (lane == 0) { upos[j * wd + w] = bp; uneg[j * wd + w] = bn; }
                        eta_p += pos ? yv[u] : 0.0;
                        eta_p -= neg ? yv[u] : 0.0;
                        cnt_p += __popc(bp) + __popc(bn);
                    }
                }
            }
            if (live && mv == MOVE_PRUNE) {
                // eta = sum of y over the right leaf = its entry of b; counted once per slot (its part-0 warp), and the
                // per-lane popcounts above are per-lane partial counts: fold them over the warp
                eta_p = (part == 0 && lane == 0) ? __ldcg(cv.b + pb) : 0.0;
            }
            if (!live || mv == MOVE_PRUNE) cnt_p = warp_sum_int(cnt_p);  // (the ballot path counts warp-wide already)
            SB_PROBE_END(1);  // gathers + masks
            // per-slot totals: eta (fixed tree over lanes, fixed order over the slot's warps), n_u (exact integers)
            const double e = warp_sum(eta_p);
            if (lane == 0) { red[wid] = e; red[SB_WARPS + wid] = (double)cnt_p; }
            __syncthreads();
            if (tid < KS) {
                double es = 0.0, cn = 0.0;
#pragma unroll
                for (int pp = 0; pp < nparts; ++pp) { es += red[pp * KS + tid]; cn += red[SB_WARPS + pp * KS + tid]; }
                ctl->eta[tid] = es;
                ctl->nu[tid] = cn;
            }
        }
        SB_MARK(2);

        // ------------------------------------------------------------------ phase 2: V = Z^T U (AND + POPC), G = U^T U
        {
            const int share = E / R;  // E is a multiple of 16 >= R
            const int r0 = cr * share, r1 = r0 + share;
            // a prune's u is a whole leaf, so its v is that leaf's column of the exact integer A: no scan
            unsigned negmask = 0, valmask = 0, scanmask = 0;
            for (int j = 0; j < nb; ++j) {
                if (ctl->prop[j].valid) {
                    valmask |= 1u << j;
                    if (ctl->prop[j].move != MOVE_PRUNE) scanmask |= 1u << j;
                    if (ctl->prop[j].move == MOVE_CHANGE) negmask |= 1u << j;
                }
            }
            const int nv4 = wd >> 2;  // 16-byte vectors per bitset (wd is a multiple of 4)
            for (int base = r0; base < r1; base += SB_THREADS / 2) {
                const int q = base + (tid >> 1), part = tid & 1;
                int cnt[KS];
#pragma unroll
                for (int j = 0; j < KS; ++j) cnt[j] = 0;
                if (q < r1 && q < p_hi && scanmask) {
                    const uint4* bq = reinterpret_cast<const uint4*>(cv.bits + (size_t)q * wd);
                    for (int v0 = part; v0 < nv4; v0 += 8) {
                        uint4 x[4];
#pragma unroll
                        for (int g = 0; g < 4; ++g) x[g] = (v0 + 2 * g < nv4) ? __ldcg(bq + v0 + 2 * g) : make_uint4(0, 0, 0, 0);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (v0 + 2 * g >= nv4) break;
                            if ((x[g].x | x[g].y | x[g].z | x[g].w) == 0u) continue;
                            const int wo = (v0 + 2 * g) * 4;
#pragma unroll
                            for (int j = 0; j < KS; ++j) {
                                if (!((scanmask >> j) & 1u)) continue;
                                const uint4 up = *reinterpret_cast<const uint4*>(upos + j * wd + wo);
                                cnt[j] += __popc(x[g].x & up.x) + __popc(x[g].y & up.y) + __popc(x[g].z & up.z) + __popc(x[g].w & up.w);
                                if ((negmask >> j) & 1u) {
                                    const uint4 un = *reinterpret_cast<const uint4*>(uneg + j * wd + wo);
                                    cnt[j] -= __popc(x[g].x & un.x) + __popc(x[g].y & un.y) + __popc(x[g].z & un.z) + __popc(x[g].w & un.w);
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < KS; ++j) cnt[j] += __shfl_xor_sync(0xffffffffu, cnt[j], 1);
                if (q < r1) {
                    // a prune's column of A: all loads before the first (distributed) shared-memory store
                    const unsigned prmask = valmask & ~scanmask;
                    if (prmask) {
#pragma unroll
                        for (int j = 0; j < KS; ++j)
                            if (((prmask >> j) & 1u) && ((j & 1) == part || KS == 1))
                                cnt[j] = (q < p_hi) ? __ldcg(cv.A + (size_t)q * P + ctl->prop[j].b) : 0;
                    }
                    // the two threads of a column write its KS values to every CTA of the cluster
#pragma unroll
                    for (int j = 0; j < KS; ++j) {
                        if (((j & 1) == part || KS == 1) && (KS > 1 || part == 0)) {
                            const double val = (double)cnt[j];
                            for (int r = 0; r < R; ++r) cluster.map_shared_rank(V, r)[(size_t)j * PS + q] = val;
                        }
                    }
                }
            }
            // g_ij = u_i^T u_j, i < j: one warp per pair
            constexpr int npair = KS * (KS - 1) / 2;
            for (int pr = wid; pr < npair; pr += SB_WARPS) {
                int i = 0, rem = pr;
                while (rem >= KS - 1 - i) { rem -= KS - 1 - i; ++i; }
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
#ifdef BARK_PHASE_TIMING
        long long pr3__ = clock64();
#endif

        // ------------------------------------------------------------------ phase 3: Wv = Binv V on the FP64 tensor pipe
        // Unit I = the 8 rows [8I, 8I+8) of the result: row part (row block I of the lower triangle times V) plus
        // column part (column block I below the diagonal, transposed, times V) -- (nb8 + 1) 8x8 blocks whatever I is,
        // so every unit costs the same.  One warp per unit; C fragment = Y[8I + lq][2 lk, 2 lk + 1].  In both parts
        // the k index of DMMA step s is row / column 2 lk + s of the block (any bijection of the eight works), so the B
        // operands of the two steps are ONE 16-byte load from the vector-major V.  Eight blocks' loads are in flight
        // at a time, four independent accumulator chains.
        for (int I = gw; I < nb8; I += ngw) {
            double acc[4][2];
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u][0] = acc[u][1] = 0.0;
            const int row = 8 * I + lq;
            const bool vec_on = lq < KS;
            const double2 zero2 = make_double2(0.0, 0.0);
            const double* vp = V + (size_t)(vec_on ? lq : 0) * PS + 2 * lk;  // B operands: *(double2*)(vp + 8 * block)
            // ---- row part, blocks 0 .. I-1 (strictly below the diagonal block): A[lq][k] = Binv[8I + lq][8 kb + 2 lk + s]
            {
                const double* rp = cv.Binv + (size_t)row * P + 2 * lk;
                int kb = 0;
                for (; kb + 8 <= I; kb += 8) {
                    double2 x[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) x[u] = sb_ldcg2(rp + 8 * (kb + u));
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const double2 bb = vec_on ? *reinterpret_cast<const double2*>(vp + 8 * (kb + u)) : zero2;
                        la::dmma_m8n8k4(acc[u & 3][0], acc[u & 3][1], x[u].x, bb.x);
                        la::dmma_m8n8k4(acc[u & 3][0], acc[u & 3][1], x[u].y, bb.y);
                    }
                }
                // remainder (fewer than eight blocks) and the diagonal block, of which only columns <= row are current
                double2 x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = (kb + u <= I) ? sb_ldcg2(rp + 8 * (kb + u)) : zero2;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (kb + u > I) break;
                    if (kb + u == I) {
                        if (2 * lk > lq) x[u].x = 0.0;
                        if (2 * lk + 1 > lq) x[u].y = 0.0;
                    }
                    const double2 bb = vec_on ? *reinterpret_cast<const double2*>(vp + 8 * (kb + u)) : zero2;
                    la::dmma_m8n8k4(acc[u & 3][0], acc[u & 3][1], x[u].x, bb.x);
                    la::dmma_m8n8k4(acc[u & 3][0], acc[u & 3][1], x[u].y, bb.y);
                }
            }
            // ---- column part, blocks I (strictly lower part) .. nb8-1: A[lq][k] = Binv[8 ib + 2 lk + s][8I + lq]
            {
                const double* cp = cv.Binv + (size_t)(8 * I + 2 * lk) * P + 8 * I + lq;  // block ib: cp + (ib - I) * 8 P
                const size_t bstep = (size_t)8 * P;
                int ib = I;
                {   // diagonal block
                    double x0 = __ldcg(cp), x1 = __ldcg(cp + P);
                    if (2 * lk <= lq) x0 = 0.0;
                    if (2 * lk + 1 <= lq) x1 = 0.0;
                    const double2 bb = vec_on ? *reinterpret_cast<const double2*>(vp + 8 * ib) : zero2;
                    la::dmma_m8n8k4(acc[0][0], acc[0][1], x0, bb.x);
                    la::dmma_m8n8k4(acc[0][0], acc[0][1], x1, bb.y);
                    ++ib;
                    cp += bstep;
                }
                for (; ib + 8 <= nb8; ib += 8) {
                    double x0[8], x1[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        x0[u] = __ldcg(cp + u * bstep);
                        x1[u] = __ldcg(cp + u * bstep + P);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const double2 bb = vec_on ? *reinterpret_cast<const double2*>(vp + 8 * (ib + u)) : zero2;
                        la::dmma_m8n8k4(acc[u & 3][0], acc[u & 3][1], x0[u], bb.x);
                        la::dmma_m8n8k4(acc[u & 3][0], acc[u & 3][1], x1[u], bb.y);
                    }
                    cp += 8 * bstep;
                }
                double x0[8], x1[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    x0[u] = x1[u] = 0.0;
                    if (ib + u < nb8) {
                        x0[u] = __ldcg(cp + u * bstep);
                        x1[u] = __ldcg(cp + u * bstep + P);
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (ib + u >= nb8) break;
                    const double2 bb = vec_on ? *reinterpret_cast<const double2*>(vp + 8 * (ib + u)) : zero2;
                    la::dmma_m8n8k4(acc[u & 3][0], acc[u & 3][1], x0[u], bb.x);
                    la::dmma_m8n8k4(acc[u & 3][0], acc[u & 3][1], x1[u], bb.y);
                }
            }
            const double c0 = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
            const double c1 = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
            if (2 * lk < KS) {
                for (int r = 0; r < R; ++r) {
                    double* dst = cluster.map_shared_rank(Wv, r) + (size_t)(2 * lk) * PS + row;
                    dst[0] = c0;
                    if (2 * lk + 1 < KS) dst[PS] = c1;
                }
            }
        }
#ifdef BARK_PHASE_TIMING
        if (lane == 0 && chain == 0 && t0 == 80 && sweep_in_call == 0 && sweep_offset >= 100)
            printf("probe ph=3 R=%d cr=%d wid=%d own=%lld\n", R, cr, wid, (long long)(clock64() - pr3__));
#endif
        SB_MARK(5);
        csync();  // (B2) Wv complete on every CTA
        SB_MARK(6);

        // ------------------------------------------------------------------ phase 4: decide the proposals in order
        // (identical on every CTA of the cluster).  Thread mapping of the O(P) vector work: slot my_j = tid / TPS,
        // rows k = tid % TPS + TPS * it (consecutive threads, consecutive rows of one vector).
        const int my_j = tid / TPS, kf = tid - my_j * TPS;
        const double* Vj = V + (size_t)my_j * PS;
        double* Wdj = Wd + (size_t)my_j * PS;
        double* Wvj = Wv + (size_t)my_j * PS;
        auto refresh_scalars = [&](int from, double* buf) {  // vWv_j, vw_j for the pending slots j >= from
            double x = 0.0, y = 0.0;
            if (my_j >= from && my_j < nb) {
                for (int k = kf; k < E; k += TPS) {
                    const double vv = Vj[k];
                    x = fma(vv, Wvj[k], x);
                    y = fma(vv, w_s[k], y);
                }
            }
            sb_slot_sum2<KS>(x, y, buf);
            if (kf == 0) { ctl->vWv[my_j] = x; ctl->vw[my_j] = y; }
        };
        refresh_scalars(0, red2);
        __syncthreads();
        while (true) {
            if (wid == 0) {
                // ---- the walk (bark_sampler.py:257-264).  Lane j evaluates slot j against the CURRENT state; the slots
                // up to the first accepted one are thereby decided exactly as a one-at-a-time walk would decide them
                // (nothing changes before the first accept), the later ones are re-evaluated after the update.
                const int from = ctl->walk_from;
                const int j = lane;
                const bool in_blk = j < nb;
                const bool act = in_blk && j >= from && ctl->prop[in_blk ? j : 0].valid;
                const double cur_mll = ctl->mll;
                double M00 = 0.0, M01 = 0.0, M11 = 0.0, det = -1.0, Ur0 = 0.0, Ur1 = 0.0, new_res = 0.0, new_ldt = 0.0;
                double new_mll = cur_mll, lqp = -INFINITY;
                int a = 0, b = 0, mv = 0;
                bool accept = false;
                if (act) {
                    a = ctl->prop[j].a; b = ctl->prop[j].b; mv = ctl->prop[j].move; lqp = ctl->prop[j].lqp;
                    // a grow's column is still unallocated: Binv e_f = e_f / c, v_f = 0, w_f = 0 whichever f it will be
                    const double Wd_a = (a >= 0) ? Wd[(size_t)j * PS + a] : inv_c;
                    const double Wv_a = (a >= 0) ? Wv[(size_t)j * PS + a] : 0.0;
                    const double w_a = (a >= 0) ? w_s[a] : 0.0;
                    const double eta = ctl->eta[j], n_u = ctl->nu[j];
                    const double dWd = Wd_a - Wd[(size_t)j * PS + b];
                    const double dWv = Wv_a - Wv[(size_t)j * PS + b];
                    const double dw = w_a - w_s[b];
                    const double vWv = ctl->vWv[j], vw = ctl->vw[j];
                    M00 = dWd; M01 = 1.0 + dWv; M11 = -n_u + vWv;
                    det = M00 * M11 - M01 * M01;  // < 0 for an SPD B'
                    new_ldt = ctl->ldt + log(-det);
                    // q = b^T Binv b changes by dq = (b'^T Binv b' - q) - [d v]^T-part; the residual y^T y - q takes -dq
                    const double dq_b = 2.0 * eta * dw + eta * eta * dWd;  // b'^T Binv b' - b^T Binv b
                    Ur0 = dw + eta * dWd;                                   // d^T r,  r = Binv b'
                    Ur1 = vw + eta * dWv;                                   // v^T r
                    new_res = ctl->res - (dq_b - (M11 * Ur0 * Ur0 - 2.0 * M01 * Ur0 * Ur1 + M00 * Ur1 * Ur1) / det);
                    new_mll = 0.5 * (-new_res / sig - nlogsig - new_ldt);
                    const double log_alpha = lqp + (new_mll - cur_mll);
                    accept = ctl->luacc[j] <= fmin(log_alpha, 0.0);
                }
                const unsigned accm = __ballot_sync(0xffffffffu, accept);
                int first = accm ? (__ffs(accm) - 1) : -1;
                int f = -1;
                if (first >= 0 && __shfl_sync(0xffffffffu, mv, first) == MOVE_GROW) {
                    f = warp_find_free_col(used_s, P);  // lowest free column; columns pruned in this block stay reserved
                    if (f < 0 || f >= E) {
                        // no column left (f >= E cannot happen: E >= p_hi + nb, at most nb columns are taken per block):
                        // the proposal becomes invalid and the walk is repeated from the same slot
                        if (lane == first) {
                            atomicOr(&sc->status, BARK_ST_COL_OVERFLOW);
                            ctl->prop[j].valid = 0;
                            ctl->acc_slot = -2;
                        }
                        first = -2;
                    }
                }
                if (first != -2) {
                    const int last = (first >= 0) ? first : nb - 1;  // slots from..last are decided by this round
                    const bool decided = in_blk && j >= from && j <= last;
                    if (decided && trace_base && cr == 0) {
                        trace_base[(t0 + j) * 3 + 0] = lqp;  // -inf for an invalid proposal
                        trace_base[(t0 + j) * 3 + 1] = new_mll;
                        trace_base[(t0 + j) * 3 + 2] = (j == first) ? 1.0 : 0.0;
                    }
                    const unsigned dm = __ballot_sync(0xffffffffu, decided && act);
                    const unsigned m0 = __ballot_sync(0xffffffffu, decided && act && mv == MOVE_GROW);
                    const unsigned m1 = __ballot_sync(0xffffffffu, decided && act && mv == MOVE_PRUNE);
                    if (lane == 0) {
                        n_valid += __popc(dm);
                        n_valid_move[0] += __popc(m0);
                        n_valid_move[1] += __popc(m1);
                        n_valid_move[2] += __popc(dm) - __popc(m0) - __popc(m1);
                    }
                    const int mv_first = (first >= 0) ? __shfl_sync(0xffffffffu, mv, first) : 0;
                    if (lane == 0 && first >= 0) {
                        ++n_acc_tot;
                        ++n_acc_move[mv_first];
                    }
                    if (first >= 0 && lane == first) {
                        if (mv == MOVE_GROW) {
                            a = f;
                            ctl->prop[j].a = f;
                            Wd[(size_t)j * PS + f] += inv_c;  // the e_f / c term of Wd = Binv (e_f - e_b)
                            used_s[f >> 5] |= 1u << (f & 31);
                            if (f + 1 > ctl->p_hi) ctl->p_hi = f + 1;
                        }
                        SbAccepted& A = ctl->acc[ctl->n_acc];
                        A.al = M11 / det; A.be = -M01 / det; A.ga = M00 / det;
                        A.eta = ctl->eta[j]; A.nu = ctl->nu[j]; A.slot = j; A.a = a; A.b = b; A.move = mv;
                        ctl->cw_d = A.al * Ur0 + A.be * Ur1;
                        ctl->cw_v = A.be * Ur0 + A.ga * Ur1;
                        ctl->n_acc += 1;
                        ctl->res = new_res; ctl->ldt = new_ldt; ctl->mll = new_mll;
                        ctl->walk_from = j + 1;
                        ctl->acc_slot = j;
                    }
                    if (first < 0 && lane == 0) {
                        ctl->walk_from = nb;
                        ctl->acc_slot = -1;
                    }
                }
            }
            __syncthreads();
            const int i = ctl->acc_slot;
            if (i == -2) {  // a grow found no free column and was invalidated: walk again
                __syncthreads();
                continue;
            }
            if (i < 0) break;
            // ---- accepted slot i: bring the pending proposals j > i, and w, up to date
            const SbAccepted A = ctl->acc[ctl->n_acc - 1];
            const int ai = A.a, bi = A.b;
            const double* Wdi = Wd + (size_t)i * PS;
            const double* Wvi = Wv + (size_t)i * PS;
            const bool pend = my_j > i && my_j < nb && ctl->prop[my_j].valid;
            const double g = pend ? (double)ctl->G[i][my_j] : 0.0;
            if (kf == 0 && pend && g != 0.0) {  // v_j += g d_i
                V[(size_t)my_j * PS + ai] += g;
                V[(size_t)my_j * PS + bi] -= g;
            }
            __syncthreads();
            double t0s = 0.0, t1s = 0.0;
            if (pend) {
                for (int k = kf; k < E; k += TPS) {
                    const double vv = Vj[k];
                    t0s = fma(Wdi[k], vv, t0s);
                    t1s = fma(Wvi[k], vv, t1s);
                }
            }
            sb_slot_sum2<KS>(t0s, t1s, red);
            {
                double ddj = 0.0, dvj = 0.0;
                if (pend) {
                    const int aj = ctl->prop[my_j].a, bj = ctl->prop[my_j].b;  // aj < 0: a grow's column, not yet allocated
                    ddj = ((aj >= 0) ? Wdi[aj] : 0.0) - Wdi[bj];
                    dvj = ((aj >= 0) ? Wvi[aj] : 0.0) - Wvi[bj];
                }
                const double cd_d = A.al * ddj + A.be * dvj, cd_v = A.be * ddj + A.ga * dvj;    // M^-1 W^T d_j
                const double cv_d = A.al * t0s + A.be * t1s, cv_v = A.be * t0s + A.ga * t1s;    // M^-1 W^T v_j
                const double cw_d = ctl->cw_d, cw_v = ctl->cw_v;
                if (pend) {
                    for (int k = kf; k < E; k += TPS) {
                        const double wd_i = Wdi[k], wv_i = Wvi[k];
                        Wdj[k] -= wd_i * cd_d + wv_i * cd_v;
                        Wvj[k] += g * wd_i - wd_i * cv_d - wv_i * cv_v;
                    }
                }
                // w' = Binv' b' (slot i's own threads are idle in this pass: they take w)
                if (my_j == i) {
                    for (int k = kf; k < E; k += TPS) {
                        const double wd_i = Wdi[k], wv_i = Wvi[k];
                        w_s[k] = w_s[k] + A.eta * wd_i - wd_i * cw_d - wv_i * cw_v;
                    }
                }
            }
            __syncthreads();
            if (tid == 0 && A.move == MOVE_PRUNE) w_s[bi] = 0.0;  // column b is an empty leaf from now on
            __syncthreads();
            refresh_scalars(i + 1, red2);
            __syncthreads();
        }
        SB_MARK(7);
#ifdef BARK_PHASE_TIMING
        long long pr5__ = clock64();
#endif

        // ------------------------------------------------------------------ phase 5: block-end updates of the global state
        const int na = ctl->n_acc;
        if (tid == 0) {
            blk_eval += (unsigned long long)E * E;
            cols_scanned += (unsigned long long)E;
            if (na) blk_upd += (unsigned long long)E * E;
            blk_upd_rank += (unsigned long long)E * E * (unsigned long long)na;
        }
        if (na > 0) {
            // Binv -= sum_s W_s M_s^-1 W_s^T on the lower triangle: C (8x8) += A (8 x 4) B (4 x 8), two accepted
            // proposals per DMMA step (k index = 2 * (accepted s mod 2) + component); eight tiles' loads in flight, then
            // their stores (a load -> store loop on the same array is serialised by possible aliasing).
            constexpr int MAXST = (KS + 1) / 2;
            const int nsteps = (na + 1) >> 1;
            const int s_of = lk >> 1, comp = lk & 1;
            // this lane's B-operand recipe per step: bfr = c1 * Wd_s[col] + c2 * Wv_s[col]; A operand = -W_s,comp[row]
            double c1[MAXST], c2[MAXST];
            int woff[MAXST];
            int n_prune = 0;
#pragma unroll
            for (int st = 0; st < MAXST; ++st) {
                const int s = 2 * st + s_of;
                c1[st] = c2[st] = 0.0;
                woff[st] = 0;
                if (s < na) {
                    const SbAccepted& A = ctl->acc[s];
                    c1[st] = comp ? A.be : A.al;
                    c2[st] = comp ? A.ga : A.be;
                    woff[st] = A.slot * PS;
                }
            }
            for (int s = 0; s < na; ++s) n_prune += (ctl->acc[s].move == MOVE_PRUNE);
            if (ngw <= 2 * SB_WARPS) {
                // up to two CTAs per chain: row blocks paired (I, nb8-1-I), every unit nb8 + 1 tiles, about one unit per warp
                // (the leanest inner loop: the A fragment and the row pointer are per unit, not per tile)
                for (int pu = gw; pu < (nb8 + 1) / 2; pu += ngw) {
                    for (int half = 0; half < 2; ++half) {
                        const int I = half ? nb8 - 1 - pu : pu;
                        if (half && I == pu) break;
                        const int row = 8 * I + lq;
                        double afr[MAXST];
    #pragma unroll
                        for (int st = 0; st < MAXST; ++st) {
                            afr[st] = 0.0;
                            if (2 * st + s_of < na) afr[st] = -(comp ? Wv[woff[st] + row] : Wd[woff[st] + row]);
                        }
                        double* rowp = cv.Binv + (size_t)row * P + 2 * lk;
                        for (int J0 = 0; J0 <= I; J0 += 8) {
                            double2 cc[8];
    #pragma unroll
                            for (int u = 0; u < 8; ++u) cc[u] = (J0 + u <= I) ? sb_ldcg2(rowp + 8 * (J0 + u)) : make_double2(0.0, 0.0);
    #pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                if (J0 + u > I) break;
                                const int jc = 8 * (J0 + u) + lq;  // B fragment column
    #pragma unroll
                                for (int st = 0; st < MAXST; ++st) {
                                    if (st >= nsteps) break;
                                    const double bfr = fma(c1[st], Wd[woff[st] + jc], __dmul_rn(c2[st], Wv[woff[st] + jc]));  // explicit: both code paths round alike
                                    la::dmma_m8n8k4(cc[u].x, cc[u].y, afr[st], bfr);
                                }
                                if (n_prune) {  // a pruned column becomes an empty leaf: its row / column is exactly e_b / c
                                    const int col = 8 * (J0 + u) + 2 * lk;
                                    for (int s = 0; s < na; ++s) {
                                        if (ctl->acc[s].move != MOVE_PRUNE) continue;
                                        const int pb = ctl->acc[s].b;
                                        if (row == pb || col == pb) cc[u].x = (row == col) ? inv_c : 0.0;
                                        if (row == pb || col + 1 == pb) cc[u].y = (row == col + 1) ? inv_c : 0.0;
                                    }
                                }
                            }
    #pragma unroll
                            for (int u = 0; u < 8; ++u)
                                if (J0 + u <= I) __stcg(reinterpret_cast<double2*>(rowp + 8 * (J0 + u)), cc[u]);
                        }
                    }
                }
            } else {
                // The nb8 (nb8 + 1) / 2 tiles of the lower triangle, row-major (I, J <= I), are dealt out in equal contiguous
                // ranges to the cluster's warps (whatever the cluster size, every warp has work; a tile's result does not
                // depend on who computes it).
                const int T = nb8 * (nb8 + 1) / 2;
                const int t_begin = (int)((long long)T * gw / ngw), t_end = (int)((long long)T * (gw + 1) / ngw);
                int I0 = (int)((sqrtf(8.0f * (float)t_begin + 1.0f) - 1.0f) * 0.5f);
                while (I0 * (I0 + 1) / 2 > t_begin) --I0;
                while ((I0 + 1) * (I0 + 2) / 2 <= t_begin) ++I0;
                int J0 = t_begin - I0 * (I0 + 1) / 2;
                for (int tb = t_begin; tb < t_end; tb += 8) {
                    double2 cc[8];
                    {
                        int I = I0, J = J0;
    #pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            cc[u] = (tb + u < t_end) ? sb_ldcg2(cv.Binv + (size_t)(8 * I + lq) * P + 8 * J + 2 * lk) : make_double2(0.0, 0.0);
                            if (J == I) { ++I; J = 0; } else { ++J; }
                        }
                    }
                    {
                        int I = I0, J = J0;
    #pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            if (tb + u >= t_end) break;
                            const int row = 8 * I + lq, jc = 8 * J + lq;  // A fragment row, B fragment column
    #pragma unroll
                            for (int st = 0; st < MAXST; ++st) {
                                if (st >= nsteps) break;
                                const bool on = 2 * st + s_of < na;
                                const double* Wc = comp ? Wv : Wd;
                                const double afr = on ? -Wc[woff[st] + row] : 0.0;
                                const double bfr = fma(c1[st], Wd[woff[st] + jc], __dmul_rn(c2[st], Wv[woff[st] + jc]));  // explicit: both code paths round alike
                                la::dmma_m8n8k4(cc[u].x, cc[u].y, afr, bfr);
                            }
                            if (n_prune) {  // a pruned column becomes an empty leaf: its row / column is exactly e_b / c
                                const int col = 8 * J + 2 * lk;
                                for (int s = 0; s < na; ++s) {
                                    if (ctl->acc[s].move != MOVE_PRUNE) continue;
                                    const int pb = ctl->acc[s].b;
                                    if (row == pb || col == pb) cc[u].x = (row == col) ? inv_c : 0.0;
                                    if (row == pb || col + 1 == pb) cc[u].y = (row == col + 1) ? inv_c : 0.0;
                                }
                            }
                            if (J == I) { ++I; J = 0; } else { ++J; }
                        }
                    }
                    {
    #pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            if (tb + u < t_end) __stcg(reinterpret_cast<double2*>(cv.Binv + (size_t)(8 * I0 + lq) * P + 8 * J0 + 2 * lk), cc[u]);
                            if (J0 == I0) { ++I0; J0 = 0; } else { ++J0; }
                        }
                    }
                }
            }
#ifdef BARK_PHASE_TIMING
            if (lane == 0 && chain == 0 && t0 == 80 && sweep_in_call == 0 && sweep_offset >= 100)
                printf("probe ph=5 R=%d cr=%d wid=%d own=%lld na=%d\n", R, cr, wid, (long long)(clock64() - pr5__), na);
#endif
            // A' = A + v d^T + d v^T + n_u d d^T for every accepted proposal (exact integers; atomics make the order
            // irrelevant).  v is the slot's V column as it stood at the acceptance.
            {
                const int share = E / R, r0 = cr * share, r1 = r0 + share;
                for (int s = 0; s < na; ++s) {
                    const int slot = ctl->acc[s].slot, a = ctl->acc[s].a, b = ctl->acc[s].b;
                    for (int k = r0 + tid; k < r1; k += SB_THREADS) {
                        const int vk = (int)V[(size_t)slot * PS + k];
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
                // one warp per accepted proposal (they edit different trees and different columns, so they are independent)
                for (int s = wid; s < na; s += SB_WARPS) {
                    const SbAccepted A = ctl->acc[s];
                    const int a = A.a, b = A.b, slot = A.slot;
                    // leaf bitsets
                    for (int w = lane; w < wd; w += 32) {
                        const uint32_t ba = __ldcg(cv.bits + (size_t)a * wd + w), bb = __ldcg(cv.bits + (size_t)b * wd + w);
                        const uint32_t up = upos[slot * wd + w], un = uneg[slot * wd + w];
                        __stcg(cv.bits + (size_t)a * wd + w, (ba | up) & ~un);
                        __stcg(cv.bits + (size_t)b * wd + w, (bb & ~up) | un);
                    }
                    if (lane == 0) {
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
            sc->res = ctl->res; sc->ldt = ctl->ldt; sc->mll = ctl->mll; sc->p_hi = ctl->p_hi;
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
            sc->counters[14] += blk_upd_rank;  // sum over blocks of extent^2 x accepted proposals (rank-2 updates applied)
        }
    }
}

}  // namespace bark
