// Tree sweep: the m tree MH steps of one sweep of one chain (bark_sampler.py:233-264) in leaf space.
//
// A thread-block CLUSTER of R CTAs (R = 1, 2 or 4; 512 threads each) owns one chain: 64 chains fill 128 of the 148
// SMs with R = 2, at most 37 chains use R = 4.  Only the LOWER TRIANGLE of the symmetric B^-1 is kept current (row q
// holds columns 0..q), which halves the bytes per pass and lets 64 chains' states stay L2-resident.  Rows are
// interleaved across the cluster (row q belongs to CTA q % R):
//   phase 1  moved-point masks u+/u-, eta = u^T y, n_u           (redundant on every CTA; N bits)
//   phase 2  v = Z^T u by AND+POPC over the leaf bitsets         (columns split; shares exchanged through DSMEM)
//            and Wd = Binv d (two rows of the symmetric matrix)
//   phase 3  Wv = Binv v: every streamed row prefix yields a dot product (row part) and an axpy into per-lane
//            column accumulators (column part); partial vectors summed over the cluster through DSMEM
//   phase 4  rank 0 alone: 2x2 capacitance matrix, proposed log-MLL, MH accept, broadcast to the cluster --
//            while the last rank generates the NEXT tree's proposal and publishes it speculatively
//   accept   symmetric rank-2 update of the CTA's rows of Binv; w; integer A / bitsets / forest edits
// Cluster barriers per proposal: after phase 2, after phase 3, after phase 4, and -- only after an accepted
// proposal -- at the end.
//
// The two passes over B^-1 (matvec, rank-2 update) stream the CTA's rows, paired short + long per slot, through a
// shared-memory ring with the bulk-copy engine (TMA, cp.async.bulk + mbarrier full/empty pairs): one producer warp
// keeps up to ~170 KB of row copies in flight, 15 consumer warps reduce one row pair each, or update it and store
// it straight from registers.  A ring slot is always consumed by the same warp (mbarrier waits carry one parity bit).
// State written by a peer CTA is only read after a cluster barrier (release/acquire) and through L2 (.cg).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "mcmc_state.cuh"
#include "proposal_device.cuh"

namespace bark {
namespace cg = cooperative_groups;

constexpr int SW_THREADS = 512;
constexpr int SW_WARPS = SW_THREADS / 32;
constexpr int SW_MAX_R = 4;  // CTAs per chain: 1, 2 or 4 (the used extent is a multiple of 16, so every share is a multiple of 4)
constexpr int SW_PROD_WARP = SW_WARPS - 1;  // bulk-copy producer
constexpr int SW_NCW = SW_WARPS - 1;        // consumer warps
constexpr int SW_NSLOT_MAX = 60;  // multiple of SW_NCW

#ifdef BARK_PHASE_TIMING
#define PHASE_MARK(i)                                        \
    do {                                                     \
        if (tid == 0) {                                      \
            const long long now__ = clock64();               \
            ph_acc[i] += (unsigned long long)(now__ - ph_t); \
            ph_t = now__;                                    \
        }                                                    \
    } while (0)
#else
#define PHASE_MARK(i) do { } while (0)
#endif

// ---------------------------------------------------------------- mbarrier / bulk-copy (TMA) primitives
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.b32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Warp-uniform wait: every lane of the calling warp waits on the SAME barrier (no divergence inside).
// Bounded: a wait that does not complete within ~2^22 probes flags BARK_ST_TIMEOUT and gives up instead of
// hanging the GPU (results of that chain are then invalid and the host raises).
constexpr unsigned SW_WAIT_LIMIT = 1u << 22;
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, unsigned* status) {
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > SW_WAIT_LIMIT) {
            atomicOr(status, BARK_ST_TIMEOUT);
            break;
        }
    }
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// Rank 0 of the cluster is the single decision maker: it takes the MH decision and broadcasts it (with the update
// coefficients) to the other CTAs through DSMEM, as the generator rank does with the proposals, so the CTAs of a
// cluster can never disagree on control flow (a disagreement would dead-lock the cluster barriers).
struct Decision {
    double eta, al, be, ga, cw_d, cw_v, new_q, new_ldt, new_mll;
    int accept, pad;
};
struct SweepCtl {  // small shared control block
    Prop prop[2];  // double-buffered by tree parity: the generator rank publishes t+1 while t is in use
    Prop next;     // generator rank: proposal of the next tree before its free column is fixed up
    Decision dec;
    double q, ldt, mll;
    int p_hi;
};

struct SweepSmemLayout {
    size_t off_ctl, off_leaf, off_u32, off_cm, off_box, off_ft, off_vd, off_wd, off_wv, off_ws, off_upos, off_uneg, off_red,
        off_logtab, off_priortab, off_colused, off_bars, off_ydot, off_parts, off_ypart, off_ring, ring_bytes, total;
};
__host__ __device__ inline SweepSmemLayout sweep_smem_layout(int L, int d, int P, int wd, size_t budget) {
    SweepSmemLayout s;
    size_t o = 0;
    s.off_ctl = o;      o += align256(sizeof(SweepCtl));
    s.off_leaf = o;     o += align256((size_t)L * 2);          // is_leaf, active
    s.off_u32 = o;      o += align256((size_t)L * 4 * 6);      // feat,left,right,parent,depth,thr
    s.off_cm = o;       o += align256((size_t)L * 2);          // this tree's leaf -> column map
    s.off_box = o;      o += align256((size_t)d * 2 * 8);
    s.off_ft = o;       o += align256((size_t)d * 4);
    s.off_vd = o;       o += align256((size_t)P * 8);
    s.off_wd = o;       o += align256((size_t)P * 8);
    s.off_wv = o;       o += align256((size_t)P * 8);
    s.off_ws = o;       o += align256((size_t)P * 8);          // w = Binv b (full copy per CTA)
    s.off_upos = o;     o += align256((size_t)wd * 4);
    s.off_uneg = o;     o += align256((size_t)wd * 4);
    s.off_red = o;      o += align256(80 * 8);
    s.off_logtab = o;   o += align256((size_t)(L + 2) * 8);
    s.off_priortab = o; o += align256((size_t)(L + 1) * 8);
    s.off_colused = o;  o += align256((size_t)(P / 32) * 4);
    s.off_bars = o;     o += align256((size_t)SW_NSLOT_MAX * 2 * 8);
    s.off_ydot = o;     o += align256((size_t)P * 8);                    // row-dot part of the symmetric matvec
    s.off_parts = o;    o += align256((size_t)P * 8) * SW_MAX_R;         // per-CTA partial vectors (DSMEM targets)
    s.off_ypart = o;    o += align256((size_t)SW_NCW * 256 * 8);         // per-warp column accumulators, half a panel at a time
    o = (o + 1023) & ~(size_t)1023;
    s.off_ring = o;
    s.ring_bytes = (budget > o + 1024) ? ((budget - o) & ~(size_t)1023) : 0;
    if (s.ring_bytes > 176 * 1024) s.ring_bytes = 176 * 1024;
    s.total = o + s.ring_bytes;
    return s;
}

__device__ __forceinline__ double2 ldcg2(const double* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }

// deterministic block-wide sum of two values at once (one barrier pair)
__device__ __forceinline__ void block_sum2(double& a, double& b, double* scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    warp_sum2(a, b);
    __syncthreads();
    if (lane == 0) { scratch[wid] = a; scratch[32 + wid] = b; }
    __syncthreads();
    const double ra = (lane < SW_WARPS) ? scratch[lane] : 0.0;
    const double rb = (lane < SW_WARPS) ? scratch[32 + lane] : 0.0;
    a = ra;
    b = rb;
    warp_sum2(a, b);
}

__global__ void __launch_bounds__(SW_THREADS, 1)
sweep_trees_kernel(WsLayout lay, void* ws, bark_nodes_soa forest, bark_params prm, int64_t sweep_in_call,
                   int64_t n_sweeps_call, uint64_t seed, int64_t chain_offset, int64_t sweep_offset,
                   const double* __restrict__ tape, double* __restrict__ trace, size_t smem_budget) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int R = (int)cluster.num_blocks(), cr = (int)cluster.block_rank();
    auto csync = [&]() {
        if (R > 1) {
            cluster.sync();
            fence_proxy_async();  // reader side: peers' generic / bulk writes before our bulk / generic reads
        } else {
            __syncthreads();
        }
    };

    const int P = (int)lay.P, L = (int)lay.L, m = (int)lay.m, n = (int)lay.n, wd = (int)lay.wd, npad = (int)lay.npad;
    const int d = (int)lay.d;
    const SweepSmemLayout sl = sweep_smem_layout(L, d, P, wd, smem_budget);
    SweepCtl* ctl = (SweepCtl*)(smem_raw + sl.off_ctl);
    TreeSmem T;
    T.is_leaf = smem_raw + sl.off_leaf; T.active = T.is_leaf + L;
    T.feat = (uint32_t*)(smem_raw + sl.off_u32); T.left = T.feat + L; T.right = T.left + L; T.parent = T.right + L;
    T.depth = T.parent + L; T.thr = (float*)(T.depth + L);
    uint16_t* cm_s = (uint16_t*)(smem_raw + sl.off_cm);
    double* box = (double*)(smem_raw + sl.off_box);
    int32_t* ftc = (int32_t*)(smem_raw + sl.off_ft);
    double* vd = (double*)(smem_raw + sl.off_vd);
    double* Wd = (double*)(smem_raw + sl.off_wd);
    double* Wv = (double*)(smem_raw + sl.off_wv);
    double* w_s = (double*)(smem_raw + sl.off_ws);
    uint32_t* upos = (uint32_t*)(smem_raw + sl.off_upos);
    uint32_t* uneg = (uint32_t*)(smem_raw + sl.off_uneg);
    double* red = (double*)(smem_raw + sl.off_red);
    double* logtab = (double*)(smem_raw + sl.off_logtab);
    double* priortab = (double*)(smem_raw + sl.off_priortab);
    uint32_t* colused_s = (uint32_t*)(smem_raw + sl.off_colused);
    uint64_t* full_bar = (uint64_t*)(smem_raw + sl.off_bars);
    uint64_t* empty_bar = full_bar + SW_NSLOT_MAX;
    double* ydot = (double*)(smem_raw + sl.off_ydot);
    double* parts = (double*)(smem_raw + sl.off_parts);
    double* ypart = (double*)(smem_raw + sl.off_ypart);
    unsigned char* ring = smem_raw + sl.off_ring;
    const size_t parts_stride = align256((size_t)P * 8) / 8;

    // peers' copies of the exchanged vectors (distributed shared memory)
    double* vd_peer[SW_MAX_R];
    double* parts_peer[SW_MAX_R];  // this CTA's slot (index cr) inside every CTA's `parts`
#pragma unroll
    for (int r = 0; r < SW_MAX_R; ++r) {
        vd_peer[r] = (r < R) ? cluster.map_shared_rank(vd, r) : vd;
        parts_peer[r] = ((r < R) ? cluster.map_shared_rank(parts, r) : parts) + (size_t)cr * parts_stride;
    }

    const int64_t chain = blockIdx.x / R;
    ChainView cv = chain_view(lay, ws, chain);
    SharedView sv = shared_view(lay, ws);
    ChainScalars* sc = cv.sc;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    // a dead chain (status set by an earlier launch) is skipped by the whole cluster
    if (__ldcg(&sc->status) & (BARK_ST_COL_OVERFLOW | BARK_ST_TREE_OVERFLOW | BARK_ST_HYPER_MODE)) return;

    const double sig = sc->sig, c = sc->c, yy = sc->yy;
    const double nlogsig = (double)n * log(sig);
    const int p_hi_start = sc->p_hi;
    if (tid == 0) {
        ctl->q = sc->q; ctl->ldt = sc->ldt; ctl->mll = sc->mll; ctl->p_hi = p_hi_start;
    }
    for (int e = tid; e < d; e += SW_THREADS) ftc[e] = sv.ft[e];
    for (int e = tid; e < L + 2; e += SW_THREADS) logtab[e] = log((double)e);
    for (int e = tid; e < L + 1; e += SW_THREADS) priortab[e] = log_prior_ratio_at_depth((uint32_t)e, prm.alpha, prm.beta);
    for (int e = tid; e < P / 32; e += SW_THREADS) colused_s[e] = __ldcg(cv.colused + e);
    for (int e = tid; e < P; e += SW_THREADS) w_s[e] = __ldcg(cv.w + e);

    // ring geometry for this launch: one slot per row segment, sized for the extent at launch + a little head-room
    // (a proposal whose segments would not fit falls back to direct loads).  A slot is always consumed by the
    // same warp (slot = use % nslot, warp = use % SW_NCW, nslot a multiple of SW_NCW): mbarrier waits only carry
    // one parity bit, so the successive laps of a slot must be observed in order by one waiter.
    // (+16: a slot also takes a PAIR of complementary row prefixes, short + long, total <= extent + 4)
    const int slot_cols = min(528, min(P, ((p_hi_start + 16) + 15) & ~15) + 16);
    const uint32_t slot_bytes = (uint32_t)slot_cols * 8u;
    const int nslot = (int)(min((size_t)SW_NSLOT_MAX, sl.ring_bytes / slot_bytes) / SW_NCW) * SW_NCW;
    const bool ring_ok = nslot >= SW_NCW;
    const int npl = max(1, min(32, nslot / 2));  // producer lanes
    if (tid == 0 && ring_ok) {
        for (int i = 0; i < nslot; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); }
        fence_mbar_init();
    }
    fence_proxy_async();  // state written with generic stores by earlier kernels -> later bulk (async proxy) reads
    uint32_t ring_base = 0;  // rows streamed so far (identical on every thread)

    unsigned long long n_valid = 0, n_acc = 0, n_acc_move[3] = {0, 0, 0}, n_valid_move[3] = {0, 0, 0};  // thread 0
    unsigned long long blk_eval = 0, blk_upd = 0, cols_scanned = 0;

    const uint32_t g_chain = (uint32_t)(chain_offset + chain), g_sweep = (uint32_t)(sweep_offset + sweep_in_call);
    const size_t tape_base =
        tape ? ((size_t)(chain * n_sweeps_call + sweep_in_call)) * (size_t)(m * TAPE_PER_TREE + TAPE_PER_HYPER) : 0;
    double* trace_base = trace ? trace + ((size_t)(chain * n_sweeps_call + sweep_in_call)) * (size_t)(m + 1) * 3 : nullptr;

#ifdef BARK_PHASE_TIMING
    unsigned long long ph_acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long ph_t = clock64();
#endif
    // The proposal of tree t+1 is generated by the LAST rank of the cluster while rank 0 takes the MH decision of tree t
    // (with one CTA per chain the two simply run back to back).  Only the free column of a grow depends on that
    // decision; it is filled in afterwards.  gen_rank stages the tree, warp 0 proposes into ctl->next.
    const int gen_rank = R - 1;
    auto generate = [&](int tn) {  // executed by every thread of the generator CTA
        double un[6];
        if (tape) {
            for (int k = 0; k < TAPE_PER_TREE; ++k) un[k] = tape[tape_base + (size_t)tn * TAPE_PER_TREE + k];
        } else {
            rng_uniforms(seed, g_chain, g_sweep, (uint32_t)tn, TAPE_PER_TREE, un);
        }
        const int64_t gn = (chain * (int64_t)m + tn) * L;
        __syncthreads();  // previous users of the staging buffers are done
        for (int e = tid; e < L; e += SW_THREADS) {
            T.is_leaf[e] = __ldcg(forest.is_leaf + gn + e);
            T.active[e] = __ldcg(forest.active + gn + e);
            T.feat[e] = __ldcg(forest.feature + gn + e);
            T.left[e] = __ldcg(forest.left + gn + e);
            T.right[e] = __ldcg(forest.right + gn + e);
            T.parent[e] = __ldcg(forest.parent + gn + e);
            T.depth[e] = __ldcg(forest.depth + gn + e);
            T.thr[e] = __ldcg(forest.threshold + gn + e);
            cm_s[e] = __ldcg(cv.colmap + (size_t)tn * L + e);
        }
        for (int e = tid; e < 2 * d; e += SW_THREADS) box[e] = sv.bounds[e];
        __syncthreads();
        if (wid == 0) {
            Prop pp = propose_tree_warp(T, L, box, ftc, d, cm_s, colused_s, P, prm, un, &sc->status, logtab, priortab, true);
            if (lane == 0) ctl->next = pp;
        }
        __syncthreads();
    };
    auto publish = [&](int tn) {  // generator CTA, after colused_s reflects every earlier decision
        if (wid == 0) {
            Prop pn = ctl->next;
            if (pn.valid && pn.move == MOVE_GROW) {
                const int fcol = warp_find_free_col(colused_s, P);
                if (fcol < 0) {
                    pn.valid = 0;
                    pn.lqp = -INFINITY;
                    if (lane == 0) atomicOr(&sc->status, BARK_ST_COL_OVERFLOW);
                } else {
                    pn.a = fcol;
                }
            }
            if (lane == 0) {
#pragma unroll
                for (int r = 0; r < SW_MAX_R; ++r)
                    if (r < R) *cluster.map_shared_rank(&ctl->prop[tn & 1], r) = pn;
            }
        }
    };
    __syncthreads();
    if (cr == gen_rank) {
        generate(0);
        publish(0);
    }
    csync();

    for (int t = 0; t < m; ++t) {
        PHASE_MARK(11);
        double u[6];
        if (tape) {
            u[4] = tape[tape_base + (size_t)t * TAPE_PER_TREE + 4];
        } else {
            rng_uniforms(seed, g_chain, g_sweep, (uint32_t)t, TAPE_PER_TREE, u);
        }
        const int64_t g0 = (chain * (int64_t)m + t) * L;
        PHASE_MARK(1);
        const Prop p = ctl->prop[t & 1];
        const double cur_q = ctl->q, cur_ldt = ctl->ldt, cur_mll = ctl->mll;
        const int p_hi = ctl->p_hi;

        double new_q = cur_q, new_ldt = cur_ldt, new_mll = cur_mll;
        double eta = 0.0, n_u = 0.0, al = 0.0, be = 0.0, ga = 0.0, cw_d = 0.0, cw_v = 0.0;
        int pe16 = 0, r0 = 0, r1 = 0;
        bool accept = false;
        bool use_ring = false;

        if (!p.valid) {
            // nothing to evaluate (the reference would still compute an MLL it can never accept): next proposal
            if (tid == 0) {
                if (trace_base && cr == 0) {
                    trace_base[t * 3 + 0] = -INFINITY;
                    trace_base[t * 3 + 1] = cur_mll;
                    trace_base[t * 3 + 2] = 0.0;
                }
            }
            if (t + 1 < m) {
                if (cr == gen_rank) {
                    generate(t + 1);
                    publish(t + 1);
                }
                csync();
            }
            continue;
        }

        if (p.valid) {
            const int a = p.a, b = p.b;
            const int pe = max(p_hi, max(a, b) + 1);
            pe16 = min(P, (pe + 15) & ~15);  // used extent, rounded to 16 columns
            const int share = pe16 / R;      // multiple of 4
            r0 = cr * share;
            r1 = r0 + share;
            use_ring = ring_ok && min(pe16, 512) <= slot_cols;
            const bool paired = use_ring && pe16 <= 512 && pe16 + 4 <= slot_cols;  // single panel, two rows per slot
            // ---- phase 1: moved-point masks u+ / u-, eta = u^T y, n_u = u^T u
            double eta_part = 0.0, cnt_part = 0.0;
            const uint32_t* bits_a = cv.bits + (size_t)a * wd;
            const uint32_t* bits_b = cv.bits + (size_t)b * wd;
            const double* xf = sv.Xt + (size_t)p.feat * npad;
            const int ftype = ftc[p.feat];
            // four points per thread at a time: all global loads are issued before the first shared-memory store
            for (int base = 0; base < npad; base += 4 * SW_THREADS) {
                uint32_t wa[4], wb[4];
                double xv[4], yv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = base + u * SW_THREADS + tid;
                    wa[u] = wb[u] = 0u;
                    xv[u] = yv[u] = 0.0;
                    if (i < npad) {
                        wb[u] = __ldcg(bits_b + (i >> 5));
                        if (p.move == MOVE_CHANGE) wa[u] = __ldcg(bits_a + (i >> 5));
                        if (i < n) {
                            xv[u] = xf[i];
                            yv[u] = sv.y[i];
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = base + u * SW_THREADS + tid;
                    if (i >= npad) break;  // warp-uniform (npad is a multiple of 32)
                    bool pos = false, neg = false;
                    if (i < n) {
                        const bool in_b = (wb[u] >> lane) & 1u;
                        if (p.move == MOVE_GROW) {
                            if (in_b) pos = !goes_left(xv[u], p.thr, ftype);
                        } else if (p.move == MOVE_PRUNE) {
                            pos = in_b;
                        } else {  // change: b = left child's column, a = right child's column
                            const bool in_a = (wa[u] >> lane) & 1u;
                            if (in_a || in_b) {
                                const bool gl = goes_left(xv[u], p.thr, ftype);
                                pos = in_b && !gl;
                                neg = in_a && gl;
                            }
                        }
                        if (pos) { eta_part += yv[u]; cnt_part += 1.0; }
                        if (neg) { eta_part -= yv[u]; cnt_part += 1.0; }
                    }
                    const unsigned bp = __ballot_sync(0xffffffffu, pos), bn = __ballot_sync(0xffffffffu, neg);
                    if (lane == 0) { upos[i >> 5] = bp; uneg[i >> 5] = bn; }
                }
            }
            block_sum2(eta_part, cnt_part, red);  // counts are exact (integers < 2^53)
            eta = eta_part;
            n_u = cnt_part;
            PHASE_MARK(2);

            // ---- phase 2: v = Z^T u for this CTA's columns [r0, r1): four threads per column, words striped
            for (int base = r0; base < r1; base += SW_THREADS / 4) {
                const int q = base + (tid >> 2), part = tid & 3;
                int cnt = 0;
                if (q < r1 && q < pe) {
                    const uint32_t* bq = cv.bits + (size_t)q * wd;
                    // wd is a multiple of 4: 16-byte loads, up to four in flight per thread
                    for (int w0 = part * 4; w0 < wd; w0 += 64) {
                        uint4 x[4];
#pragma unroll
                        for (int g = 0; g < 4; ++g)
                            x[g] = (w0 + 16 * g < wd) ? __ldcg(reinterpret_cast<const uint4*>(bq + w0 + 16 * g)) : make_uint4(0, 0, 0, 0);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (w0 + 16 * g < wd) {
                                const uint4 up = *reinterpret_cast<const uint4*>(upos + w0 + 16 * g);
                                const uint4 un = *reinterpret_cast<const uint4*>(uneg + w0 + 16 * g);
                                cnt += __popc(x[g].x & up.x) + __popc(x[g].y & up.y) + __popc(x[g].z & up.z) + __popc(x[g].w & up.w);
                                cnt -= __popc(x[g].x & un.x) + __popc(x[g].y & un.y) + __popc(x[g].z & un.z) + __popc(x[g].w & un.w);
                            }
                        }
                    }
                }
                cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
                cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
                if (q < r1 && part == 0) {
                    const double val = (double)cnt;
#pragma unroll
                    for (int r = 0; r < SW_MAX_R; ++r)
                        if (r < R) vd_peer[r][q] = val;
                }
            }
            // Wd = Binv d: rows a and b of the symmetric matrix (row prefix + column below the diagonal);
            // prune also takes its closed form Wv = e_b - c Binv[:,b] from the same loads
            for (int k = tid; k < pe16; k += SW_THREADS) {
                const double sa = __ldcg(cv.Binv + ((k <= a) ? ((size_t)a * P + k) : ((size_t)k * P + a)));
                const double sb = __ldcg(cv.Binv + ((k <= b) ? ((size_t)b * P + k) : ((size_t)k * P + b)));
                Wd[k] = sa - sb;
                if (p.move == MOVE_PRUNE) Wv[k] = ((k == b) ? 1.0 : 0.0) - c * sb;
                ydot[k] = 0.0;
            }
            PHASE_MARK(3);
            csync();  // (1) all columns of v present everywhere
            PHASE_MARK(4);

            // ---- phase 3: Wv = Binv v.  This CTA streams the prefixes of its rows (q % R == cr) in panels of 512
            // columns: row-dot part into ydot[q], column (axpy) part into per-lane accumulators, reduced over the
            // warps through shared memory; the CTA's partial vector goes to every CTA of the cluster.
            if (p.move != MOVE_PRUNE && paired) {
                // Single panel, rows paired short + long (i-th and (nrows-1-i)-th of this CTA's rows): every slot
                // carries ~extent doubles, so the ring holds twice the bytes in flight and half the hand-offs.
                const int qfirst = cr;
                const int nrows = (qfirst < pe16) ? (pe16 - 1 - qfirst) / R + 1 : 0;
                const int npairs = (nrows + 1) / 2;
                double yacc[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) yacc[j] = 0.0;
                if (wid == SW_PROD_WARP) {
                    for (int base = 0; base < npairs; base += npl) {
                        const int j = base + lane;
                        const bool active = lane < npl && j < npairs;
                        const int iA = j, iB = nrows - 1 - j;
                        const int qA = qfirst + iA * R, qB = qfirst + iB * R;
                        const uint32_t use = ring_base + (uint32_t)j;
                        const int slot = (int)(use % (uint32_t)nslot);
                        const uint32_t par = (use / (uint32_t)nslot) & 1u;
                        bool ready = !active;
                        unsigned spins = 0;
                        while (!__all_sync(0xffffffffu, ready)) {
                            if (!ready) ready = mbar_try_wait(empty_bar + slot, par ^ 1u);
                            if (++spins > SW_WAIT_LIMIT) {
                                atomicOr(&sc->status, BARK_ST_TIMEOUT);
                                break;
                            }
                        }
                        if (active) {
                            const int lenA = (qA + 2) & ~1, lenB = (iB > iA) ? ((qB + 2) & ~1) : 0;
                            unsigned char* dst = ring + (size_t)slot * slot_bytes;
                            mbar_expect_tx(full_bar + slot, (uint32_t)(lenA + lenB) * 8u);
                            bulk_g2s(dst, cv.Binv + (size_t)qA * P, (uint32_t)lenA * 8u, full_bar + slot);
                            if (lenB) bulk_g2s(dst + (size_t)lenA * 8, cv.Binv + (size_t)qB * P, (uint32_t)lenB * 8u, full_bar + slot);
                        }
                    }
                } else {
                    const int jfirst = (int)((wid + SW_NCW - (ring_base % SW_NCW)) % SW_NCW);
                    for (int j = jfirst; j < npairs; j += SW_NCW) {
                        const int iA = j, iB = nrows - 1 - j;
                        const int qA = qfirst + iA * R, qB = qfirst + iB * R;
                        const int lenA = (qA + 2) & ~1;
                        const bool hasB = iB > iA;
                        const uint32_t use = ring_base + (uint32_t)j;
                        const int slot = (int)(use % (uint32_t)nslot);
                        mbar_wait(full_bar + slot, (use / (uint32_t)nslot) & 1u, &sc->status);
                        const double* rowA = reinterpret_cast<const double*>(ring + (size_t)slot * slot_bytes);
                        const double* rowB = rowA + lenA;
                        const double vqA = vd[qA], vqB = hasB ? vd[qB] : 0.0;
                        double dotA = 0.0, dotB = 0.0;
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) {
                            const int k = lane * 2 + 64 * jj;
                            if (k <= qA) {
                                const double2 x = *reinterpret_cast<const double2*>(rowA + k);
                                dotA = fma(x.x, vd[k], dotA);
                                if (k < qA) yacc[2 * jj] = fma(x.x, vqA, yacc[2 * jj]);
                                if (k + 1 <= qA) {
                                    dotA = fma(x.y, vd[k + 1], dotA);
                                    if (k + 1 < qA) yacc[2 * jj + 1] = fma(x.y, vqA, yacc[2 * jj + 1]);
                                }
                            }
                            if (hasB && k <= qB) {
                                const double2 x = *reinterpret_cast<const double2*>(rowB + k);
                                dotB = fma(x.x, vd[k], dotB);
                                if (k < qB) yacc[2 * jj] = fma(x.x, vqB, yacc[2 * jj]);
                                if (k + 1 <= qB) {
                                    dotB = fma(x.y, vd[k + 1], dotB);
                                    if (k + 1 < qB) yacc[2 * jj + 1] = fma(x.y, vqB, yacc[2 * jj + 1]);
                                }
                            }
                        }
                        warp_sum2(dotA, dotB);  // every lane's slot reads are complete here
                        if (lane == 0) {
                            mbar_arrive(empty_bar + slot);
                            ydot[qA] += dotA;
                            if (hasB) ydot[qB] += dotB;
                        }
                    }
                }
                ring_base += (uint32_t)npairs;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if (wid < SW_NCW) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            *reinterpret_cast<double2*>(ypart + (size_t)wid * 256 + lane * 2 + 64 * j) =
                                make_double2(yacc[2 * (4 * half + j)], yacc[2 * (4 * half + j) + 1]);
                    }
                    __syncthreads();
                    for (int kk = tid; kk < 256; kk += SW_THREADS) {
                        const int k = 256 * half + kk;
                        if (k < pe16) {
                            double sacc = 0.0;
#pragma unroll
                            for (int w = 0; w < SW_NCW; ++w) sacc += ypart[(size_t)w * 256 + kk];
                            Wv[k] = sacc;
                        }
                    }
                    __syncthreads();
                }
                for (int k = tid; k < pe16; k += SW_THREADS) {
                    const double mine = Wv[k] + ydot[k];
#pragma unroll
                    for (int r = 0; r < SW_MAX_R; ++r)
                        if (r < R) parts_peer[r][k] = mine;
                }
            } else if (p.move != MOVE_PRUNE) {
                const int npanel = (pe16 + 511) / 512;
                for (int pc = 0; pc < npanel; ++pc) {
                    const int c0 = pc * 512, c1 = min(pe16, c0 + 512);
                    const int qfirst = c0 + ((cr - c0 % R) + R) % R;
                    const int nrows = (qfirst < pe16) ? (pe16 - 1 - qfirst) / R + 1 : 0;
                    double yacc[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) yacc[j] = 0.0;
                    if (use_ring && wid == SW_PROD_WARP) {
                        // npl lanes of the producer warp issue one row each per round (one thread alone caps the
                        // issue rate).  The whole warp stays convergent: a round issues when all of its slots are
                        // free, and npl <= nslot / 2 guarantees a round never waits on a slot it fills itself.
                        for (int base = 0; base < nrows; base += npl) {
                            const int i = base + lane;
                            const bool active = lane < npl && i < nrows;
                            const int q = qfirst + i * R;
                            const uint32_t use = ring_base + (uint32_t)i;
                            const int slot = (int)(use % (uint32_t)nslot);
                            const uint32_t par = (use / (uint32_t)nslot) & 1u;
                            bool ready = !active;
                            unsigned spins = 0;
                            while (!__all_sync(0xffffffffu, ready)) {
                                if (!ready) ready = mbar_try_wait(empty_bar + slot, par ^ 1u);
                                if (++spins > SW_WAIT_LIMIT) {
                                    atomicOr(&sc->status, BARK_ST_TIMEOUT);
                                    break;
                                }
                            }
                            if (active) {
                                const uint32_t bytes = (uint32_t)(((min(q + 1, c1) - c0) + 1) & ~1) * 8u;
                                mbar_expect_tx(full_bar + slot, bytes);
                                bulk_g2s(ring + (size_t)slot * slot_bytes, cv.Binv + (size_t)q * P + c0, bytes, full_bar + slot);
                            }
                        }
                    } else if (wid < SW_NCW) {
                        // row i of this pass is ring use (ring_base + i): consumed by warp (use % SW_NCW)
                        const int ifirst = use_ring ? (int)((wid + SW_NCW - (ring_base % SW_NCW)) % SW_NCW) : wid;
                        for (int i = ifirst; i < nrows; i += SW_NCW) {
                            const int q = qfirst + i * R;
                            const double* row;
                            int slot = 0;
                            if (use_ring) {
                                const uint32_t use = ring_base + (uint32_t)i;
                                slot = (int)(use % (uint32_t)nslot);
                                mbar_wait(full_bar + slot, (use / (uint32_t)nslot) & 1u, &sc->status);
                                row = reinterpret_cast<const double*>(ring + (size_t)slot * slot_bytes);
                            } else {
                                row = cv.Binv + (size_t)q * P + c0;
                            }
                            const double vq = vd[q];
                            double dot = 0.0;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int kk = lane * 2 + 64 * j, k = c0 + kk;
                                if (k <= q && k < c1) {
                                    const double2 x = use_ring ? *reinterpret_cast<const double2*>(row + kk) : ldcg2(row + kk);
                                    dot = fma(x.x, vd[k], dot);
                                    if (k < q) yacc[2 * j] = fma(x.x, vq, yacc[2 * j]);
                                    if (k + 1 <= q && k + 1 < c1) {
                                        dot = fma(x.y, vd[k + 1], dot);
                                        if (k + 1 < q) yacc[2 * j + 1] = fma(x.y, vq, yacc[2 * j + 1]);
                                    }
                                }
                            }
                            dot = warp_sum(dot);  // every lane's slot reads are complete here
                            if (lane == 0) {
                                if (use_ring) mbar_arrive(empty_bar + slot);
                                ydot[q] += dot;
                            }
                        }
                    }
                    if (use_ring) ring_base += (uint32_t)nrows;
                    // reduce the per-lane column accumulators over the consumer warps, 256 columns at a time
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        if (wid < SW_NCW) {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                *reinterpret_cast<double2*>(ypart + (size_t)wid * 256 + lane * 2 + 64 * j) =
                                    make_double2(yacc[2 * (4 * half + j)], yacc[2 * (4 * half + j) + 1]);
                        }
                        __syncthreads();
                        for (int kk = tid; kk < 256; kk += SW_THREADS) {
                            const int k = c0 + 256 * half + kk;
                            if (k < c1) {
                                double sacc = 0.0;
#pragma unroll
                                for (int w = 0; w < SW_NCW; ++w) sacc += ypart[(size_t)w * 256 + kk];
                                Wv[k] = sacc;  // column part of this CTA (Wv doubles as scratch until the cluster sum)
                            }
                        }
                        __syncthreads();
                    }
                }
                for (int k = tid; k < pe16; k += SW_THREADS) {
                    const double mine = Wv[k] + ydot[k];
#pragma unroll
                    for (int r = 0; r < SW_MAX_R; ++r)
                        if (r < R) parts_peer[r][k] = mine;
                }
            }
            PHASE_MARK(5);
            csync();  // (2) every CTA's partial vector present everywhere
            if (p.move != MOVE_PRUNE) {
                for (int k = tid; k < pe16; k += SW_THREADS) {
                    double sacc = parts[k];
#pragma unroll
                    for (int r = 1; r < SW_MAX_R; ++r)
                        if (r < R) sacc += parts[r * parts_stride + k];
                    Wv[k] = sacc;
                }
                __syncthreads();
            }
            PHASE_MARK(6);

            // ---- phase 4 (rank 0): 2x2 capacitance matrix, proposed log-MLL, MH decision (bark_sampler.py:257-264)
            if (cr == 0) {
                double pvv = 0.0, pvw = 0.0;
                for (int k = tid; k < pe16; k += SW_THREADS) {
                    pvv = fma(vd[k], Wv[k], pvv);
                    pvw = fma(vd[k], w_s[k], pvw);
                }
                block_sum2(pvv, pvw, red);
                if (tid == 0) {
                    const double vWv = pvv, vw = pvw;
                    const double dWd = Wd[a] - Wd[b];
                    const double dWv = Wv[a] - Wv[b];
                    const double dw = w_s[a] - w_s[b];
                    const double M00 = dWd, M01 = 1.0 + dWv, M11 = -n_u + vWv;
                    const double det = M00 * M11 - M01 * M01;              // < 0 for an SPD B'
                    Decision dd;
                    dd.new_ldt = cur_ldt + log(-det);
                    const double bq = cur_q + 2.0 * eta * dw + eta * eta * dWd;  // b'^T Binv b'
                    const double Ur0 = dw + eta * dWd;                           // d^T r,  r = Binv b'
                    const double Ur1 = vw + eta * dWv;                           // v^T r
                    dd.new_q = bq - (M11 * Ur0 * Ur0 - 2.0 * M01 * Ur0 * Ur1 + M00 * Ur1 * Ur1) / det;
                    dd.new_mll = 0.5 * (-(yy - dd.new_q) / sig - nlogsig - dd.new_ldt);
                    const double log_alpha = p.lqp + (dd.new_mll - cur_mll);
                    dd.accept = (log(u[4]) <= fmin(log_alpha, 0.0)) ? 1 : 0;
                    // M^-1 = [[al, be],[be, ga]] and the coefficients of the w update
                    dd.al = M11 / det; dd.be = -M01 / det; dd.ga = M00 / det;
                    dd.cw_d = dd.al * Ur0 + dd.be * Ur1;
                    dd.cw_v = dd.be * Ur0 + dd.ga * Ur1;
                    dd.eta = eta;
                    dd.pad = 0;
#pragma unroll
                    for (int r = 0; r < SW_MAX_R; ++r)
                        if (r < R) *cluster.map_shared_rank(&ctl->dec, r) = dd;
                }
            }
            // meanwhile (or right after, with one CTA per chain) the generator rank prepares the next proposal and
            // publishes it SPECULATIVELY (its free column assumes a rejection, the common case): barrier (2b) then
            // makes both the decision and the next proposal visible, and a rejected proposal needs no barrier (3)
            if (cr == gen_rank && t + 1 < m) {
                generate(t + 1);
                publish(t + 1);
            }
            csync();  // (2b) decision (and the speculative next proposal) visible on every CTA
            const Decision dec = ctl->dec;
            accept = dec.accept != 0;
            new_q = dec.new_q; new_ldt = dec.new_ldt; new_mll = dec.new_mll;
            eta = dec.eta; al = dec.al; be = dec.be; ga = dec.ga; cw_d = dec.cw_d; cw_v = dec.cw_v;
            // column allocator / extent follow the decision on every CTA, then the next proposal gets its column
            if (accept && tid == 0) {
                if (p.move == MOVE_GROW) {
                    colused_s[a >> 5] |= (1u << (a & 31));
                    if (a + 1 > ctl->p_hi) ctl->p_hi = a + 1;
                } else if (p.move == MOVE_PRUNE) {
                    colused_s[b >> 5] &= ~(1u << (b & 31));
                }
            }
            if (accept) {  // the allocator changed: the next proposal gets its column again
                __syncthreads();
                if (cr == gen_rank && t + 1 < m) publish(t + 1);
            }
        }

        if (tid == 0) {
            if (trace_base && cr == 0) {
                trace_base[t * 3 + 0] = p.lqp;
                trace_base[t * 3 + 1] = new_mll;
                trace_base[t * 3 + 2] = accept ? 1.0 : 0.0;
            }
            {
                ++n_valid;
                ++n_valid_move[p.move];
                const unsigned long long ext = (unsigned long long)pe16;
                if (p.move != MOVE_PRUNE) blk_eval += ext * ext;
                if (accept) blk_upd += ext * ext;
                cols_scanned += ext;
            }
        }
        PHASE_MARK(7);

        if (accept) {
            __syncthreads();  // every thread of this CTA has finished reading w_s / Wd / Wv for the evaluation
            const int a = p.a, b = p.b;
            // w' = (w + eta Wd) - Wd cw_d - Wv cw_v           (full copy, identical on every CTA)
            for (int k = tid; k < pe16; k += SW_THREADS) w_s[k] = w_s[k] + eta * Wd[k] - Wd[k] * cw_d - Wv[k] * cw_v;
            // Binv' = Binv - [Wd Wv] M^-1 [Wd Wv]^T on the lower triangle: prefixes of this CTA's rows
            const bool paired_u = use_ring && pe16 <= 512 && pe16 + 4 <= slot_cols;
            if (paired_u) {
                const int qfirst = cr;
                const int nrows = (qfirst < pe16) ? (pe16 - 1 - qfirst) / R + 1 : 0;
                const int npairs = (nrows + 1) / 2;
                if (wid == SW_PROD_WARP) {
                    for (int base = 0; base < npairs; base += npl) {
                        const int j = base + lane;
                        const bool active = lane < npl && j < npairs;
                        const int iA = j, iB = nrows - 1 - j;
                        const int qA = qfirst + iA * R, qB = qfirst + iB * R;
                        const uint32_t use = ring_base + (uint32_t)j;
                        const int slot = (int)(use % (uint32_t)nslot);
                        const uint32_t par = (use / (uint32_t)nslot) & 1u;
                        bool ready = !active;
                        unsigned spins = 0;
                        while (!__all_sync(0xffffffffu, ready)) {
                            if (!ready) ready = mbar_try_wait(empty_bar + slot, par ^ 1u);
                            if (++spins > SW_WAIT_LIMIT) {
                                atomicOr(&sc->status, BARK_ST_TIMEOUT);
                                break;
                            }
                        }
                        if (active) {
                            const int lenA = (qA + 2) & ~1, lenB = (iB > iA) ? ((qB + 2) & ~1) : 0;
                            unsigned char* dst = ring + (size_t)slot * slot_bytes;
                            mbar_expect_tx(full_bar + slot, (uint32_t)(lenA + lenB) * 8u);
                            bulk_g2s(dst, cv.Binv + (size_t)qA * P, (uint32_t)lenA * 8u, full_bar + slot);
                            if (lenB) bulk_g2s(dst + (size_t)lenA * 8, cv.Binv + (size_t)qB * P, (uint32_t)lenB * 8u, full_bar + slot);
                        }
                    }
                } else {
                    const int jfirst = (int)((wid + SW_NCW - (ring_base % SW_NCW)) % SW_NCW);
                    for (int j = jfirst; j < npairs; j += SW_NCW) {
                        const int iA = j, iB = nrows - 1 - j;
                        const int qA = qfirst + iA * R, qB = qfirst + iB * R;
                        const int lenA = (qA + 2) & ~1;
                        const bool hasB = iB > iA;
                        const int lenB = hasB ? ((qB + 2) & ~1) : 0;
                        const double adA = al * Wd[qA] + be * Wv[qA], avA = be * Wd[qA] + ga * Wv[qA];
                        const double adB = hasB ? al * Wd[qB] + be * Wv[qB] : 0.0, avB = hasB ? be * Wd[qB] + ga * Wv[qB] : 0.0;
                        const uint32_t use = ring_base + (uint32_t)j;
                        const int slot = (int)(use % (uint32_t)nslot);
                        mbar_wait(full_bar + slot, (use / (uint32_t)nslot) & 1u, &sc->status);
                        const double* rowA = reinterpret_cast<const double*>(ring + (size_t)slot * slot_bytes);
                        const double* rowB = rowA + lenA;
                        double* gA = cv.Binv + (size_t)qA * P;
                        double* gB = cv.Binv + (size_t)qB * P;
                        // updated rows go straight from registers to global memory (L2): the slot is free as soon as
                        // it has been read, and the store traffic does not queue behind the bulk loads
                        for (int kk = lane * 2; kk < lenB || kk < lenA; kk += 64) {
                            const double2 dd = *reinterpret_cast<const double2*>(Wd + kk);
                            const double2 vv = *reinterpret_cast<const double2*>(Wv + kk);
                            if (kk < lenA) {
                                double2 x = *reinterpret_cast<const double2*>(rowA + kk);
                                x.x -= adA * dd.x + avA * vv.x;
                                x.y -= adA * dd.y + avA * vv.y;
                                __stcg(reinterpret_cast<double2*>(gA + kk), x);
                            }
                            if (kk < lenB) {
                                double2 x = *reinterpret_cast<const double2*>(rowB + kk);
                                x.x -= adB * dd.x + avB * vv.x;
                                x.y -= adB * dd.y + avB * vv.y;
                                __stcg(reinterpret_cast<double2*>(gB + kk), x);
                            }
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(empty_bar + slot);
                    }
                    fence_proxy_async();  // generic stores of the rows -> later bulk loads (async proxy)
                }
                ring_base += (uint32_t)npairs;
            } else {
                const int npanel = (pe16 + 511) / 512;
                for (int pc = 0; pc < npanel; ++pc) {
                    const int c0 = pc * 512, c1 = min(pe16, c0 + 512);
                    const int qfirst = c0 + ((cr - c0 % R) + R) % R;
                    const int nrows = (qfirst < pe16) ? (pe16 - 1 - qfirst) / R + 1 : 0;
                    if (use_ring && wid == SW_PROD_WARP) {
                        // npl lanes of the producer warp issue one row each per round (one thread alone caps the
                        // issue rate).  The whole warp stays convergent: a round issues when all of its slots are
                        // free, and npl <= nslot / 2 guarantees a round never waits on a slot it fills itself.
                        for (int base = 0; base < nrows; base += npl) {
                            const int i = base + lane;
                            const bool active = lane < npl && i < nrows;
                            const int q = qfirst + i * R;
                            const uint32_t use = ring_base + (uint32_t)i;
                            const int slot = (int)(use % (uint32_t)nslot);
                            const uint32_t par = (use / (uint32_t)nslot) & 1u;
                            bool ready = !active;
                            unsigned spins = 0;
                            while (!__all_sync(0xffffffffu, ready)) {
                                if (!ready) ready = mbar_try_wait(empty_bar + slot, par ^ 1u);
                                if (++spins > SW_WAIT_LIMIT) {
                                    atomicOr(&sc->status, BARK_ST_TIMEOUT);
                                    break;
                                }
                            }
                            if (active) {
                                const uint32_t bytes = (uint32_t)(((min(q + 1, c1) - c0) + 1) & ~1) * 8u;
                                mbar_expect_tx(full_bar + slot, bytes);
                                bulk_g2s(ring + (size_t)slot * slot_bytes, cv.Binv + (size_t)q * P + c0, bytes, full_bar + slot);
                            }
                        }
                    } else if (wid < SW_NCW) {
                        const int ifirst = use_ring ? (int)((wid + SW_NCW - (ring_base % SW_NCW)) % SW_NCW) : wid;
                        for (int i = ifirst; i < nrows; i += SW_NCW) {
                            const int q = qfirst + i * R;
                            const double ad = al * Wd[q] + be * Wv[q], av = be * Wd[q] + ga * Wv[q];
                            const int len2 = ((min(q + 1, c1) - c0) + 1) & ~1;
                            double* grow = cv.Binv + (size_t)q * P + c0;
                            const double* row = grow;
                            int slot = 0;
                            if (use_ring) {
                                const uint32_t use = ring_base + (uint32_t)i;
                                slot = (int)(use % (uint32_t)nslot);
                                mbar_wait(full_bar + slot, (use / (uint32_t)nslot) & 1u, &sc->status);
                                row = reinterpret_cast<const double*>(ring + (size_t)slot * slot_bytes);
                            }
                            // updated rows go straight from registers to global memory: the slot is free as soon as it
                            // has been read (a deferred release would dead-lock when a warp owns a single slot)
                            for (int kk = lane * 2; kk < len2; kk += 64) {
                                double2 x = use_ring ? *reinterpret_cast<const double2*>(row + kk) : ldcg2(row + kk);
                                const double2 dd = *reinterpret_cast<const double2*>(Wd + c0 + kk);
                                const double2 vv = *reinterpret_cast<const double2*>(Wv + c0 + kk);
                                x.x -= ad * dd.x + av * vv.x;
                                x.y -= ad * dd.y + av * vv.y;
                                __stcg(reinterpret_cast<double2*>(grow + kk), x);
                            }
                            if (use_ring) {
                                __syncwarp();
                                if (lane == 0) mbar_arrive(empty_bar + slot);
                            }
                        }
                        fence_proxy_async();  // generic stores of the rows -> later bulk loads (async proxy)
                    }
                    if (use_ring) ring_base += (uint32_t)nrows;
                }
            }
            // A' = A + v d^T + d v^T + n_u d d^T   (exact integers; atomics make the cross-CTA order irrelevant)
            for (int k = r0 + tid; k < r1; k += SW_THREADS) {
                const int vk = (int)vd[k];
                if (vk != 0) {
                    atomicAdd(cv.A + (size_t)k * P + a, vk);
                    atomicAdd(cv.A + (size_t)k * P + b, -vk);
                    atomicAdd(cv.A + (size_t)a * P + k, vk);
                    atomicAdd(cv.A + (size_t)b * P + k, -vk);
                }
            }
            if (cr == 0) {
                // leaf bitsets
                for (int w = tid; w < wd; w += SW_THREADS) {
                    const uint32_t ba = __ldcg(cv.bits + (size_t)a * wd + w), bb = __ldcg(cv.bits + (size_t)b * wd + w);
                    __stcg(cv.bits + (size_t)a * wd + w, (ba | upos[w]) & ~uneg[w]);
                    __stcg(cv.bits + (size_t)b * wd + w, (bb & ~upos[w]) | uneg[w]);
                }
                if (tid == 0) {
                    const int nuu = (int)n_u;  // corner term n_u d d^T
                    atomicAdd(cv.A + (size_t)a * P + a, nuu);
                    atomicAdd(cv.A + (size_t)b * P + b, nuu);
                    atomicAdd(cv.A + (size_t)a * P + b, -nuu);
                    atomicAdd(cv.A + (size_t)b * P + a, -nuu);
                    cv.b[a] += eta;
                    cv.b[b] -= eta;
                    // forest edit (tree_proposals.py:146-183) + column bookkeeping in global memory
                    uint16_t* cm = cv.colmap + (size_t)t * L;
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
                        cv.colused[a >> 5] = colused_s[a >> 5];  // already updated above
                    } else if (p.move == MOVE_PRUNE) {
                        forest.active[g0 + p.sl] = 0;
                        forest.active[g0 + p.sr] = 0;
                        forest.is_leaf[g0 + p.node] = 1;
                        cm[p.node] = (uint16_t)a;  // merged leaf keeps the left child's column
                        cm[p.sl] = NO_COL;
                        cm[p.sr] = NO_COL;
                        cv.colused[b >> 5] = colused_s[b >> 5];  // already updated above
                        cv.b[b] = 0.0;
                    } else {
                        forest.feature[g0 + p.node] = (uint32_t)p.feat;
                        forest.threshold[g0 + p.node] = p.thr;
                    }
                }
            }
            __syncthreads();  // Binv row stores of this CTA complete; colused_s / w_s readers above are done
            if (tid == 0) {
                if (p.move == MOVE_PRUNE) w_s[b] = 0.0;
                ctl->q = new_q; ctl->ldt = new_ldt; ctl->mll = new_mll;
                ++n_acc;
                ++n_acc_move[p.move];
            }
            if (p.move == MOVE_PRUNE) {
                // column b is now an empty leaf: make its row / column of Binv exactly (1/c) e_b
                for (int k = tid; k < pe16; k += SW_THREADS) {
                    if (k <= b) {
                        if (b % R == cr) __stcg(cv.Binv + (size_t)b * P + k, (k == b) ? 1.0 / c : 0.0);  // row b prefix
                    } else if (k % R == cr) {
                        __stcg(cv.Binv + (size_t)k * P + b, 0.0);                                        // column b below
                    }
                }
                fence_proxy_async();  // generic stores above -> later bulk loads of these rows
            }
        }
        PHASE_MARK(8);
        // (3) after an accepted proposal: peer's global-memory edits and the re-published next proposal visible,
        // exchanged vectors free for reuse.  A rejected one wrote nothing that the next iteration reads before its
        // own barrier (1), and nobody reads the exchanged vectors after (2b).
        if (accept) csync();
    }
    PHASE_MARK(9);
    __syncthreads();
#ifdef BARK_PHASE_TIMING
    if (tid == 0 && cr == 0)
        for (int i = 0; i < 12; ++i) sc->phase_cycles[i] += ph_acc[i];
#endif
    if (cr == 0) {
        for (int e = tid; e < P; e += SW_THREADS) cv.w[e] = w_s[e];
        if (tid == 0) {
            sc->q = ctl->q; sc->ldt = ctl->ldt; sc->mll = ctl->mll; sc->p_hi = ctl->p_hi;
            sc->counters[0] += (unsigned long long)m;
            sc->counters[1] += n_valid;
            sc->counters[2] += n_acc;
            sc->counters[5] += n_acc_move[0];
            sc->counters[6] += n_acc_move[1];
            sc->counters[7] += n_acc_move[2];
            sc->counters[8] += n_valid_move[0];
            sc->counters[9] += n_valid_move[1];
            sc->counters[10] += n_valid_move[2];
            sc->counters[11] += blk_eval;      // sum over matvec evaluations of extent^2
            sc->counters[12] += blk_upd;       // sum over accepted updates of extent^2
            sc->counters[13] += cols_scanned;  // leaf-bitset columns scanned for v = Z^T u
        }
    }
}

}  // namespace bark
