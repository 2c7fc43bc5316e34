// a8-a13: device-resident BARK MCMC in leaf space (see mcmc_state.cuh for the algebra).
//
//   bark_mcmc_init    chain state from (forest, noise, scale)          bark_sampler.py:147-162
//   bark_mcmc_sweeps  m tree MH steps + 1 noise/scale MH step / sweep  bark_sampler.py:216-284
//
// Every tree move is the rank-one change  Z' = Z + u d^T  of the leaf-indicator matrix
// (u in {-1,0,1}^n marks the training points that change leaf, d = e_a - e_b names the two columns):
//     grow   leaf (col p) -> L keeps p, R gets a free column f :  u = 1[R],          d = e_f  - e_p
//     prune  children (pL, pR) merge into pL                    :  u = 1[R],          d = e_pL - e_pR
//     change children (pL, pR) re-split                         :  u = 1[L->R]-1[R->L], d = e_pR - e_pL
// hence  B' = B + [d v] [[n_u,1],[1,0]] [d v]^T  with  v = Z^T u  (AND+POPC on the leaf bitsets),
// n_u = u^T u, b' = b + (u^T y) d, and the proposal's log-MLL follows from ONE matvec  Binv v  (none for
// prune, where Binv v = e_b - c Binv[:,b]) plus O(P) dot products (2x2 capacitance matrix).  On accept the
// state takes a symmetric rank-2 update.  One CTA owns one chain for the whole sweep.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "forest_device.cuh"
#include "linalg.cuh"
#include "mcmc_state.cuh"
#include "proposal_device.cuh"
#include "sweep_block.cuh"

namespace bark {

// =====================================================================================================
// setup: shared inputs into the workspace (X transposed to feature-major, y, bounds, feat_types)
// =====================================================================================================
__global__ void ws_setup_kernel(WsLayout lay, void* ws, const double* __restrict__ X, const double* __restrict__ y,
                                const double* __restrict__ bounds, const int32_t* __restrict__ ft) {
    unsigned char* base = (unsigned char*)ws;
    double* Xt = (double*)(base + lay.off_xt);
    double* yy = (double*)(base + lay.off_y);
    double* bb = (double*)(base + lay.off_bounds);
    int32_t* ff = (int32_t*)(base + lay.off_ft);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = tid; e < lay.d * lay.npad; e += nth) {
        const int64_t f = e / lay.npad, i = e % lay.npad;
        Xt[e] = (i < lay.n) ? X[i * lay.d + f] : 0.0;
    }
    for (int64_t e = tid; e < lay.npad; e += nth) yy[e] = (e < lay.n) ? y[e] : 0.0;
    for (int64_t e = tid; e < lay.d * 2; e += nth) bb[e] = bounds ? bounds[e] : 0.0;
    for (int64_t e = tid; e < lay.d; e += nth) ff[e] = ft[e];
}

// =====================================================================================================
// helpers shared by init / hyper kernels (512 threads)
// =====================================================================================================
// w = Binv b over [0, ph); returns q = b^T w (to all threads).  Deterministic.
__device__ double matvec_w_q(const double* __restrict__ Binv, int P, int ph, const double* __restrict__ b, double* w,
                             double* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int r = wid; r < ph; r += nw) {
        const double* row = Binv + (size_t)r * P;
        double acc = 0.0;
        for (int k = lane; k < ph; k += 32) acc = fma(row[k], b[k], acc);
        acc = warp_sum(acc);
        if (lane == 0) w[r] = acc;
    }
    __syncthreads();
    double part = 0.0;
    for (int k = threadIdx.x; k < ph; k += blockDim.x) part = fma(b[k], w[k], part);
    return block_sum(part, red);
}

__device__ __forceinline__ double mll_from(double yy, double q, double sig, double n, double ldt) {
    return 0.5 * (-(yy - q) / sig - n * log(sig) - ldt);
}

// Fill W[0:ph,0:ph] lower triangle with A + c I.
__device__ void fill_B_lower(double* W, const int32_t* __restrict__ A, int P, int ph, double c) {
    for (int64_t e = threadIdx.x; e < (int64_t)ph * ph; e += blockDim.x) {
        const int r = (int)(e / ph), k = (int)(e % ph);
        if (k <= r) W[(size_t)r * P + k] = (double)A[(size_t)r * P + k] + (r == k ? c : 0.0);
    }
}

// =====================================================================================================
// chain init: columns, bitsets, A, b, Binv, scalars.  One CTA (512 threads) per chain.
// =====================================================================================================
__global__ void __launch_bounds__(la::THREADS, 1)
chain_init_kernel(WsLayout lay, void* ws, bark_nodes_soa forest, const double* __restrict__ noise,
                  const double* __restrict__ scale, int skip_null) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    la::Smem& s = *reinterpret_cast<la::Smem*>(smem_raw);
    WalkNode* wn = reinterpret_cast<WalkNode*>(smem_raw + sizeof(la::Smem));  // [L]
    int* ftc = reinterpret_cast<int*>(wn + lay.L);                             // [d]
    int* tcount = ftc + lay.d;                                                 // [m + 1]

    const int64_t chain = blockIdx.x;
    ChainView cv = chain_view(lay, ws, chain);
    SharedView sv = shared_view(lay, ws);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = la::THREADS >> 5;
    const int P = (int)lay.P, L = (int)lay.L, m = (int)lay.m, n = (int)lay.n, wd = (int)lay.wd, npad = (int)lay.npad;
    const int64_t nb = chain * (int64_t)m * L;

    // ---- 1. one column per active leaf, numbered in (tree, slot) order
    for (int t = tid; t < m; t += la::THREADS) {
        int c = 0;
        for (int sl = 0; sl < L; ++sl) c += (forest.active[nb + (int64_t)t * L + sl] && forest.is_leaf[nb + (int64_t)t * L + sl]);
        // skip_null (forest.py:101-111, the acquisition model's kernel): a root-only tree gets no column at all
        if (skip_null && forest.is_leaf[nb + (int64_t)t * L]) c = 0;
        tcount[t + 1] = c;
    }
    for (int e = tid; e < lay.d; e += la::THREADS) ftc[e] = sv.ft[e];
    __syncthreads();
    if (tid == 0) {
        tcount[0] = 0;
        for (int t = 0; t < m; ++t) tcount[t + 1] += tcount[t];
    }
    __syncthreads();
    const int ptot = tcount[m];
    if (ptot > P) {
        if (tid == 0) {
            cv.sc->status = BARK_ST_COL_OVERFLOW;
            cv.sc->p_hi = 0;
        }
        return;
    }
    int n_null_part = 0;
    for (int t = tid; t < m; t += la::THREADS) {
        int c = tcount[t];
        const bool null_tree = skip_null && forest.is_leaf[nb + (int64_t)t * L];
        n_null_part += null_tree;
        for (int sl = 0; sl < L; ++sl) {
            const int64_t g = nb + (int64_t)t * L + sl;
            cv.colmap[t * L + sl] = (!null_tree && forest.active[g] && forest.is_leaf[g]) ? (uint16_t)(c++) : NO_COL;
        }
    }
    const int m_eff = max(1, m - (int)block_sum((double)n_null_part, s.red));  // trees that count (m unless skip_null)
    for (int w = tid; w < P / 32; w += la::THREADS) {
        const int lo = w * 32;
        cv.colused[w] = (ptot >= lo + 32) ? 0xFFFFFFFFu : (ptot <= lo ? 0u : ((1u << (ptot - lo)) - 1u));
    }
    for (int64_t e = tid; e < (int64_t)P * wd; e += la::THREADS) cv.bits[e] = 0u;
    __syncthreads();

    // ---- 2. leaf bitsets: walk every training point down every tree (tree staged in shared memory)
    for (int t = 0; t < m; ++t) {
        __syncthreads();
        for (int e = tid; e < L; e += la::THREADS) {
            const int64_t g = nb + (int64_t)t * L + e;
            wn[e] = make_walk_node(forest.is_leaf[g], forest.feature[g], forest.threshold[g], forest.left[g],
                                   forest.right[g]);
        }
        __syncthreads();
        for (int w = wid; w < wd; w += nw) {
            const int i = w * 32 + lane;
            unsigned col = 0xFFFFFFFFu;
            if (i < n) col = cv.colmap[t * L + walk_tree(wn, sv.Xt + i, npad, ftc, L)];
            const unsigned grp = __match_any_sync(0xffffffffu, col);
            if (i < n && col != NO_COL && lane == __ffs(grp) - 1) cv.bits[(size_t)col * wd + w] = grp;
        }
    }
    __syncthreads();

    // ---- 3. A = Z^T Z (AND + POPC), b = Z^T y
    for (int64_t e = tid; e < (int64_t)P * P; e += la::THREADS) {
        const int r = (int)(e / P), k = (int)(e % P);
        int cnt = 0;
        if (r < ptot && k < ptot) {
            const uint32_t* br = cv.bits + (size_t)r * wd;
            const uint32_t* bk = cv.bits + (size_t)k * wd;
            for (int w = 0; w < wd; ++w) cnt += __popc(br[w] & bk[w]);
        }
        cv.A[e] = cnt;
    }
    for (int p = wid; p < P; p += nw) {
        double acc = 0.0;
        if (p < ptot)
            for (int w = lane; w < wd; w += 32) {
                uint32_t bw = cv.bits[(size_t)p * wd + w];
                while (bw) {
                    const int bit = __ffs(bw) - 1;
                    bw &= bw - 1;
                    acc += sv.y[w * 32 + bit];
                }
            }
        acc = warp_sum(acc);
        if (lane == 0) {
            cv.b[p] = acc;
            cv.w[p] = 0.0;
        }
    }
    double part = 0.0;
    for (int i = tid; i < n; i += la::THREADS) part = fma(sv.y[i], sv.y[i], part);
    const double yy = block_sum(part, s.red);

    // ---- 4. Binv = (c I + A)^-1, ldt, q, mll
    const double sig = noise[chain] + 1e-6;
    const double c = sig * (double)m_eff / scale[chain];
    for (int64_t e = tid; e < (int64_t)P * P; e += la::THREADS) {
        const int r = (int)(e / P), k = (int)(e % P);
        cv.Binv[e] = (r == k) ? 1.0 / c : 0.0;
    }
    __syncthreads();
    fill_B_lower(cv.Binv, cv.A, P, ptot, c);
    __syncthreads();
    unsigned* stp = &cv.sc->status;
    if (tid == 0) *stp = 0;
    __syncthreads();
    const double logdet = la::block_sweep<true>(cv.Binv, P, ptot, cv.CK, cv.GK, cv.DG, nullptr, nullptr, s, stp, la::SoloTeam());
    if (((double)n + c) / c > la::REFINE_COND) la::refine_inverse(cv.Binv, cv.Wk, cv.S2, cv.A, c, P, ptot, s, la::SoloTeam());
    double q = matvec_w_q(cv.Binv, P, ptot, cv.b, cv.w, s.red);
    for (int pass = 0; pass < 2; ++pass) {  // iterative refinement of w against the exact B = c I + A (see hyper_refresh_kernel)
        __syncthreads();
        for (int r = wid; r < ptot; r += nw) {
            double acc = 0.0;
            for (int k = lane; k < ptot; k += 32) acc = fma((double)cv.A[(size_t)r * P + k], cv.w[k], acc);
            acc = warp_sum(acc);
            if (lane == 0) cv.yv[r] = cv.b[r] - fma(c, cv.w[r], acc);
        }
        __syncthreads();
        for (int r = wid; r < ptot; r += nw) {
            double acc = 0.0;
            for (int k = lane; k < ptot; k += 32) acc = fma(cv.Binv[(size_t)r * P + k], cv.yv[k], acc);
            acc = warp_sum(acc);
            if (lane == 0) cv.w[r] += acc;
        }
        __syncthreads();
    }
    {
        double part = 0.0;
        for (int k = tid; k < ptot; k += la::THREADS) part = fma(cv.b[k], cv.w[k], part);
        q = block_sum(part, s.red);
    }
    const double ldt = logdet - (double)ptot * log(c);
    if (tid == 0) {
        ChainScalars* sc = cv.sc;
        sc->noise = noise[chain];
        sc->scale = scale[chain];
        sc->sig = sig;
        sc->c = c;
        sc->res = yy - q;
        sc->ldt = ldt;
        sc->yy = yy;
        sc->mll = mll_from(yy, q, sig, (double)n, ldt);
        for (int k = 0; k < 16; ++k) sc->counters[k] = 0ull;
        for (int k = 0; k < 12; ++k) sc->phase_cycles[k] = 0ull;
        sc->p_hi = ptot;
        sc->hyper_accept = 0;
    }
}

// =====================================================================================================
// hyper step: noise/scale MH with a full re-evaluation (bark_sampler.py:266-282).
//   hyper_eval_kernel    one cluster per chain: proposal, B' = c'I + A, forward block sweep (tiles spread over
//                        the cluster) -> MLL', MH decision on rank 0
//   hyper_refresh_kernel one 8-CTA cluster per chain, only for accepted chains: exact B'^-1 (full block sweep
//                        with tiles spread over the cluster), w = B'^-1 b, q, ldt, mll  (the reference
//                        refreshes K^-1 at the same point, :276-282)
// =====================================================================================================
constexpr int HYPER_CLUSTER = 8;
constexpr int HYPER_REFRESH_EVERY = 16;  // sweeps between forced exact refreshes (0 = only on accept, as the reference)

// Forced exact refresh of B^-1 (and with it w, the residual and log|B|) every `refresh_every` sweeps, staggered over
// the chains; more often for an ill-conditioned B = c I + Z^T Z (cond <= (c + n) / c), whose rank-2 updates drift
// faster: the period shrinks in proportion once the bound passes REFRESH_COND, down to every sweep.  The sensitive
// half of the log-MLL, the quadratic form, does not wait for it: hyper_eval_kernel refines w against the exact B every
// sweep.  Measured at BASELINE config 4 (posterior noise ~0.004, cond ~2500): running vs from-scratch log-MLL
// 1.7e-10 worst / 1.5e-11 median over 64 chains (scripts/diag_running_vs_scratch.py).
constexpr double REFRESH_COND = 2000.0;
__device__ __forceinline__ bool refresh_due(int refresh_every, double n, double c, int64_t tick) {
    if (refresh_every <= 0) return false;
    const double cond = (n + c) / c;
    int every = refresh_every;
    if (cond > REFRESH_COND) every = max(1, (int)((double)refresh_every * REFRESH_COND / cond));
    return (tick % every) == 0;
}

__global__ void __launch_bounds__(la::THREADS, 1)
hyper_eval_kernel(WsLayout lay, void* ws, bark_params prm, int64_t sweep_in_call, int64_t n_sweeps_call, uint64_t seed,
                  int64_t chain_offset, int64_t sweep_offset, const double* __restrict__ tape, double* __restrict__ trace,
                  int refresh_every) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    la::Smem& s = *reinterpret_cast<la::Smem*>(smem_raw);
    __shared__ HyperProp hp;
    __shared__ int s_phi;
    __shared__ double s_uacc;
    la::ClusterTeam team{cooperative_groups::this_cluster()};
    const int trank = team.rank(), tsize = team.size();
    const int64_t chain = blockIdx.x / tsize;
    ChainView cv = chain_view(lay, ws, chain);
    ChainScalars* sc = cv.sc;
    const int tid = threadIdx.x;
    const int P = (int)lay.P, m = (int)lay.m, n = (int)lay.n;
    // status bits that stop a chain are only ever set by the previous kernels or by the (identical) proposal
    // below, so every CTA of the cluster takes the same exit
    // and the CTAs reach no barrier before they have all either passed or left
    if (__ldcg(&sc->status) & (BARK_ST_COL_OVERFLOW | BARK_ST_TREE_OVERFLOW | BARK_ST_HYPER_MODE)) return;

    const size_t per = (size_t)(m * TAPE_PER_TREE + TAPE_PER_HYPER);
    if (tid == 0) {
        double zn, zs, ua;
        if (tape) {
            const double* tp = tape + ((size_t)(chain * n_sweeps_call + sweep_in_call)) * per + (size_t)m * TAPE_PER_TREE;
            zn = tp[0]; zs = tp[1]; ua = tp[2];
        } else {
            double u[6];
            rng_uniforms(seed, (uint32_t)(chain_offset + chain), (uint32_t)(sweep_offset + sweep_in_call), (uint32_t)m, 5, u);
            zn = std_normal_from(u[0], u[1]);
            zs = std_normal_from(u[2], u[3]);
            ua = u[4];
        }
        hp = propose_noise_scale(sc->noise, sc->scale, prm, zn, zs);
        s_uacc = ua;
        // tighten the used extent: highest allocated column + 1
        int hi = 0;
        for (int w = P / 32 - 1; w >= 0; --w)
            if (cv.colused[w]) { hi = w * 32 + 32 - __clz(cv.colused[w]); break; }
        s_phi = hi;
        if (trank == 0) {
            sc->hyper_accept = 0;
            if (hp.status) atomicOr(&sc->status, hp.status);
        }
    }
    __syncthreads();
    if (hp.status) return;
    const int ph = s_phi;
    const double sig2 = hp.noise + 1e-6;
    const double c2 = sig2 * (double)m / hp.scale;
    const double yy = sc->yy;

    // ---- once per sweep: iterative refinement of w = B^-1 b against the EXACT B = c I + A (A: integers), then the
    // residual y^T y - b^T w and the running log-MLL from it.  The quadratic form is the sensitive half of the log-MLL
    // (amplified by 1 / sig); carried through hundreds of rank-2 updates of an inverse that itself drifts it was off
    // by up to 8e-9 relative at cond(B) ~2500 (scripts/diag_running_vs_scratch.py), while two P^2 passes with the
    // drifted inverse as a preconditioner bring it back to ~cond eps.  (log|B| is updated by log|det M| per accepted
    // proposal, which is benign, and re-anchored by the exact refresh.)
    {
        const int lane = tid & 31, wid = tid >> 5, nw = la::THREADS >> 5;
        const double c_cur = sc->c;
        for (int r = trank * nw + wid; r < ph; r += tsize * nw) {  // yv = b - B w
            const int32_t* arow = cv.A + (size_t)r * P;
            double acc = 0.0;
            for (int k = lane; k < ph; k += 32) acc = fma((double)__ldcg(arow + k), __ldcg(cv.w + k), acc);
            acc = warp_sum(acc);
            if (lane == 0) __stcg(cv.yv + r, cv.b[r] - fma(c_cur, __ldcg(cv.w + r), acc));
        }
        team.sync();
        for (int r = trank * nw + wid; r < ph; r += tsize * nw) {  // w += Binv yv (lower triangle kept by the sweep)
            double acc = 0.0;
            for (int k = lane; k < ph; k += 32) {
                const double e = (k <= r) ? __ldcg(cv.Binv + (size_t)r * P + k) : __ldcg(cv.Binv + (size_t)k * P + r);
                acc = fma(e, __ldcg(cv.yv + k), acc);
            }
            acc = warp_sum(acc);
            if (lane == 0) __stcg(cv.w + r, __ldcg(cv.w + r) + acc);
        }
        team.sync();
        if (trank == 0) {
            double part = 0.0;
            for (int k = tid; k < ph; k += la::THREADS) part = fma(cv.b[k], __ldcg(cv.w + k), part);
            const double q = block_sum(part, s.red);
            if (tid == 0) {
                sc->res = yy - q;
                sc->mll = mll_from(yy, q, sc->sig, (double)n, sc->ldt);
            }
            __syncthreads();
        }
    }
    const double cur_mll = sc->mll;  // (used on rank 0 only)

    // B' lower triangle, rows split over the cluster
    for (int r = trank; r < ph; r += tsize)
        for (int k = tid; k <= r; k += la::THREADS)
            __stcg(cv.Wk + (size_t)r * P + k, (double)__ldcg(cv.A + (size_t)r * P + k) + (r == k ? c2 : 0.0));
    team.sync();  // every CTA is done with yv as the refinement residual
    if (trank == 0)
        for (int k = tid; k < ph; k += la::THREADS) __stcg(cv.yv + k, cv.b[k]);
    team.sync();
    double quad = 0.0;
    const double logdet = la::block_sweep<false>(cv.Wk, P, ph, cv.CK, cv.GK, cv.DG, cv.yv, &quad, s, &sc->status, team);
    if (trank != 0) return;  // log|B'| and the quadratic form live on rank 0
    const double ldt2 = logdet - (double)ph * log(c2);
    const double mll2 = mll_from(yy, quad, sig2, (double)n, ldt2);
    const double log_alpha = hp.lqp + (mll2 - cur_mll);
    const bool accept = log(s_uacc) <= fmin(log_alpha, 0.0);
    if (tid == 0) {
        if (trace) {
            double* tr = trace + (((size_t)(chain * n_sweeps_call + sweep_in_call)) * (size_t)(m + 1) + m) * 3;
            tr[0] = hp.lqp; tr[1] = mll2; tr[2] = accept ? 1.0 : 0.0;
        }
        sc->counters[3] += 1ull;
        sc->p_hi = ph;
        if (accept) {
            sc->prop_noise = hp.noise;
            sc->prop_scale = hp.scale;
            sc->hyper_accept = 1;
            sc->counters[4] += 1ull;
        } else if (refresh_due(refresh_every, (double)n, sc->c, sweep_offset + sweep_in_call + 1 + chain_offset + chain)) {
            // periodic exact refresh of the running state, staggered over the chains so that every launch of the
            // refresh kernel carries about chains / refresh_every clusters (bounds the drift of the rank-2 updates between accepted
            // noise/scale moves; the reference only refreshes on accept, bark_sampler.py:276-282)
            sc->prop_noise = sc->noise;
            sc->prop_scale = sc->scale;
            sc->hyper_accept = 2;
        }
    }
}

__global__ void __launch_bounds__(la::THREADS, 1)
hyper_refresh_kernel(WsLayout lay, void* ws) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    la::Smem& s = *reinterpret_cast<la::Smem*>(smem_raw);
    la::ClusterTeam team{cooperative_groups::this_cluster()};
    const int trank = team.rank(), tsize = team.size();
    const int64_t chain = blockIdx.x / tsize;
    ChainView cv = chain_view(lay, ws, chain);
    ChainScalars* sc = cv.sc;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = la::THREADS >> 5;
    const int P = (int)lay.P, m = (int)lay.m, n = (int)lay.n;
    if (__ldcg(&sc->hyper_accept) == 0) return;  // uniform over the cluster (written by the previous kernel)
    const int ph = sc->p_hi;
    const double noise = sc->prop_noise, scale = sc->prop_scale;
    const double sig2 = noise + 1e-6;
    const double c2 = sig2 * (double)m / scale;

    // B' lower triangle, rows split over the cluster
    for (int r = trank; r < ph; r += tsize)
        for (int k = tid; k <= r; k += la::THREADS)
            __stcg(cv.Binv + (size_t)r * P + k, (double)__ldcg(cv.A + (size_t)r * P + k) + (r == k ? c2 : 0.0));
    team.sync();
    const double logdet_f = la::block_sweep<true>(cv.Binv, P, ph, cv.CK, cv.GK, cv.DG, nullptr, nullptr, s, &sc->status, team);
    if (((double)n + c2) / c2 > la::REFINE_COND) la::refine_inverse(cv.Binv, cv.Wk, cv.S2, cv.A, c2, P, ph, s, team);
    for (int k = ph + trank * la::THREADS + tid; k < P; k += tsize * la::THREADS) __stcg(cv.Binv + (size_t)k * P + k, 1.0 / c2);
    // w = Binv b, rows split over the cluster (hyper_eval_kernel refines w against the exact B once per sweep)
    for (int r = trank * nw + wid; r < ph; r += tsize * nw) {
        const double* row = cv.Binv + (size_t)r * P;
        double acc = 0.0;
        for (int k = lane; k < ph; k += 32) acc = fma(__ldcg(row + k), cv.b[k], acc);
        acc = warp_sum(acc);
        if (lane == 0) __stcg(cv.w + r, acc);
    }
    team.sync();
    if (trank == 0) {
        double part = 0.0;
        for (int k = tid; k < ph; k += la::THREADS) part = fma(cv.b[k], __ldcg(cv.w + k), part);
        const double q = block_sum(part, s.red);
        if (tid == 0) {
            const double ldt = logdet_f - (double)ph * log(c2);
            sc->noise = noise; sc->scale = scale; sc->sig = sig2; sc->c = c2;
            sc->res = sc->yy - q; sc->ldt = ldt;
            sc->mll = mll_from(sc->yy, q, sig2, (double)n, ldt);
            sc->hyper_accept = 0;
            sc->counters[15] += 1ull;  // exact refreshes of this chain (accepted noise/scale moves + forced ones)
        }
    }
}

// =====================================================================================================
// read-out / export
// =====================================================================================================
__global__ void mcmc_read_kernel(WsLayout lay, const void* ws, double* noise, double* scale, double* mll,
                                 uint32_t* status, uint64_t* counters, int32_t* p_used) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= lay.chains) return;
    ChainView cv = chain_view(lay, const_cast<void*>(ws), c);
    const ChainScalars* sc = cv.sc;
    if (noise) noise[c] = sc->noise;
    if (scale) scale[c] = sc->scale;
    if (mll) mll[c] = sc->mll;
    if (status) status[c] = sc->status;
    if (counters) for (int k = 0; k < 16; ++k) counters[c * 16 + k] = sc->counters[k];
#ifdef BARK_PHASE_TIMING
    if (counters && c == 0) {  // debug build: chain 0's phase totals overwrite the counters of the LAST chain slot
        printf("phase_cycles chain0:");
        for (int k = 0; k < 12; ++k) printf(" %llu", sc->phase_cycles[k]);
        printf("\n");
    }
#endif
    if (p_used) {
        int cnt = 0;
        for (int w = 0; w < lay.P / 32; ++w) cnt += __popc(cv.colused[w]);
        p_used[c] = cnt;
    }
}

__global__ void mcmc_export_kernel(WsLayout lay, const void* ws, int64_t chain, int32_t* A, double* Binv,
                                   int32_t* colmap, uint32_t* bits) {
    ChainView cv = chain_view(lay, const_cast<void*>(ws), chain);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t PP = lay.P * lay.P;
    for (int64_t e = tid; e < PP; e += nth) {
        if (A) A[e] = cv.A[e];
        if (Binv) {  // the sampler keeps only the lower triangle current: export the symmetric matrix
            const int64_t r = e / lay.P, c = e % lay.P;
            Binv[e] = (c <= r) ? cv.Binv[e] : cv.Binv[c * lay.P + r];
        }
    }
    if (colmap)
        for (int64_t e = tid; e < lay.m * lay.L; e += nth) colmap[e] = cv.colmap[e] == NO_COL ? -1 : (int32_t)cv.colmap[e];
    if (bits)
        for (int64_t e = tid; e < lay.P * lay.wd; e += nth) bits[e] = cv.bits[e];
}

static int check_dims(const bark_mcmc_dims* dm) {
    if (!dm) return 0;
    if (dm->chains < 1 || dm->chains > 65535) return 0;
    if (dm->n < 1 || dm->n > (1 << 22)) return 0;
    if (dm->d < 1 || dm->d > 32767) return 0;
    if (dm->m < 1 || dm->m > 4096) return 0;
    if (dm->node_limit < 3 || dm->node_limit > 255) return 0;
    if (dm->p_cap < 64 || dm->p_cap % 64 != 0 || dm->p_cap > 8192) return 0;
    return 1;
}

// Launch geometry of the tree sweep: proposals per block (ks: 8, or a smaller power of two when the capacity is so
// large that eight P-vectors per matrix do not fit shared memory) and CTAs per chain (R: the largest power of two
// <= 16 with every chain's cluster resident at once; 16 is a non-portable cluster size, opted into at launch).
static size_t sweep_smem_budget() {
    int dev = 0, optin = 227 * 1024;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    return (size_t)std::min(optin, 227 * 1024);
}
static int sweep_block_size(const WsLayout& lay) {
    int ks = sb_pick_ks((int)lay.L, (int)lay.d, (int)lay.P, (int)lay.wd, sweep_smem_budget());
    const char* env = getenv("BARK_SWEEP_KB");  // testing: force a smaller block
    if (env && ks > 0) {
        const int k = atoi(env);
        if ((k == 1 || k == 2 || k == 4 || k == 8) && k <= ks) ks = k;
    }
    return ks;
}
static void sweep_launch_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int64_t chains, int R, size_t smem,
                                cudaStream_t st) {
    *cfg = cudaLaunchConfig_t{};
    cfg->gridDim = dim3((unsigned)(chains * R));
    cfg->blockDim = dim3(SB_THREADS);
    cfg->dynamicSmemBytes = smem;
    cfg->stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)R;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg->attrs = attr;
    cfg->numAttrs = 1;
}
using SweepKernel = void (*)(WsLayout, SbLayout, void*, bark_nodes_soa, bark_params, int64_t, int64_t, uint64_t, int64_t,
                             int64_t, const double*, double*);
static SweepKernel sweep_kernel_for(int ks) {
    switch (ks) {
        case 8: return sweep_block_kernel<8>;
        case 4: return sweep_block_kernel<4>;
        case 2: return sweep_block_kernel<2>;
        default: return sweep_block_kernel<1>;
    }
}
static int pick_cluster_size(int64_t chains, size_t smem, SweepKernel kern) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int R = 1;
    const char* env = getenv("BARK_SWEEP_CLUSTER");
    const int forced = env ? atoi(env) : 0;
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8 || forced == 16) {
        R = forced;
    } else {
        while (R < SB_MAX_R && chains * (R * 2) <= sms) R *= 2;
    }
    // every chain's cluster should be resident at once; the driver knows how many clusters of this size fit
    while (R > 1) {
        cudaLaunchConfig_t cfg;
        cudaLaunchAttribute attr[1];
        sweep_launch_config(&cfg, attr, chains, R, smem, nullptr);
        int nclusters = 0;
        const cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
        if (e == cudaSuccess && (nclusters >= chains || forced)) break;
        if (e != cudaSuccess) cudaGetLastError();  // e.g. cluster size not supported: clear and try a smaller one
        R /= 2;
    }
    return R;
}

static cudaError_t launch_hyper(cudaStream_t st, const WsLayout& lay, void* ws, const bark_params& prm, int64_t sidx,
                                int64_t n_sweeps, uint64_t seed, int64_t chain_offset, int64_t sweep_offset,
                                const double* tape, double* trace, int refresh_every, cudaEvent_t mid = nullptr) {
    {
        // forward sweep: as many CTAs per chain (1, 2, 4 or 8) as fit on the GPU in one wave (BARK_EVAL_CLUSTER caps it)
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        int rmax = 8;
        if (const char* e = getenv("BARK_EVAL_CLUSTER")) rmax = std::max(1, std::min(8, atoi(e)));
        // ... and every chain's cluster must be resident at once: 16 clusters of 8 do not all fit the GPCs of a B200 (a
        // second wave doubles the step), so the driver is asked (once per chain count)
        static int64_t cached_chains = -1;
        static int cached_rmax = 0, cached_R = 1;
        cudaLaunchConfig_t ecfg = {};
        cudaLaunchAttribute eattr[1];
        auto configure = [&](int r) {
            ecfg = cudaLaunchConfig_t{};
            ecfg.gridDim = dim3((unsigned)(lay.chains * r));
            ecfg.blockDim = dim3(la::THREADS);
            ecfg.dynamicSmemBytes = sizeof(la::Smem);
            ecfg.stream = st;
            eattr[0].id = cudaLaunchAttributeClusterDimension;
            eattr[0].val.clusterDim.x = (unsigned)r;
            eattr[0].val.clusterDim.y = 1;
            eattr[0].val.clusterDim.z = 1;
            ecfg.attrs = eattr;
            ecfg.numAttrs = 1;
        };
        if (cached_chains != lay.chains || cached_rmax != rmax) {
            int R0 = 1;
            while (R0 < rmax && lay.chains * (R0 * 2) <= sms) R0 *= 2;
            while (R0 > 1) {
                configure(R0);
                int nclusters = 0;
                const cudaError_t oe = cudaOccupancyMaxActiveClusters(&nclusters, hyper_eval_kernel, &ecfg);
                if (oe == cudaSuccess && nclusters >= lay.chains) break;
                if (oe != cudaSuccess) cudaGetLastError();
                R0 /= 2;
            }
            cached_chains = lay.chains;
            cached_rmax = rmax;
            cached_R = R0;
        }
        const int R = cached_R;
        configure(R);
        cudaError_t e = cudaLaunchKernelEx(&ecfg, hyper_eval_kernel, lay, ws, prm, sidx, n_sweeps, seed, chain_offset,
                                           sweep_offset, tape, trace, refresh_every);
        if (e != cudaSuccess) return e;
    }
    if (mid) {
        const cudaError_t e = cudaEventRecord(mid, st);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(lay.chains * HYPER_CLUSTER));
    cfg.blockDim = dim3(la::THREADS);
    cfg.dynamicSmemBytes = sizeof(la::Smem);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = HYPER_CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, hyper_refresh_kernel, lay, ws);
}

struct SweepGeom {
    int R, ks;
    size_t smem;
    SbLayout sl;
    SweepKernel kern;
};
// 0 on success; the geometry depends only on the dimensions, so it is computed once per C-ABI call
static cudaError_t sweep_geometry(const WsLayout& lay, SweepGeom* g) {
    g->ks = sweep_block_size(lay);
    if (g->ks <= 0) return cudaErrorInvalidValue;
    g->sl = sb_layout((int)lay.L, (int)lay.d, (int)lay.P, (int)lay.wd, g->ks);
    g->smem = g->sl.total;
    g->kern = sweep_kernel_for(g->ks);
    cudaError_t e = cudaFuncSetAttribute((const void*)g->kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute((const void*)g->kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    g->R = pick_cluster_size(lay.chains, g->smem, g->kern);
    return cudaSuccess;
}

static cudaError_t launch_sweep_trees(const SweepGeom& g, cudaStream_t st, const WsLayout& lay, void* ws,
                                      bark_nodes_soa forest, const bark_params& prm, int64_t sidx, int64_t n_sweeps,
                                      uint64_t seed, int64_t chain_offset, int64_t sweep_offset, const double* tape,
                                      double* trace) {
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    sweep_launch_config(&cfg, attr, lay.chains, g.R, g.smem, st);
    return cudaLaunchKernelEx(&cfg, g.kern, lay, g.sl, ws, forest, prm, sidx, n_sweeps, seed, chain_offset, sweep_offset,
                              tape, trace);
}

}  // namespace bark

using namespace bark;

extern "C" {

size_t bark_mcmc_workspace_bytes(const bark_mcmc_dims* dims) {
    if (!check_dims(dims)) return 0;
    return make_layout(*dims).total;
}

int64_t bark_mcmc_max_p_cap(const bark_mcmc_dims* dims) {
    if (!dims || dims->n < 1 || dims->d < 1 || dims->node_limit < 3) return 0;
    bark_mcmc_dims dm = *dims;
    int64_t best = 0;
    for (int64_t p = 64; p <= 8192; p += 64) {
        dm.p_cap = p;
        if (!check_dims(&dm)) continue;
        const WsLayout lay = make_layout(dm);
        if (sb_pick_ks((int)lay.L, (int)lay.d, (int)lay.P, (int)lay.wd, 227 * 1024) > 0) best = p;
    }
    return best;
}

int bark_mcmc_init(const bark_mcmc_dims* dims, void* workspace, bark_nodes_soa forest, const double* X,
                   const double* y, const double* bounds, const int32_t* feat_types, const double* noise,
                   const double* scale, void* stream) {
    return bark_mcmc_init_ex(dims, workspace, forest, X, y, bounds, feat_types, noise, scale, 0, stream);
}

int bark_mcmc_init_ex(const bark_mcmc_dims* dims, void* workspace, bark_nodes_soa forest, const double* X,
                      const double* y, const double* bounds, const int32_t* feat_types, const double* noise,
                      const double* scale, int32_t flags, void* stream) {
    BARK_CHECK_ARG(check_dims(dims), "bad dims (chains 1..65535, node_limit 3..255, p_cap multiple of 64 in 64..8192)");
    BARK_CHECK_ARG(workspace && X && y && feat_types && noise && scale && forest.is_leaf, "null pointer");
    const WsLayout lay = make_layout(*dims);
    cudaStream_t st = (cudaStream_t)stream;
    ws_setup_kernel<<<148, 256, 0, st>>>(lay, workspace, X, y, bounds, feat_types);
    BARK_LAUNCH_CHECK();
    const size_t smem = sizeof(la::Smem) + (size_t)lay.L * sizeof(WalkNode) + (size_t)(lay.d + lay.m + 2) * sizeof(int);
    BARK_CHECK_ARG(smem <= 227 * 1024, "d + m too large for the init kernel's shared memory");
    BARK_CUDA(cudaFuncSetAttribute(chain_init_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    chain_init_kernel<<<(unsigned)dims->chains, la::THREADS, smem, st>>>(lay, workspace, forest, noise, scale,
                                                                         (flags & BARK_INIT_SKIP_NULL) ? 1 : 0);
    BARK_LAUNCH_CHECK();
    return BARK_OK;
}

int bark_mcmc_sweeps(const bark_mcmc_dims* dims, void* workspace, bark_nodes_soa forest, const bark_params* params,
                     int64_t n_sweeps, uint64_t seed, int64_t chain_offset, int64_t sweep_offset, const double* tape,
                     double* trace, void* stream) {
    return bark_mcmc_sweeps_ex(dims, workspace, forest, params, n_sweeps, seed, chain_offset, sweep_offset, tape, trace, -1,
                               stream);
}

int bark_mcmc_sweeps_ex(const bark_mcmc_dims* dims, void* workspace, bark_nodes_soa forest, const bark_params* params,
                        int64_t n_sweeps, uint64_t seed, int64_t chain_offset, int64_t sweep_offset, const double* tape,
                        double* trace, int32_t refresh_every, void* stream) {
    BARK_CHECK_ARG(check_dims(dims), "bad dims");
    BARK_CHECK_ARG(workspace && params && forest.is_leaf, "null pointer");
    BARK_CHECK_ARG(n_sweeps >= 0, "n_sweeps < 0");
    BARK_CHECK_ARG(refresh_every >= -1 && refresh_every <= (1 << 20), "refresh_every out of range");
    const WsLayout lay = make_layout(*dims);
    cudaStream_t st = (cudaStream_t)stream;
    BARK_CHECK_ARG(sweep_block_size(lay) > 0, "p_cap / n / d too large for the sweep kernel's shared memory");
    BARK_CUDA(cudaFuncSetAttribute(hyper_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(la::Smem)));
    BARK_CUDA(cudaFuncSetAttribute(hyper_refresh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(la::Smem)));
    SweepGeom geom;
    BARK_CUDA(sweep_geometry(lay, &geom));
    // default policy: forced refresh every HYPER_REFRESH_EVERY sweeps; a replayed run (tape) refreshes only on an
    // accepted noise/scale move, exactly where the reference does (bark_sampler.py:276-282), unless told otherwise
    const int re = (refresh_every >= 0) ? refresh_every : (tape ? 0 : HYPER_REFRESH_EVERY);
    for (int64_t sidx = 0; sidx < n_sweeps; ++sidx) {
        BARK_CUDA(launch_sweep_trees(geom, st, lay, workspace, forest, *params, sidx, n_sweeps, seed, chain_offset,
                                     sweep_offset, tape, trace));
        BARK_CUDA(launch_hyper(st, lay, workspace, *params, sidx, n_sweeps, seed, chain_offset, sweep_offset, tape, trace, re));
    }
    BARK_LAUNCH_CHECK();
    return BARK_OK;
}

int bark_mcmc_sweeps_timed(const bark_mcmc_dims* dims, void* workspace, bark_nodes_soa forest, const bark_params* params,
                           int64_t n_sweeps, uint64_t seed, int64_t chain_offset, int64_t sweep_offset,
                           float* ms_trees_host, float* ms_hyper_host, void* stream) {
    float ms3[3] = {0.f, 0.f, 0.f};
    BARK_CHECK_ARG(ms_trees_host && ms_hyper_host, "null pointer");
    const int rc = bark_mcmc_sweeps_timed3(dims, workspace, forest, params, n_sweeps, seed, chain_offset, sweep_offset, ms3, stream);
    *ms_trees_host = ms3[0];
    *ms_hyper_host = ms3[1] + ms3[2];
    return rc;
}

int bark_mcmc_sweeps_timed3(const bark_mcmc_dims* dims, void* workspace, bark_nodes_soa forest, const bark_params* params,
                            int64_t n_sweeps, uint64_t seed, int64_t chain_offset, int64_t sweep_offset, float* ms3_host,
                            void* stream) {
    BARK_CHECK_ARG(check_dims(dims), "bad dims");
    BARK_CHECK_ARG(workspace && params && forest.is_leaf && ms3_host, "null pointer");
    BARK_CHECK_ARG(n_sweeps >= 1 && n_sweeps <= 4096, "n_sweeps out of range (1..4096)");
    const WsLayout lay = make_layout(*dims);
    cudaStream_t st = (cudaStream_t)stream;
    BARK_CHECK_ARG(sweep_block_size(lay) > 0, "p_cap / n / d too large for the sweep kernel's shared memory");
    BARK_CUDA(cudaFuncSetAttribute(hyper_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(la::Smem)));
    BARK_CUDA(cudaFuncSetAttribute(hyper_refresh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(la::Smem)));
    SweepGeom geom;
    BARK_CUDA(sweep_geometry(lay, &geom));
    std::vector<cudaEvent_t> ev((size_t)n_sweeps * 4);
    for (auto& e : ev) BARK_CUDA(cudaEventCreate(&e));
    for (int64_t sidx = 0; sidx < n_sweeps; ++sidx) {
        BARK_CUDA(cudaEventRecord(ev[sidx * 4 + 0], st));
        BARK_CUDA(launch_sweep_trees(geom, st, lay, workspace, forest, *params, sidx, n_sweeps, seed, chain_offset,
                                     sweep_offset, nullptr, nullptr));
        BARK_CUDA(cudaEventRecord(ev[sidx * 4 + 1], st));
        BARK_CUDA(launch_hyper(st, lay, workspace, *params, sidx, n_sweeps, seed, chain_offset, sweep_offset, nullptr, nullptr,
                               HYPER_REFRESH_EVERY, ev[sidx * 4 + 2]));
        BARK_CUDA(cudaEventRecord(ev[sidx * 4 + 3], st));
    }
    BARK_LAUNCH_CHECK();
    BARK_CUDA(cudaStreamSynchronize(st));
    float tt = 0.f, te = 0.f, tr = 0.f;
    for (int64_t sidx = 0; sidx < n_sweeps; ++sidx) {
        float a = 0.f, b = 0.f, c = 0.f;
        BARK_CUDA(cudaEventElapsedTime(&a, ev[sidx * 4 + 0], ev[sidx * 4 + 1]));
        BARK_CUDA(cudaEventElapsedTime(&b, ev[sidx * 4 + 1], ev[sidx * 4 + 2]));
        BARK_CUDA(cudaEventElapsedTime(&c, ev[sidx * 4 + 2], ev[sidx * 4 + 3]));
        tt += a; te += b; tr += c;
    }
    for (auto& e : ev) cudaEventDestroy(e);
    ms3_host[0] = tt;
    ms3_host[1] = te;
    ms3_host[2] = tr;
    return BARK_OK;
}

int bark_mcmc_read(const bark_mcmc_dims* dims, const void* workspace, double* noise, double* scale, double* mll,
                   uint32_t* status, uint64_t* counters, int32_t* p_used, void* stream) {
    BARK_CHECK_ARG(check_dims(dims) && workspace, "bad dims / null workspace");
    const WsLayout lay = make_layout(*dims);
    mcmc_read_kernel<<<(unsigned)ceil_div(dims->chains, 128), 128, 0, (cudaStream_t)stream>>>(
        lay, workspace, noise, scale, mll, status, counters, p_used);
    BARK_LAUNCH_CHECK();
    return BARK_OK;
}

int bark_mcmc_export(const bark_mcmc_dims* dims, const void* workspace, int64_t chain, int32_t* A, double* Binv,
                     int32_t* colmap, uint32_t* bits, void* stream) {
    BARK_CHECK_ARG(check_dims(dims) && workspace, "bad dims / null workspace");
    BARK_CHECK_ARG(chain >= 0 && chain < dims->chains, "chain out of range");
    const WsLayout lay = make_layout(*dims);
    mcmc_export_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(lay, workspace, chain, A, Binv, colmap, bits);
    BARK_LAUNCH_CHECK();
    return BARK_OK;
}

}  // extern "C"
