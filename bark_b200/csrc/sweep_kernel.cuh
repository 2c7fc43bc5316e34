// Tree sweep: the m tree MH steps of one sweep of one chain (bark_sampler.py:233-264) in leaf space.
//
// A thread-block CLUSTER of R CTAs (R = 1, 2 or 4; 512 threads each) owns one chain, so that 64 chains fill
// 128 of the 148 SMs.  B^-1 stays in global memory (L2/HBM) and is split by rows across the cluster:
//   phase 1  moved-point masks u+/u-, eta = u^T y, n_u           (redundant on every CTA; N bits)
//   phase 2  v = Z^T u by AND+POPC over the leaf bitsets         (columns split; halves exchanged through DSMEM)
//   phase 3  Wd = Binv d (two rows), Wv = Binv v                 (rows split; halves exchanged through DSMEM)
//   phase 4  2x2 capacitance matrix, proposed log-MLL, MH accept (redundant, bitwise identical on every CTA)
//   accept   symmetric rank-2 update of the CTA's rows of Binv, w; integer A / bitsets / forest edits
// Memory phases keep 16 independent 16-byte loads in flight per lane (two rows x eight column chunks).
// State written by a peer CTA is only read after a cluster barrier (release/acquire) and through L2 (.cg).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "mcmc_state.cuh"
#include "proposal_device.cuh"

namespace bark {
namespace cg = cooperative_groups;

constexpr int SW_THREADS = 512;

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
constexpr int SW_MAX_R = 4;

struct SweepCtl {  // small shared control block (kept identical on every CTA of the cluster)
    Prop prop;
    double q, ldt, mll;
    int p_hi;
};

__host__ __device__ inline size_t sweep_smem_bytes(int L, int d, int P, int wd) {
    size_t o = 0;
    o += align256(sizeof(SweepCtl));
    o += align256((size_t)L * 2);        // is_leaf, active
    o += align256((size_t)L * 4 * 6);    // feat,left,right,parent,depth,thr
    o += align256((size_t)d * 2 * 8);    // box
    o += align256((size_t)d * 4);        // ft
    o += align256((size_t)P * 8) * 3;    // vd, Wd, Wv
    o += align256((size_t)wd * 4) * 2;   // upos, uneg
    o += align256(64 * 8);               // red
    return o;
}

__device__ __forceinline__ double2 ldcg2(const double* p) {
    return __ldcg(reinterpret_cast<const double2*>(p));
}

__global__ void __launch_bounds__(SW_THREADS, 1)
sweep_trees_kernel(WsLayout lay, void* ws, bark_nodes_soa forest, bark_params prm, int64_t sweep_in_call,
                   int64_t n_sweeps_call, uint64_t seed, int64_t chain_offset, int64_t sweep_offset,
                   const double* __restrict__ tape, double* __restrict__ trace) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const int R = (int)cluster.num_blocks(), cr = (int)cluster.block_rank();

    const int P = (int)lay.P, L = (int)lay.L, m = (int)lay.m, n = (int)lay.n, wd = (int)lay.wd, npad = (int)lay.npad;
    const int d = (int)lay.d;
    unsigned char* sp = smem_raw;
    SweepCtl* ctl = (SweepCtl*)sp;           sp += align256(sizeof(SweepCtl));
    TreeSmem T;
    T.is_leaf = sp; T.active = sp + L;       sp += align256((size_t)L * 2);
    T.feat = (uint32_t*)sp; T.left = T.feat + L; T.right = T.left + L; T.parent = T.right + L; T.depth = T.parent + L;
    T.thr = (float*)(T.depth + L);           sp += align256((size_t)L * 4 * 6);
    double* box = (double*)sp;               sp += align256((size_t)d * 2 * 8);
    int32_t* ftc = (int32_t*)sp;             sp += align256((size_t)d * 4);
    double* vd = (double*)sp;                sp += align256((size_t)P * 8);
    double* Wd = (double*)sp;                sp += align256((size_t)P * 8);
    double* Wv = (double*)sp;                sp += align256((size_t)P * 8);
    uint32_t* upos = (uint32_t*)sp;          sp += align256((size_t)wd * 4);
    uint32_t* uneg = (uint32_t*)sp;          sp += align256((size_t)wd * 4);
    double* red = (double*)sp;

    // peers' copies of the exchanged vectors (distributed shared memory)
    double* vd_peer[SW_MAX_R];
    double* Wv_peer[SW_MAX_R];
#pragma unroll
    for (int r = 0; r < SW_MAX_R; ++r) {
        vd_peer[r] = (r < R) ? cluster.map_shared_rank(vd, r) : vd;
        Wv_peer[r] = (r < R) ? cluster.map_shared_rank(Wv, r) : Wv;
    }

    const int64_t chain = blockIdx.x / R;
    ChainView cv = chain_view(lay, ws, chain);
    SharedView sv = shared_view(lay, ws);
    ChainScalars* sc = cv.sc;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = SW_THREADS >> 5;

    // a dead chain (status set by an earlier launch) is skipped by the whole cluster
    if (__ldcg(&sc->status) & (BARK_ST_COL_OVERFLOW | BARK_ST_TREE_OVERFLOW | BARK_ST_HYPER_MODE)) return;

    const double sig = sc->sig, c = sc->c, yy = sc->yy;
    const double nlogsig = (double)n * log(sig);
    if (tid == 0) {
        ctl->q = sc->q; ctl->ldt = sc->ldt; ctl->mll = sc->mll; ctl->p_hi = sc->p_hi;
    }
    for (int e = tid; e < d; e += SW_THREADS) ftc[e] = sv.ft[e];
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
    for (int t = 0; t < m; ++t) {
        __syncthreads();
        PHASE_MARK(11);
        // ---- stage the tree and the root box
        const int64_t g0 = (chain * (int64_t)m + t) * L;
        for (int e = tid; e < L; e += SW_THREADS) {
            T.is_leaf[e] = __ldcg(forest.is_leaf + g0 + e);
            T.active[e] = __ldcg(forest.active + g0 + e);
            T.feat[e] = __ldcg(forest.feature + g0 + e);
            T.left[e] = __ldcg(forest.left + g0 + e);
            T.right[e] = __ldcg(forest.right + g0 + e);
            T.parent[e] = __ldcg(forest.parent + g0 + e);
            T.depth[e] = __ldcg(forest.depth + g0 + e);
            T.thr[e] = __ldcg(forest.threshold + g0 + e);
        }
        for (int e = tid; e < 2 * d; e += SW_THREADS) box[e] = sv.bounds[e];
        __syncthreads();
        PHASE_MARK(0);

        // ---- proposal (warp 0 of every CTA; identical inputs -> identical proposal)
        double u[6];
        if (tape) {
            for (int k = 0; k < TAPE_PER_TREE; ++k) u[k] = tape[tape_base + (size_t)t * TAPE_PER_TREE + k];
        } else {
            rng_uniforms(seed, g_chain, g_sweep, (uint32_t)t, TAPE_PER_TREE, u);
        }
        if (wid == 0) {
            Prop p = propose_tree_warp(T, L, box, ftc, d, cv.colmap + (size_t)t * L, cv.colused, P, prm, u, &sc->status);
            if (lane == 0) ctl->prop = p;
        }
        __syncthreads();
        PHASE_MARK(1);
        const Prop p = ctl->prop;
        const double cur_q = ctl->q, cur_ldt = ctl->ldt, cur_mll = ctl->mll;
        const int p_hi = ctl->p_hi;

        double new_q = cur_q, new_ldt = cur_ldt, new_mll = cur_mll;
        double eta = 0.0, n_u = 0.0, M00 = 0.0, M01 = 0.0, M11 = 0.0, det = -1.0, Ur0 = 0.0, Ur1 = 0.0;
        int pe64 = 0, r0 = 0, r1 = 0;
        bool accept = false;

        if (p.valid) {
            const int a = p.a, b = p.b;
            const int pe = max(p_hi, max(a, b) + 1);
            pe64 = min(P, (pe + 15) & ~15);  // used extent, rounded to 16 columns
            const int share = pe64 / R;      // multiple of 4
            r0 = cr * share;
            r1 = r0 + share;
            // ---- phase 1: moved-point masks u+ / u-, eta = u^T y, n_u = u^T u
            double eta_part = 0.0;
            int cnt_part = 0;
            const uint32_t* bits_a = cv.bits + (size_t)a * wd;
            const uint32_t* bits_b = cv.bits + (size_t)b * wd;
            const double* xf = sv.Xt + (size_t)p.feat * npad;
            const int ftype = ftc[p.feat];
            for (int i = tid; i < npad; i += SW_THREADS) {
                const int w = i >> 5;
                bool pos = false, neg = false;
                if (i < n) {
                    const bool in_b = (__ldcg(bits_b + w) >> lane) & 1u;
                    if (p.move == MOVE_GROW) {
                        if (in_b) pos = !goes_left(xf[i], p.thr, ftype);
                    } else if (p.move == MOVE_PRUNE) {
                        pos = in_b;
                    } else {  // change: b = left child's column, a = right child's column
                        const bool in_a = (__ldcg(bits_a + w) >> lane) & 1u;
                        if (in_a || in_b) {
                            const bool gl = goes_left(xf[i], p.thr, ftype);
                            pos = in_b && !gl;
                            neg = in_a && gl;
                        }
                    }
                    if (pos) { eta_part += sv.y[i]; ++cnt_part; }
                    if (neg) { eta_part -= sv.y[i]; ++cnt_part; }
                }
                const unsigned bp = __ballot_sync(0xffffffffu, pos), bn = __ballot_sync(0xffffffffu, neg);
                if (lane == 0) { upos[w] = bp; uneg[w] = bn; }
            }
            eta = block_sum(eta_part, red);
            n_u = block_sum((double)cnt_part, red);  // exact (integers < 2^53)
            PHASE_MARK(2);

            // ---- phase 2: v = Z^T u for this CTA's columns [r0, r1): four threads per column, words striped
            for (int base = r0; base < r1; base += SW_THREADS / 4) {
                const int q = base + (tid >> 2), part = tid & 3;
                int cnt = 0;
                if (q < r1 && q < pe) {
                    const uint32_t* bq = cv.bits + (size_t)q * wd;
                    for (int w0 = part * 4; w0 < wd; w0 += 16) {
                        uint32_t x[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) x[j] = (w0 + j < wd) ? __ldcg(bq + w0 + j) : 0u;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (w0 + j < wd) cnt += __popc(x[j] & upos[w0 + j]) - __popc(x[j] & uneg[w0 + j]);
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
            // Wd = Binv d: rows a and b (every CTA reads the two full rows)
            const double* row_a = cv.Binv + (size_t)a * P;
            const double* row_b = cv.Binv + (size_t)b * P;
            for (int k = tid; k < pe64; k += SW_THREADS) Wd[k] = __ldcg(row_a + k) - __ldcg(row_b + k);
            PHASE_MARK(3);
            cluster.sync();  // (1) all columns of v present everywhere
            PHASE_MARK(4);

            // ---- phase 3: Wv = Binv v for this CTA's rows (closed form for prune: e_b - c Binv[:,b])
            if (p.move == MOVE_PRUNE) {
                for (int k = tid; k < pe64; k += SW_THREADS) Wv[k] = ((k == b) ? 1.0 : 0.0) - c * __ldcg(row_b + k);
            } else {
                for (int q = r0 + 2 * wid; q < r1; q += 2 * nw) {
                    const double* rowA = cv.Binv + (size_t)q * P;
                    const double* rowB = rowA + P;
                    double accA = 0.0, accB = 0.0;
                    for (int kc = 0; kc < pe64; kc += 512) {
                        double2 xa[8], xb[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int k = kc + j * 64 + lane * 2;
                            if (k < pe64) { xa[j] = ldcg2(rowA + k); xb[j] = ldcg2(rowB + k); }
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int k = kc + j * 64 + lane * 2;
                            if (k < pe64) {
                                const double2 vv = *reinterpret_cast<const double2*>(vd + k);
                                accA = fma(xa[j].x, vv.x, accA); accA = fma(xa[j].y, vv.y, accA);
                                accB = fma(xb[j].x, vv.x, accB); accB = fma(xb[j].y, vv.y, accB);
                            }
                        }
                    }
                    accA = warp_sum(accA);
                    accB = warp_sum(accB);
                    if (lane == 0) {
#pragma unroll
                        for (int r = 0; r < SW_MAX_R; ++r)
                            if (r < R) { Wv_peer[r][q] = accA; Wv_peer[r][q + 1] = accB; }
                    }
                }
            }
            PHASE_MARK(5);
            cluster.sync();  // (2) all rows of Wv present everywhere
            PHASE_MARK(6);

            // ---- phase 4: 2x2 capacitance matrix and the proposed log-MLL (identical on every CTA)
            double pvv = 0.0, pvw = 0.0;
            for (int k = tid; k < pe64; k += SW_THREADS) {
                pvv = fma(vd[k], Wv[k], pvv);
                pvw = fma(vd[k], __ldcg(cv.w + k), pvw);
            }
            const double vWv = block_sum(pvv, red);
            const double vw = block_sum(pvw, red);
            const double dWd = Wd[a] - Wd[b];
            const double dWv = Wv[a] - Wv[b];
            const double dw = __ldcg(cv.w + a) - __ldcg(cv.w + b);
            M00 = dWd;
            M01 = 1.0 + dWv;
            M11 = -n_u + vWv;
            det = M00 * M11 - M01 * M01;                 // < 0 for an SPD B'
            new_ldt = cur_ldt + log(-det);
            const double bq = cur_q + 2.0 * eta * dw + eta * eta * dWd;  // b'^T Binv b'
            Ur0 = dw + eta * dWd;                                         // d^T r,  r = Binv b'
            Ur1 = vw + eta * dWv;                                         // v^T r
            new_q = bq - (M11 * Ur0 * Ur0 - 2.0 * M01 * Ur0 * Ur1 + M00 * Ur1 * Ur1) / det;
            new_mll = 0.5 * (-(yy - new_q) / sig - nlogsig - new_ldt);
        }

        // ---- MH accept (bark_sampler.py:257-264); every thread of every CTA evaluates the same scalars
        {
            const double u_acc = u[4];
            if (p.valid) {
                const double log_alpha = p.lqp + (new_mll - cur_mll);
                accept = log(u_acc) <= fmin(log_alpha, 0.0);
            }
            if (tid == 0) {
                if (trace_base && cr == 0) {
                    trace_base[t * 3 + 0] = p.valid ? p.lqp : -INFINITY;
                    trace_base[t * 3 + 1] = new_mll;
                    trace_base[t * 3 + 2] = accept ? 1.0 : 0.0;
                }
                if (p.valid) {
                    ++n_valid;
                    ++n_valid_move[p.move];
                    const unsigned long long ext = (unsigned long long)pe64;
                    if (p.move != MOVE_PRUNE) blk_eval += ext * ext;
                    if (accept) blk_upd += ext * ext;
                    cols_scanned += (unsigned long long)pe64;
                }
            }
        }

        PHASE_MARK(7);
        if (accept) {
            __syncthreads();  // every thread of this CTA has finished reading w / Binv for the evaluation
            const int a = p.a, b = p.b;
            // M^-1 = [[al, be],[be, ga]]
            const double al = M11 / det, be = -M01 / det, ga = M00 / det;
            const double cw_d = al * Ur0 + be * Ur1, cw_v = be * Ur0 + ga * Ur1;
            // w' = (w + eta Wd) - Wd cw_d - Wv cw_v           (this CTA's rows)
            for (int k = r0 + tid; k < r1; k += SW_THREADS)
                __stcg(cv.w + k, __ldcg(cv.w + k) + eta * Wd[k] - Wd[k] * cw_d - Wv[k] * cw_v);
            // Binv' = Binv - [Wd Wv] M^-1 [Wd Wv]^T              (this CTA's rows; two rows per warp pass)
            for (int q = r0 + 2 * wid; q < r1; q += 2 * nw) {
                double* rowA = cv.Binv + (size_t)q * P;
                double* rowB = rowA + P;
                const double ad = al * Wd[q] + be * Wv[q], av = be * Wd[q] + ga * Wv[q];
                const double bd = al * Wd[q + 1] + be * Wv[q + 1], bv = be * Wd[q + 1] + ga * Wv[q + 1];
                for (int kc = 0; kc < pe64; kc += 512) {
                    double2 xa[8], xb[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int k = kc + j * 64 + lane * 2;
                        if (k < pe64) { xa[j] = ldcg2(rowA + k); xb[j] = ldcg2(rowB + k); }
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int k = kc + j * 64 + lane * 2;
                        if (k < pe64) {
                            const double2 dd = *reinterpret_cast<const double2*>(Wd + k);
                            const double2 vv = *reinterpret_cast<const double2*>(Wv + k);
                            xa[j].x -= ad * dd.x + av * vv.x; xa[j].y -= ad * dd.y + av * vv.y;
                            xb[j].x -= bd * dd.x + bv * vv.x; xb[j].y -= bd * dd.y + bv * vv.y;
                            __stcg(reinterpret_cast<double2*>(rowA + k), xa[j]);
                            __stcg(reinterpret_cast<double2*>(rowB + k), xb[j]);
                        }
                    }
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
                    // forest edit (tree_proposals.py:146-183) + column bookkeeping
                    uint16_t* cm = cv.colmap + (size_t)t * L;
                    if (p.move == MOVE_GROW) {
                        const uint32_t dep = T.depth[p.node];
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
                        cv.colused[a >> 5] |= (1u << (a & 31));
                    } else if (p.move == MOVE_PRUNE) {
                        forest.active[g0 + p.sl] = 0;
                        forest.active[g0 + p.sr] = 0;
                        forest.is_leaf[g0 + p.node] = 1;
                        cm[p.node] = (uint16_t)a;  // merged leaf keeps the left child's column
                        cm[p.sl] = NO_COL;
                        cm[p.sr] = NO_COL;
                        cv.colused[b >> 5] &= ~(1u << (b & 31));
                        cv.b[b] = 0.0;
                    } else {
                        forest.feature[g0 + p.node] = (uint32_t)p.feat;
                        forest.threshold[g0 + p.node] = p.thr;
                    }
                }
            }
            if (tid == 0) {
                if (p.move == MOVE_GROW && a + 1 > ctl->p_hi) ctl->p_hi = a + 1;
                ctl->q = new_q; ctl->ldt = new_ldt; ctl->mll = new_mll;
                ++n_acc;
                ++n_acc_move[p.move];
            }
            if (p.move == MOVE_PRUNE) {
                // column b is now an empty leaf: make its row / column of Binv exactly (1/c) e_b
                __syncthreads();  // this CTA's rank-2 update of its rows is complete
                for (int k = r0 + tid; k < r1; k += SW_THREADS) __stcg(cv.Binv + (size_t)k * P + b, (k == b) ? 1.0 / c : 0.0);
                if (b >= r0 && b < r1) {
                    for (int k = tid; k < pe64; k += SW_THREADS) __stcg(cv.Binv + (size_t)b * P + k, (k == b) ? 1.0 / c : 0.0);
                    if (tid == 0) __stcg(cv.w + b, 0.0);
                }
            }
        }
        PHASE_MARK(8);
        if (p.valid) cluster.sync();  // (3) peer's global-memory edits visible; exchanged vectors free for reuse
    }
    PHASE_MARK(9);
    __syncthreads();
#ifdef BARK_PHASE_TIMING
    if (tid == 0 && cr == 0)
        for (int i = 0; i < 12; ++i) sc->phase_cycles[i] += ph_acc[i];
#endif
    if (tid == 0 && cr == 0) {
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

}  // namespace bark
