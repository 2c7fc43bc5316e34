// CTA-cooperative FP64 dense linear algebra for SPD matrices resident in global memory (L2/HBM).
//
// One building block -- a blocked symmetric sweep (block Gauss-Jordan on the lower triangle, pivot blocks
// of NB = 64) -- serves both uses on the BARK path:
//   * FULL = false : forward elimination only (a square-root-free block Cholesky, i.e. block LDL^T):
//                    yields log|W| and y^T W^-1 y in n^3/3 flops.  W is pure workspace.
//   * FULL = true  : sweep every pivot block over the whole lower triangle: W <- W^-1 (symmetric, both
//                    triangles written) and log|W| in n^3 flops.
// The trailing updates are NT GEMMs on 128x128 output tiles on the FP64 tensor pipe (mma.sync m8n8k4, 16 warps,
// 32x32 warp tiles), operands staged through shared memory; the 64x64 pivot blocks are inverted in shared
// memory by a scalar sweep that also produces the pivots for the log-determinant.
//
// Replaces np.linalg.inv / np.linalg.slogdet of src/bark/fitting/bark_sampler.py:160-161,269-270 and
// src/bark/tree_kernels/tree_gps.py:102.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace bark {
namespace la {

constexpr int THREADS = 512;
constexpr int TILE = 128;  // CTA output tile edge
constexpr int KB = 32;     // k-step staged in shared memory
constexpr int NB = 64;     // pivot block size

struct __align__(16) Smem {
    double As[TILE][KB + 4];  // row stride 36 doubles: conflict-free 8 x 4 DMMA fragment loads
    double Bs[TILE][KB + 4];
    double D[NB][NB + 1];
    double colv[NB];
    double rowv[NB];
    double piv[NB];
    double tv[NB];
    double yacc[TILE];
    double red[32];
};

enum { ACC_SUB = 0, ACC_SET = 1, ACC_ADD = 2 };

// C[0:mr, 0:nc] (-)= A[0:mr, 0:K] * B[0:nc, 0:K]^T.  Row-major, leading dimensions lda/ldb/ldc.
// lower_only: write only entries with column <= row (diagonal tiles of a symmetric update).
// C may alias A (same rows): all global reads of A complete before the first write of C.
// D (8x8) += A (8x4, row) * B (4x8, col) on the FP64 tensor pipe; lane l holds A[l/4][l%4], B[l%4][l/4],
// D[l/4][2(l%4)], D[l/4][2(l%4)+1]
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// C[0:mr, 0:nc] (-)= A[0:mr, 0:K] * B[0:nc, 0:K]^T.  Row-major, leading dimensions lda/ldb/ldc.
// lower_only: write only entries with column <= row (diagonal tiles of a symmetric update).
// C may alias A (same rows): all global reads of A complete before the first write of C.
// 16 warps in a 4 x 4 grid of 32 x 32 warp tiles, each 4 x 4 m8n8k4 DMMA tiles; operands staged through shared
// memory with a row stride of KB + 4 doubles (conflict-free fragment loads).  8 x 8 sub-tiles that are outside the
// ragged edge, or strictly above the diagonal of a lower_only tile, are skipped.
template <int MODE>
__device__ __forceinline__ void gemm_nt_tile(double* C, int64_t ldc, const double* A, int64_t lda, const double* B,
                                             int64_t ldb, int mr, int nc, int K, bool lower_only, Smem& s) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wr = warp & 3, wc = warp >> 2;  // consecutive warps (one per scheduler) differ in their row block
    const int lr = lane >> 2, lk = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    unsigned on = 0;  // bit a*4+b: sub-tile (a, b) of this warp is needed (warp-uniform)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int rlo = 32 * wr + 8 * a, clo = 32 * wc + 8 * b;
            if (rlo < mr && clo < nc && (!lower_only || clo <= rlo + 7)) on |= 1u << (a * 4 + b);
        }
    const int nrow_stage = max(mr, nc);

    for (int k0 = 0; k0 < K; k0 += KB) {
        // all global loads first (8 + 8 per thread in flight), then the shared-memory stores
        double va[TILE * KB / THREADS], vb[TILE * KB / THREADS];
#pragma unroll
        for (int it = 0; it < TILE * KB / THREADS; ++it) {
            const int e = tid + it * THREADS;
            const int r = e >> 5, k = e & 31;
            va[it] = 0.0;
            vb[it] = 0.0;
            if (k0 + k < K) {
                if (r < mr) va[it] = __ldcg(A + (int64_t)r * lda + k0 + k);
                if (r < nc) vb[it] = __ldcg(B + (int64_t)r * ldb + k0 + k);
            }
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < TILE * KB / THREADS; ++it) {
            const int e = tid + it * THREADS;
            const int r = e >> 5, k = e & 31;
            if (r < nrow_stage) {
                s.As[r][k] = va[it];
                s.Bs[r][k] = vb[it];
            }
        }
        __syncthreads();
        if (on) {
            const int kend = min(KB, (K - k0 + 3) & ~3);
#pragma unroll 2
            for (int kk = 0; kk < kend; kk += 4) {
                double a[4], b[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    a[x] = s.As[32 * wr + 8 * x + lr][kk + lk];
                    b[x] = s.Bs[32 * wc + 8 * x + lr][kk + lk];
                }
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y)
                        if (on & (1u << (x * 4 + y))) dmma_m8n8k4(acc[x][y][0], acc[x][y][1], a[x], b[y]);
            }
        }
    }
    // epilogue in two halves of 8 sub-tiles: all loads of C first, then the stores (C is only ever touched by its
    // owner thread, so the order is free)
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(C) | (uintptr_t)(ldc * 8)) & 15) == 0;
#pragma unroll
    for (int xh = 0; xh < 4; xh += 2) {
        double old[2][4][2];
        if (MODE != ACC_SET) {
#pragma unroll
            for (int xx = 0; xx < 2; ++xx)
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const int x = xh + xx;
                    old[xx][y][0] = old[xx][y][1] = 0.0;
                    if (!(on & (1u << (x * 4 + y)))) continue;
                    const int r = 32 * wr + 8 * x + lr, c = 32 * wc + 8 * y + 2 * lk;
                    if (r >= mr) continue;
                    const double* p = C + (int64_t)r * ldc + c;
                    const bool ok0 = c < nc && (!lower_only || c <= r);
                    const bool ok1 = c + 1 < nc && (!lower_only || c + 1 <= r);
                    if (ok0 && ok1 && vec_ok) {
                        const double2 o = __ldcg(reinterpret_cast<const double2*>(p));
                        old[xx][y][0] = o.x;
                        old[xx][y][1] = o.y;
                    } else {
                        if (ok0) old[xx][y][0] = __ldcg(p);
                        if (ok1) old[xx][y][1] = __ldcg(p + 1);
                    }
                }
        }
#pragma unroll
        for (int xx = 0; xx < 2; ++xx)
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const int x = xh + xx;
                if (!(on & (1u << (x * 4 + y)))) continue;
                const int r = 32 * wr + 8 * x + lr, c = 32 * wc + 8 * y + 2 * lk;
                if (r >= mr) continue;
                double* p = C + (int64_t)r * ldc + c;
                const bool ok0 = c < nc && (!lower_only || c <= r);
                const bool ok1 = c + 1 < nc && (!lower_only || c + 1 <= r);
                const double v0 = (MODE == ACC_SUB) ? old[xx][y][0] - acc[x][y][0] : (MODE == ACC_ADD) ? old[xx][y][0] + acc[x][y][0] : acc[x][y][0];
                const double v1 = (MODE == ACC_SUB) ? old[xx][y][1] - acc[x][y][1] : (MODE == ACC_ADD) ? old[xx][y][1] + acc[x][y][1] : acc[x][y][1];
                if (ok0 && ok1 && vec_ok) {
                    __stcg(reinterpret_cast<double2*>(p), make_double2(v0, v1));
                } else {
                    if (ok0) __stcg(p, v0);
                    if (ok1) __stcg(p + 1, v1);
                }
            }
    }
}

// Load the bs x bs diagonal block at G (lower triangle valid) into s.D, symmetrised, identity-padded to NB.
__device__ __forceinline__ void load_pivot_block(const double* G, int64_t ld, int bs, Smem& s) {
    double v[NB * NB / THREADS];  // all loads in flight before the first shared-memory store
#pragma unroll
    for (int it = 0; it < NB * NB / THREADS; ++it) {
        const int e = threadIdx.x + it * THREADS;
        const int r = e / NB, c = e % NB;
        if (r < bs && c < bs)
            v[it] = (c <= r) ? __ldcg(G + (int64_t)r * ld + c) : __ldcg(G + (int64_t)c * ld + r);
        else
            v[it] = (r == c) ? 1.0 : 0.0;
    }
#pragma unroll
    for (int it = 0; it < NB * NB / THREADS; ++it) {
        const int e = threadIdx.x + it * THREADS;
        s.D[e / NB][e % NB] = v[it];
    }
    __syncthreads();
}

// (A two-level variant -- sweep a 64 x 32 panel, Schur complement, sweep the 32 x 32 rest -- halves the serial
// cost but was measured 5-100x less accurate for the in-place inverse at bench scale, so the direct sweep stays.)
// Scalar symmetric sweep of s.D over all NB pivots: s.D <- -D^-1, s.piv[j] <- j-th pivot (successive
// Schur complements; their product is det D).  Returns false (uniformly) if a pivot is not positive.
// Register-resident: every thread owns 8 fixed entries (rows r0 + 8 i, column c) of the 64 x 64 block and only
// the next pivot row / column travel through shared memory (double-buffered), one barrier per pivot.
__device__ __forceinline__ bool sweep_pivot_block(Smem& s) {
    const int tid = threadIdx.x;
    const int c = tid & (NB - 1), r0 = tid >> 6;  // THREADS / NB = 8 row phases
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = s.D[r0 + 8 * i][c];
    // pivot row / column buffers: [2][NB] each, carved from s.As (free while the pivot block is processed)
    double* colb = &s.As[0][0];
    double* rowb = colb + 2 * NB;
    if (tid < NB) {
        colb[tid] = s.D[tid][0];
        rowb[tid] = s.D[0][tid];
    }
    __syncthreads();
    bool ok = true;
    const int par = (tid >> 5) & 1;  // which half of the columns this warp holds
    // The pivot loop is unrolled over blocks of 8 pivots so that the register index (j >> 3) of the pivot row is a
    // compile-time constant; the pivot column / row fix-ups are warp-uniform guards around the general FMA update
    // (scripts/pivot_bench.cu: 750 cycles per pivot against 867 for the select-only form, identical results; a
    // run-time register index costs local memory and 2400 cycles).
#pragma unroll
    for (int jb = 0; jb < 8; ++jb) {
#pragma unroll 1
        for (int jj = 0; jj < 8; ++jj) {
            const int j = 8 * jb + jj;
            const double* colv = colb + (j & 1) * NB;
            const double* rowv = rowb + (j & 1) * NB;
            double* coln = colb + ((j + 1) & 1) * NB;
            double* rown = rowb + ((j + 1) & 1) * NB;
            double cr[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) cr[i] = colv[r0 + 8 * i];
            const double p = colv[j];
            const double rowc = rowv[c];
            if (tid == 0) s.piv[j] = p;
            if (!(p > 0.0)) ok = false;
            const double ip = 1.0 / p;
            const double rc = rowc * ip;
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fma(-cr[i], rc, v[i]);  // general entry
            if (par == (j >> 5)) {                                       // pivot column
                const bool cj = (c == j);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = cj ? cr[i] * ip : v[i];
            }
            if (r0 == jj) v[jb] = (c == j) ? -ip : rc;                   // pivot row j = r0 + 8 jb
            if (par == ((j + 1) >> 5)) {
                if (c == j + 1) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) coln[r0 + 8 * i] = v[i];
                }
            }
            if (jj < 7) {
                if (r0 == jj + 1) rown[c] = v[jb];
            } else if (jb < 7) {
                if (r0 == 0) rown[c] = v[jb + 1];
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s.D[r0 + 8 * i][c] = v[i];
    __syncthreads();
    return ok;
}

// A "team" is the set of CTAs that cooperate on one matrix: a single CTA (SoloTeam) or a thread-block cluster
// (ClusterTeam, used for the exact refresh of B^-1).  Data handed between CTAs goes through global memory with
// .cg accesses and a team barrier (cluster barrier = release/acquire at cluster scope).
struct SoloTeam {
    __device__ __forceinline__ int rank() const { return 0; }
    __device__ __forceinline__ int size() const { return 1; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
};
struct ClusterTeam {
    cooperative_groups::cluster_group c;
    __device__ __forceinline__ int rank() const { return (int)c.block_rank(); }
    __device__ __forceinline__ int size() const { return (int)c.num_blocks(); }
    __device__ __forceinline__ void sync() const { c.sync(); }
};

// Blocked symmetric sweep of the n x n SPD matrix W (lower triangle valid on entry, leading dim ld).
//   CK, GK : scratch panels, n x NB doubles each (row-major, ld = NB);  DG : NB x NB scratch (pivot inverse).
//   yv     : optional (FULL == false) mutable copy of a right-hand side; on exit *quad = y^T W^-1 y.
// Returns log|W| and *quad (valid on team rank 0, every thread).  *status |= BARK_ST_NOT_SPD on a non-positive pivot.
#ifdef BARK_PHASE_TIMING
#define LA_MARK(i) do { long long t_ = clock64(); la_ph[i] += (unsigned long long)(t_ - la_t); la_t = t_; } while (0)
#else
#define LA_MARK(i) do { } while (0)
#endif

template <bool FULL, class Team>
__device__ double block_sweep(double* W, int64_t ld, int n, double* CK, double* GK, double* DG, double* yv,
                              double* quad, Smem& s, uint32_t* status, const Team& team) {
    const int tid = threadIdx.x;
    const int nblk = (n + NB - 1) / NB;
    const int trank = team.rank(), tsize = team.size();
    double logdet = 0.0, q = 0.0;
#ifdef BARK_PHASE_TIMING
    unsigned long long la_ph[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long la_t = clock64();
#endif
    for (int kb = 0; kb < nblk; ++kb) {
        const int k0 = kb * NB;
        const int bs = min(NB, n - k0);
        const int rlo = FULL ? 0 : k0 + bs;  // first row taking part in the update
        // ---- P0 (rank 0): invert the pivot block in shared memory, publish Dinv, eliminate the right-hand side
        if (trank == 0) {
            load_pivot_block(W + (int64_t)k0 * ld + k0, ld, bs, s);
            LA_MARK(10);
            const bool ok = sweep_pivot_block(s);  // s.D = -Dinv
            LA_MARK(11);
            if (!ok && tid == 0 && status) atomicOr(status, BARK_ST_NOT_SPD);
            {
                double lg = (tid < NB) ? log(s.piv[tid]) : 0.0;
                logdet += block_sum(lg, s.red);
            }
            for (int e = tid; e < NB * NB; e += THREADS) __stcg(DG + e, -s.D[e / NB][e % NB]);
            LA_MARK(12);
            if (!FULL && yv) {
                if (tid < NB) s.colv[tid] = (tid < bs) ? __ldcg(yv + k0 + tid) : 0.0;
                __syncthreads();
                if (tid < NB) {
                    double t = 0.0;
                    if (tid < bs)
                        for (int c = 0; c < bs; ++c) t -= s.D[tid][c] * s.colv[c];
                    s.tv[tid] = t;
                }
                __syncthreads();
                double part = (tid < bs) ? s.tv[tid] * s.colv[tid] : 0.0;
                q += block_sum(part, s.red);
            }
        }
        LA_MARK(0);
        if (rlo >= n && !FULL) break;
        team.sync();
        LA_MARK(1);
        if (!FULL && yv && trank != 0) {
            // the other CTAs of the team rebuild t = Dinv y_k from the published Dinv (same products, same order)
            if (tid < NB) s.colv[tid] = (tid < bs) ? __ldcg(yv + k0 + tid) : 0.0;
            __syncthreads();
            if (tid < NB) {
                double t = 0.0;
                if (tid < bs) {
#pragma unroll 16
                    for (int c = 0; c < bs; ++c) t += __ldcg(DG + tid * NB + c) * s.colv[c];
                }
                s.tv[tid] = t;
            }
            __syncthreads();
        }
        // ---- P1: pivot column panel CK (gathered) and GK = CK * Dinv, by 128-row tiles round-robin over the team
        {
            int rt = 0;
            for (int ti = rlo; ti < n; ti += TILE, ++rt) {
                if (rt % tsize != trank) continue;
                const int mr = min(TILE, n - ti);
                const bool do_y = !FULL && yv;
                if (do_y && tid < TILE) s.yacc[tid] = 0.0;
                if (do_y) __syncthreads();
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    // 8 gathered entries per thread in flight, then the stores; a warp covers half a panel row,
                    // so the row's dot product with t is a warp sum
                    double v[8];
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int e = tid + (half * 8 + it) * THREADS;
                        const int i = ti + e / NB, c = e % NB;
                        v[it] = 0.0;
                        if (e < mr * NB && c < bs) {
                            if (i >= k0 + bs)
                                v[it] = __ldcg(W + (int64_t)i * ld + k0 + c);
                            else if (i < k0)
                                v[it] = __ldcg(W + (int64_t)(k0 + c) * ld + i);
                        }
                    }
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int e = tid + (half * 8 + it) * THREADS;
                        const int i = ti + e / NB, c = e % NB;
                        if (e < mr * NB) __stcg(CK + (int64_t)i * NB + c, v[it]);
                        if (do_y) {
                            const double part = warp_sum(v[it] * s.tv[c]);
                            if ((tid & 31) == 0 && e < mr * NB) atomicAdd(&s.yacc[e / NB], part);
                        }
                    }
                }
                __syncthreads();
                if (do_y) {
                    // y_r -= G_r . y_k = C_r . t, by the owner of the row tile
                    for (int i = tid; i < mr; i += THREADS) __stcg(yv + ti + i, __ldcg(yv + ti + i) - s.yacc[i]);
                }
                gemm_nt_tile<ACC_SET>(GK + (int64_t)ti * NB, NB, CK + (int64_t)ti * NB, NB, DG, NB, mr, NB, bs, false, s);
            }
        }
        LA_MARK(2);
        team.sync();
        LA_MARK(3);
        // ---- P2: trailing update on the lower triangle, W_ij -= G_i C_j^T, tiles round-robin over the team
        {
            int idx = 0;
            for (int ti = rlo; ti < n; ti += TILE) {
                const int mr = min(TILE, n - ti);
                for (int tj = rlo; tj <= ti; tj += TILE, ++idx) {
                    if (idx % tsize != trank) continue;
                    const int nc = min(TILE, n - tj);
                    gemm_nt_tile<ACC_SUB>(W + (int64_t)ti * ld + tj, ld, GK + (int64_t)ti * NB, NB,
                                          CK + (int64_t)tj * NB, NB, mr, nc, bs, ti == tj, s);
                }
            }
        }
        LA_MARK(4);
        if (FULL) {
            team.sync();
            LA_MARK(5);
            // ---- P3: write the swept pivot column/row back: W_ik = G_i (below), W_kj = G_j^T (above), W_kk = -Dinv
            int rt = 0;
            for (int ti = 0; ti < n; ti += TILE, ++rt) {
                if (rt % tsize != trank) continue;
                const int mr = min(TILE, n - ti);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    double v[8];
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int e = tid + (half * 8 + it) * THREADS;
                        const int i = ti + e / NB, c = e % NB;
                        v[it] = 0.0;
                        if (e < mr * NB && c < bs) {
                            if (i >= k0 + bs || i < k0)
                                v[it] = __ldcg(GK + (int64_t)i * NB + c);
                            else if (c <= i - k0)
                                v[it] = -__ldcg(DG + (i - k0) * NB + c);
                        }
                    }
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int e = tid + (half * 8 + it) * THREADS;
                        const int i = ti + e / NB, c = e % NB;
                        if (e >= mr * NB || c >= bs) continue;
                        if (i >= k0 + bs)
                            __stcg(W + (int64_t)i * ld + k0 + c, v[it]);
                        else if (i < k0)
                            __stcg(W + (int64_t)(k0 + c) * ld + i, v[it]);
                        else if (c <= i - k0)
                            __stcg(W + (int64_t)i * ld + k0 + c, v[it]);
                    }
                }
            }
        }
        LA_MARK(6);
        team.sync();
        LA_MARK(7);
    }
    if (FULL) {
        // W currently holds -W^-1 on the lower triangle: negate and mirror (32x32 tiles through shared memory).
        double(*T)[33] = reinterpret_cast<double(*)[33]>(&s.As[0][0]);  // 32 x 33 tile
        const int tx = tid & 31, ty = tid >> 5;                            // 32 x 16
        const int nt = (n + 31) / 32;
        int idx = 0;
        for (int bi = 0; bi < nt; ++bi) {
            for (int bj = 0; bj <= bi; ++bj, ++idx) {
                if (idx % tsize != trank) continue;
                __syncthreads();
                for (int r = ty; r < 32; r += 16) {
                    const int gi = bi * 32 + r, gj = bj * 32 + tx;
                    double v = 0.0;
                    if (gi < n && gj < n && gj <= gi) {
                        v = -__ldcg(W + (int64_t)gi * ld + gj);
                        __stcg(W + (int64_t)gi * ld + gj, v);
                    }
                    T[r][tx] = v;
                }
                __syncthreads();
                for (int r = ty; r < 32; r += 16) {
                    const int gi = bj * 32 + r, gj = bi * 32 + tx;  // row in block bj, column in block bi
                    if (gi < n && gj < n && gj > gi) __stcg(W + (int64_t)gi * ld + gj, T[tx][r]);
                }
            }
        }
        team.sync();
    }
    LA_MARK(8);
#ifdef BARK_PHASE_TIMING
    if (tid == 0 && trank == 0 && (blockIdx.x / tsize) == 0)
        printf("block_sweep full=%d n=%d team=%d: P0 %llu s %llu P1 %llu s %llu P2 %llu s %llu P3 %llu s %llu mirror %llu | P0: load %llu sweep %llu logdet+DG %llu rest %llu\n", (int)FULL,
               n, tsize, la_ph[0], la_ph[1], la_ph[2], la_ph[3], la_ph[4], la_ph[5], la_ph[6], la_ph[7], la_ph[8], la_ph[10], la_ph[11], la_ph[12], la_ph[0]);
#endif
    if (quad) *quad = q;
    return logdet;
}


// One Newton-Schulz step on an approximate inverse X of the SPD matrix B = c I + A (A integer, lower triangle valid):
//     X <- X + (I - X B) X
// squares the residual |I - X B| (the block Gauss-Jordan inverse leaves ~cond(B)^2 eps, which is no longer small
// against the 1e-9 parity bar once cond(B) reaches ~1e4, i.e. for very small noise).  X: n x n, both triangles
// valid on entry and on exit; Bf, R: n x n scratch (leading dimension ld each).  Tiles are shared by the team.
template <class Team>
__device__ void refine_inverse(double* X, double* Bf, double* R, const int32_t* A, double c, int64_t ld, int n, Smem& s,
                               const Team& team) {
    const int tid = threadIdx.x;
    const int trank = team.rank(), tsize = team.size();
    // Bf = c I + A (both triangles), R = I
    for (int r = trank; r < n; r += tsize)
        for (int k = tid; k < n; k += THREADS) {
            const int32_t a = (k <= r) ? __ldcg(A + (int64_t)r * ld + k) : __ldcg(A + (int64_t)k * ld + r);
            __stcg(Bf + (int64_t)r * ld + k, (double)a + (r == k ? c : 0.0));
            __stcg(R + (int64_t)r * ld + k, r == k ? 1.0 : 0.0);
        }
    team.sync();
    const int nt = (n + TILE - 1) / TILE;
    // R -= X Bf^T  (Bf symmetric)  ->  R = I - X B
    for (int idx = trank; idx < nt * nt; idx += tsize) {
        const int ti = (idx / nt) * TILE, tj = (idx % nt) * TILE;
        gemm_nt_tile<ACC_SUB>(R + (int64_t)ti * ld + tj, ld, X + (int64_t)ti * ld, ld, Bf + (int64_t)tj * ld, ld,
                              min(TILE, n - ti), min(TILE, n - tj), n, false, s);
    }
    team.sync();
    // Bf <- X (Bf is dead), then Bf += R X^T (X symmetric)  ->  Bf = X + (I - X B) X
    for (int r = trank; r < n; r += tsize)
        for (int k = tid; k < n; k += THREADS) __stcg(Bf + (int64_t)r * ld + k, __ldcg(X + (int64_t)r * ld + k));
    team.sync();
    for (int idx = trank; idx < nt * nt; idx += tsize) {
        const int ti = (idx / nt) * TILE, tj = (idx % nt) * TILE;
        gemm_nt_tile<ACC_ADD>(Bf + (int64_t)ti * ld + tj, ld, R + (int64_t)ti * ld, ld, X + (int64_t)tj * ld, ld,
                              min(TILE, n - ti), min(TILE, n - tj), n, false, s);
    }
    team.sync();
    // X <- symmetrised refined inverse
    for (int r = trank; r < n; r += tsize)
        for (int k = tid; k < n; k += THREADS)
            __stcg(X + (int64_t)r * ld + k, 0.5 * (__ldcg(Bf + (int64_t)r * ld + k) + __ldcg(Bf + (int64_t)k * ld + r)));
    team.sync();
}

// cond(B) <= (c + n_points) / c for B = c I + Z^T Z: above this bound the Gauss-Jordan inverse is refined
constexpr double REFINE_COND = 3000.0;

}  // namespace la
}  // namespace bark
