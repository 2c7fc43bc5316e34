// CTA-cooperative FP64 dense linear algebra for SPD matrices resident in global memory (L2/HBM).
//
// One building block -- a blocked symmetric sweep (block Gauss-Jordan on the lower triangle, pivot blocks
// of NB = 64) -- serves both uses on the BARK path:
//   * FULL = false : forward elimination only (a square-root-free block Cholesky, i.e. block LDL^T):
//                    yields log|W| and y^T W^-1 y in n^3/3 flops.  W is pure workspace.
//   * FULL = true  : sweep every pivot block over the whole lower triangle: W <- W^-1 (symmetric, both
//                    triangles written) and log|W| in n^3 flops.
// The trailing updates are NT GEMMs on 128x128 output tiles with an 8x4 register micro-tile per thread
// (512 threads), operands staged through shared memory; the 64x64 pivot blocks are inverted in shared
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
    double As[TILE][KB + 1];
    double Bs[TILE][KB + 1];
    double D[NB][NB + 1];
    double colv[NB];
    double rowv[NB];
    double piv[NB];
    double tv[NB];
    double red[32];
};

enum { ACC_SUB = 0, ACC_SET = 1 };

// C[0:mr, 0:nc] (-)= A[0:mr, 0:K] * B[0:nc, 0:K]^T.  Row-major, leading dimensions lda/ldb/ldc.
// lower_only: write only entries with column <= row (diagonal tiles of a symmetric update).
// C may alias A (same rows): all global reads of A complete before the first write of C.
// inner product step over one staged k-panel for the first J column blocks of this thread's micro-tile
template <int J>
__device__ __forceinline__ void gemm_panel(double (&acc)[8][4], const Smem& s, int ty, int tx) {
#pragma unroll 4
    for (int kk = 0; kk < KB; ++kk) {
        double a[8], b[J];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = s.As[ty * 8 + i][kk];
#pragma unroll
        for (int j = 0; j < J; ++j) b[j] = s.Bs[tx + 32 * j][kk];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < J; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
}

template <int MODE>
__device__ __forceinline__ void gemm_nt_tile(double* C, int64_t ldc, const double* A, int64_t lda, const double* B,
                                             int64_t ldb, int mr, int nc, int K, bool lower_only, Smem& s) {
    const int tid = threadIdx.x, ty = tid >> 5, tx = tid & 31;
    double acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    // column blocks (of 32) this warp really needs: ragged tiles and the upper part of diagonal tiles are skipped
    int jmax = (nc + 31) >> 5;
    if (lower_only) jmax = min(jmax, ((ty * 8 + 7) >> 5) + 1);
    if (ty * 8 >= mr) jmax = 0;
    const int nrow_stage = max(mr, nc);

    for (int k0 = 0; k0 < K; k0 += KB) {
        __syncthreads();
        for (int e = tid; e < nrow_stage * KB; e += THREADS) {
            const int r = e >> 5, k = e & 31;
            double va = 0.0, vb = 0.0;
            if (k0 + k < K) {
                if (r < mr) va = __ldcg(A + (int64_t)r * lda + k0 + k);
                if (r < nc) vb = __ldcg(B + (int64_t)r * ldb + k0 + k);
            }
            s.As[r][k] = va;
            s.Bs[r][k] = vb;
        }
        __syncthreads();
        switch (jmax) {  // warp-uniform
            case 4: gemm_panel<4>(acc, s, ty, tx); break;
            case 3: gemm_panel<3>(acc, s, ty, tx); break;
            case 2: gemm_panel<2>(acc, s, ty, tx); break;
            case 1: gemm_panel<1>(acc, s, ty, tx); break;
            default: break;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = ty * 8 + i;
        if (r < mr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = tx + 32 * j;
                if (j < jmax && c < nc && (!lower_only || c <= r)) {
                    double* p = C + (int64_t)r * ldc + c;
                    if (MODE == ACC_SUB)
                        __stcg(p, __ldcg(p) - acc[i][j]);
                    else
                        __stcg(p, acc[i][j]);
                }
            }
        }
    }
}

// Load the bs x bs diagonal block at G (lower triangle valid) into s.D, symmetrised, identity-padded to NB.
__device__ __forceinline__ void load_pivot_block(const double* G, int64_t ld, int bs, Smem& s) {
    for (int e = threadIdx.x; e < NB * NB; e += THREADS) {
        const int r = e / NB, c = e % NB;
        double v;
        if (r < bs && c < bs)
            v = (c <= r) ? __ldcg(G + (int64_t)r * ld + c) : __ldcg(G + (int64_t)c * ld + r);
        else
            v = (r == c) ? 1.0 : 0.0;
        s.D[r][c] = v;
    }
    __syncthreads();
}

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
    for (int j = 0; j < NB; ++j) {
        const double* colv = colb + (j & 1) * NB;
        const double* rowv = rowb + (j & 1) * NB;
        double* coln = colb + ((j + 1) & 1) * NB;
        double* rown = rowb + ((j + 1) & 1) * NB;
        // branch-free body: all shared-memory loads up front, selects instead of control flow
        double cr[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) cr[i] = colv[r0 + 8 * i];
        const double p = colv[j];
        const double rowc = rowv[c];
        if (tid == 0) s.piv[j] = p;
        if (!(p > 0.0)) ok = false;
        const double ip = 1.0 / p;
        const double rc = rowc * ip;
        const bool cj = (c == j);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = r0 + 8 * i;
            const double xg = fma(-cr[i], rc, v[i]);  // general entry
            const double xc = cr[i] * ip;             // pivot column
            double x = cj ? xc : xg;
            if (r == j) x = cj ? -ip : rc;            // pivot row
            v[i] = x;
            if (c == j + 1) coln[r] = x;
            if (r == j + 1) rown[c] = x;
        }
        __syncthreads();
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
    unsigned long long la_ph[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long la_t = clock64();
#endif
    for (int kb = 0; kb < nblk; ++kb) {
        const int k0 = kb * NB;
        const int bs = min(NB, n - k0);
        const int rlo = FULL ? 0 : k0 + bs;  // first row taking part in the update
        // ---- P0 (rank 0): invert the pivot block in shared memory, publish Dinv, eliminate the right-hand side
        if (trank == 0) {
            load_pivot_block(W + (int64_t)k0 * ld + k0, ld, bs, s);
            const bool ok = sweep_pivot_block(s);  // s.D = -Dinv
            if (!ok && tid == 0 && status) atomicOr(status, BARK_ST_NOT_SPD);
            {
                double lg = (tid < NB) ? log(s.piv[tid]) : 0.0;
                logdet += block_sum(lg, s.red);
            }
            for (int e = tid; e < NB * NB; e += THREADS) __stcg(DG + e, -s.D[e / NB][e % NB]);
            if (!FULL && yv) {
                if (tid < NB) {
                    double t = 0.0;
                    if (tid < bs)
                        for (int c = 0; c < bs; ++c) t -= s.D[tid][c] * __ldcg(yv + k0 + c);
                    s.tv[tid] = t;
                }
                __syncthreads();
                double part = (tid < bs) ? s.tv[tid] * __ldcg(yv + k0 + tid) : 0.0;
                q += block_sum(part, s.red);
            }
        }
        LA_MARK(0);
        if (rlo >= n && !FULL) break;
        team.sync();
        LA_MARK(1);
        if (!FULL && yv && trank != 0) {
            // the other CTAs of the team rebuild t = Dinv y_k from the published Dinv (same products, same order)
            if (tid < NB) {
                double t = 0.0;
                if (tid < bs)
                    for (int c = 0; c < bs; ++c) t += __ldcg(DG + tid * NB + c) * __ldcg(yv + k0 + c);
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
                for (int e = tid; e < mr * NB; e += THREADS) {
                    const int i = ti + e / NB, c = e % NB;
                    double v = 0.0;
                    if (c < bs) {
                        if (i >= k0 + bs)
                            v = __ldcg(W + (int64_t)i * ld + k0 + c);
                        else if (i < k0)
                            v = __ldcg(W + (int64_t)(k0 + c) * ld + i);
                    }
                    __stcg(CK + (int64_t)i * NB + c, v);
                }
                __syncthreads();
                gemm_nt_tile<ACC_SET>(GK + (int64_t)ti * NB, NB, CK + (int64_t)ti * NB, NB, DG, NB, mr, NB, bs, false, s);
                if (!FULL && yv) {
                    // y_r -= G_r . y_k = C_r . t, by the owner of the row tile
                    for (int i = ti + tid; i < ti + mr; i += THREADS) {
                        const double* ck = CK + (int64_t)i * NB;
                        double a = 0.0;
                        for (int c = 0; c < bs; ++c) a += __ldcg(ck + c) * s.tv[c];
                        __stcg(yv + i, __ldcg(yv + i) - a);
                    }
                }
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
                for (int e = tid; e < mr * NB; e += THREADS) {
                    const int i = ti + e / NB, c = e % NB;
                    if (c >= bs) continue;
                    if (i >= k0 + bs)
                        __stcg(W + (int64_t)i * ld + k0 + c, __ldcg(GK + (int64_t)i * NB + c));
                    else if (i < k0)
                        __stcg(W + (int64_t)(k0 + c) * ld + i, __ldcg(GK + (int64_t)i * NB + c));
                    else if (c <= i - k0)
                        __stcg(W + (int64_t)i * ld + k0 + c, -__ldcg(DG + (i - k0) * NB + c));
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
        printf("block_sweep full=%d n=%d team=%d: P0 %llu s %llu P1 %llu s %llu P2 %llu s %llu P3 %llu s %llu mirror %llu\n", (int)FULL,
               n, tsize, la_ph[0], la_ph[1], la_ph[2], la_ph[3], la_ph[4], la_ph[5], la_ph[6], la_ph[7], la_ph[8]);
#endif
    if (quad) *quad = q;
    return logdet;
}

}  // namespace la
}  // namespace bark
