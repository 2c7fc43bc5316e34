// Per-device workspace layout of the leaf-space ("P-space") BARK sampler state.
//
// Model algebra.  With Z (n x P) the one-hot leaf-indicator matrix of the forest (one
// column per active leaf, P = total leaves over the m trees), sig = noise + 1e-6, s = scale, c = sig*m/s:
//     K      = (s/m) Z Z^T + sig I                                  (bark_sampler.py:153-156)
//     B      = c I_P + Z^T Z,      A = Z^T Z (exact integers),      b = Z^T y
//     y^T K^-1 y = (y^T y - b^T B^-1 b) / sig
//     log|K|     = n log sig + log|I + A/c|  =: n log sig + ldt
//     mll        = 0.5 * ( -(yy - q)/sig - n log sig - ldt ),       q = b^T B^-1 b
// so a chain carries B^-1 (P x P, P ~ 2.5 m) instead of the reference's K^-1 (n x n), plus the leaf
// bitsets bits[p] = { i : point i falls in leaf-column p } from which every proposal's column of A is an
// AND + POPC.  Unused columns are kept as empty leaves (A row/col 0, B_pp = c, Binv_pp = 1/c), which
// contribute nothing to ldt or q, so the capacity p_cap is fixed while the used extent p_hi moves.
#pragma once
#include "common.cuh"

namespace bark {

struct ChainScalars {
    double noise, scale, sig, c;
    double res, ldt, mll, yy;  // res = y^T y - b^T Binv b (= sig * y^T K^-1 y), ldt = log|I + A/c|
    unsigned long long counters[16];
    unsigned long long phase_cycles[12];  // BARK_PHASE_TIMING builds only: per-phase clock64 totals of the tree sweep
    double prop_noise, prop_scale;  // accepted-but-not-yet-refreshed hyper proposal (hyper_eval -> hyper_refresh)
    int p_hi;           // used column extent (columns >= p_hi are free and identity-like)
    unsigned status;    // BARK_ST_* bits
    int hyper_accept;   // set by hyper_eval_kernel, consumed by hyper_refresh_kernel
    int pad1;
};

struct WsLayout {
    int64_t chains, n, d, m, L, P, wd, npad;
    size_t off_xt, off_y, off_bounds, off_ft;  // shared
    size_t off_chain0, chain_stride;           // per chain block
    size_t off_binv, off_wk, off_s2, off_a, off_bits, off_ck, off_gk, off_dg, off_b, off_w, off_yv, off_colmap, off_colused, off_sc;
    size_t total;
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

__host__ __device__ inline WsLayout make_layout(const bark_mcmc_dims& dm) {
    WsLayout w;
    w.chains = dm.chains; w.n = dm.n; w.d = dm.d; w.m = dm.m; w.L = dm.node_limit; w.P = dm.p_cap;
    w.wd = (((dm.n + 31) / 32) + 3) & ~(int64_t)3;  // words per leaf bitset, padded to 16 bytes (uint4 loads)
    w.npad = w.wd * 32;
    size_t o = 0;
    w.off_xt = o;     o = align256(o + (size_t)w.d * w.npad * sizeof(double));
    w.off_y = o;      o = align256(o + (size_t)w.npad * sizeof(double));
    w.off_bounds = o; o = align256(o + (size_t)w.d * 2 * sizeof(double));
    w.off_ft = o;     o = align256(o + (size_t)w.d * sizeof(int32_t));
    w.off_chain0 = o;
    size_t c = 0;
    const size_t P = (size_t)w.P;
    w.off_binv = c;    c = align256(c + P * P * sizeof(double));
    w.off_wk = c;      c = align256(c + P * P * sizeof(double));
    w.off_a = c;       c = align256(c + P * P * sizeof(int32_t));
    w.off_bits = c;    c = align256(c + P * (size_t)w.wd * sizeof(uint32_t));
    w.off_ck = c;      c = align256(c + P * 64 * sizeof(double));
    w.off_gk = c;      c = align256(c + P * 64 * sizeof(double));
    w.off_dg = c;      c = align256(c + 64 * 64 * sizeof(double));
    w.off_b = c;       c = align256(c + P * sizeof(double));
    w.off_w = c;       c = align256(c + P * sizeof(double));
    w.off_yv = c;      c = align256(c + P * sizeof(double));
    w.off_colmap = c;  c = align256(c + (size_t)w.m * w.L * sizeof(uint16_t));
    w.off_colused = c; c = align256(c + (P / 32) * sizeof(uint32_t));
    w.off_sc = c;      c = align256(c + sizeof(ChainScalars));
    w.chain_stride = c;
    // second scratch matrix per chain (inverse refinement): kept OUTSIDE the per-chain blocks so that the layout of the
    // hot state (and with it the L2 behaviour of the sweep kernel) does not depend on it
    w.off_s2 = w.off_chain0 + (size_t)w.chains * w.chain_stride;
    w.total = w.off_s2 + (size_t)w.chains * align256(P * P * sizeof(double));
    return w;
}

struct ChainView {
    double* Binv; double* Wk; double* S2; int32_t* A; uint32_t* bits; double* CK; double* GK; double* DG;
    double* b; double* w; double* yv; uint16_t* colmap; uint32_t* colused; ChainScalars* sc;
};

__host__ __device__ inline ChainView chain_view(const WsLayout& w, void* ws, int64_t chain) {
    unsigned char* base = (unsigned char*)ws + w.off_chain0 + (size_t)chain * w.chain_stride;
    ChainView v;
    v.Binv = (double*)(base + w.off_binv);
    v.Wk = (double*)(base + w.off_wk);
    v.S2 = (double*)((unsigned char*)ws + w.off_s2 + (size_t)chain * align256((size_t)w.P * w.P * sizeof(double)));
    v.A = (int32_t*)(base + w.off_a);
    v.bits = (uint32_t*)(base + w.off_bits);
    v.CK = (double*)(base + w.off_ck);
    v.GK = (double*)(base + w.off_gk);
    v.DG = (double*)(base + w.off_dg);
    v.b = (double*)(base + w.off_b);
    v.w = (double*)(base + w.off_w);
    v.yv = (double*)(base + w.off_yv);
    v.colmap = (uint16_t*)(base + w.off_colmap);
    v.colused = (uint32_t*)(base + w.off_colused);
    v.sc = (ChainScalars*)(base + w.off_sc);
    return v;
}

struct SharedView {
    const double* Xt;  // [d][npad] feature-major
    const double* y;   // [npad]
    const double* bounds;
    const int32_t* ft;
};
__host__ __device__ inline SharedView shared_view(const WsLayout& w, const void* ws) {
    const unsigned char* base = (const unsigned char*)ws;
    SharedView s;
    s.Xt = (const double*)(base + w.off_xt);
    s.y = (const double*)(base + w.off_y);
    s.bounds = (const double*)(base + w.off_bounds);
    s.ft = (const int32_t*)(base + w.off_ft);
    return s;
}

constexpr uint16_t NO_COL = 0xFFFFu;

}  // namespace bark
