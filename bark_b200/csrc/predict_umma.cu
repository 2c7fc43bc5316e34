// a14 on the tensor cores: posterior-predictive variance as an EXACT int8-sliced one-hot GEMM.
//
//   var_s(x) = sig_s * z^T Binv_s z,   z = one-hot leaf indicator of the candidate (m ones among P columns).
// Binv_s is turned once per posterior sample into 7 signed base-256 digit planes of a 54-bit fixed-point
// representation (|Binv| <= 1/c bounds the scale):  Binv = 2^-shift * sum_k 256^k D_k,  D_k int8.
// For a tile of 128 candidates the CTA walks the trees with three threads per candidate (leaf columns collected as
// a bit mask per row, mean = sum_t w[col_t] on the way), expands the masks into the one-hot int8 A operand in
// shared memory, then T_k = Zc * D_k runs on tcgen05 (kind::i8, s32 accumulators in TMEM, M = 128, N = 256 per
// instruction: with both operands in shared memory a small N is bound by re-reading A), one (column tile, digit
// plane) item at a time into one of two 256-column TMEM buffers, the digit tiles streamed by the bulk-copy engine.
// The epilogue never needs T itself: z^T Binv z = sum_k 256^k * (sum over the candidate's own columns of T_k),
// i.e. 7 masked int32 row sums (exact), combined once in FP64; it drains one TMEM buffer while the tensor core
// fills the other.  No n_c x P matrix ever exists in memory.
//
// Replaces  scale - diag(K_xX K^-1 K_Xx)  of src/bark/tree_kernels/tree_gps.py:103-112.
#include <algorithm>

#include "common.cuh"
#include "forest_device.cuh"
#include "mcmc_state.cuh"

namespace bark {

constexpr int PU_ROWS = 128;        // candidates per CTA (UMMA M)
constexpr int PU_N = 256;           // Binv columns per accumulator tile (UMMA N); two TMEM buffers of 256 columns
constexpr int PU_KB = 128;          // K bytes per operand tile (one SWIZZLE_128B atom row)
constexpr int PU_SLICES = 7;        // base-256 digit planes
constexpr int PU_MAX_STAGES = 4;    // ring stages; one stage = one K tile of one (column tile, digit plane)
constexpr int PU_WALK_GROUPS = 4;   // threads per candidate: walk (trees t = g mod 4) and epilogue (column chunks j = g mod 4);
                                    // 6 groups measured the same (the walk is issue / LSU bound, not latency bound)
constexpr int PU_THREADS = 64 + 128 * PU_WALK_GROUPS;  // warps 0-3 and 6..: walk + epilogue groups, 4: MMA issue, 5: TMA producer
constexpr int PU_A_TILE = PU_ROWS * PU_KB;  // 16 KB
constexpr int PU_B_TILE = PU_N * PU_KB;     // 32 KB
constexpr int PU_RING_MAX = 128 * 1024;
constexpr int PU_MAX_P = 768;

__device__ __forceinline__ uint32_t pu_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ __forceinline__ uint32_t pu_swizzle(uint32_t r, uint32_t kb) {
    const uint32_t g = r >> 3, rr = r & 7, chunk = kb >> 4, b = kb & 15;
    return g * 1024u + rr * 128u + ((chunk ^ rr) << 4) + b;
}

struct PrepLayout {
    size_t off_table, off_tiles, off_scale, total;
    int hi, kt, nt;
};
__host__ __device__ inline PrepLayout prep_layout(int64_t samples, int64_t m, int slots, int p_max) {
    PrepLayout l;
    l.hi = slots;
    l.kt = (p_max + PU_KB - 1) / PU_KB;
    l.nt = (l.kt * PU_KB + PU_N - 1) / PU_N;  // the last column tile may be half full (N = 128)
    size_t o = 0;
    l.off_table = o; o = align256(o + (size_t)samples * m * slots * sizeof(WalkNode));
    l.off_tiles = o; o = align256(o + (size_t)samples * l.nt * PU_SLICES * l.kt * PU_B_TILE);
    l.off_scale = o; o = align256(o + (size_t)samples * sizeof(double));
    l.total = o;
    return l;
}

// compact walk table [sample][tree][slot < hi]; leaves carry their leaf-space column in the threshold bits
__global__ void pu_table_kernel(WsLayout lay, const void* ws, bark_nodes_soa forest, int hi, WalkNode* __restrict__ table) {
    const int64_t total = lay.chains * lay.m * hi;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t chain = e / (lay.m * hi), rem = e % (lay.m * hi), t = rem / hi, sl = rem % hi;
        const int64_t g = (chain * lay.m + t) * lay.L + sl;
        ChainView cv = chain_view(lay, const_cast<void*>(ws), chain);
        WalkNode w = make_walk_node(forest.is_leaf[g], forest.feature[g], forest.threshold[g], forest.left[g], forest.right[g]);
        if (forest.is_leaf[g]) w.thr = __int_as_float((int)cv.colmap[t * lay.L + sl]);
        table[e] = w;
    }
}

// Binv (lower triangle current) -> 7 digit planes, tiled [sample][nt][slice][kt] and pre-swizzled (K-major SW128)
__global__ void pu_slice_kernel(WsLayout lay, const void* ws, int kt_n, int nt_n, uint8_t* __restrict__ tiles,
                                double* __restrict__ scale_out) {
    const int64_t sample = blockIdx.y;
    ChainView cv = chain_view(lay, const_cast<void*>(ws), sample);
    const double c = cv.sc->c;
    // |Binv_ij| <= lambda_max(Binv) <= 1/c < 2^e  ->  x = Binv * 2^(53 - e) fits 54 signed bits
    int e;
    frexp(1.0 / c, &e);
    const int shift = 53 - e;
    if (blockIdx.x == 0 && threadIdx.x == 0) scale_out[sample] = ldexp(1.0, -shift);
    const int64_t K = (int64_t)kt_n * PU_KB, Q = K, P = lay.P;
    uint8_t* base = tiles + (size_t)sample * nt_n * PU_SLICES * kt_n * PU_B_TILE;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < Q * K; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = idx / K, k = idx % K;
        double v = 0.0;
        if (q < P && k < P) v = (k <= q) ? cv.Binv[q * P + k] : cv.Binv[k * P + q];
        long long x = llrint(ldexp(v, shift));
        const int64_t nt = q / PU_N, kt = k / PU_KB;
        const uint32_t off = pu_swizzle((uint32_t)(q % PU_N), (uint32_t)(k % PU_KB));
#pragma unroll
        for (int s = 0; s < PU_SLICES; ++s) {
            const long long dgt = ((x + 128) & 255) - 128;  // balanced digit in [-128, 127]
            x = (x - dgt) >> 8;
            base[(((size_t)nt * PU_SLICES + s) * kt_n + kt) * PU_B_TILE + off] = (uint8_t)(int8_t)dgt;
        }
    }
}

// ---- PTX wrappers (same encodings as gram_umma.cu) ------------------------------------------------------------
__device__ __forceinline__ void pu_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pu_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void pu_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pu_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pu_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pu_smem(bar)) : "memory");
}
__device__ __forceinline__ bool pu_mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.b32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(pu_smem(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: gives up after ~2^26 probes and flags BARK_ST_TIMEOUT on the sample's chain state (the host raises)
// instead of hanging the GPU.
__device__ __forceinline__ void pu_mbar_wait(uint64_t* bar, uint32_t parity, unsigned* status) {
    unsigned spins = 0;
    while (!pu_mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) {
            atomicOr(status, BARK_ST_TIMEOUT);
            break;
        }
    }
}
__device__ __forceinline__ void pu_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(pu_smem(dst)),
                 "l"(src), "r"(bytes), "r"(pu_smem(bar))
                 : "memory");
}
__device__ __forceinline__ uint64_t pu_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t pu_idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void pu_umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void pu_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(pu_smem(bar)) : "memory");
}
__device__ __forceinline__ void pu_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct PuSmem {
    size_t off_a, off_ring, off_table, off_zmask, off_w, off_xs, off_ft, off_mean, off_part, off_bars, total;
};
__host__ __device__ inline PuSmem pu_smem_layout(int m, int hi, int d, int kt, int ring_bytes) {
    PuSmem s;
    size_t o = 0;
    s.off_a = o;      o += (size_t)kt * PU_A_TILE;
    s.off_ring = o;   o += (size_t)ring_bytes;
    s.off_table = o;  o = align256(o + (size_t)m * hi * sizeof(WalkNode));
    s.off_zmask = o;  o = align256(o + (size_t)PU_ROWS * kt * 2 * 8);  // [64-column word][row]
    s.off_w = o;      o = align256(o + (size_t)kt * PU_KB * 8);
    s.off_xs = o;     o = align256(o + (size_t)d * (PU_ROWS + 1) * 8);
    s.off_ft = o;     o = align256(o + (size_t)d * 4);
    s.off_mean = o;   o = align256(o + (size_t)PU_WALK_GROUPS * PU_ROWS * 8);
    s.off_part = o;   o = align256(o + (size_t)(PU_WALK_GROUPS - 1) * PU_SLICES * PU_ROWS * 4);
    s.off_bars = o;   o = align256(o + (size_t)(2 * PU_MAX_STAGES + 4) * 8 + 16);
    s.total = o;
    return s;
}

// 16 mask bits -> 16 bytes of 0 / 1 (x * 0x00204081 spreads 4 bits over the low bits of 4 bytes)
__device__ __forceinline__ uint4 pu_expand16(uint32_t b) {
    uint4 r;
    r.x = ((b & 0xFu) * 0x00204081u) & 0x01010101u;
    r.y = (((b >> 4) & 0xFu) * 0x00204081u) & 0x01010101u;
    r.z = (((b >> 8) & 0xFu) * 0x00204081u) & 0x01010101u;
    r.w = (((b >> 12) & 0xFu) * 0x00204081u) & 0x01010101u;
    return r;
}

__global__ void __launch_bounds__(PU_THREADS, 1)
predict_umma_kernel(WsLayout lay, const void* ws, const WalkNode* __restrict__ table, const uint8_t* __restrict__ tiles,
                    const double* __restrict__ scales, int hi, int kt_n, int nt_n, int ring_bytes,
                    const double* __restrict__ cand,
                    int64_t n_c, double* __restrict__ mu, double* __restrict__ var) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int d = (int)lay.d, m = (int)lay.m;
    const PuSmem sl = pu_smem_layout(m, hi, d, kt_n, ring_bytes);
    unsigned char* a_tiles = smem_raw + sl.off_a;
    unsigned char* ring = smem_raw + sl.off_ring;
    WalkNode* tb = reinterpret_cast<WalkNode*>(smem_raw + sl.off_table);
    unsigned long long* zmask = reinterpret_cast<unsigned long long*>(smem_raw + sl.off_zmask);  // [word][row]
    double* w_s = reinterpret_cast<double*>(smem_raw + sl.off_w);
    double* xs = reinterpret_cast<double*>(smem_raw + sl.off_xs);  // [d][PU_ROWS + 1]
    int* ftc = reinterpret_cast<int*>(smem_raw + sl.off_ft);
    double* meanp = reinterpret_cast<double*>(smem_raw + sl.off_mean);  // [group][row]
    int* accp = reinterpret_cast<int*>(smem_raw + sl.off_part);          // [group - 1][slice][row]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + sl.off_bars);
    uint64_t* empty_bar = full_bar + PU_MAX_STAGES;
    uint64_t* acc_full = empty_bar + PU_MAX_STAGES;  // [2]
    uint64_t* acc_free = acc_full + 2;               // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);
    const int nstages = min(PU_MAX_STAGES, ring_bytes / PU_B_TILE);
    const int nwords = kt_n * 2;  // 64-column mask words per candidate row

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t sample = blockIdx.y;
    const int64_t p0 = (int64_t)blockIdx.x * PU_ROWS;
    const int np = (int)min((int64_t)PU_ROWS, n_c - p0);
    ChainView cv = chain_view(lay, const_cast<void*>(ws), sample);
    unsigned* pu_status = &cv.sc->status;  // bounded pipeline waits flag BARK_ST_TIMEOUT here
    SharedView sv = shared_view(lay, ws);

    // ---- setup: barriers, TMEM, staging of the sample's trees / w and of the candidate tile
    if (tid == 0) {
        for (int s = 0; s < PU_MAX_STAGES; ++s) { pu_mbar_init(full_bar + s, 1); pu_mbar_init(empty_bar + s, 1); }
        for (int b = 0; b < 2; ++b) { pu_mbar_init(acc_full + b, 1); pu_mbar_init(acc_free + b, PU_ROWS * PU_WALK_GROUPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(pu_smem(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {
        const WalkNode* src = table + (size_t)sample * m * hi;
        for (int e = tid; e < m * hi; e += PU_THREADS) tb[e] = src[e];
        for (int e = tid; e < kt_n * PU_KB; e += PU_THREADS) w_s[e] = (e < lay.P) ? cv.w[e] : 0.0;
        for (int e = tid; e < np * d; e += PU_THREADS) xs[(size_t)(e % d) * (PU_ROWS + 1) + e / d] = cand[p0 * d + e];
        for (int e = tid; e < d; e += PU_THREADS) ftc[e] = sv.ft[e];
        for (int e = tid; e < PU_ROWS * nwords; e += PU_THREADS) zmask[e] = 0ull;
    }
    // ---- warp 5: TMA producer.  One stage = one K tile (n_cols x 128 B) of one (column tile, digit plane) item.
    // The first ring-full needs no free-slot wait: it is issued now, concurrently with the walk (the digit stream
    // does not depend on the candidates).
    const uint8_t* src_tiles = tiles + (size_t)sample * nt_n * PU_SLICES * kt_n * PU_B_TILE;
    const int items = nt_n * PU_SLICES;
    const int loads = items * kt_n;  // global tile index == load index: tiles are stored [nt][slice][kt]
    const int last_cols = kt_n * PU_KB - (nt_n - 1) * PU_N;  // columns of the last column tile (128 or 256)
    __syncthreads();  // barriers initialised
    if (warp == 5 && lane == 0) {
        for (int ld = 0; ld < nstages && ld < loads; ++ld) {
            const int nt = ld / (PU_SLICES * kt_n);
            const uint32_t bytes = (uint32_t)((nt == nt_n - 1) ? last_cols : PU_N) * PU_KB;
            pu_mbar_expect_tx(full_bar + ld, bytes);
            pu_bulk_g2s(ring + (size_t)ld * PU_B_TILE, src_tiles + (size_t)ld * PU_B_TILE, bytes, full_bar + ld);
        }
    }

    // ---- walk: PU_WALK_GROUPS threads per candidate (group g takes the trees t = g mod groups): leaf columns into
    // the row's bit mask, partial means per group
    const int wg = (warp < 4) ? 0 : (warp >= 6 ? 1 + (warp - 6) / 4 : -1);  // walk / epilogue group of this warp
    {
        const int row = (warp < 4) ? tid : (warp >= 6 ? (tid - 6 * 32) % PU_ROWS : 0);
        if (wg >= 0) {
            double mean = 0.0;
            if (row < np) {
                const double* xp = xs + row;
                for (int t = wg; t < m; t += PU_WALK_GROUPS) {
                    const WalkNode* wn = tb + (size_t)t * hi;
                    WalkNode nd = wn[0];
                    for (int it = 0; it < hi && !(nd.feat_leaf & 0x8000u); ++it) {
                        const int f = nd.feat_leaf & 0x7fffu;
                        const uint32_t at = goes_left(xp[(size_t)f * (PU_ROWS + 1)], nd.thr, ftc[f]) ? nd.left : nd.right;
                        nd = wn[min(at, (uint32_t)(hi - 1))];
                    }
                    const int col = __float_as_int(nd.thr);
                    mean += w_s[col];
                    atomicOr(zmask + (size_t)(col >> 6) * PU_ROWS + row, 1ull << (col & 63));
                }
            }
            meanp[wg * PU_ROWS + row] = mean;
        }
    }
    __syncthreads();
    // ---- one-hot A operand (K-major, SWIZZLE_128B) from the masks: one 16-byte chunk per thread and step
    {
        const int chunks_per_row = kt_n * 8;
        for (int e = tid; e < PU_ROWS * chunks_per_row; e += PU_THREADS) {
            const int row = e / chunks_per_row, ch = e % chunks_per_row;
            const uint32_t bits = (uint32_t)(zmask[(size_t)(ch >> 2) * PU_ROWS + row] >> ((ch & 3) * 16)) & 0xFFFFu;
            const uint32_t r = (uint32_t)row, c = (uint32_t)(ch & 7);
            unsigned char* dst = a_tiles + (size_t)(ch >> 3) * PU_A_TILE + (r >> 3) * 1024u + (r & 7) * 128u + ((c ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(dst) = pu_expand16(bits);
        }
    }
    asm volatile("fence.proxy.async;" ::: "memory");  // generic writes of the A operand -> tensor-core (async proxy) reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 5) {
        if (lane == 0) {
            for (int ld = nstages; ld < loads; ++ld) {
                const int s = ld % nstages;
                const int nt = ld / (PU_SLICES * kt_n);
                const uint32_t bytes = (uint32_t)((nt == nt_n - 1) ? last_cols : PU_N) * PU_KB;
                pu_mbar_wait(empty_bar + s, (uint32_t)((ld / nstages - 1) & 1), pu_status);
                pu_mbar_expect_tx(full_bar + s, bytes);
                pu_bulk_g2s(ring + (size_t)s * PU_B_TILE, src_tiles + (size_t)ld * PU_B_TILE, bytes, full_bar + s);
            }
        }
        __syncwarp();
    } else if (warp == 4) {
        if (lane == 0) {
            // ---- MMA issuer: item = (column tile, digit plane) into TMEM buffer item & 1
            int ld = 0;
            for (int it = 0; it < items; ++it) {
                const int nt = it / PU_SLICES, buf = it & 1, use = it >> 1;
                const int ncols = (nt == nt_n - 1) ? last_cols : PU_N;
                const uint32_t idesc = pu_idesc_i8(PU_ROWS, ncols);
                if (use > 0) pu_mbar_wait(acc_free + buf, (uint32_t)((use - 1) & 1), pu_status);  // epilogue has drained the buffer
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kt = 0; kt < kt_n; ++kt, ++ld) {
                    const int s = ld % nstages;
                    pu_mbar_wait(full_bar + s, (uint32_t)((ld / nstages) & 1), pu_status);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = pu_smem(a_tiles + (size_t)kt * PU_A_TILE);
                    const uint32_t b_addr = pu_smem(ring + (size_t)s * PU_B_TILE);
#pragma unroll
                    for (int k4 = 0; k4 < PU_KB / 32; ++k4)
                        pu_umma_i8(tmem_d + (uint32_t)(buf * PU_N), pu_desc_sw128(a_addr + k4 * 32),
                                   pu_desc_sw128(b_addr + k4 * 32), idesc, (kt > 0 || k4 > 0) ? 1u : 0u);
                    pu_commit(empty_bar + s);
                }
                pu_commit(acc_full + buf);
            }
        }
        __syncwarp();
    } else {
        // ---- epilogue: masked int32 row sums of every digit plane.  A warp reads the TMEM lanes of its quarter
        // (warp % 4), thread = candidate row; the four groups share the 32-column chunks of every item.
        const int row = 32 * (warp & 3) + lane;
        int acc[PU_SLICES];
#pragma unroll
        for (int s = 0; s < PU_SLICES; ++s) acc[s] = 0;
        int it = 0;
        for (int nt = 0; nt < nt_n; ++nt) {
            const int ncols = (nt == nt_n - 1) ? last_cols : PU_N;
            unsigned long long zm[PU_N / 64];
#pragma unroll
            for (int q = 0; q < PU_N / 64; ++q) zm[q] = (q * 64 < ncols) ? zmask[(size_t)(nt * (PU_N / 64) + q) * PU_ROWS + row] : 0ull;
#pragma unroll
            for (int s = 0; s < PU_SLICES; ++s, ++it) {
                const int buf = it & 1, use = it >> 1;
                pu_mbar_wait(acc_full + buf, (uint32_t)(use & 1), pu_status);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                int a = acc[s];
#pragma unroll
                for (int j = 0; j < PU_N / 32; ++j) {
                    if ((j % PU_WALK_GROUPS) != wg) continue;  // warp-uniform
                    const uint32_t bits = (uint32_t)(zm[j >> 1] >> (32 * (j & 1)));
                    if (__any_sync(0xffffffffu, bits != 0u)) {
                        uint32_t v[32];
                        pu_tmem_ld32(tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(buf * PU_N + 32 * j), v);
#pragma unroll
                        for (int c = 0; c < 32; ++c) a += (int)v[c] * (int)((bits >> c) & 1u);
                    }
                }
                acc[s] = a;
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                pu_mbar_arrive(acc_free + buf);
            }
        }
        if (wg > 0) {
#pragma unroll
            for (int s = 0; s < PU_SLICES; ++s) accp[((wg - 1) * PU_SLICES + s) * PU_ROWS + row] = acc[s];
        }
        asm volatile("bar.sync 1, %0;" ::"r"(PU_WALK_GROUPS * PU_ROWS) : "memory");  // the epilogue warps only
        if (wg == 0 && row < np) {
            // z^T Binv z = 2^-shift * sum_k 256^k acc_k   (each acc_k exact)
            double tsum = 0.0;
#pragma unroll
            for (int s = PU_SLICES - 1; s >= 0; --s) {
                int a = acc[s];
#pragma unroll
                for (int g = 1; g < PU_WALK_GROUPS; ++g) a += accp[((g - 1) * PU_SLICES + s) * PU_ROWS + row];
                tsum = tsum * 256.0 + (double)a;
            }
            double mean = 0.0;
#pragma unroll
            for (int g = 0; g < PU_WALK_GROUPS; ++g) mean += meanp[g * PU_ROWS + row];
            const int64_t o = sample * n_c + p0 + row;
            mu[o] = mean;
            var[o] = cv.sc->sig * (tsum * scales[sample]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512u) : "memory");
}

}  // namespace bark

using namespace bark;

extern "C" {

size_t bark_predict_prep_bytes(const bark_mcmc_dims* dims, int32_t slots, int32_t p_max) {
    if (!dims || slots < 1 || slots > 255 || p_max < 1 || p_max > PU_MAX_P) return 0;
    return prep_layout(dims->chains, dims->m, slots, p_max).total;
}

int bark_predict_prepare(const bark_mcmc_dims* dims, const void* workspace, bark_nodes_soa forest, int32_t slots,
                         int32_t p_max, void* prep, void* stream) {
    BARK_CHECK_ARG(dims && workspace && prep && forest.is_leaf, "null pointer");
    BARK_CHECK_ARG(slots >= 1 && slots <= 255, "slots out of range");
    BARK_CHECK_ARG(p_max >= 1 && p_max <= PU_MAX_P && p_max <= dims->p_cap, "p_max out of range (<= 768)");
    const WsLayout lay = make_layout(*dims);
    const PrepLayout pl = prep_layout(dims->chains, dims->m, slots, p_max);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* base = (unsigned char*)prep;
    pu_table_kernel<<<148 * 2, 256, 0, st>>>(lay, workspace, forest, slots, (WalkNode*)(base + pl.off_table));
    dim3 grid(148, (unsigned)dims->chains);
    pu_slice_kernel<<<grid, 256, 0, st>>>(lay, workspace, pl.kt, pl.nt, base + pl.off_tiles, (double*)(base + pl.off_scale));
    BARK_LAUNCH_CHECK();
    return BARK_OK;
}

// per-sample moments (samples, n_c) into mu / var
int bark_predict_umma(const bark_mcmc_dims* dims, const void* workspace, const void* prep, int32_t slots, int32_t p_max,
                      const double* candidates, int64_t n_c, double* mu, double* var, void* stream) {
    BARK_CHECK_ARG(dims && workspace && prep, "null pointer");
    BARK_CHECK_ARG(n_c >= 0, "n_c < 0");
    if (n_c == 0) return BARK_OK;
    BARK_CHECK_ARG(candidates && mu && var, "null pointer");
    BARK_CHECK_ARG(slots >= 1 && slots <= 255 && p_max >= 1 && p_max <= PU_MAX_P, "slots / p_max out of range");
    BARK_CHECK_ARG(dims->chains <= 65535, "too many samples per call");
    const WsLayout lay = make_layout(*dims);
    const PrepLayout pl = prep_layout(dims->chains, dims->m, slots, p_max);
    const int stage_bytes = PU_B_TILE;
    const size_t fixed = pu_smem_layout((int)lay.m, slots, (int)lay.d, pl.kt, 0).total;
    BARK_CHECK_ARG(fixed + 2 * (size_t)stage_bytes <= 227 * 1024, "m * slots / d / p_max too large for the predict kernel's shared memory");
    const int ring_bytes = (int)(std::min<size_t>(PU_RING_MAX, 227 * 1024 - fixed) / stage_bytes) * stage_bytes;
    const PuSmem sl = pu_smem_layout((int)lay.m, slots, (int)lay.d, pl.kt, ring_bytes);
    BARK_CUDA(cudaFuncSetAttribute(predict_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sl.total));
    const unsigned char* base = (const unsigned char*)prep;
    dim3 grid((unsigned)ceil_div(n_c, PU_ROWS), (unsigned)dims->chains);
    predict_umma_kernel<<<grid, PU_THREADS, sl.total, (cudaStream_t)stream>>>(
        lay, workspace, (const WalkNode*)(base + pl.off_table), base + pl.off_tiles, (const double*)(base + pl.off_scale),
        slots, pl.kt, pl.nt, ring_bytes, candidates, n_c, mu, var);
    BARK_LAUNCH_CHECK();
    return BARK_OK;
}

}  // extern "C"
