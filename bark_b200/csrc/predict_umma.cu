// a14 on the tensor cores: posterior-predictive variance as an EXACT int8-sliced one-hot GEMM.
//
//   var_s(x) = sig_s * z^T Binv_s z,   z = one-hot leaf indicator of the candidate (m ones among P columns).
// Binv_s is turned once per posterior sample into 7 signed base-256 digit planes of a fixed-point representation
// (|Binv| <= 1/c bounds the scale) of the triangle { diag, 2 x strictly lower }:  z^T Binv z = 2^-shift * sum_k 256^k z^T D_k z.
//
// One persistent CTA per tile of 128 candidates loops over the posterior samples; its warps have fixed roles that run
// concurrently, one sample apart (laid out by warp scheduler, see PU_MMA_WARP):
//   walkers (12 warps)    stage the sample's trees, walk them with three threads per candidate (four trees in flight per
//                         thread, numeric splits decided in FP32 on candidates rounded up -- exact, and off the FP64
//                         pipe), leaf columns -> the row's bit mask, partial means on the way;
//   epilogue (8 warps)    expand the masks into the one-hot int8 A operand IN TENSOR MEMORY (tcgen05.st; row = lane, four
//                         K bytes per column) once the previous sample's MMAs are done; drain the accumulators:
//                         z^T D_k z is the masked row sum of T_k = Z D_k, read as packed int16 pairs
//                         (tcgen05.ld .pack::16b) and reduced with dp2a against the mask bytes;
//   MMA issuer (1 thread) tcgen05.mma kind::i8 with A from TMEM, M = 128, N = 192 (two accumulators beside the A columns),
//                         one (column tile, digit plane) item at a time, only the K tiles on or below the diagonal band;
//   producer (1 thread)   streams the 24 KB digit tiles with the bulk-copy engine through a six-stage shared-memory ring,
//                         free-running across samples.
// The walk of sample s + 1 overlaps the MMAs of sample s.  No n_c x P matrix ever exists in memory.
// Measured building blocks: scripts/umma_shapes.cu (MMA rate per shape, TMEM read rate), scripts/tma_feed.cu (bulk-copy feed).
//
// Replaces  scale - diag(K_xX K^-1 K_Xx)  of src/bark/tree_kernels/tree_gps.py:103-112.
#include <algorithm>

#include "common.cuh"
#include "forest_device.cuh"
#include "mcmc_state.cuh"

namespace bark {

constexpr int PU_ROWS = 128;        // candidates per CTA (UMMA M)
constexpr int PU_N_MAX = 192;       // Binv columns per accumulator tile (UMMA N).  TMEM holds the one-hot A operand (K_pad / 4
                                    // columns) and two accumulators: N = 192 for K_pad <= 512, 160 above (pu_ntile)
constexpr int PU_KB = 128;          // K bytes per operand tile (one SWIZZLE_128B atom row)
constexpr int PU_SLICES = 7;        // base-256 digit planes
constexpr int PU_MAX_STAGES = 8;    // ring stages; one stage = up to PU_GROUP consecutive K tiles of one (column tile, digit plane)
constexpr int PU_GROUP = 2;         // K tiles per stage: one bulk copy, one barrier hand-shake and one commit per group
constexpr int PU_WALK_GROUPS = 3;   // walker threads per candidate (trees t = g mod 3)
constexpr int PU_WALK_WARPS = 4 * PU_WALK_GROUPS;
constexpr int PU_EPI_WARPS = 8;     // two per TMEM lane quarter: each takes half of an item's columns
// Warp roles are laid out by warp scheduler (warp % 4, which is also the TMEM lane quarter a warp may touch): the
// MMA-issuing thread runs a latency-bound scalar loop, so its scheduler (0) hosts no walker -- only the producer and the two
// light epilogue warps of quarter 0.
//   warp 0: MMA issue      warp 4: bulk-copy producer      warps 16, 20: unused
//   epilogue: warps 1-3, 5-7 (quarters 1-3, column halves 0 / 1) and 8, 12 (quarter 0)
//   walkers: the twelve warps 9-11, 13-15, 17-19, 21-23
constexpr int PU_MMA_WARP = 0, PU_PRODUCER_WARP = 4;
constexpr int PU_THREADS = 32 * 24;
__device__ __forceinline__ int pu_epilogue_half(int warp) {  // -1: not an epilogue warp
    if ((warp & 3) == 0) return warp == 8 ? 0 : (warp == 12 ? 1 : -1);
    return warp < 4 ? 0 : (warp < 8 ? 1 : -1);
}
__device__ __forceinline__ int pu_walker_index(int warp) {  // 0 .. 11, -1: not a walker
    return (warp >= 9 && warp < 24 && (warp & 3) != 0) ? 3 * ((warp - 8) >> 2) + (warp & 3) - 1 : -1;
}
constexpr int PU_RING_MAX = 4 * PU_GROUP * PU_N_MAX * PU_KB;
constexpr int PU_MAX_P = 768;

__device__ __forceinline__ uint32_t pu_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ __forceinline__ uint32_t pu_swizzle(uint32_t r, uint32_t kb) {
    const uint32_t g = r >> 3, rr = r & 7, chunk = kb >> 4, b = kb & 15;
    return g * 1024u + rr * 128u + ((chunk ^ rr) << 4) + b;
}

// Accumulator width for kt K tiles: the A operand takes 32 kt TMEM columns, two accumulators share the other 512 - 32 kt.
__host__ __device__ __forceinline__ int pu_ntile(int kt) { return kt <= 4 ? 192 : 160; }
// First K tile of column tile nt that can hold an entry of the triangle k >= q (columns q of tile nt start at nt * ntile).
__host__ __device__ __forceinline__ int pu_kt_lo(int nt, int ntile) { return (nt * ntile) / PU_KB; }

struct PrepLayout {
    size_t off_table, off_tiles, off_scale, total;
    int hi, kt, nt, ntile;
};
__host__ __device__ inline PrepLayout prep_layout(int64_t samples, int64_t m, int slots, int p_max) {
    PrepLayout l;
    l.hi = slots;
    l.kt = (p_max + PU_KB - 1) / PU_KB;
    l.ntile = pu_ntile(l.kt);
    l.nt = (l.kt * PU_KB + l.ntile - 1) / l.ntile;  // the last column tile may be narrower
    size_t o = 0;
    l.off_table = o; o = align256(o + (size_t)samples * m * slots * sizeof(WalkNode));
    l.off_tiles = o; o = align256(o + (size_t)samples * l.nt * PU_SLICES * l.kt * l.ntile * PU_KB);
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

// Binv (lower triangle current) -> 7 digit planes, tiled [sample][nt][slice][kt] and pre-swizzled (K-major SW128);
// only the K tiles kt >= pu_kt_lo(nt, ntile) of a column tile are written and later streamed
__global__ void pu_slice_kernel(WsLayout lay, const void* ws, int kt_n, int nt_n, int ntile, uint8_t* __restrict__ tiles,
                                double* __restrict__ scale_out) {
    const size_t tile_bytes = (size_t)ntile * PU_KB;
    const int64_t sample = blockIdx.y;
    ChainView cv = chain_view(lay, const_cast<void*>(ws), sample);
    const double c = cv.sc->c;
    // |Binv_ij| <= lambda_max(Binv) <= 1/c < 2^e  ->  x = Binv * 2^(53 - e) fits 54 signed bits
    int e;
    frexp(1.0 / c, &e);
    const int shift = 53 - e;
    if (blockIdx.x == 0 && threadIdx.x == 0) scale_out[sample] = ldexp(1.0, -shift);
    const int64_t K = (int64_t)kt_n * PU_KB, Q = K, P = lay.P;
    uint8_t* base = tiles + (size_t)sample * nt_n * PU_SLICES * kt_n * tile_bytes;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < Q * K; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = idx / K, k = idx % K;
        const int64_t nt = q / ntile, kt = k / PU_KB;
        if (kt < pu_kt_lo((int)nt, ntile)) continue;  // tile above the diagonal band: never loaded
        // z^T Binv z = sum_i Binv_ii z_i + 2 sum_{k > q} Binv_kq z_k z_q: the operand is the triangle k >= q with the
        // off-diagonal entries doubled (|2 Binv_kq| < 2^(e+1): 55 signed bits, inside the 7 balanced digits' +-2^55)
        double v = 0.0;
        if (q < P && k < P && k >= q) v = (k == q) ? cv.Binv[q * P + q] : 2.0 * cv.Binv[k * P + q];
        long long x = llrint(ldexp(v, shift));
        const uint32_t off = pu_swizzle((uint32_t)(q % ntile), (uint32_t)(k % PU_KB));
#pragma unroll
        for (int s = 0; s < PU_SLICES; ++s) {
            const long long dgt = ((x + 128) & 255) - 128;  // balanced digit in [-128, 127]
            x = (x - dgt) >> 8;
            base[(((size_t)nt * PU_SLICES + s) * kt_n + kt) * tile_bytes + off] = (uint8_t)(int8_t)dgt;
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
// Latency-critical hand-offs (accumulator full / free) poll with the non-suspending test_wait: try_wait may park the
// thread for a hardware-chosen interval before it re-checks.
__device__ __forceinline__ void pu_mbar_spin(uint64_t* bar, uint32_t parity, unsigned* status) {
    unsigned spins = 0;
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.b32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(pu_smem(bar)), "r"(parity)
            : "memory");
        if (ok) break;
        if (++spins > (1u << 28)) {
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
}
// 64 accumulator columns as 32 registers: .pack::16b keeps the low 16 bits of two adjacent 32-bit columns (|T| <= 128 m fits
// int16 for m <= 255 trees; larger forests take the 32-bit path)
__device__ __forceinline__ void pu_tmem_ld64_pack16(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void pu_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {  // 16 columns, 32-bit
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void pu_tmem_ld32_pack16(uint32_t taddr, uint32_t (&v)[16]) {  // 32 columns -> 16 registers
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void pu_tmem_ld16_pack16(uint32_t taddr, uint32_t (&v)[8]) {  // 16 columns -> 8 registers
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void pu_tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand of an M = 128 kind::i8 MMA is 128 lanes x 8 columns (32 K bytes per row,
// four per 32-bit column)
__device__ __forceinline__ void pu_umma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void pu_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct PuSmem {
    size_t off_ring, off_table, off_zmask, off_w, off_xs, off_xf, off_ft, off_mean, off_part, off_bars, total;
};
__host__ __device__ inline PuSmem pu_smem_layout(int m, int hi, int d, int kt, int ring_bytes, int nb) {
    PuSmem s;
    size_t o = 0;
    s.off_ring = o;   o += (size_t)ring_bytes;
    s.off_table = o;  o = align256(o + (size_t)m * hi * sizeof(WalkNode));
    s.off_zmask = o;  o = align256(o + (size_t)nb * PU_ROWS * kt * 2 * 8);  // [mask buffer][64-column word][row]
    s.off_w = o;      o = align256(o + (size_t)kt * PU_KB * 8);
    s.off_xs = o;     o = align256(o + (size_t)d * (PU_ROWS + 1) * 8);
    s.off_xf = o;     o = align256(o + (size_t)d * (PU_ROWS + 1) * 4);
    s.off_ft = o;     o = align256(o + (size_t)d * 4);
    s.off_mean = o;   o = align256(o + (size_t)nb * PU_WALK_GROUPS * PU_ROWS * 8);  // [mask buffer][group][row]
    s.off_part = o;   o = align256(o + (size_t)PU_SLICES * PU_ROWS * 4);           // second column half's row sums
    s.off_bars = o;   o = align256(o + (size_t)(2 * PU_MAX_STAGES + 9) * 8 + 16);
    s.total = o;
    return s;
}

#ifdef BARK_PHASE_TIMING
// per-role wait / work cycle totals of one CTA in the middle of the grid (debug build only)
#define PU_T0() do { pu_c = clock64(); } while (0)
#define PU_ACC(i) do { long long t_ = clock64(); pu_w[i] += t_ - pu_c; pu_c = t_; } while (0)
#define PU_REPORT(...) do { if (blockIdx.x == gridDim.x / 2 && blockIdx.y == 0) printf(__VA_ARGS__); } while (0)
#else
#define PU_T0() do { } while (0)
#define PU_ACC(i) do { } while (0)
#define PU_REPORT(...) do { } while (0)
#endif

__device__ __forceinline__ void pu_named_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

__global__ void __launch_bounds__(PU_THREADS, 1)
predict_umma_kernel(WsLayout lay, const void* ws, const WalkNode* __restrict__ table, const uint8_t* __restrict__ tiles,
                    const double* __restrict__ scales, int hi, int kt_n, int nt_n, int ntile, int ring_bytes, int nb,
                    const double* __restrict__ cand,
                    int64_t n_c, double* __restrict__ mu, double* __restrict__ var, int mix, double y_mean, double y_std,
                    int add_noise) {
    // mix != 0 (only when this CTA covers every sample, gridDim.y == 1): mu / var (n_c) receive the moments of the
    // mixture over the samples after un-standardisation, accumulated in registers in sample order -- the arithmetic of
    // predict_mixture_kernel (csrc/predict.cu) without the (samples, n_c) round trip through memory
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int d = (int)lay.d, m = (int)lay.m;
    const PuSmem sl = pu_smem_layout(m, hi, d, kt_n, ring_bytes, nb);  // nb = 2: the walk runs one sample ahead; 1: tight shared memory
    unsigned char* ring = smem_raw + sl.off_ring;
    WalkNode* tb = reinterpret_cast<WalkNode*>(smem_raw + sl.off_table);
    unsigned long long* zmask = reinterpret_cast<unsigned long long*>(smem_raw + sl.off_zmask);  // [parity][word][row]
    double* w_s = reinterpret_cast<double*>(smem_raw + sl.off_w);
    double* xs = reinterpret_cast<double*>(smem_raw + sl.off_xs);  // [d][PU_ROWS + 1]
    float* xf = reinterpret_cast<float*>(smem_raw + sl.off_xf);    // [d][PU_ROWS + 1]: candidates rounded UP to f32
    int* ftc = reinterpret_cast<int*>(smem_raw + sl.off_ft);
    double* meanp = reinterpret_cast<double*>(smem_raw + sl.off_mean);  // [parity][group][row]
    int* accp = reinterpret_cast<int*>(smem_raw + sl.off_part);          // [slice][row]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + sl.off_bars);
    uint64_t* empty_bar = full_bar + PU_MAX_STAGES;
    uint64_t* acc_full = empty_bar + PU_MAX_STAGES;  // [2] MMA -> epilogue
    uint64_t* acc_free = acc_full + 2;               // [2] epilogue -> MMA
    uint64_t* mask_ready = acc_free + 2;             // [2] walkers -> epilogue: masks / partial means of a sample complete
    uint64_t* mask_free = mask_ready + 2;            // [2] epilogue -> walkers: the sample that used the buffer is finished
    uint64_t* a_ready = mask_free + 2;               // [1] epilogue -> MMA: the A operand of the next sample is in place
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 1);
    const int tile_bytes = ntile * PU_KB;  // one K tile: ntile Binv columns x 128 K bytes
    const int stage_bytes = PU_GROUP * tile_bytes;
    const int nstages = min(PU_MAX_STAGES, ring_bytes / stage_bytes);
    // TMEM columns: [0, 32 kt) the one-hot A operand (row = lane, four K bytes per column), the two accumulators at the top
    const int acc_base = 512 - 2 * ntile;
    const int nwords = kt_n * 2;  // 64-column mask words per candidate row
    const int mask_words = PU_ROWS * nwords;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef BARK_PHASE_TIMING
    long long pu_w[6] = {0, 0, 0, 0, 0, 0}, pu_c = 0;
    __shared__ long long pu_tl[6][16];  // timeline of sample 20: [event][item]
    const bool pu_trace = (blockIdx.x == gridDim.x / 2 && blockIdx.y == 0);
#define PU_TL(ev, si_, it_) do { if (pu_trace && (si_) == 20 && (it_) < 16) pu_tl[ev][it_] = clock64(); } while (0)
#else
#define PU_TL(ev, si_, it_) do { } while (0)
#endif
    const int64_t p0 = (int64_t)blockIdx.x * PU_ROWS;
    const int np = (int)min((int64_t)PU_ROWS, n_c - p0);
    // samples of this CTA: gridDim.y splits them when there are fewer candidate tiles than SMs
    const int64_t S = lay.chains;
    const int64_t s_lo = S * blockIdx.y / gridDim.y, s_hi = S * (blockIdx.y + 1) / gridDim.y;
    const int ns = (int)(s_hi - s_lo);
    SharedView sv = shared_view(lay, ws);

    // ---- setup: barriers, TMEM, the candidate tile
    if (tid == 0) {
        for (int s = 0; s < PU_MAX_STAGES; ++s) { pu_mbar_init(full_bar + s, 1); pu_mbar_init(empty_bar + s, 1); }
        for (int b = 0; b < 2; ++b) {
            pu_mbar_init(acc_full + b, 1);
            pu_mbar_init(acc_free + b, PU_EPI_WARPS);
            pu_mbar_init(mask_ready + b, PU_WALK_WARPS);
            pu_mbar_init(mask_free + b, PU_EPI_WARPS);
        }
        pu_mbar_init(a_ready, PU_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == PU_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(pu_smem(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    bool any_cat = false;
    // x <= (double)thr with an f32 threshold  <=>  ru_f32(x) <= thr  (ru = the smallest f32 >= x): the numeric splits are
    // decided in FP32 -- exactly as the reference's comparison -- and keep the walk off the FP64 pipe
    for (int e = tid; e < np * d; e += PU_THREADS) {
        const double x = cand[p0 * d + e];
        xs[(size_t)(e % d) * (PU_ROWS + 1) + e / d] = x;
        xf[(size_t)(e % d) * (PU_ROWS + 1) + e / d] = __double2float_ru(x);
    }
    for (int e = tid; e < d; e += PU_THREADS) ftc[e] = sv.ft[e];
    for (int e = 0; e < d; ++e) any_cat |= (sv.ft[e] == FEAT_CAT);  // CTA-uniform: all-numeric problems skip the bitmask test
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;
    const int items = nt_n * PU_SLICES;
    (void)items;
    const int last_cols = kt_n * PU_KB - (nt_n - 1) * ntile;  // columns of the last column tile

    const int eh = pu_epilogue_half(warp), wi = pu_walker_index(warp);
    if (warp == PU_PRODUCER_WARP) {
        // ================= producer: digit tiles of sample after sample, items in order, K tiles kt >= pu_kt_lo(nt).
        // (Both single-thread loops below keep their bookkeeping incremental -- ring position and phase, source pointer,
        // descriptors by addition: one lone thread runs ~5 cycles per dependent instruction, and an integer division by
        // the runtime stage count per tile was costing more than the tile's four MMAs.)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 1;  // parity of the empty-barrier phase that must have completed: first lap needs none
            bool first_lap = true;
            for (int si = 0; si < ns; ++si) {
                const int64_t sample = s_lo + si;
                unsigned* st = &chain_view(lay, const_cast<void*>(ws), sample).sc->status;
                const uint8_t* src_item = tiles + (size_t)sample * nt_n * PU_SLICES * kt_n * tile_bytes;
                for (int nt = 0; nt < nt_n; ++nt) {
                    const uint32_t bytes = (uint32_t)((nt == nt_n - 1) ? last_cols : ntile) * PU_KB;
                    const int kt_lo = pu_kt_lo(nt, ntile);
                    for (int sl_ = 0; sl_ < PU_SLICES; ++sl_, src_item += (size_t)kt_n * tile_bytes) {
                        const uint8_t* src = src_item + (size_t)kt_lo * tile_bytes;
                        for (int kt = kt_lo; kt < kt_n; kt += PU_GROUP, src += (size_t)PU_GROUP * tile_bytes) {
                            // a full group is one contiguous copy of PU_GROUP tile slots (for a narrow last column
                            // tile the unused rows of the first slot travel along); a single tile copies its rows only
                            const uint32_t nbytes = (kt + PU_GROUP <= kt_n) ? (uint32_t)(PU_GROUP * tile_bytes) : bytes;
                            PU_T0();
                            if (!first_lap) pu_mbar_wait(empty_bar + stage, phase, st);
                            PU_ACC(0);
                            pu_mbar_expect_tx(full_bar + stage, nbytes);
                            pu_bulk_g2s(ring + (size_t)stage * stage_bytes, src, nbytes, full_bar + stage);
                            PU_ACC(1);
                            if (++stage == nstages) {
                                stage = 0;
                                phase ^= first_lap ? 0u : 1u;
                                if (first_lap) { first_lap = false; phase = 0; }
                            }
                        }
                    }
                }
            }
            PU_REPORT("pu_producer per sample: wait_empty %lld issue %lld\n", pu_w[0] / ns, pu_w[1] / ns);
        }
        __syncwarp();
    } else if (warp == PU_MMA_WARP) {
        // ================= MMA issuer: item = (column tile, digit plane) into TMEM buffer (running item index) & 1.
        // (Tried and dropped: doing the next tile's barrier waits between the second and third MMA of the current tile --
        // 9.6 instead of 10.8 M points/s; a flattened tile iterator -- 8.0 M: this lone thread is latency-bound on its own
        // scalar instructions, anything added to the loop costs more than it hides.  Two issuing threads, one per
        // accumulator buffer: 11.9 vs 12.4 M points/s.)
        if (lane == 0) {
            int stage = 0;
            uint32_t full_phase = 0;
            uint32_t git = 0;  // running item index: TMEM buffer git & 1, its use count git >> 1
            const uint64_t desc_hi = (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);  // pu_desc_sw128 without the address
            const uint64_t b_desc0 = desc_hi | (uint64_t)((pu_smem(ring) >> 4) & 0x3FFFu);
            const uint32_t idesc_full = pu_idesc_i8(PU_ROWS, ntile), idesc_last = pu_idesc_i8(PU_ROWS, last_cols);
            const uint32_t b_step = (uint32_t)(stage_bytes >> 4), t_step = (uint32_t)(tile_bytes >> 4);
            for (int si = 0; si < ns; ++si) {
                unsigned* st = &chain_view(lay, const_cast<void*>(ws), s_lo + si).sc->status;
                PU_T0();
                pu_mbar_wait(a_ready, (uint32_t)(si & 1), st);  // A operand of this sample written into TMEM
                PU_ACC(0);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int nt = 0; nt < nt_n; ++nt) {
                    const uint32_t idesc = (nt == nt_n - 1) ? idesc_last : idesc_full;
                    const int kt_lo = pu_kt_lo(nt, ntile);
                    for (int sl_ = 0; sl_ < PU_SLICES; ++sl_) {
                        const uint32_t buf = git & 1u, use = git >> 1;
                        PU_T0();
                        if (use > 0) pu_mbar_wait(acc_free + buf, (use - 1) & 1u, st);  // epilogue has drained the buffer
                        PU_ACC(1);
                        PU_TL(0, si, nt * PU_SLICES + sl_);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t d_addr = tmem_d + (uint32_t)acc_base + buf * (uint32_t)ntile;
                        uint32_t a_addr = tmem_d + (uint32_t)(kt_lo * (PU_KB / 4));  // 32 A columns per K tile, 8 per MMA
                        for (int kt = kt_lo; kt < kt_n; kt += PU_GROUP) {
                            PU_T0();
                            pu_mbar_wait(full_bar + stage, full_phase, st);
                            PU_ACC(2);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            uint64_t b_desc = b_desc0 + (uint64_t)((uint32_t)stage * b_step);
                            const int cnt = min(PU_GROUP, kt_n - kt);
                            for (int c = 0; c < cnt; ++c, a_addr += PU_KB / 4, b_desc += t_step) {
                                pu_umma_i8_ts(d_addr, a_addr, b_desc, idesc, (kt + c > kt_lo) ? 1u : 0u);
                                pu_umma_i8_ts(d_addr, a_addr + 8, b_desc + 2, idesc, 1u);
                                pu_umma_i8_ts(d_addr, a_addr + 16, b_desc + 4, idesc, 1u);
                                pu_umma_i8_ts(d_addr, a_addr + 24, b_desc + 6, idesc, 1u);
                            }
                            // the stage of an item's last group is released by the epilogue when it sees acc_full
                            if (kt + PU_GROUP < kt_n) pu_commit(empty_bar + stage);
                            PU_ACC(3);
                            if (++stage == nstages) { stage = 0; full_phase ^= 1u; }
                        }
                        pu_commit(acc_full + buf);
                        PU_TL(1, si, nt * PU_SLICES + sl_);
                        ++git;
                    }
                }
            }
            PU_REPORT("pu_mma per sample: wait_a_ready %lld wait_acc_free %lld wait_full %lld issue %lld\n", pu_w[0] / ns, pu_w[1] / ns, pu_w[2] / ns,
                      pu_w[3] / ns);
        }
        __syncwarp();
    } else if (wi >= 0) {
        // ================= walkers: PU_WALK_GROUPS threads per candidate, one sample ahead of the MMAs
        const int wt = wi * 32 + lane;  // 0 .. 32 * PU_WALK_WARPS - 1
        const int wg = wt / PU_ROWS, row = wt % PU_ROWS;
        constexpr int WALKERS = 32 * PU_WALK_WARPS;
        for (int si = 0; si < ns; ++si) {
            const int64_t sample = s_lo + si;
            const int par = si % nb, gen = si / nb;  // mask buffer and how often it has been used before
            ChainView cv = chain_view(lay, const_cast<void*>(ws), sample);
            unsigned long long* zm = zmask + (size_t)par * mask_words;
            PU_T0();
            if (gen > 0) pu_mbar_wait(mask_free + par, (uint32_t)((gen - 1) & 1), &cv.sc->status);
            PU_ACC(0);
            {
                const WalkNode* src = table + (size_t)sample * m * hi;
                for (int e = wt; e < m * hi; e += WALKERS) tb[e] = src[e];
                for (int e = wt; e < kt_n * PU_KB; e += WALKERS) w_s[e] = (e < lay.P) ? cv.w[e] : 0.0;
                for (int e = wt; e < mask_words; e += WALKERS) zm[e] = 0ull;
            }
            pu_named_sync(2, WALKERS);
            PU_ACC(1);
            double mean = 0.0;
            if (row < np) {
                const double* xp = xs + row;
                const float* xfp = xf + row;
                unsigned* zm32 = reinterpret_cast<unsigned*>(zm);
                // four trees in flight per thread, branch-free (a leaf steps to itself) so that the four pointer chases
                // overlap; the partial mean keeps the order of t
                for (int t0 = wg; t0 < m; t0 += 4 * PU_WALK_GROUPS) {
                    WalkNode nd[4];
                    const WalkNode* wn[4];
                    uint32_t cur[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        wn[u] = tb + (size_t)min(t0 + u * PU_WALK_GROUPS, m - 1) * hi;
                        nd[u] = wn[u][0];
                        cur[u] = 0;
                    }
                    for (int it = 0; it < hi; ++it) {
                        if ((nd[0].feat_leaf & nd[1].feat_leaf & nd[2].feat_leaf & nd[3].feat_leaf) & 0x8000u) break;
                        float xv[4];
                        int fi[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            fi[u] = min((int)(nd[u].feat_leaf & 0x7fffu), d - 1);  // (a leaf's feature bits are ignored below)
                            xv[u] = xfp[(size_t)fi[u] * (PU_ROWS + 1)];
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            bool left = xv[u] <= nd[u].thr;
                            if (any_cat && ftc[fi[u]] == FEAT_CAT) left = goes_left(xp[(size_t)fi[u] * (PU_ROWS + 1)], nd[u].thr, FEAT_CAT);
                            const uint32_t at = min((uint32_t)(left ? nd[u].left : nd[u].right), (uint32_t)(hi - 1));
                            cur[u] = (nd[u].feat_leaf & 0x8000u) ? cur[u] : at;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) nd[u] = wn[u][cur[u]];
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (t0 + u * PU_WALK_GROUPS >= m) break;
                        const int col = __float_as_int(nd[u].thr);
                        mean += w_s[col];
                        atomicOr(zm32 + 2 * ((size_t)(col >> 6) * PU_ROWS + row) + ((col >> 5) & 1), 1u << (col & 31));
                    }
                }
            }
            meanp[((size_t)par * PU_WALK_GROUPS + wg) * PU_ROWS + row] = mean;
            pu_named_sync(2, WALKERS);  // every walker is done with the table / w of this sample (and its masks are written)
            if (lane == 0) pu_mbar_arrive(mask_ready + par);
            PU_ACC(2);
        }
        if (wt == 0) PU_REPORT("pu_walker per sample: wait_mask_free %lld stage %lld walk %lld\n", pu_w[0] / ns, pu_w[1] / ns, pu_w[2] / ns);
    } else if (eh >= 0) {
        // ================= epilogue warps: A operand of the next sample, masked int32 row sums of every digit plane
        const int quarter = warp & 3, half = eh;  // TMEM lane quarter of this warp; which half of an item's columns
        const int row = 32 * quarter + lane;
        const int et = (half * 4 + quarter) * 32 + lane;  // 0 .. 255
        constexpr int EPI = 32 * PU_EPI_WARPS;
        const bool narrow = (m * 128 <= 32767);  // |T_k| <= 128 m fits int16

        auto build_a = [&](int si) {
            // one-hot A operand of sample si into TMEM from the masks
            unsigned* st = &chain_view(lay, const_cast<void*>(ws), s_lo + si).sc->status;
            PU_T0();
            pu_mbar_wait(mask_ready + si % nb, (uint32_t)((si / nb) & 1), st);
            PU_ACC(0);
            const unsigned long long* zm = zmask + (size_t)(si % nb) * mask_words;
            // thread = candidate row = TMEM lane of its warp's quarter; one 64-bit mask word = 64 K bytes = 16 columns per
            // tcgen05.st; the two warps of a quarter share the words
            for (int wd = half * kt_n; wd < (half + 1) * kt_n; ++wd) {
                const unsigned long long bits = zm[(size_t)wd * PU_ROWS + row];
                uint32_t v[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) v[c] = (((uint32_t)(bits >> (4 * c)) & 0xFu) * 0x00204081u) & 0x01010101u;
                pu_tmem_st16(tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(wd * 16), v);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) pu_mbar_arrive(a_ready);
            PU_ACC(1);
        };

        if (ns > 0) build_a(0);
        double mix_sm = 0.0, mix_s2 = 0.0;
        int git = 0;
        int ep_stage = 0;  // ring position of the next item's first K tile (same walk as the producer / MMA issuer)
        for (int si = 0; si < ns; ++si) {
            const int64_t sample = s_lo + si;
            const int par = si % nb;
            ChainView cv = chain_view(lay, const_cast<void*>(ws), sample);
            unsigned* st = &cv.sc->status;
            const unsigned long long* zm = zmask + (size_t)par * mask_words;
            int acc[PU_SLICES];
#pragma unroll
            for (int s = 0; s < PU_SLICES; ++s) acc[s] = 0;
            const int hcols = ntile / 2;  // this warp's share of an item's columns: 96 (64 + 32) or 80 (64 + 16)
            for (int nt = 0; nt < nt_n; ++nt) {
                const int ncols = (nt == nt_n - 1) ? last_cols : ntile;
                // the row's mask bits of this warp's columns [o, o + hcols), those beyond the tile's width cleared
                const int o = nt * ntile + half * hcols;
                const int valid = max(0, min(hcols, ncols - half * hcols));
                const int item_tiles = (kt_n - pu_kt_lo(nt, ntile) + PU_GROUP - 1) / PU_GROUP;  // ring stages of the item
                unsigned long long bits0, bits1;
                {
                    const int wi = o >> 6, sh = o & 63;
                    const unsigned long long a = (wi < nwords) ? zm[(size_t)wi * PU_ROWS + row] : 0ull;
                    const unsigned long long b = (wi + 1 < nwords) ? zm[(size_t)(wi + 1) * PU_ROWS + row] : 0ull;
                    const unsigned long long c = (wi + 2 < nwords) ? zm[(size_t)(wi + 2) * PU_ROWS + row] : 0ull;
                    bits0 = sh ? ((a >> sh) | (b << (64 - sh))) : a;
                    bits1 = sh ? ((b >> sh) | (c << (64 - sh))) : b;
                    if (valid < 64) { bits0 = valid > 0 ? (bits0 & ((1ull << valid) - 1ull)) : 0ull; bits1 = 0ull; }
                    else bits1 = (valid > 64) ? (bits1 & ((1ull << (valid - 64)) - 1ull)) : 0ull;
                }
#pragma unroll
                for (int s = 0; s < PU_SLICES; ++s, ++git) {
                    const int buf = git & 1, use = git >> 1;
                    PU_T0();
                    pu_mbar_wait(acc_full + buf, (uint32_t)(use & 1), st);
                    PU_ACC(2);
                    if (et == 0) PU_TL(2, si, nt * PU_SLICES + s);
                    {   // the item's MMAs are complete: release the ring stage of its last K tile
                        ep_stage += item_tiles;
                        while (ep_stage >= nstages) ep_stage -= nstages;
                        if (et == 0) pu_mbar_arrive(empty_bar + (ep_stage == 0 ? nstages - 1 : ep_stage - 1));
                    }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t taddr = tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc_base + buf * ntile + half * hcols);
                    int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
                    if (__any_sync(0xffffffffu, (bits0 | bits1) != 0ull)) {
                        if (narrow) {
                            // packed loads (two int16 columns per register), all issued before the single wait; one dp2a
                            // per register against the row's 0 / 1 mask bytes
                            uint32_t v0[32], v1[16];
                            pu_tmem_ld64_pack16(taddr, v0);
                            if (hcols == 96) {
                                pu_tmem_ld32_pack16(taddr + 64, v1);
                            } else {
                                uint32_t v8[8];
                                pu_tmem_ld16_pack16(taddr + 64, v8);
#pragma unroll
                                for (int r = 0; r < 8; ++r) { v1[r] = v8[r]; v1[8 + r] = 0u; }
                            }
                            pu_tmem_wait_ld();
                            if (et == 0) PU_TL(3, si, nt * PU_SLICES + s);
#pragma unroll
                            for (int r = 0; r < 16; ++r) {
                                const uint32_t m0 = (((uint32_t)(bits0 >> (4 * r)) & 0xFu) * 0x00204081u) & 0x01010101u;
                                a0 = __dp2a_lo((int)v0[2 * r], (int)m0, a0);
                                a1 = __dp2a_hi((int)v0[2 * r + 1], (int)m0, a1);
                            }
#pragma unroll
                            for (int r = 0; r < 8; ++r) {
                                const uint32_t m1 = (((uint32_t)(bits1 >> (4 * r)) & 0xFu) * 0x00204081u) & 0x01010101u;
                                a2 = __dp2a_lo((int)v1[2 * r], (int)m1, a2);
                                a3 = __dp2a_hi((int)v1[2 * r + 1], (int)m1, a3);
                            }
                        } else {
                            // forests of more than 255 trees: 32-bit accumulators, 16 columns at a time
                            for (int j = 0; j * 16 < valid; ++j) {
                                uint32_t v[16];
                                pu_tmem_ld16(taddr + 16 * j, v);
                                pu_tmem_wait_ld();
                                const uint32_t b = (uint32_t)((j < 4 ? bits0 : bits1) >> (16 * (j & 3))) & 0xFFFFu;
#pragma unroll
                                for (int c = 0; c < 16; ++c) a0 += (int)v[c] * (int)((b >> c) & 1u);
                            }
                        }
                    }
                    acc[s] += (a0 + a1) + (a2 + a3);
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) pu_mbar_arrive(acc_free + buf);
                    if (et == 0) PU_TL(4, si, nt * PU_SLICES + s);
                    PU_ACC(3);
                }
            }
            PU_T0();
            // every MMA of this sample has completed (the last acc_full): the A operand may be rewritten.  (With a single
            // mask buffer the next sample's masks only appear after this sample's are released: build A at the end.
            // Building A before the drain of the last item instead measured no gain.)
            if (nb > 1 && si + 1 < ns) build_a(si + 1);
            if (half == 1) {
#pragma unroll
                for (int s = 0; s < PU_SLICES; ++s) accp[s * PU_ROWS + row] = acc[s];
            }
            pu_named_sync(1, EPI);
            if (half == 0 && row < np) {
                // z^T Binv z = 2^-shift * sum_k 256^k acc_k   (each acc_k exact)
                double tsum = 0.0;
#pragma unroll
                for (int s = PU_SLICES - 1; s >= 0; --s) tsum = tsum * 256.0 + (double)(acc[s] + accp[s * PU_ROWS + row]);
                double mean = 0.0;
#pragma unroll
                for (int g = 0; g < PU_WALK_GROUPS; ++g) mean += meanp[((size_t)par * PU_WALK_GROUPS + g) * PU_ROWS + row];
                const double v = cv.sc->sig * (tsum * scales[sample]);
                if (mix) {
                    const double mj = mean * y_std + y_mean;
                    double vj = v * (y_std * y_std);
                    if (add_noise) vj += cv.sc->noise;
                    mix_sm += mj;
                    mix_s2 += vj + mj * mj;
                } else {
                    const int64_t o = sample * n_c + p0 + row;
                    mu[o] = mean;
                    var[o] = v;
                }
            }
            pu_named_sync(1, EPI);  // accp / meanp / masks of this sample are no longer needed
            if (lane == 0) pu_mbar_arrive(mask_free + par);
            PU_ACC(4);
            if (nb == 1 && si + 1 < ns) build_a(si + 1);
        }
        if (mix && half == 0 && row < np) {
            const double e = mix_sm / (double)S;
            mu[p0 + row] = e;
            var[p0 + row] = mix_s2 / (double)S - e * e;
        }
        if (et == 0) PU_REPORT("pu_epilogue per sample: wait_mask_ready %lld build_a %lld wait_acc_full %lld drain %lld output %lld\n", pu_w[0] / ns,
                               pu_w[1] / ns, pu_w[2] / ns, pu_w[3] / ns, pu_w[4] / ns);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == PU_MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512u) : "memory");
#ifdef BARK_PHASE_TIMING
    if (pu_trace && tid == 0 && ns > 20)
        for (int it = 0; it < items && it < 16; ++it)
            printf("pu_tl item %2d: mma_start %6lld mma_issued %6lld | acc_full_seen %6lld loads_done %6lld arrived %6lld\n", it,
                   pu_tl[0][it] - pu_tl[0][0], pu_tl[1][it] - pu_tl[0][0], pu_tl[2][it] - pu_tl[0][0], pu_tl[3][it] - pu_tl[0][0],
                   pu_tl[4][it] - pu_tl[0][0]);
#endif
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
    pu_slice_kernel<<<grid, 256, 0, st>>>(lay, workspace, pl.kt, pl.nt, pl.ntile, base + pl.off_tiles, (double*)(base + pl.off_scale));
    BARK_LAUNCH_CHECK();
    return BARK_OK;
}

static int predict_umma_launch(const char* __func_name, const bark_mcmc_dims* dims, const void* workspace, const void* prep, int32_t slots,
                               int32_t p_max, const double* candidates, int64_t n_c, double* mu, double* var, int mix, double y_mean,
                               double y_std, int add_noise, void* stream) {
    (void)__func_name;
    const WsLayout lay = make_layout(*dims);
    const PrepLayout pl = prep_layout(dims->chains, dims->m, slots, p_max);
    const int stage_bytes = PU_GROUP * pl.ntile * PU_KB;
    // two mask buffers (the walk runs a sample ahead of the MMAs) when they fit beside a ring of >= 2 stages, else one
    int nb = 2;
    size_t fixed = pu_smem_layout((int)lay.m, slots, (int)lay.d, pl.kt, 0, nb).total;
    if (fixed + 2 * (size_t)stage_bytes > 227 * 1024) {
        nb = 1;
        fixed = pu_smem_layout((int)lay.m, slots, (int)lay.d, pl.kt, 0, nb).total;
    }
    BARK_CHECK_ARG(fixed + 2 * (size_t)stage_bytes <= 227 * 1024, "m * slots / d / p_max too large for the predict kernel's shared memory");
    const int ring_bytes = (int)(std::min<size_t>(PU_RING_MAX, 227 * 1024 - fixed) / stage_bytes) * stage_bytes;
    const PuSmem sl = pu_smem_layout((int)lay.m, slots, (int)lay.d, pl.kt, ring_bytes, nb);
    BARK_CUDA(cudaFuncSetAttribute(predict_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sl.total));
    const unsigned char* base = (const unsigned char*)prep;
    // one persistent CTA per candidate tile loops over the samples; with fewer tiles than SMs the samples are split
    const int64_t tiles_n = ceil_div(n_c, PU_ROWS);
    const int64_t ysplit = mix ? 1 : std::max<int64_t>(1, std::min<int64_t>(dims->chains, 148 / tiles_n));
    dim3 grid((unsigned)tiles_n, (unsigned)ysplit);
    predict_umma_kernel<<<grid, PU_THREADS, sl.total, (cudaStream_t)stream>>>(
        lay, workspace, (const WalkNode*)(base + pl.off_table), base + pl.off_tiles, (const double*)(base + pl.off_scale),
        slots, pl.kt, pl.nt, pl.ntile, ring_bytes, nb, candidates, n_c, mu, var, mix, y_mean, y_std, add_noise);
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
    return predict_umma_launch(__func__, dims, workspace, prep, slots, p_max, candidates, n_c, mu, var, 0, 0.0, 1.0, 0, stream);
}

// moments (n_c) of the mixture over the samples, un-standardised (bark_predict mode 1), folded inside the kernel
int bark_predict_umma_mixture(const bark_mcmc_dims* dims, const void* workspace, const void* prep, int32_t slots, int32_t p_max,
                              const double* candidates, int64_t n_c, double y_mean, double y_std, int add_noise, double* mu,
                              double* var, void* stream) {
    BARK_CHECK_ARG(dims && workspace && prep, "null pointer");
    BARK_CHECK_ARG(n_c >= 0, "n_c < 0");
    if (n_c == 0) return BARK_OK;
    BARK_CHECK_ARG(candidates && mu && var, "null pointer");
    BARK_CHECK_ARG(slots >= 1 && slots <= 255 && p_max >= 1 && p_max <= PU_MAX_P, "slots / p_max out of range");
    BARK_CHECK_ARG(ceil_div(n_c, PU_ROWS) <= 2147483647LL, "too many candidates per call");
    return predict_umma_launch(__func__, dims, workspace, prep, slots, p_max, candidates, n_c, mu, var, 1, y_mean, y_std,
                               add_noise ? 1 : 0, stream);
}

}  // extern "C"
