// a4 on the 5th-generation tensor cores: leaf co-occurrence counts as an exact int8 one-hot GEMM.
//
//   count[i,j] = #{t : leaf_t(x_i) == leaf_t(x'_j)} = sum_k Za[i,k] * Zb[j,k],   k = column of (tree t, leaf slot),
// with Z the {0,1} leaf-indicator matrix.  int8 x int8 products accumulate in s32, so the tensor-core result is the
// exact integer count (src/bark/forest.py:85-88).  Only the (tree, slot) pairs that some row actually lands in get a
// column: the K extent is the forest's leaf count (about 2.3 m), not m * slots (a 12 x shorter K loop at config 4).
//
// Pipeline
//   1. leaf_presence_kernel / leaf_columns_kernel  mark the leaf slots that occur per (forest, tree) and number them
//      consecutively over the forest (column base per tree + rank inside the tree's presence mask).
//   2. onehot_build_kernel  one CTA per (forest, 128-row tile): builds the int8 tiles in shared memory and writes them
//      pre-tiled and pre-swizzled -- every (128 rows x 128 K-bytes) tile is one contiguous 16 KB block holding exactly
//      the shared-memory image that a K-major SWIZZLE_128B UMMA operand needs (16-byte chunk c of row r stored at
//      chunk c ^ (r & 7) of its 1 KB 8-row group).  No memset, no scattered byte stores to global memory.
//   3. gram_umma_kernel     one CTA per 128 x 128 output tile: an elected thread streams the A/B tiles of every
//      K chunk with the bulk-copy engine (TMA, cp.async.bulk -> mbarrier complete_tx) through a ring and
//      issues tcgen05.mma.cta_group::1.kind::i8 (M = N = 128, K = 32, four per chunk) into a 128-column s32 TMEM
//      accumulator; tcgen05.commit releases the stages and finally signals the epilogue.
//   4. epilogue             4 warps read TMEM (tcgen05.ld 32x32b.x32), emit the int32 counts and / or the FP64
//      kernel matrix K = scale * ((1/m) * count) + (jitter + noise) I with un-fused multiplies (bit-exact vs numpy).
#include <algorithm>

#include "common.cuh"

namespace bark {

constexpr int UT = 128;                 // tile edge (UMMA M = N = 128)
constexpr int UK = 128;                 // K bytes per tile (one SWIZZLE_128B atom row)
constexpr int TILE_BYTES = UT * UK;     // 16 KB
constexpr int U_STAGES = 2;             // 64 KB of operand ring: three CTAs per SM, so one tile's epilogue overlaps the others' loads and MMAs
constexpr int U_THREADS = 128;

__device__ __forceinline__ uint32_t u_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row r in [0,128), K byte kb in [0,128)) inside a pre-swizzled 16 KB tile
__host__ __device__ __forceinline__ uint32_t swizzled_offset(uint32_t r, uint32_t kb) {
    const uint32_t g = r >> 3, rr = r & 7, chunk = kb >> 4, b = kb & 15;
    return g * 1024u + rr * 128u + ((chunk ^ rr) << 4) + b;
}

constexpr int PRES_ROWS = 32;   // rows folded into one register mask before the atomics
constexpr int OB_ROWS = 32;     // rows per CTA of onehot_build_kernel (4 swizzle groups = 4 KB of every K tile)
constexpr int OB_GROUP = 12;    // K tiles built per pass of onehot_build_kernel (48 KB of shared memory)
constexpr int OB_THREADS = 256;
constexpr int OB_MAP_WORDS = 4096;  // presence words + column bases staged per forest when they fit (16 KB)

// pres[(b * m + t) * W + w] |= bit(slot) for every leaf slot a row of forest b lands in (W = words per tree).
// One thread per (forest, tree, block of PRES_ROWS rows): coalesced over t, one atomicOr per non-zero word.
__global__ void leaf_presence_kernel(const uint32_t* __restrict__ leaves, int64_t batch, int64_t n, int64_t m, int slots, int W,
                                     uint32_t* __restrict__ pres, uint32_t* __restrict__ status) {
    const int64_t rblocks = (n + PRES_ROWS - 1) / PRES_ROWS;
    const int64_t total = batch * rblocks * m;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = e % m, rb = (e / m) % rblocks, b = e / (m * rblocks);
        const int64_t i0 = rb * PRES_ROWS;
        const int cnt = (int)min((int64_t)PRES_ROWS, n - i0);
        const uint32_t* src = leaves + (b * n + i0) * m + t;
        uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        bool bad = false;
#pragma unroll 1
        for (int i = 0; i < PRES_ROWS; i += 8) {
            uint32_t id[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) id[u] = (i + u < cnt) ? __ldg(src + (int64_t)(i + u) * m) : 0xffffffffu;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (id[u] >= (uint32_t)slots) {
                    bad |= (i + u < cnt);
                    continue;
                }
                const uint32_t bit = 1u << (id[u] & 31), wi = id[u] >> 5;
#pragma unroll
                for (int q = 0; q < 8; ++q) w[q] |= (wi == (uint32_t)q) ? bit : 0u;
            }
        }
        if (bad && status) atomicOr(status, 1u);
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (q < W && w[q]) atomicOr(pres + (b * m + t) * W + q, w[q]);
    }
}

// One warp per forest: base[b * m + t] = number of occupied (tree, slot) columns before tree t; kt[b] = K tiles used.
__global__ void leaf_columns_kernel(const uint32_t* __restrict__ pres, int64_t batch, int64_t m, int W,
                                    int32_t* __restrict__ base, int32_t* __restrict__ kt) {
    const int64_t b = blockIdx.x;
    const int lane = threadIdx.x;
    int run = 0;
    for (int64_t t0 = 0; t0 < m; t0 += 32) {
        const int64_t t = t0 + lane;
        int c = 0;
        if (t < m)
            for (int q = 0; q < W; ++q) c += __popc(pres[(b * m + t) * W + q]);
        int inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (t < m) base[b * m + t] = run + inc - c;
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) kt[b] = max(1, (run + UK - 1) / UK);
}

// One CTA per (block of OB_ROWS rows, forest): the rows' one-hot bytes for up to OB_GROUP K tiles at a time are set in
// shared memory (already in the swizzled operand image: 4 KB of every tile) and written out with coalesced 16-byte
// stores.  The forest's presence words and column bases are staged in shared memory when they fit.
__global__ void __launch_bounds__(OB_THREADS)
onehot_build_kernel(const uint32_t* __restrict__ leaves, int64_t n, int64_t m, int slots, int W, int64_t row_tiles,
                    int64_t k_tiles, const uint32_t* __restrict__ pres, const int32_t* __restrict__ base,
                    const int32_t* __restrict__ kt, uint8_t* __restrict__ Z, int stage_map) {
    extern __shared__ __align__(16) unsigned char ob_smem[];
    constexpr int PIECE = OB_ROWS * UK;  // bytes of one K tile owned by this CTA
    const int64_t rb = blockIdx.x, b = blockIdx.y;
    const int ktb = kt[b];
    const int64_t row0 = rb * OB_ROWS;
    const int rows = (int)max((int64_t)0, min((int64_t)OB_ROWS, n - row0));  // 0: padding rows of the last tile (zeroed)
    const int64_t rt = row0 / UT;
    const uint32_t rin = (uint32_t)(row0 % UT);  // first row inside the 128-row tile (a multiple of 8)
    uint8_t* out = Z + ((b * row_tiles + rt) * k_tiles) * (int64_t)TILE_BYTES + rin * UK;
    const int gmax = (int)min((int64_t)OB_GROUP, k_tiles);
    uint32_t* s_map = reinterpret_cast<uint32_t*>(ob_smem + (size_t)gmax * PIECE);
    const uint32_t* pw_all = pres + b * m * W;
    const int32_t* bs_all = base + b * m;
    if (stage_map) {
        for (int e = threadIdx.x; e < (int)m * W; e += OB_THREADS) s_map[e] = pw_all[e];
        for (int e = threadIdx.x; e < (int)m; e += OB_THREADS) s_map[(int)m * W + e] = (uint32_t)bs_all[e];
        pw_all = s_map;
        bs_all = reinterpret_cast<const int32_t*>(s_map + (int)m * W);
    }
    for (int g0 = 0; g0 < ktb; g0 += OB_GROUP) {
        const int gt = min(OB_GROUP, ktb - g0);
        uint4* z4 = reinterpret_cast<uint4*>(ob_smem);
        for (int e = threadIdx.x; e < gt * (PIECE / 16); e += OB_THREADS) z4[e] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        const uint32_t cells = (uint32_t)rows * (uint32_t)m, mu = (uint32_t)m;
        const uint32_t* src = leaves + (b * n + row0) * m;
#pragma unroll 1
        for (uint32_t e0 = threadIdx.x; e0 < cells; e0 += 4 * OB_THREADS) {
            uint32_t id[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t e = e0 + u * OB_THREADS;
                id[u] = (e < cells) ? __ldg(src + e) : 0xffffffffu;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (id[u] >= (uint32_t)slots) continue;  // out of range (flagged by leaf_presence_kernel) or past the end
                const uint32_t e = e0 + u * OB_THREADS;
                const uint32_t r = e / mu, t = e - r * mu;
                const uint32_t* pw = pw_all + t * W;
                int k = bs_all[t] + __popc(pw[id[u] >> 5] & ((1u << (id[u] & 31)) - 1u));
                for (uint32_t q = 0; q < (id[u] >> 5); ++q) k += __popc(pw[q]);
                const int tile = (k >> 7) - g0;
                // rows rin + r of the tile: this CTA's piece starts at row rin, so the swizzled offset of the local row
                // (a multiple of 8 rows = whole 1 KB groups) is the same expression on r
                if (tile >= 0 && tile < gt) ob_smem[tile * PIECE + swizzled_offset(r, (uint32_t)(k & 127))] = 1;
            }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < gt * (PIECE / 16); e += OB_THREADS) {
            const int tile = e / (PIECE / 16), o = e % (PIECE / 16);
            reinterpret_cast<uint4*>(out + (int64_t)(g0 + tile) * TILE_BYTES)[o] = z4[e];
        }
        __syncthreads();
    }
}

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ void u_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(u_smem(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void u_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(u_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool u_mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.b32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(u_smem(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a barrier that does not complete within ~2^26 probes (seconds) sets bit 1 of the status word (when the
// caller passed one) and gives up instead of hanging the GPU; the output of that tile is then invalid.
__device__ __forceinline__ void u_mbar_wait(uint64_t* bar, uint32_t parity, uint32_t* status = nullptr) {
    unsigned spins = 0;
    while (!u_mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) {
            if (status) atomicOr(status, 2u);
            break;
        }
    }
}
__device__ __forceinline__ void u_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(u_smem(dst)),
                 "l"(src), "r"(bytes), "r"(u_smem(bar))
                 : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in
// bits [0,14), LBO (ignored for swizzled K-major) = 1 in [16,30), SBO = 1024 B >> 4 in [32,46), version 1 in
// [46,48), layout type 2 (SWIZZLE_128B) in [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::i8: D = S32 (2 @ bit 4), A = B = signed int8
// (1 @ bit 7, 1 @ bit 10), both K-major (bits 15, 16 = 0), N >> 3 @ bit 17, M >> 4 @ bit 24.
__device__ __forceinline__ uint32_t umma_idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(u_smem(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
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

struct GramEpilogue {
    int32_t* counts;      // may be null
    double* K;            // may be null
    const double* scale;  // per batch (K only)
    const double* noise;  // per batch or null
    double inv_m, jitter;
    int add_diag;
    uint32_t* status;     // may be null; bit 1: a pipeline wait timed out
};

__global__ void __launch_bounds__(U_THREADS, 3)
gram_umma_kernel(const uint8_t* __restrict__ Za, const uint8_t* __restrict__ Zb, int64_t na, int64_t nb, int64_t rt_a,
                 int64_t rt_b, int64_t k_tiles, const int32_t* __restrict__ kt_used, GramEpilogue ep) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* tiles = smem_raw;  // U_STAGES x (A tile | B tile)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)U_STAGES * 2 * TILE_BYTES);
    uint64_t* empty_bar = full_bar + U_STAGES;
    uint64_t* acc_bar = empty_bar + U_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t b = blockIdx.z, tr = blockIdx.y, tc = blockIdx.x;

    if (tid == 0) {
        for (int s = 0; s < U_STAGES; ++s) { u_mbar_init(full_bar + s, 2); u_mbar_init(empty_bar + s, 1); }  // full: A and B producers
        u_mbar_init(acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // one warp allocates 128 TMEM columns (s32 accumulator, 128 lanes x 128 columns)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(u_smem(tmem_slot)), "r"(128u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    // ---- roles: thread 0 issues the MMAs, lane 0 of warps 1 and 2 stream the A and the B tiles (one thread issues one
    // bulk copy per ~620 cycles whatever its size -- scripts/tma_feed.cu -- so the two operands get a thread each)
    const int64_t KT = kt_used[b];  // occupied K tiles of this forest (tile stride stays k_tiles)
    if ((warp == 1 || warp == 2) && lane == 0) {
        const int which = warp - 1;  // 0: A, 1: B
        const uint8_t* src = which == 0 ? Za + ((b * rt_a + tr) * k_tiles) * (int64_t)TILE_BYTES
                                        : Zb + ((b * rt_b + tc) * k_tiles) * (int64_t)TILE_BYTES;
        for (int64_t it = 0; it < KT; ++it) {
            const int s = (int)(it % U_STAGES);
            if (it >= U_STAGES) u_mbar_wait(empty_bar + s, (uint32_t)((it / U_STAGES - 1) & 1), ep.status);
            u_mbar_expect_tx(full_bar + s, TILE_BYTES);
            u_bulk_g2s(tiles + (size_t)s * 2 * TILE_BYTES + (size_t)which * TILE_BYTES, src + it * TILE_BYTES, TILE_BYTES, full_bar + s);
        }
    } else if (tid == 0) {
        const uint32_t idesc = umma_idesc_i8(UT, UT);
        for (int64_t kc = 0; kc < KT; ++kc) {
            const int s = (int)(kc % U_STAGES);
            u_mbar_wait(full_bar + s, (uint32_t)((kc / U_STAGES) & 1), ep.status);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_addr = u_smem(tiles + (size_t)s * 2 * TILE_BYTES);
            const uint32_t b_addr = a_addr + TILE_BYTES;
#pragma unroll
            for (int k4 = 0; k4 < UK / 32; ++k4)
                umma_i8(tmem_d, umma_desc_sw128(a_addr + k4 * 32), umma_desc_sw128(b_addr + k4 * 32), idesc, (kc > 0 || k4 > 0) ? 1u : 0u);
            umma_commit(empty_bar + s);  // stage free once these MMAs have read it
        }
        umma_commit(acc_bar);  // accumulator complete
    }
    __syncwarp();

    // ---- epilogue: TMEM -> registers -> shared-memory transpose -> global.  tcgen05.ld hands every thread ONE
    // accumulator row (32 consecutive columns); stored as such, a warp's store instruction would touch 32 different
    // rows (32 sectors for 256 bytes).  Each warp therefore turns its 32 x 32 block through a padded shared-memory
    // tile, so that one store instruction writes 32 consecutive columns of one row (one 256-byte / 128-byte segment).
    u_mbar_wait(acc_bar, 0, ep.status);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // (the operand ring is idle once the accumulator is complete: it doubles as the transpose buffer)
    double* stage = reinterpret_cast<double*>(tiles) + (size_t)warp * 32 * 33;
    int32_t* stage_i = reinterpret_cast<int32_t*>(stage);
    const int64_t row0 = tr * UT + warp * 32;
    const double sc = ep.K ? ep.scale[b] : 0.0;
    const double dg = (ep.K && ep.add_diag) ? __dadd_rn(ep.jitter, ep.noise[b]) : 0.0;
#pragma unroll 1
    for (int c0 = 0; c0 < UT; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        const int64_t col = tc * UT + c0 + lane;  // this lane's column in the transposed phase
        if (ep.counts) {
#pragma unroll
            for (int j = 0; j < 32; ++j) stage_i[lane * 33 + j] = (int32_t)v[j];
            __syncwarp();
            if (col < nb) {
#pragma unroll 4
                for (int r = 0; r < 32; ++r)
                    if (row0 + r < na) ep.counts[(b * na + row0 + r) * nb + col] = stage_i[r * 33 + lane];
            }
            __syncwarp();
        }
        if (ep.K) {
#pragma unroll
            for (int j = 0; j < 32; ++j) stage[lane * 33 + j] = __dmul_rn(sc, __dmul_rn(ep.inv_m, (double)(int32_t)v[j]));
            __syncwarp();
            if (col < nb) {
#pragma unroll 4
                for (int r = 0; r < 32; ++r) {
                    if (row0 + r < na) {
                        double x = stage[r * 33 + lane];
                        if (ep.add_diag && col == row0 + r) x = __dadd_rn(x, dg);
                        ep.K[(b * na + row0 + r) * nb + col] = x;
                    }
                }
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(128u) : "memory");
}

// Tile counts.  The K extent allocated is the largest number of distinct leaf columns the rows can occupy:
// at most `slots` per tree, and at most one per row and tree.
static void gram_dims(int64_t na, int64_t nb, int64_t m, int slots, int64_t* rt_a, int64_t* rt_b, int64_t* k_tiles) {
    *rt_a = ceil_div(na, (int64_t)UT);
    *rt_b = ceil_div(nb, (int64_t)UT);
    *k_tiles = ceil_div(m * std::min<int64_t>(slots, na + nb), (int64_t)UK);
}

}  // namespace bark

using namespace bark;

extern "C" {

static size_t gram_aux_bytes(int64_t batch, int64_t m, int W) {
    // presence words + column bases + K tiles per forest
    return (((size_t)batch * m * (W + 1) + batch) * 4 + 255) & ~(size_t)255;
}

size_t bark_gram_workspace_bytes(int64_t batch, int64_t na, int64_t nb, int64_t m, int32_t slots) {
    if (batch <= 0 || na <= 0 || nb <= 0 || m <= 0 || slots <= 0) return 0;
    int64_t rta, rtb, kt;
    gram_dims(na, nb, m, slots, &rta, &rtb, &kt);
    return (size_t)batch * (size_t)(rta + rtb) * (size_t)kt * TILE_BYTES + gram_aux_bytes(batch, m, (slots + 31) / 32) + 256;
}

int bark_gram_umma(const uint32_t* leaves_a, const uint32_t* leaves_b, int64_t batch, int64_t na, int64_t nb, int64_t m,
                   int32_t slots, int32_t* counts, double* K, const double* scale, const double* noise, double jitter,
                   int add_diag, uint32_t* status, void* workspace, void* stream) {
    BARK_CHECK_ARG(batch >= 0 && na >= 0 && nb >= 0 && m >= 1, "bad size");
    BARK_CHECK_ARG(slots >= 1 && slots <= 256, "slots out of range (1..256)");
    if (batch == 0 || na == 0 || nb == 0) return BARK_OK;
    BARK_CHECK_ARG(leaves_a && leaves_b && workspace && (counts || K), "null pointer");
    BARK_CHECK_ARG(!K || scale, "K needs scale");
    BARK_CHECK_ARG(!(K && add_diag) || (noise && na == nb), "add_diag needs noise and a square matrix");
    BARK_CHECK_ARG(batch <= 65535 && ceil_div(na, UT) <= 65535 && m <= (1 << 23), "grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t rta, rtb, kt;
    gram_dims(na, nb, m, slots, &rta, &rtb, &kt);
    const int W = (slots + 31) / 32;
    const bool same = (leaves_a == leaves_b) && (na == nb);
    unsigned char* wsp = (unsigned char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    uint32_t* pres = (uint32_t*)wsp;
    int32_t* base = (int32_t*)(pres + (size_t)batch * m * W);
    int32_t* kt_used = base + (size_t)batch * m;
    uint8_t* Za = wsp + gram_aux_bytes(batch, m, W);
    const size_t za_bytes = (size_t)batch * rta * kt * TILE_BYTES;
    uint8_t* Zb = same ? Za : Za + za_bytes;

    // 1. occupied leaf slots -> consecutive columns (the union over both operands; a column one side never hits adds 0)
    BARK_CUDA(cudaMemsetAsync(pres, 0, (size_t)batch * m * W * 4, st));
    auto pres_grid = [&](int64_t n) { return (unsigned)std::min<int64_t>(148 * 16, ceil_div(batch * ceil_div(n, (int64_t)PRES_ROWS) * m, (int64_t)128)); };
    leaf_presence_kernel<<<pres_grid(na), 128, 0, st>>>(leaves_a, batch, na, m, slots, W, pres, status);
    if (!same) leaf_presence_kernel<<<pres_grid(nb), 128, 0, st>>>(leaves_b, batch, nb, m, slots, W, pres, status);
    leaf_columns_kernel<<<(unsigned)batch, 32, 0, st>>>(pres, batch, m, W, base, kt_used);
    // 2. operand tiles
    const int stage_map = (m * (W + 1) <= OB_MAP_WORDS) ? 1 : 0;
    const size_t ob_smem = (size_t)std::min<int64_t>(OB_GROUP, kt) * OB_ROWS * UK + (stage_map ? (size_t)m * (W + 1) * 4 : 0);
    BARK_CUDA(cudaFuncSetAttribute(onehot_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ob_smem));
    onehot_build_kernel<<<dim3((unsigned)(rta * (UT / OB_ROWS)), (unsigned)batch), OB_THREADS, ob_smem, st>>>(
        leaves_a, na, m, slots, W, rta, kt, pres, base, kt_used, Za, stage_map);
    if (!same)
        onehot_build_kernel<<<dim3((unsigned)(rtb * (UT / OB_ROWS)), (unsigned)batch), OB_THREADS, ob_smem, st>>>(
            leaves_b, nb, m, slots, W, rtb, kt, pres, base, kt_used, Zb, stage_map);
    BARK_LAUNCH_CHECK();
    // 3. counts
    const size_t smem = (size_t)U_STAGES * 2 * TILE_BYTES + 256;
    BARK_CUDA(cudaFuncSetAttribute(gram_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GramEpilogue ep{counts, K, scale, noise, 1.0 / (double)m, jitter, add_diag, status};
    dim3 grid((unsigned)rtb, (unsigned)rta, (unsigned)batch);
    gram_umma_kernel<<<grid, U_THREADS, smem, st>>>(Za, Zb, na, nb, rta, rtb, kt, kt_used, ep);
    BARK_LAUNCH_CHECK();
    return BARK_OK;
}

}  // extern "C"
