// a4 on the 5th-generation tensor cores: leaf co-occurrence counts as an exact int8 one-hot GEMM.
//
//   count[i,j] = #{t : leaf_t(x_i) == leaf_t(x'_j)} = sum_k Za[i,k] * Zb[j,k],   k = t * S + leaf slot,
// with Z the {0,1} leaf-indicator matrix (S = slots per tree).  int8 x int8 products accumulate in s32, so the
// tensor-core result is the exact integer count (src/bark/forest.py:85-88).
//
// Pipeline
//   1. onehot_build_kernel  writes Za / Zb as int8, pre-tiled and pre-swizzled: every (128 rows x 128 K-bytes)
//      tile is one contiguous 16 KB block holding exactly the shared-memory image that a K-major SWIZZLE_128B
//      UMMA operand needs (16-byte chunk c of row r stored at chunk c ^ (r & 7) of its 1 KB 8-row group).
//   2. gram_umma_kernel     one CTA per 128 x 128 output tile: an elected thread streams the A/B tiles of every
//      K chunk with the bulk-copy engine (TMA, cp.async.bulk -> mbarrier complete_tx) through a 4-stage ring and
//      issues tcgen05.mma.cta_group::1.kind::i8 (M = N = 128, K = 32, four per chunk) into a 128-column s32 TMEM
//      accumulator; tcgen05.commit releases the stages and finally signals the epilogue.
//   3. epilogue             4 warps read TMEM (tcgen05.ld 32x32b.x32), emit the int32 counts and / or the FP64
//      kernel matrix K = scale * ((1/m) * count) + (jitter + noise) I with un-fused multiplies (bit-exact vs numpy).
#include <algorithm>

#include "common.cuh"

namespace bark {

constexpr int UT = 128;                 // tile edge (UMMA M = N = 128)
constexpr int UK = 128;                 // K bytes per tile (one SWIZZLE_128B atom row)
constexpr int TILE_BYTES = UT * UK;     // 16 KB
constexpr int U_STAGES = 3;             // 96 KB of operand ring: two CTAs per SM, so one tile's epilogue overlaps the other's MMAs
constexpr int U_THREADS = 128;

__device__ __forceinline__ uint32_t u_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row r in [0,128), K byte kb in [0,128)) inside a pre-swizzled 16 KB tile
__host__ __device__ __forceinline__ uint32_t swizzled_offset(uint32_t r, uint32_t kb) {
    const uint32_t g = r >> 3, rr = r & 7, chunk = kb >> 4, b = kb & 15;
    return g * 1024u + rr * 128u + ((chunk ^ rr) << 4) + b;
}

__global__ void onehot_build_kernel(const uint32_t* __restrict__ leaves, int64_t batch, int64_t n, int64_t m, int slots,
                                    int64_t row_tiles, int64_t k_tiles, uint8_t* __restrict__ Z,
                                    uint32_t* __restrict__ status) {
    const int64_t total = batch * n * m;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = e / (n * m), rem = e % (n * m), i = rem / m, t = rem % m;
        const uint32_t id = leaves[e];
        if (id >= (uint32_t)slots) {
            if (status) atomicOr(status, 1u);
            continue;
        }
        const int64_t k = t * slots + id;
        const int64_t tile = (b * row_tiles + (i >> 7)) * k_tiles + (k >> 7);
        Z[tile * TILE_BYTES + swizzled_offset((uint32_t)(i & 127), (uint32_t)(k & 127))] = 1;
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

__global__ void __launch_bounds__(U_THREADS, 2)
gram_umma_kernel(const uint8_t* __restrict__ Za, const uint8_t* __restrict__ Zb, int64_t na, int64_t nb, int64_t rt_a,
                 int64_t rt_b, int64_t k_tiles, GramEpilogue ep) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* tiles = smem_raw;  // U_STAGES x (A tile | B tile)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)U_STAGES * 2 * TILE_BYTES);
    uint64_t* empty_bar = full_bar + U_STAGES;
    uint64_t* acc_bar = empty_bar + U_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t b = blockIdx.z, tr = blockIdx.y, tc = blockIdx.x;

    if (tid == 0) {
        for (int s = 0; s < U_STAGES; ++s) { u_mbar_init(full_bar + s, 1); u_mbar_init(empty_bar + s, 1); }
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

    if (tid == 0) {
        // ---- producer + MMA issuer (single elected thread)
        const uint8_t* a_src = Za + ((b * rt_a + tr) * k_tiles) * (int64_t)TILE_BYTES;
        const uint8_t* b_src = Zb + ((b * rt_b + tc) * k_tiles) * (int64_t)TILE_BYTES;
        const uint32_t idesc = umma_idesc_i8(UT, UT);
        const int64_t KT = k_tiles;
        for (int64_t it = 0; it < KT + U_STAGES - 1; ++it) {
            if (it < KT) {
                const int s = (int)(it % U_STAGES);
                if (it >= U_STAGES) u_mbar_wait(empty_bar + s, (uint32_t)((it / U_STAGES - 1) & 1), ep.status);
                u_mbar_expect_tx(full_bar + s, 2 * TILE_BYTES);
                u_bulk_g2s(tiles + (size_t)s * 2 * TILE_BYTES, a_src + it * TILE_BYTES, TILE_BYTES, full_bar + s);
                u_bulk_g2s(tiles + (size_t)s * 2 * TILE_BYTES + TILE_BYTES, b_src + it * TILE_BYTES, TILE_BYTES, full_bar + s);
            }
            const int64_t kc = it - (U_STAGES - 1);
            if (kc >= 0) {
                const int s = (int)(kc % U_STAGES);
                u_mbar_wait(full_bar + s, (uint32_t)((kc / U_STAGES) & 1), ep.status);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = u_smem(tiles + (size_t)s * 2 * TILE_BYTES);
                const uint32_t b_addr = a_addr + TILE_BYTES;
#pragma unroll
                for (int k4 = 0; k4 < UK / 32; ++k4)
                    umma_i8(tmem_d, umma_desc_sw128(a_addr + k4 * 32), umma_desc_sw128(b_addr + k4 * 32), idesc,
                            (kc > 0 || k4 > 0) ? 1u : 0u);
                umma_commit(empty_bar + s);  // stage free once these MMAs have read it
            }
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

static void gram_dims(int64_t n, int64_t m, int slots, int64_t* row_tiles, int64_t* k_tiles) {
    *row_tiles = ceil_div(n, UT);
    *k_tiles = ceil_div(m * (int64_t)slots, UK);
}

}  // namespace bark

using namespace bark;

extern "C" {

size_t bark_gram_workspace_bytes(int64_t batch, int64_t na, int64_t nb, int64_t m, int32_t slots) {
    if (batch <= 0 || na <= 0 || nb <= 0 || m <= 0 || slots <= 0) return 0;
    int64_t rta, rtb, kt;
    gram_dims(na, m, slots, &rta, &kt);
    gram_dims(nb, m, slots, &rtb, &kt);
    return (size_t)batch * (size_t)(rta + rtb) * (size_t)kt * TILE_BYTES + 256;
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
    BARK_CHECK_ARG(batch <= 65535 && ceil_div(na, UT) <= 65535, "grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t rta, rtb, kt;
    gram_dims(na, m, slots, &rta, &kt);
    gram_dims(nb, m, slots, &rtb, &kt);
    const bool same = (leaves_a == leaves_b) && (na == nb);
    uint8_t* Za = (uint8_t*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    const size_t za_bytes = (size_t)batch * rta * kt * TILE_BYTES;
    uint8_t* Zb = same ? Za : Za + za_bytes;
    const size_t zb_bytes = same ? 0 : (size_t)batch * rtb * kt * TILE_BYTES;
    BARK_CUDA(cudaMemsetAsync(Za, 0, za_bytes + zb_bytes, st));
    const int bgrid = 148 * 8;
    onehot_build_kernel<<<bgrid, 256, 0, st>>>(leaves_a, batch, na, m, slots, rta, kt, Za, status);
    if (!same) onehot_build_kernel<<<bgrid, 256, 0, st>>>(leaves_b, batch, nb, m, slots, rtb, kt, Zb, status);
    BARK_LAUNCH_CHECK();
    const size_t smem = (size_t)U_STAGES * 2 * TILE_BYTES + 256;
    BARK_CUDA(cudaFuncSetAttribute(gram_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GramEpilogue ep{counts, K, scale, noise, 1.0 / (double)m, jitter, add_diag, status};
    dim3 grid((unsigned)rtb, (unsigned)rta, (unsigned)batch);
    gram_umma_kernel<<<grid, U_THREADS, smem, st>>>(Za, Zb, na, nb, rta, rtb, kt, ep);
    BARK_LAUNCH_CHECK();
    return BARK_OK;
}

}  // extern "C"
