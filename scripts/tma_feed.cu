// Micro-benchmark: how fast can one SM pull L2-resident operand tiles into shared memory when every SM does the same
// (the B-operand stream of predict_umma_kernel)?  Variants: cp.async.bulk of 16 / 32 KB through a ring of D stages,
// the same with a 2-CTA cluster where each CTA fetches half of every tile and multicasts it to both, and plain
// ld.global.cg.v4 + st.shared by 16 warps.  Output: bytes per clock per SM landed in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/tma_feed_bin scripts/tma_feed.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    unsigned spins = 0;
    while (!ok && ++spins < (1u << 26))
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

constexpr int REGION = 1344 * 1024;  // the 42 x 32 KB digit tiles of one posterior sample

// mode 0: bulk copies, one CTA;  mode 1: 2-CTA cluster, each CTA fetches half a tile and multicasts to both
template <int MODE>
__global__ void __launch_bounds__(128, 1) bulk_kernel(const unsigned char* src, int tile, int stages, int tiles, long long* out, int issuers = 1) {
    extern __shared__ __align__(1024) unsigned char ring[];
    __shared__ uint64_t full[8];
    const int tid = threadIdx.x;
    uint32_t rank = 0;
    if (MODE == 1) rank = cg::this_cluster().block_rank();
    if (tid == 0) {
        for (int s = 0; s < 8; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(full + s)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (MODE == 1) cg::this_cluster().sync(); else __syncthreads();
    if ((tid & 31) == 0 && (tid >> 5) < issuers) {
        const long long t0 = clock64();
        for (int i = tid >> 5; i < tiles + stages; i += issuers) {
            const int s = i % stages;
            if (i >= stages) wait_bar(full + s, (uint32_t)((i / stages - 1) & 1));  // tile i - stages landed: slot reusable
            if (MODE == 1 && i >= stages) {
                // both CTAs must have consumed the slot before either overwrites it in both: cluster-wide hand-shake
                // is what a real kernel needs; here the two run in lock step through their own barriers (a lower bound
                // on the cost), so only the local wait is used
            }
            if (i < tiles) {
                const size_t off = ((size_t)i * tile) % REGION;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(full + s)), "r"((uint32_t)tile) : "memory");
                if (MODE == 0) {
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring + (size_t)s * tile)),
                                 "l"(src + off), "r"((uint32_t)tile), "r"(smem_u32(full + s))
                                 : "memory");
                } else {
                    const uint32_t half = (uint32_t)tile / 2;
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
                            smem_u32(ring + (size_t)s * tile + rank * half)),
                        "l"(src + off + rank * half), "r"(half), "r"(smem_u32(full + s)), "h"((uint16_t)3)
                        : "memory");
                }
            }
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0 && tid == 0) out[0] = t1 - t0;
    }
    if (MODE == 1) cg::this_cluster().sync(); else __syncthreads();
}

__global__ void __launch_bounds__(512, 1) ldg_kernel(const unsigned char* src, int tile, int tiles, long long* out) {
    extern __shared__ __align__(1024) unsigned char ring[];
    const int tid = threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < tiles; ++i) {
        const uint4* s4 = reinterpret_cast<const uint4*>(src + ((size_t)i * tile) % REGION);
        uint4* d4 = reinterpret_cast<uint4*>(ring + (size_t)(i & 1) * tile);
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (tid + u * 512 < tile / 16) v[u] = __ldcg(s4 + tid + u * 512);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (tid + u * 512 < tile / 16) d4[tid + u * 512] = v[u];
        __syncthreads();
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && tid == 0) out[0] = t1 - t0;
}

int main() {
    unsigned char* src;
    long long* d;
    cudaMalloc(&src, REGION);
    cudaMemset(src, 1, REGION);
    cudaMalloc(&d, 8);
    const int tiles = 2048;
    auto report = [&](const char* name, int tile, int stages, int per_sm_tiles, cudaError_t e) {
        long long c = 0;
        cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("{\"variant\": \"%s\", \"tile_kb\": %d, \"stages\": %d, \"bytes_per_clk_per_sm\": %.1f, \"aggregate_tbs_148sm_1965mhz\": %.2f, \"cuda\": \"%s\"}\n", name,
               tile / 1024, stages, (double)per_sm_tiles * tile / (double)c, (double)per_sm_tiles * tile / (double)c * 148 * 1.965e9 / 1e12,
               cudaGetErrorString(e));
    };
    cudaFuncSetAttribute(bulk_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(bulk_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(ldg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int tile : {4096, 8192, 16384, 32768, 65536})
        for (int stages : {2, 4}) {
            if ((size_t)tile * stages > 200 * 1024) continue;
            if (stages % 2) continue;
            for (int issuers : {1, 2}) {
                for (int rep = 0; rep < 2; ++rep) bulk_kernel<0><<<148, 128, (size_t)tile * stages>>>(src, tile, stages, tiles, d, issuers);
                report(issuers == 1 ? "cp.async.bulk, one issuing thread" : "cp.async.bulk, two issuing threads (two warps)", tile, stages, tiles,
                       cudaDeviceSynchronize());
            }
        }
    // one SM alone (no contention from the other 147)
    for (int rep = 0; rep < 2; ++rep) bulk_kernel<0><<<1, 128, (size_t)32768 * 4>>>(src, 32768, 4, tiles, d, 1);
    report("cp.async.bulk, one issuing thread, ONE SM only", 32768, 4, tiles, cudaDeviceSynchronize());
    for (int stages : {4}) {
        const int tile = 32768;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(148);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = (size_t)tile * stages;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, bulk_kernel<1>, (const unsigned char*)src, (int)tile, (int)stages, (int)tiles, d, 1);
        report("cp.async.bulk multicast, 2-CTA cluster, half a tile fetched per CTA", tile, stages, tiles, cudaDeviceSynchronize());
    }
    for (int rep = 0; rep < 2; ++rep) ldg_kernel<<<148, 512, 64 * 1024>>>(src, 32768, tiles, d);
    report("ld.global.cg.v4 + st.shared, 16 warps", 32768, 2, tiles, cudaDeviceSynchronize());
    return 0;
}
