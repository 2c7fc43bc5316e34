// Micro-benchmark: cycles per tcgen05.mma kind::i8 instruction (M = 128, K = 32) for the operand placements and N the
// predict kernel can choose from -- A in shared memory (SS) or in TMEM (TS), N = 128 / 256 -- with resident operands,
// and the round trip of a commit -> mbarrier wait after every group of G instructions (the per-item hand-off).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/umma_shapes_bin scripts/umma_shapes.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u)
                 : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(d),
                 "r"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u)
                 : "memory");
}
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    unsigned spins = 0;
    while (!ok && ++spins < (1u << 26))
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// TS: A in TMEM columns [0, 32) (one 128-byte K tile), D at column 128.  group > 0: commit + wait after every `group` MMAs.
template <bool TS, int N>
__global__ void __launch_bounds__(128, 1) k(int iters, int group, long long* out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + 256) * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x01010101u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_slot;
    if (TS) {  // fill the A columns with ones
        uint32_t v = 0x01010101u;
        for (int c = 0; c < 32; ++c)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem_d + ((uint32_t)(warp * 32) << 16) + c), "r"(v) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a_addr = smem_u32(sm), b_addr = a_addr + 128 * 128;
        uint32_t phase = 0;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t acc = (it > 0) ? 1u : 0u;
            const int k4 = it & 3;
            const uint32_t dcol = tmem_d + 128 + ((group > 0) ? (uint32_t)(((it / group) & 1) * (N == 128 ? 128 : 0)) : 0u);
            if (TS) mma_ts(dcol, tmem_d + k4 * 8, desc_sw128(b_addr + k4 * 32), idesc, acc);
            else mma_ss(dcol, desc_sw128(a_addr + k4 * 32), desc_sw128(b_addr + k4 * 32), idesc, acc);
            if (group > 0 && (it + 1) % group == 0) {
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                wait_bar(&bar, phase);
                phase ^= 1;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
        }
        if (group == 0 || iters % group != 0) {
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            wait_bar(&bar, phase);
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512u) : "memory");
}

// Fresh operands: B rotates through 4 x 32 KB stages and A through 4 tiles (no operand-collector reuse), and the
// accumulator alternates between NACC TMEM buffers (consecutive MMAs into one accumulator form a dependent chain).
template <bool TS, int N, int NACC, int CE = 0>
__global__ void __launch_bounds__(128, 1) k_rot(int iters, long long* out, int commit_every = 0) {
    extern __shared__ __align__(1024) unsigned char sm[];  // A: 4 x 16 KB, B: 4 x 32 KB
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (4 * 16384 + 4 * 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x01010101u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar2)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_slot;
    if (TS) {
        uint32_t v = 0x01010101u;
        for (int c = 0; c < 128; ++c)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem_d + ((uint32_t)(warp * 32) << 16) + c), "r"(v) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a_base = smem_u32(sm), b_base = a_base + 4 * 16384;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const int k4 = it & 3, stage = (it >> 2) & 3, accb = it % NACC;
            const uint32_t dcol = tmem_d + 128 + (uint32_t)(accb * N);
            const uint32_t acc = (it >= NACC) ? 1u : 0u;
            if (TS) mma_ts(dcol, tmem_d + stage * 32 + k4 * 8, desc_sw128(b_base + stage * 32768 + k4 * 32), idesc, acc);
            else mma_ss(dcol, desc_sw128(a_base + stage * 16384 + k4 * 32), desc_sw128(b_base + stage * 32768 + k4 * 32), idesc, acc);
            // a commit nobody waits for (the ring-slot release of a pipelined kernel): does it slow the MMA stream?
            if (CE > 0 && ((it + 1) & (CE - 1)) == 0)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        wait_bar(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512u) : "memory");
}

template <bool TS, int N, int NACC, int CE = 0>
static void run_rot(const char* name, long long* d, int commit_every = CE) {
    const int iters = 4096, smem = 4 * 16384 + 4 * 32768 + 1024;
    cudaFuncSetAttribute(k_rot<TS, N, NACC, CE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_rot<TS, N, NACC, CE><<<148, 128, smem>>>(iters, d, commit_every);
    cudaDeviceSynchronize();
    k_rot<TS, N, NACC, CE><<<148, 128, smem>>>(iters, d, commit_every);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("{\"variant\": \"%s, rotating operands\", \"commit_every\": %d, \"N\": %d, \"accumulators\": %d, \"cycles_per_mma\": %.1f, \"int8_tops_148sm_1965mhz\": %.0f, \"cuda\": \"%s\"}\n",
           name, commit_every, N, NACC, (double)c / iters, 2.0 * 128 * N * 32 / ((double)c / iters) * 148 * 1.965e9 / 1e12, cudaGetErrorString(e));
}

// TMEM read throughput: W warps (W % 4 == 0: every lane quarter gets W / 4 warps) read 32 lanes x 64 columns per step
// (two 32x32b.x32 loads, one wait), optionally while thread 0 of an extra warp keeps the tensor pipe busy with N = 256 MMAs
// into the other half of TMEM.
__global__ void __launch_bounds__(640, 1) ldtm_kernel(int iters, int warps, int with_mma, long long* out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ int stop;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + 256) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x01010101u;
    if (tid == 0) {
        stop = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_slot;
    if (warp < warps) {
        unsigned acc = 0;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            uint32_t v0[32], v1[32];
            const uint32_t ta = tmem_d + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp >> 2) * 64) & 255);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(v0[0]), "=r"(v0[1]), "=r"(v0[2]), "=r"(v0[3]), "=r"(v0[4]), "=r"(v0[5]), "=r"(v0[6]), "=r"(v0[7]), "=r"(v0[8]), "=r"(v0[9]), "=r"(v0[10]), "=r"(v0[11]), "=r"(v0[12]), "=r"(v0[13]), "=r"(v0[14]), "=r"(v0[15]), "=r"(v0[16]), "=r"(v0[17]), "=r"(v0[18]), "=r"(v0[19]), "=r"(v0[20]), "=r"(v0[21]), "=r"(v0[22]), "=r"(v0[23]), "=r"(v0[24]), "=r"(v0[25]), "=r"(v0[26]), "=r"(v0[27]), "=r"(v0[28]), "=r"(v0[29]), "=r"(v0[30]), "=r"(v0[31])
                         : "r"(ta) : "memory");
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(v1[0]), "=r"(v1[1]), "=r"(v1[2]), "=r"(v1[3]), "=r"(v1[4]), "=r"(v1[5]), "=r"(v1[6]), "=r"(v1[7]), "=r"(v1[8]), "=r"(v1[9]), "=r"(v1[10]), "=r"(v1[11]), "=r"(v1[12]), "=r"(v1[13]), "=r"(v1[14]), "=r"(v1[15]), "=r"(v1[16]), "=r"(v1[17]), "=r"(v1[18]), "=r"(v1[19]), "=r"(v1[20]), "=r"(v1[21]), "=r"(v1[22]), "=r"(v1[23]), "=r"(v1[24]), "=r"(v1[25]), "=r"(v1[26]), "=r"(v1[27]), "=r"(v1[28]), "=r"(v1[29]), "=r"(v1[30]), "=r"(v1[31])
                         : "r"(ta + 32) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 32; ++c) acc += v0[c] ^ v1[c];
        }
        const long long t1 = clock64();
        if (acc == 0x12345u) out[1] = acc;
        if (blockIdx.x == 0 && tid == 0) out[0] = t1 - t0;
        __syncwarp();
        if (tid == 0) *(volatile int*)&stop = 1;
    } else if (warp == 19 && (tid & 31) == 0 && with_mma) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a_addr = smem_u32(sm), b_addr = a_addr + 128 * 128;
        int it = 0;
        while (!*(volatile int*)&stop && it < (1 << 22)) {
            for (int k4 = 0; k4 < 4; ++k4, ++it) mma_ss(tmem_d + 256, desc_sw128(a_addr + k4 * 32), desc_sw128(b_addr + k4 * 32), idesc, it > 0);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        wait_bar(&bar, 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512u) : "memory");
}

static void run_ldtm(int warps, int with_mma, long long* d) {
    const int iters = 2048, smem = (128 + 256) * 128 + 1024;
    cudaFuncSetAttribute(ldtm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    ldtm_kernel<<<148, 640, smem>>>(iters, warps, with_mma, d);
    cudaDeviceSynchronize();
    ldtm_kernel<<<148, 640, smem>>>(iters, warps, with_mma, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("{\"variant\": \"tcgen05.ld 32x32b.x32 x2 per step\", \"warps\": %d, \"mma_running\": %d, \"tmem_read_bytes_per_clk_per_sm\": %.1f, \"cycles_per_step\": %.1f, \"cuda\": \"%s\"}\n",
           warps, with_mma, (double)warps * 32 * 64 * 4 * iters / (double)c, (double)c / iters, cudaGetErrorString(e));
}

template <bool TS, int N>
static void run(const char* name, int group, long long* d) {
    const int iters = 4096, smem = (128 + 256) * 128 + 1024;
    cudaFuncSetAttribute(k<TS, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<TS, N><<<148, 128, smem>>>(iters, group, d);
    cudaDeviceSynchronize();
    k<TS, N><<<148, 128, smem>>>(iters, group, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("{\"variant\": \"%s\", \"N\": %d, \"commit_every\": %d, \"cycles_per_mma\": %.1f, \"int8_tops_148sm_1965mhz\": %.0f, \"cuda\": \"%s\"}\n", name, N,
           group, (double)c / iters, 2.0 * 128 * N * 32 / ((double)c / iters) * 148 * 1.965e9 / 1e12, cudaGetErrorString(e));
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    run<false, 256>("SS (A, B in shared memory)", 0, d);
    run<false, 128>("SS (A, B in shared memory)", 0, d);
    run<true, 256>("TS (A in TMEM)", 0, d);
    run<true, 128>("TS (A in TMEM)", 0, d);
    run<true, 128>("TS, commit + wait every 16", 16, d);
    run<true, 128>("TS, commit + wait every 4", 4, d);
    run<false, 256>("SS, commit + wait every 16", 16, d);
    run<false, 256>("SS, commit + wait every 4", 4, d);
    run_rot<false, 256, 1>("SS", d);
    run_rot<false, 256, 1, 4>("SS", d);
    run_rot<false, 256, 1, 1>("SS", d);
    run_rot<true, 192, 2, 4>("TS", d);
    run_rot<true, 128, 2, 4>("TS", d);
    run_rot<true, 256, 1>("TS", d);
    run_rot<false, 128, 1>("SS", d);
    run_rot<false, 128, 2>("SS", d);
    run_rot<true, 128, 1>("TS", d);
    run_rot<true, 128, 2>("TS", d);
    run_rot<true, 128, 3>("TS", d);
    run_rot<true, 64, 4>("TS", d);
    run_rot<true, 192, 2>("TS", d);
    for (int warps : {4, 8, 16})
        for (int with_mma : {0, 1}) run_ldtm(warps, with_mma, d);
    return 0;
}
