// Micro-benchmarks of the B200 ceilings that the BARK kernels are actually bound by and that
// MEASURED_PEAKS.json does not carry: FP64 tensor pipe (DMMA m8n8k4), FP64 FMA pipe, int8 tcgen05 (UMMA kind::i8),
// L2 read bandwidth, HBM read / copy bandwidth, and the latency of the synchronisation primitives the MCMC sweep
// uses (__syncthreads, cluster barrier at cluster sizes 2/4/8).  Prints ONE JSON object.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/peaks_bin scripts/peaks.cu
//   scripts/peaks_bin > profiles/peaks_r2.json          (scripts/peaks.py does both)
//
// Every figure is the best of 5 timed launches after one warm-up launch, CUDA events on the launching stream.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

namespace cg = cooperative_groups;

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e__ = (x);                                                         \
        if (e__ != cudaSuccess) {                                                      \
            fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e__)); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

// ------------------------------------------------------------------------------------------------ FP64
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(512, 1) dmma_kernel(double* out, int iters) {
    double acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma(acc[i][0], acc[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
    if (s == 123.456) out[0] = s;
}
__global__ void __launch_bounds__(512, 1) dfma_kernel(double* out, int iters) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

// ------------------------------------------------------------------------------------------------ memory
__global__ void __launch_bounds__(512, 2) read_kernel(const uint4* __restrict__ buf, size_t n_vec, int reps, unsigned* out) {
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r) {
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n_vec; i += 4 * stride) {
            const uint4 a = __ldcg(buf + i), b = __ldcg(buf + i + stride), c = __ldcg(buf + i + 2 * stride),
                        d = __ldcg(buf + i + 3 * stride);
            acc += a.x ^ b.y ^ c.z ^ d.w;
        }
        for (; i < n_vec; i += stride) acc += __ldcg(buf + i).x;
    }
    if (acc == 0x12345u) out[0] = acc;
}
__global__ void __launch_bounds__(512, 2) copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n_vec) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------------ int8 tcgen05
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__global__ void __launch_bounds__(128, 1) umma_i8_kernel(int iters, unsigned* out) {
    extern __shared__ __align__(1024) unsigned char sm[];  // A: 128 rows x 128 B, B: 256 rows x 128 B (SWIZZLE_128B images)
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + 256) * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x01010101u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_slot;
    if (tid == 0) {
        // D = S32, A = B = s8, K-major, N = 256, M = 128
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a_addr = smem_u32(sm), b_addr = a_addr + 128 * 128;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
                const uint32_t accum = (it > 0 || k4 > 0) ? 1u : 0u;
                asm volatile(
                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(tmem_d),
                    "l"(desc_sw128(a_addr + k4 * 32)), "l"(desc_sw128(b_addr + k4 * 32)), "r"(idesc), "r"(accum), "r"(0u)
                    : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    {
        uint32_t ok = 0;
        unsigned spins = 0;
        while (!ok && ++spins < (1u << 28)) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}\n"
                         : "=r"(ok)
                         : "r"(smem_u32(&bar)), "r"(0u)
                         : "memory");
        }
        if (!ok && tid == 0) out[1] = 0xDEADu;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(256u) : "memory");
}

// ------------------------------------------------------------------------------------------------ barriers
__global__ void __launch_bounds__(512, 1) sync_latency_kernel(int iters, long long* out) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double s[32];
    // __syncthreads
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) __syncthreads();
    long long t1 = clock64();
    // cluster barrier (if launched with a cluster dimension)
    cluster.sync();
    long long t2 = clock64();
    for (int i = 0; i < iters; ++i) cluster.sync();
    long long t3 = clock64();
    // deterministic block sum: warp shuffle tree + 2 barriers (what the sweep uses for its dot products)
    double v = threadIdx.x;
    long long t4 = clock64();
    for (int i = 0; i < iters; ++i) {
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
        __syncthreads();
        v = (threadIdx.x & 31) < 16 ? s[threadIdx.x & 31] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    long long t5 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        out[0] = (t1 - t0) / iters;
        out[1] = (t3 - t2) / iters;
        out[2] = (t5 - t4) / iters;
        out[3] = (long long)v;
    }
}

// ------------------------------------------------------------------------------------------------ load latency
// One warp per CTA chases pointers (dependent 8-byte ld.global.cg) through a buffer: L2-resident (8 MB) or HBM-sized
// (2 GB, random): cycles per dependent load.  Then the round trip of a BATCH of 8 independent 16-byte loads per lane
// (what the sweep kernel's product pass issues), alone on the GPU and with every SM doing the same.
__global__ void chase_kernel(const unsigned long long* __restrict__ next, int steps, long long* out) {
    unsigned long long p = threadIdx.x + blockIdx.x * 997;
    const long long t0 = clock64();
    for (int i = 0; i < steps; ++i) p = __ldcg(next + p);
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = (t1 - t0) / steps; out[1] = (long long)p; }
}
__global__ void __launch_bounds__(512, 1) batch_kernel(const double2* __restrict__ buf, size_t n_vec, int iters, long long* out) {
    const int lane = threadIdx.x & 31, lq = lane >> 2, lk = lane & 3;
    size_t base = ((size_t)blockIdx.x * 16 + (threadIdx.x >> 5)) * 65536 % (n_vec / 2);
    double acc = 0.0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        double2 x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = __ldcg(buf + base + (size_t)lq * 416 + lk + 4 * u + (size_t)it * 3331);  // 8 rows x 64 B, like a tile batch
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += x[u].x + x[u].y;
        base = (base + (size_t)(acc != 12345.678)) % (n_vec / 2);  // dependency: the next batch waits for this one
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = (t1 - t0) / iters; out[1] = (long long)acc; }
}

template <class F>
static float best_ms(F launch, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    double* dout;
    unsigned* uout;
    long long* lout;
    CK(cudaMalloc(&dout, 64));
    CK(cudaMalloc(&uout, 64));
    CK(cudaMalloc(&lout, 64));
    CK(cudaMemset(uout, 0, 64));

    // FP64 tensor pipe: 16 warps x 8 independent DMMA chains per CTA, 1 CTA per SM
    const int it_d = 1 << 14;
    const float ms_dmma = best_ms([&] { dmma_kernel<<<sms, 512>>>(dout, it_d); });
    const double tf_dmma = (double)sms * 16 * it_d * 8 * (2.0 * 8 * 8 * 4) / (ms_dmma * 1e-3) / 1e12;
    const int it_f = 1 << 14;
    const float ms_dfma = best_ms([&] { dfma_kernel<<<sms, 512>>>(dout, it_f); });
    const double tf_dfma = (double)sms * 512 * it_f * 8 * 2.0 / (ms_dfma * 1e-3) / 1e12;

    // L2 read: 48 MB buffer (fits the 126 MB L2), 24 passes; HBM read / copy: 4 GiB / 2 GiB + 2 GiB
    const size_t l2_bytes = 48ull << 20, big = 4ull << 30;
    uint4* buf;
    CK(cudaMalloc(&buf, big));
    CK(cudaMemset(buf, 1, big));
    const int l2_reps = 24;
    const float ms_l2 = best_ms([&] { read_kernel<<<sms * 2, 512>>>(buf, l2_bytes / 16, l2_reps, uout); });
    const double gbs_l2 = (double)l2_bytes * l2_reps / (ms_l2 * 1e-3) / 1e9;
    const float ms_hr = best_ms([&] { read_kernel<<<sms * 2, 512>>>(buf, big / 16, 1, uout); });
    const double gbs_hr = (double)big / (ms_hr * 1e-3) / 1e9;
    const float ms_cp = best_ms([&] { copy_kernel<<<sms * 2, 512>>>(buf, buf + big / 32, big / 32); });
    const double gbs_cp = (double)big / (ms_cp * 1e-3) / 1e9;  // read + write bytes, as a STREAM copy counts them

    // int8 UMMA: one issuing thread per SM, M=128 N=256 K=32 per instruction, operands resident in shared memory
    const int it_u = 4096;
    CK(cudaFuncSetAttribute(umma_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (128 + 256) * 128 + 1024));
    const float ms_u = best_ms([&] { umma_i8_kernel<<<sms, 128, (128 + 256) * 128 + 1024>>>(it_u, uout); });
    const double tops_u = (double)sms * it_u * 4 * (2.0 * 128 * 256 * 32) / (ms_u * 1e-3) / 1e12;
    unsigned hu[2] = {0, 0};
    CK(cudaMemcpy(hu, uout, 8, cudaMemcpyDeviceToHost));

    // barrier latencies (cycles), cluster sizes 1/2/4/8
    long long lat[4][3];
    const int csz[4] = {1, 2, 4, 8};
    for (int k = 0; k < 4; ++k) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(csz[k] * 8);
        cfg.blockDim = dim3(512);
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = csz[k];
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, sync_latency_kernel, 2000, lout));
        CK(cudaDeviceSynchronize());
        long long h[4];
        CK(cudaMemcpy(h, lout, 32, cudaMemcpyDeviceToHost));
        lat[k][0] = h[0]; lat[k][1] = h[1]; lat[k][2] = h[2];
    }

    // load latencies
    long long lat_l2 = 0, lat_hbm = 0, bat_alone = 0, bat_all = 0;
    {
        const size_t n_small = (8ull << 20) / 8, n_big = (2ull << 30) / 8;
        std::vector<unsigned long long> h(n_big);
        unsigned long long x = 88172645463325252ull;
        for (size_t i = 0; i < n_big; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; h[i] = x % n_big; }
        unsigned long long* d_next = reinterpret_cast<unsigned long long*>(buf);
        CK(cudaMemcpy(d_next, h.data(), n_big * 8, cudaMemcpyHostToDevice));
        chase_kernel<<<1, 32>>>(d_next, 2000, lout);
        CK(cudaDeviceSynchronize());
        long long hh[2];
        CK(cudaMemcpy(hh, lout, 16, cudaMemcpyDeviceToHost));
        lat_hbm = hh[0];
        for (size_t i = 0; i < n_small; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; h[i] = x % n_small; }
        CK(cudaMemcpy(d_next, h.data(), n_small * 8, cudaMemcpyHostToDevice));
        chase_kernel<<<1, 32>>>(d_next, 20000, lout);  // warms L2
        chase_kernel<<<1, 32>>>(d_next, 20000, lout);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hh, lout, 16, cudaMemcpyDeviceToHost));
        lat_l2 = hh[0];
        CK(cudaMemset(buf, 0, 64ull << 20));
        batch_kernel<<<1, 32>>>(reinterpret_cast<const double2*>(buf), (32ull << 20) / 16, 2000, lout);
        batch_kernel<<<1, 32>>>(reinterpret_cast<const double2*>(buf), (32ull << 20) / 16, 2000, lout);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hh, lout, 16, cudaMemcpyDeviceToHost));
        bat_alone = hh[0];
        batch_kernel<<<sms, 512>>>(reinterpret_cast<const double2*>(buf), (32ull << 20) / 16, 2000, lout);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hh, lout, 16, cudaMemcpyDeviceToHost));
        bat_all = hh[0];
    }

    printf("{\"device\": \"%s\", \"sms\": %d, \"sm_clock_mhz_attr\": %.0f,\n", prop.name, sms, clk_khz / 1e3);
    printf(" \"load_latency_cycles\": {\"l2_hit_dependent_ldcg\": %lld, \"hbm_dependent_ldcg\": %lld, "
           "\"batch8x16B_one_warp\": %lld, \"batch8x16B_16_warps_every_sm\": %lld},\n", lat_l2, lat_hbm, bat_alone, bat_all);
    printf(" \"fp64_dmma_tflops\": %.2f, \"fp64_dfma_tflops\": %.2f,\n", tf_dmma, tf_dfma);
    printf(" \"int8_umma_tops\": %.1f, \"int8_umma_ok\": %s,\n", tops_u, hu[1] == 0xDEADu ? "false" : "true");
    printf(" \"l2_read_gbs\": %.0f, \"l2_buffer_mb\": %zu, \"hbm_read_gbs\": %.0f, \"hbm_copy_gbs\": %.0f,\n", gbs_l2,
           l2_bytes >> 20, gbs_hr, gbs_cp);
    printf(" \"latency_cycles\": {\"syncthreads_512\": %lld, \"block_sum_512\": %lld, \"cluster_sync_1\": %lld, "
           "\"cluster_sync_2\": %lld, \"cluster_sync_4\": %lld, \"cluster_sync_8\": %lld},\n",
           lat[0][0], lat[0][2], lat[0][1], lat[1][1], lat[2][1], lat[3][1]);
    printf(" \"method\": \"best of 5 launches after a warm-up, CUDA events; dmma/dfma: 16 warps x 8 independent chains per SM; "
           "l2: 48 MB buffer read 24x with ld.global.cg.v4; hbm: 4 GiB read once / 2+2 GiB copy; umma: M128 N256 K32 kind::i8, "
           "operands resident in shared memory, one issuing thread per SM\"}\n");
    return 0;
}
