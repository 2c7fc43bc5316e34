// Micro-benchmark of the 64 x 64 pivot-block sweep of csrc/linalg.cuh (the serial part of the block LDL^T / inverse that
// the hyper step and the batched log-MLL run P / 64 times per matrix): variants of the per-pivot step, timed in
// cycles per step with 512 threads on one CTA, each checked against variant 0.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/pivot_bench_bin scripts/pivot_bench.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int NB = 64, THREADS = 512;

__device__ __forceinline__ double fast_rcp(double x) {
    double r = (double)__frcp_rn((float)x);
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}

struct Sm {
    double D[NB][NB + 1];
    double colb[2][NB], rowb[2][NB], ipb[2];
};

// V0: the shipped select-based step (every thread divides)
template <int VAR>
__device__ void sweep(Sm& s) {
    const int tid = threadIdx.x, c = tid & (NB - 1), r0 = tid >> 6, warp = tid >> 5;
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = s.D[r0 + 8 * i][c];
    if (tid < NB) { s.colb[0][tid] = s.D[tid][0]; s.rowb[0][tid] = s.D[0][tid]; }
    if (tid == 0) s.ipb[0] = 1.0 / s.D[0][0];
    __syncthreads();
    for (int j = 0; j < NB; ++j) {
        const double* colv = s.colb[j & 1];
        const double* rowv = s.rowb[j & 1];
        double* coln = s.colb[(j + 1) & 1];
        double* rown = s.rowb[(j + 1) & 1];
        double cr[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) cr[i] = colv[r0 + 8 * i];
        const double p = colv[j];
        const double rowc = rowv[c];
        double ip;
        if (VAR == 0 || VAR == 2) ip = 1.0 / p;
        else if (VAR == 1) ip = fast_rcp(p);
        else ip = s.ipb[j & 1];
        const double rc = rowc * ip;
        if (VAR == 0 || VAR == 1) {
            const bool cj = (c == j);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = r0 + 8 * i;
                const double xg = fma(-cr[i], rc, v[i]);
                const double xc = cr[i] * ip;
                double x = cj ? xc : xg;
                if (r == j) x = cj ? -ip : rc;
                v[i] = x;
                if (c == j + 1) coln[r] = x;
                if (r == j + 1) rown[c] = x;
            }
        } else {
            // warp-uniform guards: the pivot column lives in one lane of the warps with (warp & 1) == (j >> 5), the pivot
            // row in element j >> 3 of the two warps with r0 == (j & 7)
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fma(-cr[i], rc, v[i]);
            if ((warp & 1) == (j >> 5)) {
                const bool cj = (c == j);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = cj ? cr[i] * ip : v[i];
            }
            if (r0 == (j & 7)) {
                const double x = (c == j) ? -ip : rc;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (i == (j >> 3)) v[i] = x;
            }
            if ((warp & 1) == ((j + 1) >> 5)) {
                if (c == j + 1) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) coln[r0 + 8 * i] = v[i];
                }
            }
            if (r0 == ((j + 1) & 7)) {
                double x = 0.0;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (i == ((j + 1) >> 3)) x = v[i];
                rown[c] = x;
                if (VAR == 3 && c == j + 1) s.ipb[(j + 1) & 1] = fast_rcp(x);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s.D[r0 + 8 * i][c] = v[i];
    __syncthreads();
}

// V4: as V2 but the pivot loop is unrolled over blocks of 8 pivots, so that the element index j >> 3 of the pivot row is
// a compile-time constant (V2's predicated element pick is turned into a dynamic register index = local memory)
__device__ void sweep4(Sm& s) {
    const int tid = threadIdx.x, c = tid & (NB - 1), r0 = tid >> 6, warp = tid >> 5;
    double v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = s.D[r0 + 8 * i][c];
    if (tid < NB) { s.colb[0][tid] = s.D[tid][0]; s.rowb[0][tid] = s.D[0][tid]; }
    __syncthreads();
#pragma unroll
    for (int jb = 0; jb < 8; ++jb) {
#pragma unroll 1
        for (int jj = 0; jj < 8; ++jj) {
            const int j = 8 * jb + jj;
            const double* colv = s.colb[j & 1];
            const double* rowv = s.rowb[j & 1];
            double* coln = s.colb[(j + 1) & 1];
            double* rown = s.rowb[(j + 1) & 1];
            double cr[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) cr[i] = colv[r0 + 8 * i];
            const double p = colv[j];
            const double ip = 1.0 / p;
            const double rc = rowv[c] * ip;
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fma(-cr[i], rc, v[i]);
            if ((warp & 1) == (j >> 5)) {
                const bool cj = (c == j);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = cj ? cr[i] * ip : v[i];
            }
            if (r0 == jj) v[jb] = (c == j) ? -ip : rc;  // pivot row: row j = r0 + 8 jb with r0 == jj
            if ((warp & 1) == ((j + 1) >> 5)) {
                if (c == j + 1) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) coln[r0 + 8 * i] = v[i];
                }
            }
            if (jj < 7) {
                if (r0 == jj + 1) rown[c] = v[jb];  // next pivot row j + 1 = (jj + 1) + 8 jb
            } else if (jb < 7) {
                if (r0 == 0) rown[c] = v[jb + 1];   // next pivot row 8 (jb + 1)
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s.D[r0 + 8 * i][c] = v[i];
    __syncthreads();
}

template <int VAR>
__global__ void __launch_bounds__(THREADS, 1) bench_kernel(const double* in, double* out, long long* cycles, int reps) {
    extern __shared__ __align__(16) unsigned char raw[];
    Sm& s = *reinterpret_cast<Sm*>(raw);
    long long total = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int e = threadIdx.x; e < NB * NB; e += THREADS) s.D[e / NB][e % NB] = in[e];
        __syncthreads();
        const long long t0 = clock64();
        if (VAR == 4) sweep4(s); else sweep<VAR>(s);
        total += clock64() - t0;
    }
    for (int e = threadIdx.x; e < NB * NB; e += THREADS) out[e] = s.D[e / NB][e % NB];
    if (threadIdx.x == 0) cycles[0] = total / reps;
}

template <int VAR>
static void run(const double* d_in, double* d_out, long long* d_cyc, const double* ref, double* host, const char* name) {
    cudaFuncSetAttribute(bench_kernel<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Sm));
    bench_kernel<VAR><<<148, THREADS, sizeof(Sm)>>>(d_in, d_out, d_cyc, 50);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(host, d_out, NB * NB * 8, cudaMemcpyDeviceToHost);
    double err = 0.0, mx = 0.0;
    if (ref)
        for (int i = 0; i < NB * NB; ++i) { err = fmax(err, fabs(host[i] - ref[i])); mx = fmax(mx, fabs(ref[i])); }
    printf("{\"variant\": \"%s\", \"cycles_per_block\": %lld, \"cycles_per_step\": %.1f, \"max_abs_diff_vs_v0\": %.3e, \"max_abs\": %.3e, \"cuda\": \"%s\"}\n",
           name, cyc, (double)cyc / NB, err, mx, cudaGetErrorString(e));
}

int main() {
    double h[NB * NB], ref[NB * NB], out[NB * NB];
    srand(1);
    double G[NB][8];
    for (int i = 0; i < NB; ++i)
        for (int k = 0; k < 8; ++k) G[i][k] = (double)rand() / RAND_MAX - 0.5;
    for (int i = 0; i < NB; ++i)
        for (int j = 0; j < NB; ++j) {
            double a = (i == j) ? 0.7 : 0.0;
            for (int k = 0; k < 8; ++k) a += G[i][k] * G[j][k];
            h[i * NB + j] = a;
        }
    double *d_in, *d_out;
    long long* d_cyc;
    cudaMalloc(&d_in, sizeof(h));
    cudaMalloc(&d_out, sizeof(h));
    cudaMalloc(&d_cyc, 8);
    cudaMemcpy(d_in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<0>(d_in, d_out, d_cyc, nullptr, ref, "v0 selects, 1.0/p per thread (shipped)");
    run<1>(d_in, d_out, d_cyc, ref, out, "v1 selects, Newton reciprocal per thread");
    run<2>(d_in, d_out, d_cyc, ref, out, "v2 warp-uniform guards, 1.0/p per thread");
    run<3>(d_in, d_out, d_cyc, ref, out, "v3 warp-uniform guards, reciprocal by the pivot's owner");
    run<4>(d_in, d_out, d_cyc, ref, out, "v4 warp-uniform guards, pivot loop unrolled by 8 (static element index)");
    return 0;
}
