// Microbenchmark: do the FP64 tensor pipe (DMMA.8x8x4) and the FP64 vector pipe (DFMA) share execution
// resources on sm_100a?  Runs DMMA-only, DFMA-only and mixed (half the warps each) and prints TFLOP/s.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// mode 0: all warps DMMA, 1: all DFMA, 2: even warps DMMA / odd warps DFMA
__global__ void __launch_bounds__(256) k(int mode, int iters, double* out) {
    const int warp = threadIdx.x >> 5;
    const bool use_mma = mode == 0 || (mode == 2 && (warp & 1) == 0);
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = i;
    if (use_mma) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) dmma(acc[2 * i], acc[2 * i + 1], a, b);
        }
    } else {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

int main() {
    double* out;
    cudaMalloc(&out, 8);
    const int iters = 20000, grid = 148 * 4, block = 256;
    for (int mode = 0; mode < 3; mode++) {
        k<<<grid, block>>>(mode, 100, out);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k<<<grid, block>>>(mode, iters, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double warps = (double)grid * block / 32;
        const double mma_warps = mode == 0 ? warps : mode == 1 ? 0 : warps / 2;
        const double fma_warps = warps - mma_warps;
        const double flops = 2.0 * iters * (mma_warps * 8 * 256 + fma_warps * 16 * 32);
        printf("mode %d: %.3f ms  %.2f TFLOP/s (dmma warps %.0f, dfma warps %.0f)\n", mode, ms, flops / ms / 1e9, mma_warps, fma_warps);
    }
    return 0;
}
