// gemm_dmma.cuh -- FP64 tensor-core (DMMA) GEMM used for every large contraction of the sweep.
//
//   Out[m][n] = sum_k A[k][m] * B[k][n]        ("TN": both operands stored k-major, tile dim contiguous)
//
// Three uses per Gibbs iteration, all with the chain axis as a GEMM dimension:
//   psi   [n  x C ] = U'[pq x n]'  theta[pq x C]     (X beta + K eta;   logit.py:201, 485)
//   G     [C  x M ] = Omega[n x C]' Z[n x M]         (vech of [X K]' diag(omega_b) [X K]; logit.py:230, 327-335)
//   H     [C  x pq] = kd[n x C]'    U[n x pq]        ([X K]' (z - 1/2);             logit.py:231, 481)
// G and H are produced by ONE launch: column tiles >= split_tile read A from A2 instead of A1.
//
// FP64 has no tcgen05 / wgmma kind on sm_100a; the FP64 tensor path is mma.sync m8n8k4 (SASS DMMA.8x8x4).
// Operands are staged global -> shared with a 4-deep cp.async pipeline; rows are padded by 4 doubles so
// that the per-thread fragment loads (lane = 4*row + k) are bank-conflict free.
//
// All extents must be multiples of the tile: M % 128 == 0, N % 128 == 0, K % 16 == 0 (the engine pads
// and zero-fills).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dgemm {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4, THREADS = 256;
constexpr int LDS = BM + 4;  // padded row length (doubles) of a staged tile
constexpr int STAGE_DOUBLES = 2 * BK * LDS;
constexpr int SMEM_BYTES = STAGES * STAGE_DOUBLES * 8;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

struct Args {
    const double* A1;  // [K][lda]
    const double* A2;  // [K][lda], used by column tiles >= split_tile (may equal A1)
    const double* B;   // [K][ldb]
    double* Out;       // [M][ldo]
    int M, N, K;
    int lda, ldb, ldo;
    int split_tile;  // first column tile that reads A2
};

__global__ void __launch_bounds__(THREADS, 1) gemm_tn_kernel(const Args g) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile_n = blockIdx.x, tile_m = blockIdx.y;
    const int m0 = tile_m * BM, n0 = tile_n * BN;
    const double* __restrict__ A = (tile_n >= g.split_tile) ? g.A2 : g.A1;
    const double* __restrict__ B = g.B;
    const int KT = g.K / BK;

    // 8 warps: 2 along M (64 rows each) x 4 along N (32 cols each)
    const int wm = (warp & 1) * 64, wn = (warp >> 1) * 32;
    const int fr = lane >> 2, fk = lane & 3;  // fragment row / k index of this lane

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto load_stage = [&](int kt, int slot) {
        double* As = smem + slot * STAGE_DOUBLES;
        double* Bs = As + BK * LDS;
        const int k0 = kt * BK;
#pragma unroll
        for (int i = 0; i < (BK * BM / 2) / THREADS; i++) {  // 16-byte chunks
            const int ch = tid + i * THREADS;
            const int r = ch / (BM / 2), c2 = (ch % (BM / 2)) * 2;
            cp_async16(As + r * LDS + c2, A + (size_t)(k0 + r) * g.lda + m0 + c2);
            cp_async16(Bs + r * LDS + c2, B + (size_t)(k0 + r) * g.ldb + n0 + c2);
        }
    };

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < KT) load_stage(s, s);
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; kt++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + STAGES - 1;
            if (nk < KT) load_stage(nk, nk % STAGES);
            cp_async_commit();
        }
        const double* As = smem + (kt % STAGES) * STAGE_DOUBLES;
        const double* Bs = As + BK * LDS;
#pragma unroll
        for (int ks = 0; ks < BK; ks += 4) {
            double af[8], bf[4];
#pragma unroll
            for (int i = 0; i < 8; i++) af[i] = As[(ks + fk) * LDS + wm + i * 8 + fr];
#pragma unroll
            for (int j = 0; j < 4; j++) bf[j] = Bs[(ks + fk) * LDS + wn + j * 8 + fr];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: C fragment (row = lane/4, cols = 2*(lane%4) + {0,1}) -> 16-byte stores
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int row = m0 + wm + i * 8 + fr;
        double* orow = g.Out + (size_t)row * g.ldo + n0 + wn + 2 * fk;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            *reinterpret_cast<double2*>(orow + j * 8) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
    }
}

inline cudaError_t launch(const Args& a, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    dim3 grid(a.N / BN, a.M / BM);
    gemm_tn_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace dgemm
