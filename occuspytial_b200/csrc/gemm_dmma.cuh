// gemm_dmma.cuh -- FP64 tensor-core (DMMA) GEMM used for every large contraction of the sweep.
//
//   Out[m][n] = sum_k A[k][m] * B[k][n]        ("TN": both operands stored k-major, tile dim contiguous)
//
// Three uses per Gibbs iteration, all with the chain axis as a GEMM dimension:
//   psi   [n  x C ] = U'[pq x n]'  theta[pq x C]     (X beta + K eta;   logit.py:201, 485)
//   G     [C  x M ] = Omega[n x C]' Z[n x M]         (vech of [X K]' diag(omega_b) [X K]; logit.py:230, 327-335)
//   H     [C  x pq] = kd[n x C]'    U[n x pq]        ([X K]' (z - 1/2);             logit.py:231, 481)
// G and H come from ONE launch.  Work items are ordered big-first: all full-K tiles of G, then the few
// H tiles split `ksplit` ways along K (partials summed in fixed order by the consumer).  The hardware
// dispatches CTAs in blockIdx order, so the small split items back-fill the last, partially filled wave:
// with 128 x 64 tiles and 2 CTAs/SM the headline shape (C = 4096, M = 2278 + 67, 148 SMs) runs 4.1 waves of
// work in 4.125 wave-times instead of 5.
//
// FP64 has no tcgen05 / wgmma kind on sm_100a; the FP64 tensor path is mma.sync m8n8k4 (SASS DMMA.8x8x4).
// Operands are staged global -> shared with a 4-deep cp.async pipeline; rows are padded by 4 doubles so
// that the per-thread fragment loads (lane = 4*row + k) are bank-conflict free.
//
// All extents must be multiples of the tile: M % 128 == 0, N % 64 == 0, K % 16 == 0 (the engine pads
// and zero-fills).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dgemm {

#ifndef GEMM_STAGES
#define GEMM_STAGES 4
#endif
constexpr int BM = 128, BN = 64, BK = 16, STAGES = GEMM_STAGES, THREADS = 256;
constexpr int WARPS_M = 4, WARPS_N = 2;            // 8 warps, warp tile 32 x 32
constexpr int MI = BM / WARPS_M / 8, NI = BN / WARPS_N / 8;
constexpr int LDA_S = BM + 4, LDB_S = BN + 4;      // padded row lengths (doubles): (ld mod 16) == 4
constexpr int STAGE_DOUBLES = BK * (LDA_S + LDB_S);
constexpr int SMEM_BYTES = STAGES * STAGE_DOUBLES * 8;
constexpr int CTAS_PER_SM = 2;
// WIDE instance: the LAST column tile of the second product is 72 columns wide (the warps of the upper column half take a
// fifth 8-column MMA tile), for widths n2_valid = 64 k + (1..8) -- pq = p + q = 67 is the common case -- where a separate
// 64-column tile would stream the whole A operand a second time for a handful of columns
constexpr int BNW = BN + 8, LDB_SW = BNW + 4;
constexpr int STAGE_DOUBLES_W = BK * (LDA_S + LDB_SW);
constexpr int SMEM_BYTES_W = STAGES * STAGE_DOUBLES_W * 8;
// Register cap: 96 keeps two CTAs per SM with room to spare and measured fastest on B200
// (124 regs under __launch_bounds__(256, 2): 6.04 ms; 96: 5.87 ms; 88: 5.95 ms; 80 spills: 6.40 ms).
#ifndef GEMM_MAXREG
#define GEMM_MAXREG 96
#endif
#define GEMM_BOUNDS __maxnreg__(GEMM_MAXREG)

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
    // main product: Out[M][N] = A1' B[:, 0:N]
    const double* A1;  // [K][lda]
    const double* B;   // [K][ldb]
    double* Out;       // [M][ldo]
    int M, N, K;
    int lda, ldb, ldo;
    // optional second product, split along K: Out2[s][M][ldo2] = A2[Ks]' B[Ks, col2 : col2 + N2]
    const double* A2;  // [K][lda] or nullptr
    double* Out2;      // [ksplit][M][ldo2]
    int N2, col2, ldo2, ksplit;
    int n2_valid;      // columns of the second product that are real (<= N2; 0: all): the rest of the last 64-column tile is padding
                       // whose MMAs are skipped (pq = 67 would otherwise pay for 128 columns)
};

template <bool WIDE>
__global__ void __maxnreg__(WIDE ? 128 : GEMM_MAXREG) gemm_tn_kernel(const Args g) {
    constexpr int NIX = WIDE ? NI + 1 : NI, LDB = WIDE ? LDB_SW : LDB_S, STG = WIDE ? STAGE_DOUBLES_W : STAGE_DOUBLES;
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tiles_n = g.N / BN;
    const int main_items = (g.M / BM) * tiles_n;
    const int KT_all = g.K / BK;

    const double* __restrict__ A;
    const double* __restrict__ B;
    double* __restrict__ Out;
    int m0, ldo, kt0, kt1;
    int ncols = BN;   // real columns of this CTA's tile
    int bcols = BN;   // columns of the B tile this CTA stages (72 for the wide last tile)
    int item = blockIdx.x;
    if (item < main_items) {
        const int tile_n = item % tiles_n, tile_m = item / tiles_n;
        m0 = tile_m * BM;
        A = g.A1 + m0;
        B = g.B + tile_n * BN;
        Out = g.Out + (size_t)m0 * g.ldo + tile_n * BN;
        ldo = g.ldo;
        kt0 = 0;
        kt1 = KT_all;
    } else {
        item -= main_items;
        const int tiles_n2 = WIDE ? (g.n2_valid - 8 + BN - 1) / BN : g.N2 / BN;
        const int s = item % g.ksplit;
        const int tile_n = (item / g.ksplit) % tiles_n2, tile_m = item / (g.ksplit * tiles_n2);
        const int per = (KT_all + g.ksplit - 1) / g.ksplit;
        m0 = tile_m * BM;
        A = g.A2 + m0;
        B = g.B + g.col2 + tile_n * BN;
        Out = g.Out2 + ((size_t)s * g.M + m0) * g.ldo2 + tile_n * BN;
        ldo = g.ldo2;
        kt0 = min(s * per, KT_all);
        kt1 = min(kt0 + per, KT_all);
        if (WIDE && tile_n == tiles_n2 - 1) bcols = BNW;
        if (g.n2_valid > 0) ncols = max(0, min(bcols, g.n2_valid - tile_n * BN));
    }
    const int KT = kt1 - kt0;

    const int wm = (warp % WARPS_M) * (BM / WARPS_M), wn = (warp / WARPS_M) * (BN / WARPS_N);
    const int fr = lane >> 2, fk = lane & 3;  // fragment row / k index of this lane
    // 8-column MMA tiles of this warp that hold real columns (warp-uniform); the upper half takes the wide tile's extra one
    const int nj = max(0, min((WIDE && wn > 0) ? NIX : NI, (ncols - wn + 7) / 8));

    double acc[MI][NIX][2];
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NIX; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    auto load_stage = [&](int kt, int slot) {
        double* As = smem + slot * STG;
        double* Bs = As + BK * LDA_S;
        const size_t k0 = (size_t)(kt0 + kt) * BK;
#pragma unroll
        for (int i = 0; i < (BK * BM / 2) / THREADS; i++) {  // 16-byte chunks of the A tile
            const int ch = tid + i * THREADS;
            const int r = ch / (BM / 2), c2 = (ch % (BM / 2)) * 2;
            cp_async16(As + r * LDA_S + c2, A + (k0 + r) * g.lda + c2);
        }
        if (WIDE) {   // B tile of 64 or 72 columns
            const int half = bcols / 2;
            for (int ch = tid; ch < BK * half; ch += THREADS) {
                const int r = ch / half, c2 = (ch - r * half) * 2;
                cp_async16(Bs + r * LDB + c2, B + (k0 + r) * g.ldb + c2);
            }
        } else {
#pragma unroll
            for (int i = 0; i < (BK * BN / 2) / THREADS; i++) {  // B tile
                const int ch = tid + i * THREADS;
                const int r = ch / (BN / 2), c2 = (ch % (BN / 2)) * 2;
                cp_async16(Bs + r * LDB + c2, B + (k0 + r) * g.ldb + c2);
            }
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
        const double* As = smem + (kt % STAGES) * STG;
        const double* Bs = As + BK * LDA_S;
#pragma unroll
        for (int ks = 0; ks < BK; ks += 4) {
            if (nj == 0) continue;
            double af[MI], bf[NIX];
#pragma unroll
            for (int i = 0; i < MI; i++) af[i] = As[(ks + fk) * LDA_S + wm + i * 8 + fr];
#pragma unroll
            for (int j = 0; j < NIX; j++)
                if (j < nj) bf[j] = Bs[(ks + fk) * LDB + wn + j * 8 + fr];
#pragma unroll
            for (int j = 0; j < NIX; j++) {
                if (j < nj) {
#pragma unroll
                    for (int i = 0; i < MI; i++) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
                }
            }
        }
    }
    cp_async_wait<0>();

    // epilogue: C fragment (row = lane/4, cols = 2*(lane%4) + {0,1}) -> 16-byte stores
#pragma unroll
    for (int i = 0; i < MI; i++) {
        double* orow = Out + (size_t)(wm + i * 8 + fr) * ldo + wn + 2 * fk;
#pragma unroll
        for (int j = 0; j < NIX; j++) {
            if (j < NI || (WIDE && wn > 0 && bcols == BNW))
                *reinterpret_cast<double2*>(orow + j * 8) = make_double2(acc[i][j][0], acc[i][j][1]);
        }
    }
}

inline cudaError_t launch(const Args& a, cudaStream_t stream) {
    // the attribute belongs to the DEVICE: one flag per device (a process may drive an engine on every GPU of the box)
    static bool configured[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !configured[dev]) {
        e = cudaFuncSetAttribute(gemm_tn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(gemm_tn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_W);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev] = true;
    }
    int items = (a.M / BM) * (a.N / BN);
    const int rem = a.n2_valid > 0 ? a.n2_valid % BN : 0;
    // ... and only when the grid still fills every SM twice without the extra column tile (fewer, fatter CTAs lose
    // at small chain counts: measured at 512 chains and at q = 129 with 64 chains)
    static int sms[64] = {};
    if (dev >= 0 && dev < 64 && sms[dev] == 0) {
        if (cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms[dev] < 1) sms[dev] = 148;
    }
    const int sm_count = (dev >= 0 && dev < 64) ? sms[dev] : 148;
    const int wide_tiles = a.n2_valid > 8 ? (a.n2_valid - 8 + BN - 1) / BN : 0;
    const bool wide = a.A2 && a.ksplit > 0 && rem >= 1 && rem <= 8 && wide_tiles > 0 &&
                      (wide_tiles - 1) * BN + BNW <= a.N2 &&   // the wide tile stays inside the padded width
                      items + (a.M / BM) * wide_tiles * a.ksplit >= CTAS_PER_SM * sm_count;
    if (a.A2 && a.ksplit > 0) items += (a.M / BM) * (wide ? (a.n2_valid - 8 + BN - 1) / BN : a.N2 / BN) * a.ksplit;
    if (wide) gemm_tn_kernel<true><<<items, THREADS, SMEM_BYTES_W, stream>>>(a);
    else gemm_tn_kernel<false><<<items, THREADS, SMEM_BYTES, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace dgemm
