// engine.cu -- lock-step multi-chain Gibbs sweep for OccuSpytial's samplers on B200 (sm_100a).
//
// One engine = one GPU's shard of chains.  The chain axis is the SIMD axis: every per-chain vector is
// stored chain-minor ([element][chain]) so that a warp is 32 chains looking at the same site / visit
// row, shared data (X, K, W, y) is a warp-uniform broadcast and all per-chain traffic is coalesced.
//
// One iteration (same Markov transition as <Sampler>.step(), logit.py:254-266, re-cut at z so that
// every conditional that only needs (z, alpha, psi) shares one pass over the sites):
//   1. psi   = [X K] theta                       DMMA GEMM            (logit.py:201, 485)
//   2. site pass: z | psi, alpha   (skipped at iteration 0: z comes from the data, base.py:113-118)
//                 omega_a | z, alpha -> partial sums of W_e' Om_a W_e and W_e'(y_e - 1/2)
//                 omega_b | psi                                       (logit.py:180-204, 234-252)
//   3. [G | H] = [Omega' Z | kd' U]              tcgen05 kind::i8 slice kernel (gram_tc.cuh; n >= 1024), else the
//                                                FP64 DMMA GEMM      (logit.py:222-231, 327-335, 481)
//   4. reduce the alpha partial sums (fixed order, no float atomics)
//   5. per-chain solve: tau, eta, beta, alpha    (logit.py:206-232, 478-485; distributions.pyx:42-110)
// The draw recorded after step 5 is (alpha, beta, tau) exactly as Chain.append sees it (base.py:238-239).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <type_traits>

#include "../../include/occu_b200.h"
#include "common.cuh"
#include "gemm_dmma.cuh"
#include "gram_tc.cuh"
#include "samplers.cuh"

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int set_err(const char* fmt, const char* a = "", const char* b = "") {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return -1;
}
#define CUDA_OK(call)                                                                  \
    do {                                                                               \
        cudaError_t _e = (call);                                                       \
        if (_e != cudaSuccess) return set_err("CUDA error %s at %s", cudaGetErrorString(_e), #call); \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int round_up(int x, int a) { return (x + a - 1) / a * a; }
__device__ __forceinline__ int round_up_dev(int x, int a) { return (x + a - 1) / a * a; }

// Philox stage ids (ctr[3] high byte)
enum { ST_Z = 0, ST_OMEGA_B = 1, ST_OMEGA_A = 2, ST_TAU = 3, ST_NORMALS = 4, ST_EPS = 5, ST_HOOK = 6 };

// sites per CTA in the site pass; measured on B200 (n = 10^4, 4096 chains): 16..48 -> 1.79..1.85 ms, 64 -> 2.18 ms
// (the 32 KB argument tile then costs occupancy); 32 balances the pass against the partial-sum reduction
#ifndef SITE_TILE_N
#define SITE_TILE_N 32
#endif
constexpr int SITE_TILE = SITE_TILE_N;
constexpr int CHAIN_TILE = 128;  // chains per CTA in the site pass (= padding unit of the chain axis)
constexpr int SOLVE_THREADS = 128;
constexpr int MAX_SMALL = 32;    // max p and r
constexpr int TC_BN = gtc::BN_DEFAULT;  // Z columns per tile of the tcgen05 Gram (6 level accumulators side by side in TMEM)

struct Control {
    uint32_t iter;     // iteration counter (Philox coordinate)
    int32_t first;     // 1 until the first site pass has run (z taken from the data)
    int32_t rec_index; // draw index to record at, < 0: none
    int32_t fails;     // number of chains with a numerical failure
};

struct HParams {
    double tau_rate, tau_shape;
    double a_prec[MAX_SMALL * MAX_SMALL], b_prec[MAX_SMALL * MAX_SMALL];
    double a_pm[MAX_SMALL], b_pm[MAX_SMALL];  // prec @ mu
};

// device view of an engine (passed to kernels by value)
struct View {
    int model, n, p, r, q, pq, V, C, Cp, chain_offset;
    int n_pad, kq_pad, M, M_pad, ldz, ldh, ksplit, n_tiles, NV, bw;
    uint64_t seed;
    PhiloxKeys rk;  // Philox round keys of the seed (constant-bank operands of the site pass)
    const double* W;
    const uint8_t* y;
    const int32_t* visit_ptr;
    const uint8_t* site_class;
    const double* Qrsr;
    const HParams* hp;
    double* UT;     // [kq_pad][n_pad]
    double* ZU;     // [n_pad][ldz]   Z (M_pad columns) | U (ldh columns)
    double* theta;  // [kq_pad][Cp]   beta rows then eta rows
    double* alpha;  // [r][Cp]
    double* tau;    // [Cp]
    double* psi;    // [n_pad][Cp]
    double* omega;  // [n_pad][Cp]
    double* kd;     // [n_pad][Cp]
    uint8_t* z;     // [n][Cp]
    double* G;      // [Cp][M_pad]    packed joint Gram per chain
    double* Hpart;  // [ksplit][Cp][ldh]  K-split partials of U'(z - 1/2)
    double* apart;  // [n_tiles][NV][Cp]
    double* asum;   // [NV][Cp]
    double* gam;    // [Cp]  Gamma(tau_shape, 1) variates of this iteration
    // exact-integer Gram path on the tcgen05 tensor cores (gram_tc.cuh)
    int int8;               // 1: G and H come from the tcgen05 kind::i8 slice kernel
    int tc_Kp, tc_nsplit, tc_Mrows;
    uint8_t* Aq;            // [SA + 1][Cp][tc_Kp]   base-256 digits of omega / 8, then z
    int8_t* Bq;             // [SB][tc_Mrows][tc_Kp] balanced base-256 digits of [Z | U] / column scale
    double* tscale;         // [tc_Mrows]  output scale per sliced row: 8 c_m (Gram columns), c_a (U columns)
    double* hshift;         // [ldh]       -colsum(U) / 2:  U'(z - 1/2) = U'z + hshift
    double* Gp;             // [tc_nsplit][Cp][M_pad] K-chunk partials (tc_nsplit > 1 only)
    double* Hp;             // [tc_nsplit][Cp][ldh]
    int* tc_err;            // barrier-timeout flag of the tcgen05 kernel
    double* solve_scratch;  // [Cp][solve_stride] working set of the RSR solve when it exceeds shared memory, else NULL
    size_t solve_stride;    // doubles per chain (0: shared-memory path)
    double* zprob;  // [n][Cp] or NULL
    double* eps;    // [n_pad][Cp]  probit only
    // ICAR only
    const double* X;      // [n][p]
    const double* Qband;  // [n][bw + 1]   Qband[i][d] = Q[i][i - d]
    double* eta_n;        // [n_pad][Cp]   eta (= spatial effect), chain-minor
    double* Lc;           // [Cp][n][bw + 1]  column-band Cholesky factor, Lc[j][e] = L[j + e][j]
    double* scratch;      // [Cp][8][n4]   y_b, y_1, x, zv, mean, omega | 1/L_jj, kd, spare  (n4 = n rounded up to 4)
    double* G0;     // [M_pad]      probit only: vech of [X K]'[X K], shared by all chains
    int32_t* status;  // [Cp]
    Control* ctl;
};

// kernel classes timed by the optional profiler (occu_set_profiling / occu_get_kernel_ms)
enum { KC_PSI_GEMM = 0, KC_SITE_PASS = 1, KC_GRAM_GEMM = 2, KC_REDUCE = 3, KC_SOLVE = 4, KC_COUNT = 5 };

struct occu_engine {
    View v;
    occu_problem_t prob;
    int pg_fp64, debug;
    long long launches;
    char* ws;
    size_t ws_bytes;
    gtc::Geometry geo;
    int sm_count;
    size_t smem_configured[3];  // dynamic shared memory already opted in for the ICAR / RSR / tiled solve kernels
    int tc_order;           // tile order of the tcgen05 Gram (gtc::Params::order; OCCU_TC_ORDER overrides)
    CUtensorMap tmA, tmB;
    bool slices_valid;      // Aq holds the digits of the current (omega, z)
    bool omega_in_slices;   // ... and v.omega was NOT written by the last site pass (non-debug engines)
    int profiling;
    double kernel_ms[KC_COUNT];
    long long kernel_calls[KC_COUNT];
    cudaEvent_t ev[KC_COUNT + 1];
};

// ------------------------------------------------------------------------------------------------
// layout / workspace carving (shared by occu_workspace_bytes and occu_create)
// ------------------------------------------------------------------------------------------------
struct Carver {
    char* base;
    size_t off;
    template <typename T>
    T* take(size_t count) {
        off = align_up(off, 256);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

// bytes of the RSR solve kernel's working set (shared memory, or per-chain global scratch for large q);
// must mirror the carve at the top of solve_logit_rsr_kernel
static size_t solve_bytes_for(int q, int p, int pq) {
    const size_t qb = (q + 3) / 4, q4 = qb * 4;
    const size_t nblk = (qb + 1) * (qb + 2) / 2;
    const size_t doubles = nblk * 17 + q4 * p + (size_t)p * p + 2 * (size_t)pq + 4 + 3 * q4 + 2 * MAX_SMALL + 8 + qb * 20;
    return doubles * sizeof(double) + 2 * nblk + 16;
}

// scratch / shared-memory sizes of the tiled large-q solve (solve_logit_rsr_tiled_kernel)
constexpr int TILED_QMAX = 1312;     // 41 tile rows: the per-chain vectors and 9 panel slots stay within shared memory
static size_t tiled_scratch_doubles(int q, int p);
static size_t tiled_smem_bytes(int q, int pq);

static int fill_dims(const occu_problem_t* P, View& v) {
    if (!P) return set_err("null problem");
    if (P->abi_version != OCCU_ABI_VERSION) return set_err("ABI version mismatch");
    if (P->model != OCCU_LOGIT_RSR && P->model != OCCU_PROBIT_RSR && P->model != OCCU_LOGIT_ICAR)
        return set_err("unknown model id");
    const bool icar = P->model == OCCU_LOGIT_ICAR;
    if (P->n < 1 || P->p < 1 || P->r < 1 || P->chains < 1 || P->V < 0) return set_err("bad dimensions");
    if (P->p > MAX_SMALL || P->r > 8) return set_err("p must be <= 32 and r <= 8");
    if (!icar && P->q < 1) return set_err("RSR model needs q >= 1");
    if (icar && (P->icar_bw < 1 || P->icar_bw >= P->n)) return set_err("ICAR model needs 1 <= icar_bw < n");
    if (icar && (P->icar_bw > 160 ||
                 ((size_t)(P->icar_bw + 1) * (P->icar_bw + 2) + 8 * (size_t)(P->icar_bw + 1) + 4096) * 8 > 200 * 1024))
        return set_err("ICAR half bandwidth too large for the shared-memory window (reorder the sites)");
    v.bw = icar ? P->icar_bw : 0;
    v.model = P->model;
    v.n = P->n; v.p = P->p; v.r = P->r; v.q = icar ? 0 : P->q; v.V = P->V;
    v.pq = v.p + v.q;
    v.C = P->chains;
    v.Cp = round_up(P->chains, CHAIN_TILE);
    v.chain_offset = P->chain_offset;
    v.seed = P->seed;
    philox_round_keys(P->seed, v.rk);
    v.n_pad = round_up(P->n, 128);
    v.kq_pad = round_up(v.pq, dgemm::BK);
    v.M = v.pq * (v.pq + 1) / 2;
    v.M_pad = round_up(v.M, dgemm::BN);
    v.ldh = round_up(v.pq, dgemm::BN);
    v.ldz = v.M_pad + v.ldh;
    v.ksplit = (v.n_pad / dgemm::BK >= 64) ? 8 : 1;
    v.int8 = (P->gram_int8 && P->model == OCCU_LOGIT_RSR) ? 1 : 0;
    if (v.int8) {
        v.ksplit = 1;
        const gtc::Geometry g = gtc::make_geometry(v.n, v.Cp, v.M, v.pq, TC_BN);
        v.tc_Kp = g.Kp;
        v.tc_nsplit = g.nsplit;
        v.tc_Mrows = g.Mrows;
    }
    if (P->model == OCCU_PROBIT_RSR) {
        // only the K-split product U'd runs: split it until the grid covers the GPU (2 CTAs x 148 SMs)
        const int tiles = (v.Cp / dgemm::BM) * (v.ldh / dgemm::BN);
        int ks = (296 + tiles - 1) / tiles;
        const int kt = v.n_pad / dgemm::BK;
        v.ksplit = ks < 1 ? 1 : (ks > kt ? kt : (ks > 64 ? 64 : ks));
    }
    v.n_tiles = (P->n + SITE_TILE - 1) / SITE_TILE;
    v.NV = P->r * (P->r + 1) / 2 + P->r;
    // solve kernel keeps Lambda (4x4-blocked, packed lower triangle) and a few vectors in shared memory
    if (!icar) {
        const size_t qb = (v.q + 3) / 4;
        const size_t solve_smem = solve_bytes_for(v.q, v.p, v.pq);
        // beyond ~200 the blocked factor no longer fits in shared memory: the tiled kernel then works out of a
        // per-chain global scratch block (32 x 32 tiles, DMMA updates)
        const char* tq = getenv("OCCU_TILED_MIN_Q");   // tuning knob: smallest q that takes the tiled kernel
        const bool tiled = solve_smem > 200 * 1024 || (tq && v.q >= atoi(tq));
        v.solve_stride = tiled ? (tiled_scratch_doubles(v.q, v.p) + 1) / 2 * 2 : 0;
        if (tiled ? v.q > TILED_QMAX : qb > (size_t)SOLVE_THREADS)
            return set_err("RSR rank q too large (q <= 1312 in this build)");
    }
    return 0;
}

static void carve(View& v, char* base, size_t* total, int debug) {
    Carver c{base, 0};
    const size_t Cp = v.Cp;
    v.hp = c.take<HParams>(1);
    v.ctl = c.take<Control>(1);
    v.W = c.take<double>((size_t)v.V * v.r + 1);
    v.y = c.take<uint8_t>((size_t)v.V + 1);
    v.visit_ptr = c.take<int32_t>((size_t)v.n + 1);
    v.site_class = c.take<uint8_t>(v.n);
    const bool icar = v.model == OCCU_LOGIT_ICAR;
    v.Qrsr = icar ? nullptr : c.take<double>((size_t)v.q * v.q);
    v.UT = icar ? nullptr : c.take<double>((size_t)v.kq_pad * v.n_pad);
    // the face-splitting design [Z | U] is the B operand of the FP64 DMMA Gram only; the tcgen05 path slices it from UT
    v.ZU = (icar || v.int8) ? nullptr : c.take<double>((size_t)v.n_pad * v.ldz);
    v.X = icar ? c.take<double>((size_t)v.n * v.p) : nullptr;
    v.Qband = icar ? c.take<double>((size_t)v.n * (v.bw + 1)) : nullptr;
    v.eta_n = icar ? c.take<double>((size_t)v.n_pad * Cp) : nullptr;
    v.Lc = icar ? c.take<double>(Cp * (size_t)v.n * (v.bw + 1)) : nullptr;
    v.scratch = icar ? c.take<double>(Cp * (size_t)8 * round_up(v.n, 4)) : nullptr;
    v.theta = c.take<double>((size_t)v.kq_pad * Cp);
    v.alpha = c.take<double>((size_t)v.r * Cp);
    v.tau = c.take<double>(Cp);
    v.psi = c.take<double>((size_t)v.n_pad * Cp);
    v.omega = c.take<double>((size_t)v.n_pad * Cp);
    v.kd = v.int8 ? nullptr : c.take<double>((size_t)v.n_pad * Cp);  // the tcgen05 Gram reads z as a byte plane instead
    v.z = c.take<uint8_t>((size_t)v.n * Cp);
    const bool probit = v.model == OCCU_PROBIT_RSR;
    v.G = (probit || icar) ? nullptr : c.take<double>(Cp * (size_t)v.M_pad);
    v.eps = probit ? c.take<double>((size_t)v.n_pad * Cp) : nullptr;
    v.G0 = probit ? c.take<double>(v.M_pad) : nullptr;
    v.Hpart = icar ? nullptr : c.take<double>((size_t)v.ksplit * Cp * v.ldh);
    v.apart = c.take<double>((size_t)v.n_tiles * v.NV * Cp);
    v.asum = c.take<double>((size_t)v.NV * Cp);
    v.gam = c.take<double>(Cp);
    v.solve_scratch = v.solve_stride ? c.take<double>(Cp * v.solve_stride) : nullptr;
    if (v.int8) {
        v.Aq = c.take<uint8_t>((size_t)(gtc::SA + 1) * Cp * v.tc_Kp);
        v.Bq = c.take<int8_t>((size_t)gtc::SB * v.tc_Mrows * v.tc_Kp);
        v.tscale = c.take<double>(v.tc_Mrows);
        v.hshift = c.take<double>(v.ldh);
        v.Gp = v.tc_nsplit > 1 ? c.take<double>((size_t)v.tc_nsplit * Cp * v.M_pad) : nullptr;
        v.Hp = v.tc_nsplit > 1 ? c.take<double>((size_t)v.tc_nsplit * Cp * v.ldh) : nullptr;
        v.tc_err = c.take<int>(1);
    }
    v.zprob = debug ? c.take<double>((size_t)v.n * Cp) : nullptr;
    v.status = c.take<int32_t>(Cp);
    *total = align_up(c.off, 256);
}

// ------------------------------------------------------------------------------------------------
// setup kernels
// ------------------------------------------------------------------------------------------------
// UT[a][i] = U[i][a], ZU[i][vech(a,b)] = U[i][a] U[i][b], ZU[i][M_pad + a] = U[i][a], U = [X K]
__global__ void build_design_kernel(View v, const double* __restrict__ X, const double* __restrict__ K) {
    const int i = blockIdx.x;
    extern __shared__ double u[];
    for (int a = threadIdx.x; a < v.pq; a += blockDim.x)
        u[a] = (a < v.p) ? X[(size_t)i * v.p + a] : K[(size_t)i * v.q + (a - v.p)];
    __syncthreads();
    for (int a = threadIdx.x; a < v.pq; a += blockDim.x) {
        v.UT[(size_t)a * v.n_pad + i] = u[a];
        if (v.ZU) v.ZU[(size_t)i * v.ldz + v.M_pad + a] = u[a];
    }
    if (!v.ZU) return;
    for (int m = threadIdx.x; m < v.M; m += blockDim.x) {
        int a = (int)((sqrtf(8.0f * (float)m + 1.0f) - 1.0f) * 0.5f);
        while (a * (a + 1) / 2 > m) a--;
        while ((a + 1) * (a + 2) / 2 <= m) a++;
        const int b = m - a * (a + 1) / 2;
        v.ZU[(size_t)i * v.ldz + m] = u[a] * u[b];
    }
}

// z from the data (base.py:113-118): 0 at surveyed-never-detected sites, 1 elsewhere; kd = z - 1/2
__global__ void reset_z_kernel(View v) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)v.n * v.Cp;
    if (idx >= total) return;
    const int i = (int)(idx / v.Cp);
    const uint8_t zc = v.site_class[i] == OCCU_SITE_NOT_OBS ? 0 : 1;
    v.z[idx] = zc;
    if (v.kd) v.kd[idx] = (double)zc - 0.5;
}

__global__ void set_control_kernel(Control* c, uint32_t iter, int first, int rec_index, int clear_fails) {
    c->iter = iter;
    c->first = first;
    c->rec_index = rec_index;
    if (clear_fails) c->fails = 0;
}

__global__ void set_rec_index_kernel(Control* c, int rec_index) { c->rec_index = rec_index; }

__global__ void advance_kernel(Control* c) {
    c->iter += 1;
    c->first = 0;
    if (c->rec_index >= 0) c->rec_index += 1;
}

// chain-major [C][d] -> chain-minor [d][Cp]; padded chains replicate chain (c mod C)
template <typename T>
__global__ void to_chain_minor_kernel(const T* __restrict__ src, T* __restrict__ dst, int C, int Cp, int d) {
    __shared__ T tile[32][33];
    const int c0 = blockIdx.x * 32, e0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, e = e0 + threadIdx.x;
        if (c < Cp && e < d) tile[j][threadIdx.x] = src[(size_t)(c % C) * d + e];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int e = e0 + j, c = c0 + threadIdx.x;
        if (c < Cp && e < d) dst[(size_t)e * Cp + c] = tile[threadIdx.x][j];
    }
}

template <typename T>
__global__ void to_chain_major_kernel(const T* __restrict__ src, T* __restrict__ dst, int C, int Cp, int d) {
    __shared__ T tile[32][33];
    const int c0 = blockIdx.x * 32, e0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int e = e0 + j, c = c0 + threadIdx.x;
        if (c < Cp && e < d) tile[j][threadIdx.x] = src[(size_t)e * Cp + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = c0 + j, e = e0 + threadIdx.x;
        if (c < C && e < d) dst[(size_t)c * d + e] = tile[threadIdx.x][j];
    }
}

// ------------------------------------------------------------------------------------------------
// site pass (logit): z-update, omega_a draws + alpha partial sums, omega_b draw
//
// CTA = 128 chains x SITE_TILE sites, thread = chain.
//  phase 1 (lock-step over the tile's sites, coalesced): z | psi, alpha in FP64; z, kd written; psi kept
//          in shared memory as the PG argument; z bits kept in a register.
//  phase 2 (lanes desynchronise): every lane walks its OWN stream of PG draws -- for each site the visits
//          of an occupied surveyed site (omega_a) and then the site itself (omega_b) -- running one
//          rejection round per loop trip (pg1_round).  An accepted lane commits and immediately fetches
//          its next draw, a rejected lane retries: no lane waits for another lane's rejection loop.
//  phase 3: omega_b (staged in the lane's own shared-memory column) is written out coalesced.
// Philox streams are keyed by (chain, visit row | site, iteration, stage), so the result does not
// depend on how lanes interleave.
// ------------------------------------------------------------------------------------------------
//  SLICES (tcgen05 Gram path): phase 3 writes omega as SA base-256 digits of its 48-bit fixed-point value and z as one
//          more byte plane, site-contiguous per chain (gram_tc.cuh: Aq[s][chain][site]) -- the A operand of the Gram
//          kernel, 7 bytes per site instead of the two FP64 arrays omega and kd; a debug engine also writes omega.
// resident CTAs per SM the register allocation aims at: 8 (64 registers) for the usual r <= 4, fewer when the
// r (r + 1) / 2 + r FP64 accumulators no longer fit
constexpr int site_min_blocks(int r) { return r <= 4 ? 8 : (r <= 6 ? 5 : 3); }
//  FUSE_PSI: psi = [X K] theta of the tile is computed here, on the FP64 tensor pipe (DMMA m8n8k4: 8 sites x 32 chains
//          per warp and step, operands straight from L1/L2), instead of by a separate GEMM that writes psi [n x chains]
//          to HBM for this kernel to read back: the DMMA work hides under the ALU-bound PG sampling of the other warps.
constexpr int PSI_SUB = 8;   // sites per DMMA sub-tile
template <int R, typename PGT, bool SLICES, bool FUSE_PSI>
__global__ void __launch_bounds__(CHAIN_TILE, site_min_blocks(R)) site_pass_logit_kernel(const View v) {
    constexpr int NA = R * (R + 1) / 2, NV = NA + R;
    extern __shared__ __align__(16) unsigned char site_smem[];
    PGT (*arg_s)[CHAIN_TILE] = reinterpret_cast<PGT(*)[CHAIN_TILE]>(site_smem);  // column tid is private to thread tid
    // FUSE_PSI: FP64 psi of the current sub-tile, [PSI_SUB][CHAIN_TILE]; a warp only touches its own 32 columns
    double (*psi_d)[CHAIN_TILE] = reinterpret_cast<double(*)[CHAIN_TILE]>(site_smem + sizeof(PGT) * SITE_TILE * CHAIN_TILE);
    const int tid = threadIdx.x;
    const uint32_t iter = v.ctl->iter;
    const bool first = v.ctl->first != 0;
    const size_t Cp = v.Cp;
    const int tile = blockIdx.x;
    const int c = blockIdx.y * CHAIN_TILE + tid;
    const int i0 = tile * SITE_TILE;
    const int i1 = min(i0 + SITE_TILE, v.n);
    const uint32_t gc = (uint32_t)(v.chain_offset + c);

    double al[R];
#pragma unroll
    for (int k = 0; k < R; k++) al[k] = v.alpha[(size_t)k * Cp + c];

    // ---- phase 1: z | . ----
    uint64_t zbits = 0ull;
    static_assert(SITE_TILE % PSI_SUB == 0, "whole DMMA sub-tiles");
    for (int ib = i0; ib < i1; ib += PSI_SUB) {
    if (FUSE_PSI) {
        // psi[ib .. ib + 8)[32 chains of this warp] = U theta:  D (8 sites x 8 chains) += A (8 x 4: UT'[site][k]) B (4 x 8:
        // theta[k][chain]); rows of UT and theta beyond pq are zero up to kq_pad
        const int lane = tid & 31, g = lane >> 2, t4 = lane & 3;
        const int cw = c - lane;   // first chain of this warp
        double acc[4][2];
#pragma unroll
        for (int nb = 0; nb < 4; nb++) acc[nb][0] = acc[nb][1] = 0.0;
        const double* ua = v.UT + (size_t)t4 * v.n_pad + ib + g;       // n_pad >= i0 + SITE_TILE: no bounds check
        const double* tb = v.theta + (size_t)t4 * Cp + cw + g;
        const int ksteps = (v.pq + 3) >> 2;
#pragma unroll 4
        for (int ks = 0; ks < ksteps; ks++) {
            const double a = ua[(size_t)ks * 4 * v.n_pad];
            double b[4];
#pragma unroll
            for (int nb = 0; nb < 4; nb++) b[nb] = tb[(size_t)ks * 4 * Cp + nb * 8];
#pragma unroll
            for (int nb = 0; nb < 4; nb++) dgemm::dmma884(acc[nb][0], acc[nb][1], a, b[nb]);
        }
        __syncwarp();   // the previous sub-tile's readers are done
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
            psi_d[g][(tid & ~31) + nb * 8 + 2 * t4] = acc[nb][0];
            psi_d[g][(tid & ~31) + nb * 8 + 2 * t4 + 1] = acc[nb][1];
        }
        __syncwarp();
    }
    for (int i = ib; i < min(ib + PSI_SUB, i1); i++) {
        const int cls = v.site_class[i];
        const size_t off = (size_t)i * Cp + c;
        const double psi = FUSE_PSI ? psi_d[i - ib][tid] : v.psi[off];
        if (FUSE_PSI && v.zprob) v.psi[off] = psi;   // debug engines keep psi readable
        int zc;
        if (first) {
            zc = v.z[off];
        } else {
            if (cls == OCCU_SITE_OBS) {
                zc = 1;
            } else {
                // logit.py:241-245 / 249:  pr = expit(psi) prod / (1 - expit(psi) + expit(psi) prod) with
                // prod = prod_j expit(-w_ij alpha)  ==  1 / (1 + exp(-psi) prod_j (1 + exp(w_ij alpha))): one exp per
                // visit and one division per site; not-surveyed sites have the empty product
                double grow = exp(-psi);
                if (cls == OCCU_SITE_NOT_OBS) {
                    const int v1 = v.visit_ptr[i + 1];
                    for (int vv = v.visit_ptr[i]; vv < v1; vv++) {
                        double lin = 0.0;
#pragma unroll
                        for (int k = 0; k < R; k++) lin += v.W[(size_t)vv * R + k] * al[k];
                        grow *= 1.0 + exp(lin);
                    }
                }
                const double pr = 1.0 / (1.0 + grow);
                if (v.zprob) v.zprob[off] = pr;
                Philox g;
                g.init(v.seed, gc, (uint32_t)i, iter, ST_Z);
                g.refill_rk(v.rk);
                const uint64_t bits = ((uint64_t)g.buf[0] << 21) ^ (uint64_t)(g.buf[1] >> 11);  // == g.unif_d()
                zc = ((double)bits + 0.5) * (1.0 / 9007199254740992.0) < pr ? 1 : 0;  // logit.py:246 / 250
            }
            v.z[off] = (uint8_t)zc;
            if (!SLICES) v.kd[off] = (double)zc - 0.5;  // logit.py:252
        }
        zbits |= (uint64_t)zc << (i - i0);
        arg_s[i - i0][tid] = (PGT)psi;
    }
    }  // sub-tiles

    // ---- phase 2: per-lane stream of PG draws ----
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) acc[k] = 0.0;
    int i = i0 - 1, vv = 0, vend = 0, cur = 0;
    bool site_done = true, active = true, have = false, is_visit = false;
    typename PGSel<PGT>::Task task;
    Philox g;
    double wv[R];  // design row of the visit being drawn (kept from the fetch to the commit)
#pragma unroll
    for (int k = 0; k < R; k++) wv[k] = 0.0;
    while (__any_sync(0xffffffffu, active || have)) {
        if (!have && active) {
            // next draw of this lane's stream, loop-free: the visits of an occupied surveyed site
            // (omega_a ~ PG(1, w_ij alpha), logit.py:187-193), then the site itself (omega_b ~ PG(1, psi),
            // logit.py:201-204), then the next site
            if (vv >= vend && site_done) {
                if (++i >= i1) {
                    active = false;
                } else {
                    site_done = false;
                    if (v.site_class[i] != OCCU_SITE_NOT_SURVEYED && ((zbits >> (i - i0)) & 1ull)) {
                        vv = v.visit_ptr[i];
                        vend = v.visit_ptr[i + 1];
                    }
                }
            }
            if (active) {
                PGT carg;
                is_visit = vv < vend;
                if (is_visit) {
                    cur = vv++;
                    double lin = 0.0;
#pragma unroll
                    for (int k = 0; k < R; k++) {
                        wv[k] = v.W[(size_t)cur * R + k];
                        lin += wv[k] * al[k];
                    }
                    carg = (PGT)lin;
                } else {
                    site_done = true;
                    cur = i;
                    carg = arg_s[i - i0][tid];
                }
                g.init(v.seed, gc, (uint32_t)cur, iter, is_visit ? ST_OMEGA_A : ST_OMEGA_B);
                PGSel<PGT>::setup(carg, task);  // one shared copy of the setup code for both kinds of draw
                have = true;
            }
        }
        if (have) {
            PGT ub, ux, u3, us, x;
            RT<PGT>::unif4_rk(g, v.rk, ub, ux, u3, us);
            if (PGSel<PGT>::round(task, ub, ux, u3, us, x)) {
                have = false;
                const double om = 0.25 * (double)x;
                if (is_visit) {
                    const double yk = (double)v.y[cur] - 0.5;
                    int idx = 0;
#pragma unroll
                    for (int a = 0; a < R; a++) {
                        const double owa = om * wv[a];
#pragma unroll
                        for (int b = 0; b <= a; b++, idx++) acc[idx] = fma(owa, wv[b], acc[idx]);  // logit.py:222
                    }
#pragma unroll
                    for (int a = 0; a < R; a++) acc[NA + a] = fma(wv[a], yk, acc[NA + a]);  // logit.py:223
                } else {
                    arg_s[cur - i0][tid] = (PGT)om;
                }
            }
        }
    }

    // ---- phase 3: write-out ----
    if (!SLICES || v.zprob)
        for (int ii = i0; ii < i1; ii++) v.omega[(size_t)ii * Cp + c] = (double)arg_s[ii - i0][tid];
    if (SLICES) {
        static_assert(SITE_TILE % 16 == 0, "16 sites per 128-bit store");
        const size_t plane = (size_t)Cp * v.tc_Kp;
        uint8_t* row = v.Aq + (size_t)c * v.tc_Kp + i0;
#pragma unroll 1
        for (int g = 0; g < SITE_TILE / 16; g++) {
            uint32_t w[gtc::SA + 1][4];
#pragma unroll
            for (int s = 0; s <= gtc::SA; s++)
#pragma unroll
                for (int k = 0; k < 4; k++) w[s][k] = 0u;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int ii = g * 16 + j;
                unsigned long long q = 0ull;
                if (i0 + ii < i1) {
                    const double x = (double)arg_s[ii][tid] * 35184372088832.0;  // omega 2^45 = (omega / 8) 2^48
                    q = x >= 281474976710655.0 ? 281474976710655ull : (x > 0.0 ? (unsigned long long)x : 0ull);
                }
                const int sh = 8 * (j & 3);
#pragma unroll
                for (int s = 0; s < gtc::SA; s++) w[s][j >> 2] |= (uint32_t)((q >> (8 * (gtc::SA - 1 - s))) & 0xFFull) << sh;
                w[gtc::SA][j >> 2] |= (uint32_t)((zbits >> ii) & 1ull) << sh;
            }
#pragma unroll
            for (int s = 0; s <= gtc::SA; s++)
                *reinterpret_cast<uint4*>(row + (size_t)s * plane + g * 16) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
        }
    }
#pragma unroll
    for (int k = 0; k < NV; k++) v.apart[((size_t)tile * NV + k) * Cp + c] = acc[k];
}

// Variant used by occu_debug_gaussian_half when omega_a is injected: alpha partial sums only.
template <int R>
__global__ void __launch_bounds__(CHAIN_TILE) alpha_partials_injected_kernel(const View v, const double* __restrict__ omega_a_cm) {
    const bool probit = v.model == OCCU_PROBIT_RSR;
    constexpr int NA = R * (R + 1) / 2, NV = NA + R;
    const int c = blockIdx.y * CHAIN_TILE + threadIdx.x;
    const int tile = blockIdx.x;
    const int i0 = tile * SITE_TILE;
    const int i1 = min(i0 + SITE_TILE, v.n);
    const size_t Cp = v.Cp;
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) acc[k] = 0.0;
    const double* oma = omega_a_cm + (size_t)(c % v.C) * v.V;
    for (int i = i0; i < i1; i++) {
        const int cls = v.site_class[i];
        if (cls == OCCU_SITE_NOT_SURVEYED || !v.z[(size_t)i * Cp + c]) continue;
        for (int vv = v.visit_ptr[i]; vv < v.visit_ptr[i + 1]; vv++) {
            double w[R];
#pragma unroll
            for (int k = 0; k < R; k++) w[k] = v.W[(size_t)vv * R + k];
            const double om = probit ? 1.0 : oma[vv];                               // probit.py:233 vs logit.py:222
            const double yk = probit ? oma[vv] : (double)v.y[vv] - 0.5;             // probit.py:234 vs logit.py:223
            int idx = 0;
#pragma unroll
            for (int a = 0; a < R; a++) {
#pragma unroll
                for (int b = 0; b <= a; b++) acc[idx++] += om * w[a] * w[b];
            }
#pragma unroll
            for (int a = 0; a < R; a++) acc[NA + a] += w[a] * yk;
        }
    }
#pragma unroll
    for (int k = 0; k < NV; k++) v.apart[((size_t)tile * NV + k) * Cp + c] = acc[k];
}

// G0[m] = sum_i Z[i][m]: the joint Gram of [X K] with unit weights (probit: K'K, K'X, X'X; probit.py:112, 144)
__global__ void gram_unit_kernel(View v) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= v.M) return;
    double s = 0.0;
    for (int i = 0; i < v.n; i++) s += v.ZU[(size_t)i * v.ldz + m];
    v.G0[m] = s;
}

// injected randomness for the parity tests (all chain-major, NULL = Philox)
struct SiteDebug {
    const double* u_site;   // [C][n]  uniforms of the omega_b truncated normals (probit.py:203-209)
    const double* u_visit;  // [C][V]  uniforms of the omega_a truncated normals (probit.py:191-194)
    const double* nrm_eps;  // [C][n]  normals of the eps update (probit.py:220)
    int keep_z;             // 1: do not redraw z (the reference draws omega_b from the CURRENT z)
};

// ------------------------------------------------------------------------------------------------
// site pass (probit): z | ., omega_a (truncated normal) + alpha partial sums, omega_b, eps, d = omega_b - eps
// No rejection loops (inverse-CDF draws), so the lanes stay in lock-step over the sites of the tile.
// ------------------------------------------------------------------------------------------------
template <int R, bool F64>
__global__ void __launch_bounds__(CHAIN_TILE) site_pass_probit_kernel(const View v, const SiteDebug dbg) {
    constexpr int NA = R * (R + 1) / 2, NV = NA + R;
    const int c = blockIdx.y * CHAIN_TILE + threadIdx.x;
    const int tile = blockIdx.x;
    const int i0 = tile * SITE_TILE;
    const int i1 = min(i0 + SITE_TILE, v.n);
    const uint32_t iter = v.ctl->iter;
    const bool keep_z = v.ctl->first != 0 || dbg.keep_z;
    const uint32_t gc = (uint32_t)(v.chain_offset + c);
    const size_t Cp = v.Cp;
    const int cr = c % v.C;

    double al[R];
#pragma unroll
    for (int k = 0; k < R; k++) al[k] = v.alpha[(size_t)k * Cp + c];
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) acc[k] = 0.0;

    for (int i = i0; i < i1; i++) {
        const int cls = v.site_class[i];
        const size_t off = (size_t)i * Cp + c;
        const double lp = v.psi[off];   // x_i beta + K_i eta
        const double ep = v.eps[off];
        const int v0 = v.visit_ptr[i], v1 = v.visit_ptr[i + 1];
        int zc;
        if (keep_z) {
            zc = v.z[off];
        } else {
            if (cls == OCCU_SITE_OBS) {
                zc = 1;
            } else {
                double pr = ndtr_d(lp + ep);  // probit.py:251 / 258
                if (cls == OCCU_SITE_NOT_OBS) {
                    double prod = 1.0;  // probit.py:252-253: prod_j (1 - Phi(w_ij alpha))
                    for (int vv = v0; vv < v1; vv++) {
                        double lin = 0.0;
#pragma unroll
                        for (int k = 0; k < R; k++) lin += v.W[(size_t)vv * R + k] * al[k];
                        prod *= 1.0 - ndtr_d(lin);
                    }
                    const double num = pr * prod;
                    pr = num / ((1.0 - pr) + num);  // probit.py:254-255
                }
                if (v.zprob) v.zprob[off] = pr;
                Philox g;
                g.init(v.seed, gc, (uint32_t)i, iter, ST_Z);
                zc = g.unif_d() < pr ? 1 : 0;
            }
            v.z[off] = (uint8_t)zc;
        }
        if (cls != OCCU_SITE_NOT_SURVEYED && zc) {
            // omega_a ~ TN(w_ij alpha, 1) on (0, inf) if y_ij = 1 else (-inf, 0)   (probit.py:180-194)
            for (int vv = v0; vv < v1; vv++) {
                double w[R];
                double lin = 0.0;
#pragma unroll
                for (int k = 0; k < R; k++) {
                    w[k] = v.W[(size_t)vv * R + k];
                    lin += w[k] * al[k];
                }
                const bool positive = v.y[vv] == 1;
                double oma;
                if (dbg.u_visit) {
                    oma = truncnorm_icdf(lin, positive, dbg.u_visit[(size_t)cr * v.V + vv]);
                } else {
                    Philox g;
                    g.init(v.seed, gc, (uint32_t)vv, iter, ST_OMEGA_A);
                    oma = F64 ? truncnorm_draw(lin, positive, g) : truncnorm_draw_f32(lin, positive, g);
                }
                int idx = 0;
#pragma unroll
                for (int a = 0; a < R; a++) {
#pragma unroll
                    for (int b = 0; b <= a; b++) acc[idx++] += w[a] * w[b];  // probit.py:233
                }
#pragma unroll
                for (int a = 0; a < R; a++) acc[NA + a] += w[a] * oma;  // probit.py:234
            }
        }
        // omega_b ~ TN(lp + eps, 1), sign by z (probit.py:196-209); eps | . (probit.py:216-221)
        Philox g;
        g.init(v.seed, gc, (uint32_t)i, iter, ST_OMEGA_B);
        const double loc = lp + ep;
        const double omb = dbg.u_site ? truncnorm_icdf(loc, zc == 1, dbg.u_site[(size_t)cr * v.n + i])
                                      : (F64 ? truncnorm_draw(loc, zc == 1, g) : truncnorm_draw_f32(loc, zc == 1, g));
        const double nz = dbg.nrm_eps ? dbg.nrm_eps[(size_t)cr * v.n + i] : (F64 ? g.normal_d() : normal_f32(g));
        const double ep_new = 0.5 * (omb - lp) + nz * 0.70710678118654752440;
        v.omega[off] = omb;
        v.eps[off] = ep_new;
        v.kd[off] = omb - ep_new;  // d: right-hand-side vector of both the eta and the beta systems
    }
#pragma unroll
    for (int k = 0; k < NV; k++) v.apart[((size_t)tile * NV + k) * Cp + c] = acc[k];
}

// ------------------------------------------------------------------------------------------------
// per-chain solve: tau, eta, beta, alpha
// ------------------------------------------------------------------------------------------------
struct SolveDebug {
    const double* normals;  // [C][q+p+r] or NULL
    const double* tau_in;   // [C] or NULL
    occu_debug_out_t out;
    int enabled;
};

__device__ __forceinline__ int tri(int a, int b) { return a * (a + 1) / 2 + b; }  // b <= a

// In-place Cholesky of the d x d block whose (i, j) entry (j <= i) lives at A[tri(o + i, o + j)].
// Returns false (uniformly) on a non-positive pivot.
__device__ bool block_cholesky_packed(double* A, int o, int d, int tid, int nthreads) {
    for (int j = 0; j < d; j++) {
        __syncthreads();
        const double piv = A[tri(o + j, o + j)];
        if (!(piv > 0.0)) return false;
        const double dj = sqrt(piv);
        const double inv = 1.0 / dj;
        __syncthreads();
        for (int i = j + tid; i < d; i += nthreads) {
            if (i == j) A[tri(o + j, o + j)] = dj;
            else A[tri(o + i, o + j)] *= inv;
        }
        __syncthreads();
        const int m = d - 1 - j;
        const int cnt = m * (m + 1) / 2;
        for (int t = tid; t < cnt; t += nthreads) {
            int ii = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
            while (ii * (ii + 1) / 2 > t) ii--;
            while ((ii + 1) * (ii + 2) / 2 <= t) ii++;
            const int kk = t - ii * (ii + 1) / 2;
            const int i = j + 1 + ii, k = j + 1 + kk;
            A[tri(o + i, o + k)] -= A[tri(o + i, o + j)] * A[tri(o + k, o + j)];
        }
    }
    __syncthreads();
    return true;
}

// Sequential small dense solve by one thread: x = P^-1 b + L^-T z, P = L L' (full d x d row-major in S,
// lower triangle overwritten by L).  Returns false on a non-positive pivot.  distributions.pyx:88-105.
__device__ bool small_precision_mvnorm(double* S, int d, const double* b, const double* zn, double* x, double* tmp) {
    for (int j = 0; j < d; j++) {
        double s = S[j * d + j];
        for (int k = 0; k < j; k++) s -= S[j * d + k] * S[j * d + k];
        if (!(s > 0.0)) return false;
        const double dj = sqrt(s);
        S[j * d + j] = dj;
        for (int i = j + 1; i < d; i++) {
            double t = S[i * d + j];
            for (int k = 0; k < j; k++) t -= S[i * d + k] * S[j * d + k];
            S[i * d + j] = t / dj;
        }
    }
    for (int i = 0; i < d; i++) {  // L y = b
        double t = b[i];
        for (int k = 0; k < i; k++) t -= S[i * d + k] * tmp[k];
        tmp[i] = t / S[i * d + i];
    }
    for (int i = d - 1; i >= 0; i--) {  // L' x = y + z
        double t = tmp[i] + zn[i];
        for (int k = i + 1; k < d; k++) t -= S[k * d + i] * x[k];
        x[i] = t / S[i * d + i];
    }
    return true;
}

// Per-chain scalar work that does not depend on the Gram, one THREAD per chain (chain-minor, coalesced):
//  * alpha | omega_a, z  (logit.py:219-224 / probit.py:231-235), committed and recorded here,
//  * the unit-rate Gamma variate of tau | eta (tau = gam / rate, logit.py:206-209).
// Keeping these serial FP64 chains (sqrt / divide / log) out of the one-CTA-per-chain solve kernels matters:
// there a single lane would hold 127 others at a barrier for ~10^4 cycles per chain.
__global__ void __launch_bounds__(32) alpha_tau_kernel(const View v, const SolveDebug dbg, double* rec_alpha,
                                                        int rec_capacity) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= v.Cp) return;
    const int r = v.r, NA = r * (r + 1) / 2, NV = v.NV;
    const size_t Cp = v.Cp;
    const HParams* hp = v.hp;
    const uint32_t iter = v.ctl->iter;
    const int rec_index = v.ctl->rec_index;
    const uint32_t gc = (uint32_t)(v.chain_offset + c);
    const int cr = c % v.C;
    const int d_eta = v.model == OCCU_LOGIT_ICAR ? v.n : v.q;   // normals stream: eta, beta, alpha
    double sums[8 * 9 / 2 + 8];
    for (int k = 0; k < NV; k++) sums[k] = v.asum[(size_t)k * Cp + c];  // reduced by reduce_partials_kernel
    double S[64], b[8], x[8], tmp[8], zn[8];
    for (int a = 0; a < r; a++) {
        for (int bb = 0; bb < r; bb++) S[a * r + bb] = sums[bb <= a ? tri(a, bb) : tri(bb, a)] + hp->a_prec[a * r + bb];
        b[a] = sums[NA + a] + hp->a_pm[a];
        if (dbg.normals) {
            zn[a] = dbg.normals[(size_t)cr * (d_eta + v.p + r) + d_eta + v.p + a];
        } else {
            Philox g;
            g.init(v.seed, gc, (uint32_t)(d_eta + v.p + a), iter, ST_NORMALS);
            zn[a] = g.normal_d();
        }
    }
    if (dbg.enabled && c < v.C) {
        if (dbg.out.alpha_prec)
            for (int idx = 0; idx < r * r; idx++) dbg.out.alpha_prec[(size_t)c * r * r + idx] = S[idx];
        if (dbg.out.alpha_b)
            for (int a = 0; a < r; a++) dbg.out.alpha_b[(size_t)c * r + a] = b[a];
    }
    if (!small_precision_mvnorm(S, r, b, zn, x, tmp)) {
        v.status[c] = 1;
        if (c < v.C) atomicAdd(&v.ctl->fails, 1);
    } else if (!v.status[c]) {
        for (int a = 0; a < r; a++) v.alpha[(size_t)a * Cp + c] = x[a];
        if (rec_alpha && rec_index >= 0 && rec_index < rec_capacity && c < v.C) {
            const size_t row = (size_t)c * rec_capacity + rec_index;
            for (int a = 0; a < r; a++) rec_alpha[row * r + a] = x[a];
        }
    }
    Philox g;
    g.init(v.seed, gc, 0u, iter, ST_TAU);
    v.gam[c] = gamma_draw(hp->tau_shape, g);
}

// asum[k][c] = sum_tile apart[tile][k][c], fixed order
// Block = 32 sums x 8 interleaved slices of the site tiles (8 x the loads in flight of a one-thread-per-sum loop),
// combined in slice order: the summation order is fixed, so the result is reproducible.
constexpr int RED_SPLIT = 8;
__global__ void __launch_bounds__(32 * RED_SPLIT) reduce_partials_kernel(const View v) {
    __shared__ double part[RED_SPLIT][33];
    const size_t idx = (size_t)blockIdx.x * 32 + threadIdx.x;
    const size_t per = (size_t)v.NV * v.Cp;
    double s = 0.0;
    if (idx < per) {
#pragma unroll 4
        for (int t = threadIdx.y; t < v.n_tiles; t += RED_SPLIT) s += v.apart[(size_t)t * per + idx];
    }
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && idx < per) {
        double tot = 0.0;
#pragma unroll
        for (int k = 0; k < RED_SPLIT; k++) tot += part[k][threadIdx.x];
        v.asum[idx] = tot;
    }
}

// ---- 4x4-blocked Cholesky machinery of the RSR solve ------------------------------------------------
// Lambda (q x q, padded to qb = ceil(q/4) blocks with an identity tail) plus ONE extra block row holding the
// right-hand side b (row 0 of block row qb) are kept in shared memory as packed lower-triangular 4x4 blocks,
// block (I, K) at BLK * (I (I + 1) / 2 + K), row-major inside, BLK = 17 doubles (odd stride: threads working on
// consecutive blocks hit different banks).  Factorising the augmented matrix yields y = L^-1 b in the extra
// row for free.  Per block column J: (1) every thread that owns a panel block (I, J) factorises the 4x4
// diagonal block redundantly in registers and solves its own block against it -- no serial section, no extra
// barrier; (2) the trailing blocks (I, K), J < K <= I, are updated by A_IK -= P_I P_K' (64 FMAs per block).
// Two barriers per block column instead of three per column, and no per-element index decoding.
constexpr int BLK = 17;
__device__ __forceinline__ int bidx(int I, int K) { return (I * (I + 1) / 2 + K) * BLK; }

// lower Cholesky of a 4x4 block held in registers (a[r][c], c <= r used); returns false on a bad pivot
// (inv[j] = 1 / L_jj; one rsqrt per pivot instead of a sqrt and a divide on the critical path)
__device__ __forceinline__ bool chol4(double (&a)[4][4], double (&invd)[4]) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
        double d = a[j][j];
#pragma unroll
        for (int k = 0; k < j; k++) d -= a[j][k] * a[j][k];
        if (!(d > 0.0)) return false;
        const double inv = rsqrt(d), dj = d * inv;
        invd[j] = inv;
        a[j][j] = dj;
#pragma unroll
        for (int i = j + 1; i < 4; i++) {
            double t = a[i][j];
#pragma unroll
            for (int k = 0; k < j; k++) t -= a[i][k] * a[j][k];
            a[i][j] = t * inv;
        }
    }
    return true;
}

__global__ void __launch_bounds__(SOLVE_THREADS, 6) solve_logit_rsr_kernel(const View v, const SolveDebug dbg,
                                                                         double* rec_alpha, double* rec_beta,
                                                                         double* rec_tau, int rec_capacity) {
    extern __shared__ double sm_dyn[];
    const int c = blockIdx.x, tid = threadIdx.x, T = blockDim.x;  // 64 (q <= 128) or SOLVE_THREADS
    double* sm = v.solve_stride ? v.solve_scratch + (size_t)c * v.solve_stride : sm_dyn;
    const int p = v.p, q = v.q, r = v.r, pq = v.pq;
    const int qb = (q + 3) / 4, q4 = qb * 4;
    const int nblk = (qb + 1) * (qb + 2) / 2;  // lower-triangular blocks incl. the right-hand-side block row
    const size_t Cp = v.Cp;
    const HParams* hp = v.hp;
    double* Bs = sm;                          // [nblk][BLK]
    double* gx = Bs + (size_t)nblk * BLK;     // [q4][p]  K' Om X  (rows of the joint Gram below the XX block)
    double* gxx = gx + (size_t)q4 * p;        // [p][p]   X' Om X (symmetric, full)
    double* h = gxx + p * p;                  // [pq]     U'(z - 1/2)
    double* th = h + pq;                      // [pq + 4] current beta, eta (eta padded)
    double* xm = th + pq + 4;                 // [q4] backward-solve result: mean
    double* xe = xm + q4;                     // [q4] backward-solve result: L^-T eps
    double* nrm = xe + q4;                    // [q4 + p + r] normals (eta part padded to q4)
    double* red = nrm + q4 + 2 * MAX_SMALL;   // [8]
    double* Ld = red + 8;                     // [qb][16] factored diagonal blocks (kept apart: the panel threads
                                              //          read the unfactored block while its owner factorises it)
    double* Li = Ld + qb * 16;                // [qb][4] reciprocals of the diagonal of L
    uint8_t* pair_i = reinterpret_cast<uint8_t*>(Li + qb * 4);
    uint8_t* pair_k = pair_i + nblk;
    __shared__ double small_S[MAX_SMALL * MAX_SMALL];
    __shared__ double small_b[MAX_SMALL], small_x[MAX_SMALL], small_t[MAX_SMALL];
    __shared__ int ok_flag;

    const uint32_t iter = v.ctl->iter;
    const int rec_index = v.ctl->rec_index;
    const uint32_t gc = (uint32_t)(v.chain_offset + c);
    const int cr = c % v.C;  // chain whose injected values a padded chain replicates

    // logit: per-chain omega-weighted Gram; probit: the shared unit-weight Gram (A_eta = K'K + tau Q, probit.py:226)
    const double* Grow = v.model == OCCU_PROBIT_RSR ? v.G0 : v.G + (size_t)c * v.M_pad;

    for (int a = tid; a < pq; a += T) {
        double hs = 0.0;  // K-split partials, fixed order
        for (int s = 0; s < v.ksplit; s++) hs += v.Hpart[((size_t)s * Cp + c) * v.ldh + a];
        h[a] = hs;
        th[a] = v.theta[(size_t)a * Cp + c];
    }
    // normals: eta (q), beta (p), alpha (r); stored as eta[0..q4) (zero padded), then beta, alpha
    for (int k = tid; k < q4 + p + r; k += T) {
        const int src = k < q4 ? k : k - q4 + q;  // index in the (q + p + r) stream
        double val = 0.0;
        if (k >= q4 || k < q) {
            if (dbg.normals) {
                val = dbg.normals[(size_t)cr * (pq + r) + src];
            } else {
                Philox g;
                g.init(v.seed, gc, (uint32_t)src, iter, ST_NORMALS);
                val = g.normal_d();
            }
        }
        nrm[k] = val;
    }
    {   // block-pair table: t -> (ii, kk), kk <= ii, row-major over the lower triangle
        int t = 0;
        for (int ii = 0; ii <= qb; ii++) {
            for (int kk = tid; kk <= ii; kk += T) {
                pair_i[t + kk] = (uint8_t)ii;
                pair_k[t + kk] = (uint8_t)kk;
            }
            t += ii + 1;
        }
    }
    if (tid == 0) ok_flag = 1;
    __syncthreads();

    // ---- tau | eta   (logit.py:206-209) ----
    {
        double part = 0.0;
#pragma unroll 8
        for (int idx = tid; idx < q * q; idx += T) {
            const int i = idx / q, j = idx - i * q;
            part += th[p + i] * v.Qrsr[idx] * th[p + j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if ((tid & 31) == 0) red[tid >> 5] = part;
        __syncthreads();
        if (tid == 0) {
            double quad = 0.0;
            for (int w = 0; w < T / 32; w++) quad += red[w];
            const double rate = 0.5 * quad + hp->tau_rate;
            const double tau = dbg.tau_in ? dbg.tau_in[cr] : v.gam[c] / rate;
            red[T / 32] = tau;
            if (dbg.enabled && dbg.out.tau_rate && c < v.C) dbg.out.tau_rate[c] = rate;
        }
        __syncthreads();
    }
    const double tau = red[T / 32];

    // ---- stage the joint Gram: XX block, KX rows, and Lambda = K' Om K + tau Qrsr in blocked form ----
    for (int idx = tid; idx < nblk * BLK; idx += T) Bs[idx] = 0.0;
    for (int idx = tid; idx < p * p; idx += T) {
        const int a = idx / p, b = idx - a * p;
        gxx[idx] = (b <= a) ? Grow[tri(a, b)] : Grow[tri(b, a)];
    }
    for (int idx = tid; idx < q4 * p; idx += T) {
        const int i = idx / p, a = idx - i * p;
        gx[idx] = i < q ? Grow[tri(p + i, a)] : 0.0;
    }
    __syncthreads();
    {   // Lambda = K' Om K + tau Qrsr.  One warp per row (contiguous in the packed Gram), four rows in flight per
        // warp: the loads of a batch are issued before any is consumed (a plain row-by-row loop serialises q
        // global-load latencies; decoding (i, j) from a flat triangle index costs ~30 instructions per element)
        const int warp = tid >> 5, lane = tid & 31, nwarp = T >> 5;
        for (int i0 = warp; i0 < q; i0 += 4 * nwarp) {
            double gv[4][2], qv[4][2];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + u * nwarp;
#pragma unroll
                for (int w = 0; w < 2; w++) {
                    const int j = lane + 32 * w;
                    const bool on = i < q && j <= i;
                    gv[u][w] = on ? Grow[tri(p + i, p) + j] : 0.0;
                    qv[u][w] = on ? v.Qrsr[(size_t)i * q + j] : 0.0;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + u * nwarp;
#pragma unroll
                for (int w = 0; w < 2; w++) {
                    const int j = lane + 32 * w;
                    if (i < q && j <= i) Bs[bidx(i >> 2, j >> 2) + (i & 3) * 4 + (j & 3)] = gv[u][w] + tau * qv[u][w];
                }
            }
            for (int i = i0; i < q && i < i0 + 4 * nwarp; i += nwarp)   // columns beyond 64 (q > 64)
                for (int j = 64 + lane; j <= i; j += 32)
                    Bs[bidx(i >> 2, j >> 2) + (i & 3) * 4 + (j & 3)] = Grow[tri(p + i, p) + j] + tau * v.Qrsr[(size_t)i * q + j];
        }
    }
    for (int i = q + tid; i < q4; i += T) Bs[bidx(i >> 2, i >> 2) + (i & 3) * 5] = 1.0;  // identity padding
    __syncthreads();
    // right-hand side b = K'k - (K' Om X) beta  (logit.py:481) into row 0 of block row qb
    for (int i = tid; i < q; i += T) {
        double b = h[p + i];
        for (int a = 0; a < p; a++) b -= gx[i * p + a] * th[a];
        Bs[bidx(qb, i >> 2) + (i & 3)] = b;
        if (dbg.enabled && dbg.out.eta_b && c < v.C) dbg.out.eta_b[(size_t)c * q + i] = b;
    }
    __syncthreads();
    if (dbg.enabled && dbg.out.eta_prec && c < v.C)
        for (int idx = tid; idx < q * q; idx += T) {
            const int i = idx / q, j = idx - i * q;
            const int ii = i >= j ? i : j, jj = i >= j ? j : i;
            dbg.out.eta_prec[(size_t)c * q * q + idx] = Bs[bidx(ii >> 2, jj >> 2) + (ii & 3) * 4 + (jj & 3)];
        }

    // ---- blocked right-looking Cholesky of the augmented matrix ----
    for (int J = 0; J < qb; J++) {
        // (1) panel: blocks (I, J), I = J .. qb
        for (int I = J + tid; I <= qb; I += T) {
            double d[4][4];
            const double* dj = Bs + bidx(J, J);
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) d[a][b] = dj[a * 4 + b];
            double invd[4];
            if (!chol4(d, invd)) {
                ok_flag = 0;
                continue;
            }
            double* blk = Bs + bidx(I, J);
            if (I == J) {
#pragma unroll
                for (int a = 0; a < 4; a++) {
#pragma unroll
                    for (int b = 0; b < 4; b++) Ld[J * 16 + a * 4 + b] = b <= a ? d[a][b] : 0.0;
                    Li[J * 4 + a] = invd[a];
                }
            } else {  // P = A_IJ L_JJ^-T, row by row
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    double x[4];
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        double t = blk[a * 4 + b];
#pragma unroll
                        for (int k = 0; k < b; k++) t -= x[k] * d[b][k];
                        x[b] = t * invd[b];
                    }
#pragma unroll
                    for (int b = 0; b < 4; b++) blk[a * 4 + b] = x[b];
                }
            }
        }
        __syncthreads();
        if (!ok_flag) break;
        // (2) trailing update: blocks (J+1+ii, J+1+kk), kk <= ii <= qb-J-1
        const int m1 = qb - J;  // block rows below J, including the right-hand-side row
        const int cnt = m1 * (m1 + 1) / 2;
        for (int t = tid; t < cnt; t += T) {
            const int I = J + 1 + pair_i[t], K = J + 1 + pair_k[t];
            if (K > qb - 1 + (I == qb ? 0 : 1)) continue;  // (qb, qb) does not exist
            const double* PI = Bs + bidx(I, J);
            const double* PK = Bs + bidx(K, J);
            double* A = Bs + bidx(I, K);
            double pk[4][4];
#pragma unroll
            for (int e = 0; e < 16; e++) pk[e >> 2][e & 3] = PK[e];
#pragma unroll
            for (int a = 0; a < 4; a++) {
                const double p0 = PI[a * 4], p1 = PI[a * 4 + 1], p2 = PI[a * 4 + 2], p3 = PI[a * 4 + 3];
#pragma unroll
                for (int b = 0; b < 4; b++)
                    A[a * 4 + b] -= p0 * pk[b][0] + p1 * pk[b][1] + p2 * pk[b][2] + p3 * pk[b][3];
            }
        }
        __syncthreads();
    }
    if (!ok_flag) {
        if (tid == 0) {
            v.status[c] = 1;
            if (c < v.C) atomicAdd(&v.ctl->fails, 1);
        }
        return;  // chain is frozen at its previous state; the failure is reported to the host
    }
    if (dbg.enabled && dbg.out.eta_chol && c < v.C)
        for (int idx = tid; idx < q * q; idx += T) {
            const int i = idx / q, j = idx - i * q;
            const int I = i >> 2, K = j >> 2, e = (i & 3) * 4 + (j & 3);
            dbg.out.eta_chol[(size_t)c * q * q + idx] = (j <= i) ? (I == K ? Ld[I * 16 + e] : Bs[bidx(I, K) + e]) : 0.0;
        }

    // ---- backward, right-looking over block rows: L' m = y (mean) and L' e = eps ----
    // thread J owns the two 4-vectors of block row J
    double rm[4] = {0, 0, 0, 0}, re[4] = {0, 0, 0, 0};
    if (tid < qb) {
#pragma unroll
        for (int b = 0; b < 4; b++) {
            rm[b] = Bs[bidx(qb, tid) + b];  // y = L^-1 b (row 0 of the extra block row)
            re[b] = nrm[tid * 4 + b];
        }
    }
    for (int I = qb - 1; I >= 0; I--) {
        if (tid == I) {  // x_I = L_II^-T rhs_I
            const double* L = Ld + I * 16;
#pragma unroll
            for (int a = 3; a >= 0; a--) {
                double tm = rm[a], te = re[a];
#pragma unroll
                for (int k = a + 1; k < 4; k++) {
                    tm -= L[k * 4 + a] * rm[k];
                    te -= L[k * 4 + a] * re[k];
                }
                rm[a] = tm * Li[I * 4 + a];
                re[a] = te * Li[I * 4 + a];
            }
#pragma unroll
            for (int a = 0; a < 4; a++) {
                xm[I * 4 + a] = rm[a];
                xe[I * 4 + a] = re[a];
            }
        }
        __syncthreads();
        if (tid < I) {  // rhs_J -= L_IJ' x_I
            const double* L = Bs + bidx(I, tid);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const double m = xm[I * 4 + k], e = xe[I * 4 + k];
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    rm[b] -= L[k * 4 + b] * m;
                    re[b] -= L[k * 4 + b] * e;
                }
            }
        }
    }
    __syncthreads();
    if (dbg.enabled && dbg.out.eta_mean && c < v.C)
        for (int i = tid; i < q; i += T) dbg.out.eta_mean[(size_t)c * q + i] = xm[i];
    for (int i = tid; i < q; i += T) th[p + i] = xm[i] + xe[i];  // eta = mean + L^-T eps
    __syncthreads();

    // ---- beta | .   A = X' Om X + b_prec,  b = X'k - (X' Om K) eta + b_prec b_mu   (logit.py:226-232) ----
    for (int idx = tid; idx < p * p; idx += T) small_S[idx] = gxx[idx] + hp->b_prec[idx];
    for (int a = tid; a < p; a += T) {
        double b = h[a] + hp->b_pm[a];
        for (int j = 0; j < q; j++) b -= gx[j * p + a] * th[p + j];
        small_b[a] = b;
    }
    __syncthreads();
    if (dbg.enabled && c < v.C) {
        if (dbg.out.beta_prec)
            for (int idx = tid; idx < p * p; idx += T) dbg.out.beta_prec[(size_t)c * p * p + idx] = small_S[idx];
        if (dbg.out.beta_b)
            for (int a = tid; a < p; a += T) dbg.out.beta_b[(size_t)c * p + a] = small_b[a];
    }
    __syncthreads();
    if (tid == 0) {
        if (small_precision_mvnorm(small_S, p, small_b, nrm + q4, small_x, small_t)) {
            for (int a = 0; a < p; a++) th[a] = small_x[a];
        } else {
            ok_flag = 0;
        }
    }
    __syncthreads();

    if (!ok_flag) {
        if (tid == 0) {
            v.status[c] = 1;
            if (c < v.C) atomicAdd(&v.ctl->fails, 1);
        }
        return;
    }

    // ---- commit state, record the draw (Chain.append, chain.py:56-74) ----
    for (int a = tid; a < pq; a += T) v.theta[(size_t)a * Cp + c] = th[a];
    if (tid == 0) v.tau[c] = tau;
    if (rec_beta && rec_index >= 0 && rec_index < rec_capacity && c < v.C) {
        const size_t row = (size_t)c * rec_capacity + rec_index;
        for (int a = tid; a < p; a += T) rec_beta[row * p + a] = th[a];
        if (tid == 0) rec_tau[row] = tau;
    }
}

// ------------------------------------------------------------------------------------------------
// large-q RSR solve: when Lambda no longer fits in shared memory (q > ~200) it is kept in
// per-chain global scratch as 32 x 32 tiles and factorised by a left-looking tiled Cholesky: for block column k
//   (A) every warp takes tiles (I, k), I >= k, and applies all updates  A_Ik -= sum_{j<k} L_Ij L_kj'  on the FP64
//       tensor pipe (DMMA m8n8k4, accumulators in registers, operands straight from L1/L2),
//   (B) warp 0 factorises the diagonal tile in shared memory,
//   (C) the rows of the tiles below are solved against it (one thread per row, row in registers),
//   (D) the finished panel is written back -- every tile is written once.
// The right-hand side b rides along as an extra tile row, so y = L^-1 b comes out of the factorisation; the two
// backward solves (mean and noise) are tiled the same way.  Everything around the factorisation (tau, beta, alpha,
// commit, record) is the code of solve_logit_rsr_kernel.  Reference: logit.py:325-337, 478-485 (dgemm + dgesv).
// ------------------------------------------------------------------------------------------------
constexpr int TB = 32;               // tile size
constexpr int TP = TB + 1;           // padded row length of a tile in shared memory
#ifndef TILED_THREADS_N
#define TILED_THREADS_N 256
#endif
constexpr int TILED_THREADS = TILED_THREADS_N;   // 8 warps
constexpr int TILED_WARPS = TILED_THREADS / 32;
constexpr int PANEL_SLOTS = 1 + TILED_THREADS / 32;   // diagonal tile + one tile per warp
__device__ __forceinline__ size_t tile_off(int I, int J) { return ((size_t)I * (I + 1) / 2 + J) * (TB * TB); }  // J <= I

static size_t tiled_scratch_doubles(int q, int p) {
    const size_t nbk = (q + TB - 1) / TB;
    return (nbk * (nbk + 1) / 2 + nbk) * TB * TB + nbk * TB * (size_t)p + 64;
}
// shared memory: panel slots, then h[pq], th[pq + 4], xm[q32], xe[q32], nrm[q32 + 2 MAX_SMALL]
static size_t tiled_smem_bytes(int q, int pq) {
    const size_t q32 = (size_t)(q + TB - 1) / TB * TB;
    return ((size_t)PANEL_SLOTS * TB * TP + 2 * (size_t)pq + 4 + 3 * q32 + 2 * MAX_SMALL + 8) * sizeof(double);
}

__global__ void __launch_bounds__(TILED_THREADS, 512 / TILED_THREADS) solve_logit_rsr_tiled_kernel(const View v, const SolveDebug dbg,
                                                                               double* rec_beta, double* rec_tau,
                                                                               int rec_capacity) {
    extern __shared__ double panel[];                 // [PANEL_SLOTS][TB][TP], then the per-chain vectors
    __shared__ double partial[TILED_THREADS / 32][2][TB];
    __shared__ double invd[TB], red[TILED_THREADS / 32 + 1];
    __shared__ double small_S[MAX_SMALL * MAX_SMALL];
    __shared__ double small_b[MAX_SMALL], small_x[MAX_SMALL], small_t[MAX_SMALL];
    __shared__ int ok_flag;
    const int c = blockIdx.x, tid = threadIdx.x, T = TILED_THREADS;
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int p = v.p, q = v.q, r = v.r, pq = v.pq;
    const int nbk = (q + TB - 1) / TB, q32 = nbk * TB;
    const size_t Cp = v.Cp;
    const HParams* hp = v.hp;
    double* h = panel + (size_t)PANEL_SLOTS * TB * TP;   // [pq]      U'(z - 1/2)
    double* th = h + pq;                                 // [pq + 4]  beta, eta
    double* xm = th + pq + 4;                            // [q32]     backward solve: mean
    double* xe = xm + q32;                               // [q32]     backward solve: L^-T eps
    double* nrm = xe + q32;                              // [q32 + 2 MAX_SMALL] normals
    double* tiles = v.solve_scratch + (size_t)c * v.solve_stride;
    double* gx = tiles + ((size_t)nbk * (nbk + 1) / 2 + nbk) * TB * TB;   // [q32][p]  K' Om X
    const uint32_t iter = v.ctl->iter;
    const int rec_index = v.ctl->rec_index;
    const uint32_t gc = (uint32_t)(v.chain_offset + c);
    const int cr = c % v.C;
    const double* Grow = v.model == OCCU_PROBIT_RSR ? v.G0 : v.G + (size_t)c * v.M_pad;

    for (int a = tid; a < pq; a += T) {
        double hs = 0.0;
        for (int s = 0; s < v.ksplit; s++) hs += v.Hpart[((size_t)s * Cp + c) * v.ldh + a];
        h[a] = hs;
        th[a] = v.theta[(size_t)a * Cp + c];
    }
    for (int k = tid; k < q32 + p + r; k += T) {   // normals: eta (zero padded to q32), then beta, alpha
        const int src = k < q32 ? k : k - q32 + q;
        double val = 0.0;
        if (k >= q32 || k < q) {
            if (dbg.normals) {
                val = dbg.normals[(size_t)cr * (pq + r) + src];
            } else {
                Philox gph;
                gph.init(v.seed, gc, (uint32_t)src, iter, ST_NORMALS);
                val = gph.normal_d();
            }
        }
        nrm[k] = val;
    }
    if (tid == 0) ok_flag = 1;
    __syncthreads();

    // ---- tau | eta   (logit.py:206-209) ----
    {
        double part = 0.0;
#pragma unroll 8
        for (int idx = tid; idx < q * q; idx += T) {
            const int i = idx / q, j = idx - i * q;
            part += th[p + i] * v.Qrsr[idx] * th[p + j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) red[warp] = part;
        __syncthreads();
        if (tid == 0) {
            double quad = 0.0;
            for (int w = 0; w < T / 32; w++) quad += red[w];
            const double rate = 0.5 * quad + hp->tau_rate;
            red[T / 32] = dbg.tau_in ? dbg.tau_in[cr] : v.gam[c] / rate;
            if (dbg.enabled && dbg.out.tau_rate && c < v.C) dbg.out.tau_rate[c] = rate;
        }
        __syncthreads();
    }
    const double tau = red[T / 32];

    // ---- K' Om X rows, then Lambda = K' Om K + tau Qrsr and b = K'k - (K' Om X) beta as tiles ----
    for (int idx = tid; idx < q32 * p; idx += T) {
        const int i = idx / p, a = idx - i * p;
        gx[idx] = i < q ? Grow[tri(p + i, a)] : 0.0;
    }
    __syncthreads();
    for (int I = 0; I < nbk; I++) {   // Lambda tiles: four elements per thread in flight (read-only loads first)
        const int count = (I + 1) * TB * TB;
        for (int e0 = tid; e0 < count; e0 += 4 * T) {
            double gv[4], qv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int e = e0 + u * T;
                gv[u] = 0.0;
                qv[u] = 0.0;
                if (e < count) {
                    const int J = e / (TB * TB), rc = e - J * TB * TB, rr = rc / TB, cc = rc - rr * TB;
                    const int i = I * TB + rr, j = J * TB + cc;
                    if (i < q && j < q) {
                        const int hi = i >= j ? i : j, lo = i >= j ? j : i;
                        gv[u] = __ldg(Grow + tri(p + hi, p + lo));
                        qv[u] = __ldg(v.Qrsr + (size_t)hi * q + lo);
                    } else {
                        gv[u] = i == j ? 1.0 : 0.0;   // identity padding
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int e = e0 + u * T;
                if (e < count) {
                    const int J = e / (TB * TB), rc = e - J * TB * TB;
                    tiles[tile_off(I, J) + rc] = gv[u] + tau * qv[u];
                }
            }
        }
    }
    for (int e = tid; e < nbk * TB * TB; e += T) {   // right-hand side row (logit.py:481): row 0 of tile row nbk
        const int J = e / (TB * TB), rc = e - J * TB * TB, rr = rc / TB, cc = rc - rr * TB;
        const int j = J * TB + cc;
        double val = 0.0;
        if (rr == 0 && j < q) {
            double b = h[p + j];
            for (int a = 0; a < p; a++) b -= gx[j * p + a] * th[a];
            val = b;
            if (dbg.enabled && dbg.out.eta_b && c < v.C) dbg.out.eta_b[(size_t)c * q + j] = b;
        }
        tiles[tile_off(nbk, J) + rc] = val;
    }
    __syncthreads();
    if (dbg.enabled && dbg.out.eta_prec && c < v.C)
        for (int idx = tid; idx < q * q; idx += T) {
            const int i = idx / q, j = idx - i * q;
            const int hi = i >= j ? i : j, lo = i >= j ? j : i;
            dbg.out.eta_prec[(size_t)c * q * q + idx] = tiles[tile_off(hi / TB, lo / TB) + (hi % TB) * TB + (lo % TB)];
        }

    // ---- left-looking tiled Cholesky of the augmented matrix ----
    for (int k = 0; k < nbk && ok_flag; k++) {
        // tiles of block column k are processed in groups: the first group holds the diagonal tile (warp 0, slot 0)
        // and the next TILED_WARPS - 1 tiles below it, every later group TILED_WARPS tiles; slots 1.. hold the group's
        // off-diagonal tiles
        const int nbelow = nbk - k;                       // tiles below the diagonal, incl. the right-hand-side row
        bool first = true;
        for (int g0 = 0; first || g0 < nbelow; g0 = first ? TILED_WARPS - 1 : g0 + TILED_WARPS, first = false) {
            // tile handled by this warp: first group -> warp 0: diagonal, warps 1..7: below[0..6];
            // later groups -> below[g0 + warp]
            const int below_idx = first ? warp - 1 : g0 + warp;
            const int gcount = first ? min(TILED_WARPS - 1, nbelow) : min(TILED_WARPS, nbelow - g0);   // off-diagonal tiles
            const bool has_tile = first ? (warp == 0 || below_idx < nbelow) : (below_idx < nbelow);
            const int I = (first && warp == 0) ? k : k + 1 + below_idx;
            const int slot = (first && warp == 0) ? 0 : (first ? warp : 1 + warp);
            // (A) A_Ik -= sum_{j<k} L_Ij L_kj'
            if (has_tile) {
                double acc[4][4][2];
                const double* src = tiles + tile_off(I, k);
#pragma unroll
                for (int mb = 0; mb < 4; mb++)
#pragma unroll
                    for (int nb = 0; nb < 4; nb++) {
                        const double2 v2 = *reinterpret_cast<const double2*>(src + (mb * 8 + g) * TB + nb * 8 + 2 * t4);
                        acc[mb][nb][0] = v2.x;
                        acc[mb][nb][1] = v2.y;
                    }
                for (int j = 0; j < k; j++) {
                    const double* A = tiles + tile_off(I, j);
                    const double* B = tiles + tile_off(k, j);
#pragma unroll 2
                    for (int ks = 0; ks < TB / 4; ks++) {
                        double a[4], b[4];
#pragma unroll
                        for (int mb = 0; mb < 4; mb++) a[mb] = -A[(mb * 8 + g) * TB + ks * 4 + t4];
#pragma unroll
                        for (int nb = 0; nb < 4; nb++) b[nb] = B[(nb * 8 + g) * TB + ks * 4 + t4];
#pragma unroll
                        for (int mb = 0; mb < 4; mb++)
#pragma unroll
                            for (int nb = 0; nb < 4; nb++) dgemm::dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
                    }
                }
                double* dst = panel + (size_t)slot * TB * TP;
#pragma unroll
                for (int mb = 0; mb < 4; mb++)
#pragma unroll
                    for (int nb = 0; nb < 4; nb++) {
                        dst[(mb * 8 + g) * TP + nb * 8 + 2 * t4] = acc[mb][nb][0];
                        dst[(mb * 8 + g) * TP + nb * 8 + 2 * t4 + 1] = acc[mb][nb][1];
                    }
            }
            __syncthreads();
            if (first) {
                // (B) diagonal tile, whole CTA, one barrier per column: the trailing update of column j uses the
                //     UNSCALED column (a_ic -= a_ij a_cj / a_jj, j < c <= i); column j is scaled to
                //     l_ij = a_ij / sqrt(a_jj) during step j + 1, when nobody reads it any more
                double* D = panel;
                for (int j = 0; j <= TB; j++) {
                    double piv = 1.0;
                    if (j < TB) {
                        piv = D[j * TP + j];
                        if (!(piv > 0.0)) {   // uniform: every thread reads the same pivot
                            if (tid == 0) ok_flag = 0;
                            break;
                        }
                        if (lane > j) {   // lane = row, warp = column group (columns j + 1 + warp, + 8, ... <= row)
                            const double lij = D[lane * TP + j] / piv;
                            for (int cc = j + 1 + warp; cc <= lane; cc += T / 32) D[lane * TP + cc] -= lij * D[cc * TP + j];
                        }
                    }
                    if (j > 0 && tid < TB && tid >= j - 1) D[tid * TP + j - 1] *= invd[j - 1];   // column j - 1
                    if (j < TB && tid == 0) invd[j] = rsqrt(piv);
                    __syncthreads();
                }
                __syncthreads();
                if (!ok_flag) break;
            }
            // (C) rows of the group's off-diagonal tiles: x L_kk' = a, one thread per row
            if (tid < gcount * TB) {
                double* R = panel + (size_t)(1 + tid / TB) * TB * TP + (tid % TB) * TP;   // row stride TP: conflict-free
                for (int cc = 0; cc < TB; cc++) {
                    double s0 = R[cc], s1 = 0.0;          // two accumulators: the dependent chain is cc / 2 long
                    int m = 0;
                    for (; m + 1 < cc; m += 2) {
                        s0 -= R[m] * panel[cc * TP + m];
                        s1 -= R[m + 1] * panel[cc * TP + m + 1];
                    }
                    if (m < cc) s0 -= R[m] * panel[cc * TP + m];
                    R[cc] = (s0 + s1) * invd[cc];
                }
            }
            __syncthreads();
            // (D) write the group back (the strictly upper part of the diagonal tile is never read)
            {
                const int s0 = first ? 0 : 1, ns = gcount + (first ? 1 : 0);
                for (int e = tid; e < ns * TB * TB; e += T) {
                    const int s = s0 + e / (TB * TB), rc = e % (TB * TB);
                    const int It = s == 0 ? k : k + 1 + (first ? 0 : g0) + (s - 1);
                    tiles[tile_off(It, k) + rc] = panel[(size_t)s * TB * TP + (rc / TB) * TP + (rc % TB)];
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();
    if (!ok_flag) {
        if (tid == 0) {
            v.status[c] = 1;
            if (c < v.C) atomicAdd(&v.ctl->fails, 1);
        }
        return;
    }
    if (dbg.enabled && dbg.out.eta_chol && c < v.C)
        for (int idx = tid; idx < q * q; idx += T) {
            const int i = idx / q, j = idx - i * q;
            dbg.out.eta_chol[(size_t)c * q * q + idx] = j <= i ? tiles[tile_off(i / TB, j / TB) + (i % TB) * TB + (j % TB)] : 0.0;
        }

    // ---- backward: L' m = y (mean) and L' e = eps, block rows from the bottom ----
    for (int i = tid; i < q32; i += T) {
        xm[i] = tiles[tile_off(nbk, i / TB) + (i % TB)];   // y = L^-1 b: row 0 of the extra tile row
        xe[i] = nrm[i];
    }
    __syncthreads();
    for (int k = nbk - 1; k >= 0; k--) {
        // rhs_k -= sum_{I > k} L_Ik' x_I : warps split the tiles below, lane = column
        double sm_ = 0.0, se_ = 0.0;
        for (int I = k + 1 + warp; I < nbk; I += T / 32) {
            const double* L = tiles + tile_off(I, k);
#pragma unroll 8
            for (int rr = 0; rr < TB; rr++) {
                const double l = L[rr * TB + lane];
                sm_ += l * xm[I * TB + rr];
                se_ += l * xe[I * TB + rr];
            }
        }
        partial[warp][0][lane] = sm_;
        partial[warp][1][lane] = se_;
        // diagonal tile into shared memory for the triangular solve
        for (int e = tid; e < TB * TB; e += T) panel[(e / TB) * TP + (e % TB)] = tiles[tile_off(k, k) + e];
        __syncthreads();
        if (warp < 2) {   // warp 0: mean, warp 1: noise
            double* x = warp == 0 ? xm : xe;
            double rhs = x[k * TB + lane];
            for (int w = 0; w < T / 32; w++) rhs -= partial[w][warp][lane];
            // L_kk' x = rhs, backward: x_c = (rhs_c - sum_{m > c} L[m][c] x_m) / L[c][c]
            for (int cc = TB - 1; cc >= 0; cc--) {
                const double xc = __shfl_sync(0xffffffffu, rhs, cc) / panel[cc * TP + cc];
                if (lane == cc) rhs = xc;
                if (lane < cc) rhs -= panel[cc * TP + lane] * xc;
            }
            x[k * TB + lane] = rhs;
        }
        __syncthreads();
    }
    if (dbg.enabled && dbg.out.eta_mean && c < v.C)
        for (int i = tid; i < q; i += T) dbg.out.eta_mean[(size_t)c * q + i] = xm[i];
    for (int i = tid; i < q; i += T) th[p + i] = xm[i] + xe[i];  // eta = mean + L^-T eps
    __syncthreads();

    // ---- beta | .   A = X' Om X + b_prec,  b = X'k - (X' Om K) eta + b_prec b_mu   (logit.py:226-232) ----
    for (int idx = tid; idx < p * p; idx += T) {
        const int a = idx / p, b = idx - a * p;
        small_S[idx] = ((b <= a) ? Grow[tri(a, b)] : Grow[tri(b, a)]) + hp->b_prec[idx];
    }
    for (int a = warp; a < p; a += T / 32) {   // one warp per row of the system: q-long dot product
        double b = 0.0;
        for (int j = lane; j < q; j += 32) b -= gx[j * p + a] * th[p + j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
        if (lane == 0) small_b[a] = b + h[a] + hp->b_pm[a];
    }
    __syncthreads();
    if (dbg.enabled && c < v.C) {
        if (dbg.out.beta_prec)
            for (int idx = tid; idx < p * p; idx += T) dbg.out.beta_prec[(size_t)c * p * p + idx] = small_S[idx];
        if (dbg.out.beta_b)
            for (int a = tid; a < p; a += T) dbg.out.beta_b[(size_t)c * p + a] = small_b[a];
    }
    __syncthreads();
    if (tid == 0) {
        if (small_precision_mvnorm(small_S, p, small_b, nrm + q32, small_x, small_t)) {
            for (int a = 0; a < p; a++) th[a] = small_x[a];
        } else {
            ok_flag = 0;
        }
    }
    __syncthreads();
    if (!ok_flag) {
        if (tid == 0) {
            v.status[c] = 1;
            if (c < v.C) atomicAdd(&v.ctl->fails, 1);
        }
        return;
    }
    // ---- commit state, record the draw (Chain.append, chain.py:56-74) ----
    for (int a = tid; a < pq; a += T) v.theta[(size_t)a * Cp + c] = th[a];
    if (tid == 0) v.tau[c] = tau;
    if (rec_beta && rec_index >= 0 && rec_index < rec_capacity && c < v.C) {
        const size_t row = (size_t)c * rec_capacity + rec_index;
        for (int a = tid; a < p; a += T) rec_beta[row * p + a] = th[a];
        if (tid == 0) rec_tau[row] = tau;
    }
}

// ------------------------------------------------------------------------------------------------
// stand-alone primitives
// ------------------------------------------------------------------------------------------------
template <typename PGT>
__global__ void pg1_kernel(const double* __restrict__ c, size_t count, uint64_t seed, const PhiloxKeys rk,
                           double* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Philox g;
    g.init(seed, (uint32_t)(i >> 32), (uint32_t)i, 0u, ST_HOOK);
    out[i] = (double)pg1_draw<PGT>((PGT)c[i], g, rk);
}

__global__ void truncnorm_kernel(const double* __restrict__ loc, const uint8_t* __restrict__ positive,
                                 const double* __restrict__ u, size_t count, uint64_t seed, int fp64,
                                 double* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    if (u) {
        out[i] = truncnorm_icdf(loc[i], positive[i] != 0, u[i]);
    } else {
        Philox g;
        g.init(seed, (uint32_t)(i >> 32), (uint32_t)i, 0u, ST_HOOK);
        out[i] = fp64 ? truncnorm_draw(loc[i], positive[i] != 0, g) : truncnorm_draw_f32(loc[i], positive[i] != 0, g);
    }
}

__global__ void gamma_kernel(const double* __restrict__ shape, const double* __restrict__ rate, size_t count,
                             uint64_t seed, double* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Philox g;
    g.init(seed, (uint32_t)(i >> 32), (uint32_t)i, 0u, ST_HOOK);
    out[i] = gamma_draw(shape[i], g) / rate[i];
}

// one CTA per system; P (full row-major) is staged packed in shared memory
__global__ void __launch_bounds__(SOLVE_THREADS) precision_mvnorm_kernel(int d, const double* __restrict__ b,
                                                                          double* __restrict__ P,
                                                                          const double* __restrict__ z, uint64_t seed,
                                                                          double* __restrict__ x, int* fails) {
    extern __shared__ double sm[];
    const int s = blockIdx.x, tid = threadIdx.x, T = SOLVE_THREADS;
    double* A = sm;                   // packed lower
    double* rhs = A + d * (d + 1) / 2;  // [d]
    double* Ps = P + (size_t)s * d * d;
    for (int idx = tid; idx < d * d; idx += T) {
        const int i = idx / d, j = idx - i * d;
        if (j <= i) A[tri(i, j)] = Ps[idx];
    }
    __syncthreads();
    if (!block_cholesky_packed(A, 0, d, tid, T)) {
        if (tid == 0) atomicAdd(fails, 1);
        return;
    }
    for (int idx = tid; idx < d * d; idx += T) {
        const int i = idx / d, j = idx - i * d;
        if (j <= i) Ps[idx] = A[tri(i, j)];  // reference overwrites prec with its factor (distributions.pyx:100)
    }
    for (int i = tid; i < d; i += T) rhs[i] = b[(size_t)s * d + i];
    for (int j = 0; j < d; j++) {
        __syncthreads();
        const double yj = rhs[j] / A[tri(j, j)];
        __syncthreads();
        if (tid == 0) rhs[j] = yj;
        for (int i = j + 1 + tid; i < d; i += T) rhs[i] -= A[tri(i, j)] * yj;
    }
    __syncthreads();
    for (int i = tid; i < d; i += T) {
        double zn;
        if (z) {
            zn = z[(size_t)s * d + i];
        } else {
            Philox g;
            g.init(seed, (uint32_t)s, (uint32_t)i, 0u, ST_HOOK);
            zn = g.normal_d();
        }
        rhs[i] += zn;
    }
    for (int j = d - 1; j >= 0; j--) {
        __syncthreads();
        const double xj = rhs[j] / A[tri(j, j)];
        __syncthreads();
        if (tid == 0) rhs[j] = xj;
        for (int i = tid; i < j; i += T) rhs[i] -= A[tri(j, i)] * xj;
    }
    __syncthreads();
    for (int i = tid; i < d; i += T) x[(size_t)s * d + i] = rhs[i];
}

__global__ void sums_to_zero_kernel(int d, const double* __restrict__ x, const double* __restrict__ z,
                                    double* __restrict__ out) {
    // one warp per row, fixed-order tree reduction
    const int row = blockIdx.x, lane = threadIdx.x;
    const double* xr = x + (size_t)row * d;
    const double* zr = z + (size_t)row * d;
    double sx = 0.0, sz = 0.0;
    for (int i = lane; i < d; i += 32) {
        sx += xr[i];
        sz += zr[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sx += __shfl_xor_sync(0xffffffffu, sx, o);
        sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    const double a = -sx / sz;
    for (int i = lane; i < d; i += 32) out[(size_t)row * d + i] = xr[i] + a * zr[i];
}

// per (chain, component): whole-chain and split-half moments + autocovariances
__global__ void chain_summaries_kernel(const double* __restrict__ rec, int chains, int draws, int dim, int max_lag,
                                       double* __restrict__ out) {
    const int cd = blockIdx.x;  // chain * dim + component
    const int c = cd / dim, k = cd - c * dim;
    const double* x = rec + (size_t)c * draws * dim + k;
    const int W = 6 + max_lag + 1;
    double* o = out + (size_t)cd * W;
    __shared__ double s_mean;
    __shared__ double red[32];
    const int tid = threadIdx.x, T = blockDim.x;
    auto block_sum = [&](double val) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) val += __shfl_xor_sync(0xffffffffu, val, off);
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = val;
        __syncthreads();
        double tot = 0.0;
        for (int w = 0; w < T / 32; w++) tot += red[w];
        return tot;
    };
    const int half = draws / 2;
    // ranges: whole, first half, second half
    for (int part = 0; part < 3; part++) {
        const int lo = part == 2 ? draws - half : 0;
        const int hi = part == 1 ? half : draws;
        const int cnt = hi - lo;
        double s = 0.0;
        for (int t = lo + tid; t < hi; t += T) s += x[(size_t)t * dim];
        const double mean = cnt > 0 ? block_sum(s) / cnt : 0.0;
        double m2 = 0.0;
        for (int t = lo + tid; t < hi; t += T) {
            const double dlt = x[(size_t)t * dim] - mean;
            m2 += dlt * dlt;
        }
        m2 = block_sum(m2);
        if (tid == 0) {
            o[2 * part] = mean;
            o[2 * part + 1] = m2;
            if (part == 0) s_mean = mean;
        }
        __syncthreads();
    }
    const double mean = s_mean;
    for (int lag = 0; lag <= max_lag; lag++) {
        double s = 0.0;
        for (int t = tid; t + lag < draws; t += T) s += (x[(size_t)t * dim] - mean) * (x[(size_t)(t + lag) * dim] - mean);
        s = block_sum(s);
        if (tid == 0) o[6 + lag] = s / draws;
    }
}

// ------------------------------------------------------------------------------------------------
// LogitICARGibbs ("PGGibbs"): eta | . ~ N(Lambda^-1 b, Lambda^-1) on {1'eta = 0}, Lambda = tau Q + diag(omega_b)
//
// The reference perturbs b through a dense eigen-factor of Q (2n - 1 normals) and solves two systems with
// MINRES at rtol 1e-5 (logit.py:73-99).  Here Lambda -- whose sparsity pattern is fixed by Q -- is factorised
// exactly per chain by a banded Cholesky over the fixed symbolic pattern (half bandwidth bw after the host's
// reordering), the draw is x = Lambda^-1 b + L^-T eps (n normals, same law), zv = Lambda^-1 1, and
// eta = x - (sum x / sum zv) zv (ensure_sums_to_zero, distributions.pyx:24-39).
//
// One CTA per chain.  Right-looking column Cholesky on a (bw+1) x (bw+1) ring-indexed window in shared
// memory (no data movement when the window slides), forward substitution of both right-hand sides fused
// into the factorisation, L streamed to HBM once in column-band form and read back once by the backward pass.
// ------------------------------------------------------------------------------------------------
constexpr int ICAR_THREADS = 128;

// psi = X beta + eta   (logit.py:201 with spatial = eta)
__global__ void psi_icar_kernel(const View v) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)v.n * v.Cp;
    if (idx >= total) return;
    const int i = (int)(idx / v.Cp), c = (int)(idx - (size_t)i * v.Cp);
    double s = v.eta_n[idx];
    for (int a = 0; a < v.p; a++) s += v.X[(size_t)i * v.p + a] * v.theta[(size_t)a * v.Cp + c];
    v.psi[idx] = s;
}

__device__ __forceinline__ double block_sum_128(double val, double* red, int tid) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = val;
    __syncthreads();
    return red[0] + red[1] + red[2] + red[3];
}

__global__ void __launch_bounds__(ICAR_THREADS) solve_logit_icar_kernel(const View v, const SolveDebug dbg,
                                                                         double* rec_alpha, double* rec_beta,
                                                                         double* rec_tau, int rec_capacity) {
    extern __shared__ double sm[];
    const int c = blockIdx.x, tid = threadIdx.x, T = ICAR_THREADS;
    const int n = v.n, p = v.p, r = v.r, w = v.bw, W1 = v.bw + 1;
    const size_t Cp = v.Cp;
    const HParams* hp = v.hp;
    double* Wn = sm;                     // [W1][W1] ring window, element (i, k) at Wn[(i % W1) * W1 + (k % W1)]
    double* col = Wn + W1 * W1;          // [W1] current column of L
    double* r1 = col + W1;               // [W1] ring: forward residual of b
    double* r2 = r1 + W1;                // [W1] ring: forward residual of 1
    double* xr = r2 + W1;                // [3][W1] rings of the backward solves
    double* nrm = xr + 3 * W1;           // [p + r] normals of beta, alpha
    double* red = nrm + 2 * MAX_SMALL;   // [8]
    uint8_t* pair_i = reinterpret_cast<uint8_t*>(red + 8);  // [w (w+1) / 2] trailing-update pair table
    uint8_t* pair_k = pair_i + (w * (w + 1)) / 2;
    __shared__ double small_S[MAX_SMALL * MAX_SMALL];
    __shared__ double small_b[MAX_SMALL], small_x[MAX_SMALL], small_t[MAX_SMALL], beta_s[MAX_SMALL];
    __shared__ double ybc[2];
    __shared__ int ok_flag;

    const uint32_t iter = v.ctl->iter;
    const int rec_index = v.ctl->rec_index;
    const uint32_t gc = (uint32_t)(v.chain_offset + c);
    const int cr = c % v.C;
    double* Lc = v.Lc + (size_t)c * n * W1;
    double* yb = v.scratch + (size_t)c * 8 * round_up_dev(n, 4);  // forward results / backward inputs
    double* y1 = yb + n;
    double* xs = y1 + n;
    double* zs = xs + n;
    double* ms = zs + n;

    for (int a = tid; a < p; a += T) beta_s[a] = v.theta[(size_t)a * Cp + c];
    for (int k = tid; k < p + r; k += T) {
        double val;
        if (dbg.normals) {
            val = dbg.normals[(size_t)cr * (n + p + r) + n + k];
        } else {
            Philox g;
            g.init(v.seed, gc, (uint32_t)(n + k), iter, ST_NORMALS);
            val = g.normal_d();
        }
        nrm[k] = val;
    }
    {
        int t = 0;  // pair table (ii, kk), kk <= ii < w, built redundantly but cheaply by thread 0..: parallel fill
        for (int ii = 0; ii < w; ii++) {
            for (int kk = tid; kk <= ii; kk += T) {
                pair_i[t + kk] = (uint8_t)ii;
                pair_k[t + kk] = (uint8_t)kk;
            }
            t += ii + 1;
        }
    }
    if (tid == 0) ok_flag = 1;
    __syncthreads();

    // ---- tau | eta: rate = eta' Q eta / 2 + tau_rate (logit.py:206-209), Q in lower-band form ----
    double part = 0.0;
    for (int i = tid; i < n; i += T) {
        const double ei = v.eta_n[(size_t)i * Cp + c];
        double s = 0.5 * v.Qband[(size_t)i * W1] * ei;
        const int dmax = min(w, i);
        for (int d = 1; d <= dmax; d++) s += v.Qband[(size_t)i * W1 + d] * v.eta_n[(size_t)(i - d) * Cp + c];
        part += ei * s;  // eta' Q eta = 2 * sum_i eta_i (Q_ii eta_i / 2 + sum_{k<i} Q_ik eta_k)
    }
    const double quad = 2.0 * block_sum_128(part, red, tid);
    __syncthreads();
    if (tid == 0) {
        const double rate = 0.5 * quad + hp->tau_rate;
        const double tau = dbg.tau_in ? dbg.tau_in[cr] : v.gam[c] / rate;
        red[4] = tau;
        if (dbg.enabled && dbg.out.tau_rate && c < v.C) dbg.out.tau_rate[c] = rate;
    }
    __syncthreads();
    const double tau = red[4];

    // Stage this chain's per-site inputs CONTIGUOUSLY (chain-major scratch) before the sequential loops: the
    // chain-minor state arrays would cost one uncoalesced DRAM round trip per row inside the dependency chain,
    // and the Box-Muller normals of the backward solve must not sit on its critical path either.
    double* om_c = ms + n;  // [n] omega of this chain ("spare" slot of the scratch block)
    for (int i = tid; i < n; i += T) {
        const double om = v.omega[(size_t)i * Cp + c];
        double xb = 0.0;
        for (int a = 0; a < p; a++) xb += v.X[(size_t)i * p + a] * beta_s[a];
        const double bi = v.kd[(size_t)i * Cp + c] - om * xb;  // b_i = k_i - omega_i x_i'beta   (logit.py:213)
        om_c[i] = om;
        xs[i] = bi;  // xs / zs are free until the backward pass: b and the normals live there meanwhile
        if (dbg.enabled && dbg.out.eta_b && c < v.C) dbg.out.eta_b[(size_t)c * n + i] = bi;
        double ez;
        if (dbg.normals) {
            ez = dbg.normals[(size_t)cr * (n + p + r) + i];
        } else {
            Philox g;
            g.init(v.seed, gc, (uint32_t)i, iter, ST_NORMALS);
            ez = g.normal_d();
        }
        zs[i] = ez;
    }
    __syncthreads();

    // row i of Lambda (band part) and of the right-hand sides enters ring slot i % W1
    auto load_row = [&](int i) {
        const int slot = i % W1;
        for (int d = tid; d < W1; d += T) {  // k = i - d
            double val = 0.0;
            if (i < n) {
                if (d <= i) val = tau * v.Qband[(size_t)i * W1 + d];
                if (d == 0) val += om_c[i];
            } else if (d == 0) {
                val = 1.0;  // identity padding past the last row
            }
            const int k = i - d;
            const int kslot = ((k % W1) + W1) % W1;
            Wn[slot * W1 + kslot] = val;
        }
        if (tid == 0) {
            // b_i = k_i - omega_i x_i'beta   (logit.py:213)
            double bi = 0.0, one = 0.0;
            if (i < n) {
                bi = xs[i];
                one = 1.0;
            }
            r1[slot] = bi;
            r2[slot] = one;
        }
    };
    for (int i = 0; i <= w; i++) load_row(i);
    __syncthreads();

    // ---- factorisation + fused forward substitution ----
    const int npairs = w * (w + 1) / 2;
    for (int j = 0; j < n; j++) {
        const int jm = j % W1;
        const double piv = Wn[jm * W1 + jm];
        if (!(piv > 0.0)) {
            if (tid == 0) {
                v.status[c] = 1;
                if (c < v.C) atomicAdd(&v.ctl->fails, 1);
            }
            return;
        }
        const double inv = rsqrt(piv), dj = piv * inv;  // one rsqrt on the critical path instead of sqrt + divide
        if (tid == 0) om_c[j] = inv;                    // om_c[j] is dead (row j entered long ago): keep 1 / L_jj there
        for (int e = tid; e < W1; e += T) {
            int slot = jm + e;
            if (slot >= W1) slot -= W1;
            const double le = (e == 0) ? dj : Wn[slot * W1 + jm] * inv;
            col[e] = le;
            Lc[(size_t)j * W1 + e] = le;
        }
        if (tid == T - 1) {
            const double a = r1[jm] * inv, b = r2[jm] * inv;
            ybc[0] = a;
            ybc[1] = b;
            yb[j] = a;
            y1[j] = b;
        }
        __syncthreads();
        const int j1 = (jm + 1 == W1) ? 0 : jm + 1;
        for (int t = tid; t < npairs; t += T) {
            int si = j1 + pair_i[t], sk = j1 + pair_k[t];
            if (si >= W1) si -= W1;
            if (sk >= W1) sk -= W1;
            Wn[si * W1 + sk] -= col[1 + pair_i[t]] * col[1 + pair_k[t]];
        }
        for (int e = 1 + tid; e <= w; e += T) {  // forward residuals of rows j+1 .. j+w
            int slot = jm + e;
            if (slot >= W1) slot -= W1;
            r1[slot] -= col[e] * ybc[0];
            r2[slot] -= col[e] * ybc[1];
        }
        load_row(j + W1);  // reuses slot jm: row / column j are no longer referenced
        __syncthreads();
    }
    if (dbg.enabled && dbg.out.eta_chol && c < v.C)  // column-band factor [n][bw + 1]
        for (size_t idx = tid; idx < (size_t)n * W1; idx += T) dbg.out.eta_chol[(size_t)c * n * W1 + idx] = Lc[idx];
    __threadfence_block();
    __syncthreads();

    // ---- backward: L' x = y_b + eps, L' zv = y_1, (debug) L' m = y_b; one warp, lanes over the band ----
    if (tid < 32) {
        for (int e = tid; e < 3 * W1; e += 32) xr[e] = 0.0;
        __syncwarp();
        constexpr int NL = 5;  // band entries per lane (bw <= 160)
        double lnext[NL], sc[4] = {0.0, 0.0, 0.0, 0.0};
        auto fetch = [&](int j) {  // loads of column j: independent of the recurrence, issued one step ahead
#pragma unroll
            for (int u = 0; u < NL; u++) {
                const int e = 1 + tid + 32 * u;
                lnext[u] = (e <= w && j + e < n) ? Lc[(size_t)j * W1 + e] : 0.0;
            }
            if (tid == 0) {
                sc[0] = yb[j];
                sc[1] = y1[j];
                sc[2] = zs[j];    // staged normal
                sc[3] = om_c[j];  // 1 / L_jj
            }
        };
        fetch(n - 1);
        for (int j = n - 1; j >= 0; j--) {
            const int jm = j % W1;
            double lcur[NL];
#pragma unroll
            for (int u = 0; u < NL; u++) lcur[u] = lnext[u];
            const double ybj = sc[0], y1j = sc[1], ez = sc[2], inv = sc[3];
            if (j > 0) fetch(j - 1);
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int u = 0; u < NL; u++) {
                const int e = 1 + tid + 32 * u;
                if (e <= w) {
                    int slot = jm + e;
                    if (slot >= W1) slot -= W1;
                    s0 += lcur[u] * xr[slot];
                    s1 += lcur[u] * xr[W1 + slot];
                    s2 += lcur[u] * xr[2 * W1 + slot];
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            if (tid == 0) {
                const double x = (ybj + ez - s0) * inv, z = (y1j - s1) * inv, m = (ybj - s2) * inv;
                xr[jm] = x;
                xr[W1 + jm] = z;
                xr[2 * W1 + jm] = m;
                xs[j] = x;
                zs[j] = z;
                ms[j] = m;
            }
            __syncwarp();
        }
    }
    __threadfence_block();
    __syncthreads();

    // ---- eta = x - (sum x / sum zv) zv (distributions.pyx:24-39); beta system (logit.py:226-231) ----
    double sx = 0.0, sz = 0.0;
    for (int i = tid; i < n; i += T) {
        sx += xs[i];
        sz += zs[i];
    }
    sx = block_sum_128(sx, red, tid);
    sz = block_sum_128(sz, red, tid);
    const double shift = -sx / sz;
    for (int i = tid; i < n; i += T) {
        const double e = xs[i] + shift * zs[i];
        xs[i] = e;
        v.eta_n[(size_t)i * Cp + c] = e;
        if (dbg.enabled && dbg.out.eta_mean && c < v.C) dbg.out.eta_mean[(size_t)c * n + i] = ms[i];
    }
    __syncthreads();
    for (int idx = 0; idx < p * p; idx++) {  // A = X' Om X + b_prec
        const int a = idx / p, b = idx - a * p;
        if (b > a) continue;
        double s = 0.0;
        for (int i = tid; i < n; i += T) s += v.omega[(size_t)i * Cp + c] * v.X[(size_t)i * p + a] * v.X[(size_t)i * p + b];
        s = block_sum_128(s, red, tid);
        if (tid == 0) {
            small_S[a * p + b] = s + hp->b_prec[a * p + b];
            small_S[b * p + a] = s + hp->b_prec[b * p + a];
        }
    }
    for (int a = 0; a < p; a++) {  // b = X'(k - omega * eta) + b_prec b_mu
        double s = 0.0;
        for (int i = tid; i < n; i += T)
            s += v.X[(size_t)i * p + a] * (v.kd[(size_t)i * Cp + c] - v.omega[(size_t)i * Cp + c] * xs[i]);
        s = block_sum_128(s, red, tid);
        if (tid == 0) small_b[a] = s + hp->b_pm[a];
    }
    __syncthreads();
    if (dbg.enabled && c < v.C) {
        if (dbg.out.beta_prec)
            for (int idx = tid; idx < p * p; idx += T) dbg.out.beta_prec[(size_t)c * p * p + idx] = small_S[idx];
        if (dbg.out.beta_b)
            for (int a = tid; a < p; a += T) dbg.out.beta_b[(size_t)c * p + a] = small_b[a];
    }
    __syncthreads();
    if (tid == 0) {
        if (small_precision_mvnorm(small_S, p, small_b, nrm, small_x, small_t)) {
            for (int a = 0; a < p; a++) beta_s[a] = small_x[a];
        } else {
            ok_flag = 0;
        }
    }
    __syncthreads();

    if (!ok_flag) {
        if (tid == 0) {
            v.status[c] = 1;
            if (c < v.C) atomicAdd(&v.ctl->fails, 1);
        }
        return;
    }
    for (int a = tid; a < p; a += T) v.theta[(size_t)a * Cp + c] = beta_s[a];
    if (tid == 0) v.tau[c] = tau;
    if (rec_beta && rec_index >= 0 && rec_index < rec_capacity && c < v.C) {
        const size_t row = (size_t)c * rec_capacity + rec_index;
        for (int a = tid; a < p; a += T) rec_beta[row * p + a] = beta_s[a];
        if (tid == 0) rec_tau[row] = tau;
    }
}

// ------------------------------------------------------------------------------------------------
// ICAR solve on the FP64 tensor pipe (half bandwidth bw <= 60): 4-column blocked banded Cholesky whose
// rank-4 trailing update A -= P P' is exactly the k = 4 of DMMA.8x8x4.
//
// The sliding window of the factorisation covers matrix indices [4s, 4s + W8), W8 = 8 NT >= bw + 4, and is
// held ENTIRELY IN REGISTERS as m8n8 accumulator fragments, addressed by ring slot (index mod W8) so that
// sliding the window moves no data: 4 slots leave, 4 enter, per block step.  Two extra rows (tile row NT)
// carry the right-hand sides b and 1, so the forward substitutions ride along in the same MMAs.
// Per block step (2 barriers): (1) the owners of block-column J publish their 4 raw columns to shared
// memory; (2) one thread per window row factorises the 4x4 diagonal block redundantly (rsqrt, no divide),
// solves its row against it, emits L (column-band layout) and the panel P; (3) every warp updates its tiles
// with DMMA and overwrites the 4 entering rows / columns with tau Q + diag(omega).
// The backward pass is blocked the same way: 4 warps reduce the 4 columns of a block against the ring of
// already-known x in parallel, one lane finishes the 4x4 triangular solve; L is prefetched one block ahead.
// ------------------------------------------------------------------------------------------------
template <int NT>
// 7 resident CTAs per SM (72 registers, < 32 KB of shared memory): 1036 chain slots on 148 SMs, so that the 1024 chains
// of BASELINE config 4 are one wave instead of 1.4 (measured 7.7 -> 6.6 ms per iteration); the widest bands keep 5
__global__ void __launch_bounds__(128, NT <= 7 ? 7 : 5) solve_logit_icar_mma_kernel(const View v, const SolveDebug dbg, double* rec_beta,
                                                                    double* rec_tau, int rec_capacity) {
    constexpr int W8 = 8 * NT, T = 128, RPW = (NT + 1 + 3) / 4;  // tile rows per warp (incl. the RHS tile row NT)
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int n = v.n, n4 = round_up_dev(v.n, 4), p = v.p, w = v.bw, W1 = v.bw + 1;
    const size_t Cp = v.Cp;
    const HParams* hp = v.hp;
    __shared__ double raw[W8 + 8][4];
    __shared__ double P[W8 + 8][4];
    __shared__ double xr[3][W8];          // rings of the backward solves
    __shared__ double S4[4][3];           // backward: the four column sums of a block
    constexpr int NBF = 2, NBB = 4;       // prefetch depth (block steps) of the forward / backward pass
    __shared__ double qin[NBF][4][64];    // ring: lower-band rows of Q for the 4 matrix rows entering at a block step
    __shared__ double oin[NBF][4], bin[NBF][4];  // their omega and b
    __shared__ double lring[NBB][4][64];  // backward ring: the 4 columns of L of a block (column-band layout)
    __shared__ double sring[NBB][4][4];   // backward ring: y_b, y_1, normals, 1 / L_jj of the block
    __shared__ double red[8];
    __shared__ double small_S[MAX_SMALL * MAX_SMALL];
    __shared__ double small_b[MAX_SMALL], small_x[MAX_SMALL], small_t[MAX_SMALL], beta_s[MAX_SMALL], nrm[MAX_SMALL];
    __shared__ int ok_flag;

    const uint32_t iter = v.ctl->iter;
    const int rec_index = v.ctl->rec_index;
    const uint32_t gc = (uint32_t)(v.chain_offset + c);
    const int cr = c % v.C;
    double* Lc = v.Lc + (size_t)c * n * W1;
    double* yb = v.scratch + (size_t)c * 8 * n4;
    double* y1 = yb + n4;
    double* xs = y1 + n4;    // b (forward) then x (backward)
    double* zs = xs + n4;    // normals (forward) then zv (backward)
    double* ms = zs + n4;    // L^-T L^-1 b (debug mean)
    double* om_c = ms + n4;  // omega of this chain
    double* inv_c = om_c + n4;  // 1 / L_jj
    double* kd_c = inv_c + n4;  // z - 1/2 of this chain

    for (int a = tid; a < p; a += T) beta_s[a] = v.theta[(size_t)a * Cp + c];
    for (int k = tid; k < p; k += T) {
        double val;
        if (dbg.normals) {
            val = dbg.normals[(size_t)cr * (n + p + v.r) + n + k];
        } else {
            Philox gph;
            gph.init(v.seed, gc, (uint32_t)(n + k), iter, ST_NORMALS);
            val = gph.normal_d();
        }
        nrm[k] = val;
    }
    if (tid == 0) ok_flag = 1;
    __syncthreads();

    // ---- tau | eta (logit.py:206-209); eta is first copied chain-major so that the band products hit L1 ----
    for (int i = tid; i < n; i += T) ms[i] = v.eta_n[(size_t)i * Cp + c];
    __syncthreads();
    double part = 0.0;
    for (int i = tid; i < n; i += T) {
        const double ei = ms[i];
        const double* qrow = v.Qband + (size_t)i * W1;
        double sacc = 0.5 * qrow[0] * ei;
        const int dmax = min(w, i);
        for (int d = 1; d <= dmax; d++) sacc += qrow[d] * ms[i - d];
        part += ei * sacc;
    }
    const double quad = 2.0 * block_sum_128(part, red, tid);
    __syncthreads();
    if (tid == 0) {
        const double rate = 0.5 * quad + hp->tau_rate;
        red[4] = dbg.tau_in ? dbg.tau_in[cr] : v.gam[c] / rate;
        if (dbg.enabled && dbg.out.tau_rate && c < v.C) dbg.out.tau_rate[c] = rate;
    }
    __syncthreads();
    const double tau = red[4];

    // ---- stage this chain's per-site inputs contiguously (see solve_logit_icar_kernel) ----
    for (int i = tid; i < n4; i += T) {
        double om = 0.0, bi = 0.0, ez = 0.0, kk = 0.0;
        if (i < n) {
            om = v.omega[(size_t)i * Cp + c];
            kk = v.kd[(size_t)i * Cp + c];
            double xb = 0.0;
            for (int a = 0; a < p; a++) xb += v.X[(size_t)i * p + a] * beta_s[a];
            bi = kk - om * xb;  // logit.py:213
            if (dbg.enabled && dbg.out.eta_b && c < v.C) dbg.out.eta_b[(size_t)c * n + i] = bi;
            if (dbg.normals) {
                ez = dbg.normals[(size_t)cr * (n + p + v.r) + i];
            } else {
                Philox gph;
                gph.init(v.seed, gc, (uint32_t)i, iter, ST_NORMALS);
                ez = gph.normal_d();
            }
        }
        om_c[i] = om;
        kd_c[i] = kk;
        xs[i] = bi;
        zs[i] = ez;
    }
    __syncthreads();

    // entering rows i0 .. i0 + 3 -> staging buffer `buf` (plain loads issued a whole block step before use)
    // (each thread stages at most two band entries; their (row, offset) coordinates never change)
    const int st_a0 = tid / W1, st_d0 = tid - st_a0 * W1;
    const int st_a1 = (tid + T) / W1, st_d1 = (tid + T) - st_a1 * W1;
    // 8-byte cp.async with zero fill when `ok` is false: the copies land while the block steps in between run
    auto cp8 = [&](double* dst, const double* src, bool ok) {
        const uint32_t sdst = (uint32_t)__cvta_generic_to_shared(dst);
        const int bytes = ok ? 8 : 0;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sdst), "l"(src), "r"(bytes));
    };
    auto stage_rows = [&](int i0, int buf) {
        {
            const int i = i0 + st_a0;
            const bool ok = st_a0 < 4 && i < n && st_d0 <= i;
            if (st_a0 < 4) cp8(&qin[buf][st_a0][st_d0], ok ? v.Qband + (size_t)i * W1 + st_d0 : v.Qband, ok);
        }
        if (st_a1 < 4) {
            const int i = i0 + st_a1;
            const bool ok = i < n && st_d1 <= i;
            cp8(&qin[buf][st_a1][st_d1], ok ? v.Qband + (size_t)i * W1 + st_d1 : v.Qband, ok);
        }
        if (tid < 8) {
            const int a = tid & 3, i = i0 + a;
            const double* srcv = (tid < 4) ? om_c : xs;
            cp8(tid < 4 ? &oin[buf][a] : &bin[buf][a], i < n ? srcv + i : srcv, i < n);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    // Lambda[i][k] for an entering row or column (max(i, k) in [i0, i0 + 4)), from the staging buffer
    auto lam_in = [&](int i, int k, int i0, int buf) -> double {
        if (i < k) {
            const int t2 = i;
            i = k;
            k = t2;
        }
        const int d = i - k;
        if (i >= n) return d == 0 ? 1.0 : 0.0;
        if (d > w) return 0.0;
        double val = tau * qin[buf][i - i0][d];
        if (d == 0) val += oin[buf][i - i0];
        return val;
    };
    // Lambda[i][k] (identity beyond n) and the right-hand-side rows
    auto lam = [&](int i, int k) -> double {
        if (i < k) {
            const int t2 = i;
            i = k;
            k = t2;
        }
        const int d = i - k;
        if (i >= n) return d == 0 ? 1.0 : 0.0;
        if (d > w) return 0.0;
        double val = tau * v.Qband[(size_t)i * W1 + d];
        if (d == 0) val += om_c[i];
        return val;
    };
    auto rhs = [&](int row, int k) -> double { return (k < n) ? (row == 0 ? xs[k] : (row == 1 ? 1.0 : 0.0)) : 0.0; };

    // ---- window in registers ----
    double acc[RPW][NT][2];
#pragma unroll
    for (int ri = 0; ri < RPW; ri++) {
        const int I = warp + 4 * ri;
#pragma unroll
        for (int K = 0; K < NT; K++) {
            const int sc = 8 * K + 2 * tq;
            if (I < NT) {
                acc[ri][K][0] = lam(8 * I + g, sc);
                acc[ri][K][1] = lam(8 * I + g, sc + 1);
            } else if (I == NT) {
                acc[ri][K][0] = rhs(g, sc);
                acc[ri][K][1] = rhs(g, sc + 1);
            } else {
                acc[ri][K][0] = acc[ri][K][1] = 0.0;
            }
        }
    }

    // run f(K) with K a compile-time constant equal to the (warp-uniform) run-time tile column kc
    auto with_tile_col = [&](int kc, auto f) {
        switch (kc) {
            case 0: f(std::integral_constant<int, 0>{}); break;
            case 1: if constexpr (NT > 1) f(std::integral_constant<int, 1>{}); break;
            case 2: if constexpr (NT > 2) f(std::integral_constant<int, 2>{}); break;
            case 3: if constexpr (NT > 3) f(std::integral_constant<int, 3>{}); break;
            case 4: if constexpr (NT > 4) f(std::integral_constant<int, 4>{}); break;
            case 5: if constexpr (NT > 5) f(std::integral_constant<int, 5>{}); break;
            case 6: if constexpr (NT > 6) f(std::integral_constant<int, 6>{}); break;
            case 7: if constexpr (NT > 7) f(std::integral_constant<int, 7>{}); break;
        }
    };

    // ---- blocked factorisation + fused forward substitutions ----
    int jm4 = 0;
    for (int pre = 0; pre < NBF - 1; pre++) stage_rows(W8 + 4 * pre, pre);  // rows entering at steps 0 .. NBF-2
    for (int j0 = 0; j0 < n4; j0 += 4) {
        const int Kc = jm4 >> 3, h = (jm4 >> 2) & 1;
        const int buf = (j0 >> 2) % NBF, inew = j0 + W8;  // rows inew .. inew + 3 enter at the end of this step
        // (1) publish the raw block column
        with_tile_col(Kc, [&](auto kc) {
            constexpr int K = decltype(kc)::value;
            if ((tq >> 1) == h) {
#pragma unroll
                for (int ri = 0; ri < RPW; ri++) {
                    const int I = warp + 4 * ri;
                    if (I <= NT) {
                        const int row = (I < NT ? 8 * I : W8) + g;
                        raw[row][2 * (tq & 1)] = acc[ri][K][0];
                        raw[row][2 * (tq & 1) + 1] = acc[ri][K][1];
                    }
                }
            }
        });
        __syncthreads();
        // rows entering NBF-1 steps from now -> the ring buffer read last at the previous step (everyone is past
        // it); then wait for THIS step's rows (issued NBF-1 steps ago), published by the barrier below
        stage_rows(inew + 4 * (NBF - 1), ((j0 >> 2) + NBF - 1) % NBF);
        asm volatile("cp.async.wait_group %0;\n" ::"n"(NBF - 1));
        // (2) one thread per window row: factorise the diagonal block, solve the row, emit L and P
        if (tid < W8 + 8) {
            double d[4][4], invd[4];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) d[a][b] = raw[jm4 + a][b];
            if (!chol4(d, invd)) {
                ok_flag = 0;
            } else {
                double x[4];
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    double t = raw[tid][b];
#pragma unroll
                    for (int k = 0; k < b; k++) t -= x[k] * d[b][k];
                    x[b] = t * invd[b];
                }
                if (tid < W8) {
                    int dt = tid - jm4;
                    if (dt < 0) dt += W8;
                    const bool leaving = dt < 4;  // rows of the diagonal block: x_b = L_JJ[dt][b] for b <= dt
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const int e = dt - b;
                        if (e >= 0 && e <= w && j0 + b < n) Lc[(size_t)(j0 + b) * W1 + e] = x[b];
                        P[tid][b] = leaving ? 0.0 : x[b];
                    }
                    if (dt == 0) {
#pragma unroll
                        for (int b = 0; b < 4; b++) inv_c[j0 + b] = invd[b];
                    }
                } else {
                    const int rr = tid - W8;
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        P[tid][b] = rr < 2 ? x[b] : 0.0;
                        if (rr == 0) yb[j0 + b] = x[b];  // y_b = L^-1 b
                        if (rr == 1) y1[j0 + b] = x[b];  // y_1 = L^-1 1
                    }
                }
            }
        }
        __syncthreads();
        if (!ok_flag) break;
        // (3) rank-4 trailing update on the tensor pipe
        const int base = jm4 + 4;  // ring slot of matrix index j0 + 4 after the slide
#pragma unroll
        for (int ri = 0; ri < RPW; ri++) {
            const int I = warp + 4 * ri;
            if (I <= NT) {
                const int prow = (I < NT ? 8 * I : W8) + g;
                const double aF = -P[prow][tq];
#pragma unroll
                for (int K = 0; K < NT; K++) dgemm::dmma884(acc[ri][K][0], acc[ri][K][1], aF, P[8 * K + g][tq]);
            }
        }
        // (4a) the 4 entering ROWS live in tile row Kc: only its owner warp, only the lanes on those rows
#pragma unroll
        for (int ri = 0; ri < RPW; ri++) {
            if (warp + 4 * ri == Kc && (g >> 2) == h) {
                const int ir = inew + (g & 3);
#pragma unroll
                for (int K = 0; K < NT; K++) {
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        int dts = 8 * K + 2 * tq + u - base;
                        if (dts < 0) dts += W8;
                        acc[ri][K][u] = lam_in(ir, j0 + 4 + dts, inew, buf);
                    }
                }
            }
        }
        // (4b) the 4 entering COLUMNS live in tile column Kc: the lanes holding those columns, every tile row
        if ((tq >> 1) == h) {
            const int ic = inew + 2 * (tq & 1);
#pragma unroll
            for (int ri = 0; ri < RPW; ri++) {
                const int I = warp + 4 * ri;
                if (I <= NT) {
                    double v0, v1;
                    if (I < NT) {
                        int dts = 8 * I + g - base;
                        if (dts < 0) dts += W8;
                        const int ir = j0 + 4 + dts;
                        v0 = lam_in(ir, ic, inew, buf);
                        v1 = lam_in(ir, ic + 1, inew, buf);
                    } else {
                        v0 = (ic < n) ? (g == 0 ? bin[buf][ic - inew] : (g == 1 ? 1.0 : 0.0)) : 0.0;
                        v1 = (ic + 1 < n) ? (g == 0 ? bin[buf][ic + 1 - inew] : (g == 1 ? 1.0 : 0.0)) : 0.0;
                    }
                    with_tile_col(Kc, [&](auto kc) {
                        constexpr int K = decltype(kc)::value;
                        acc[ri][K][0] = v0;
                        acc[ri][K][1] = v1;
                    });
                }
            }
        }
        jm4 = base == W8 ? 0 : base;
    }
    if (!ok_flag) {
        if (tid == 0) {
            v.status[c] = 1;
            if (c < v.C) atomicAdd(&v.ctl->fails, 1);
        }
        return;
    }
    if (dbg.enabled && dbg.out.eta_chol && c < v.C)  // column-band factor [n][bw + 1]
        for (size_t idx = tid; idx < (size_t)n * W1; idx += T) {
            const int j = (int)(idx / W1), e = (int)(idx - (size_t)j * W1);
            dbg.out.eta_chol[(size_t)c * n * W1 + idx] = (j + e < n) ? Lc[idx] : 0.0;
        }
    for (int e = tid; e < 3 * W8; e += T) (&xr[0][0])[e] = 0.0;
    __syncthreads();

    // ---- blocked backward: L' x = y_b + eps, L' zv = y_1, L' m = y_b ----
    {
        asm volatile("cp.async.wait_group 0;\n" ::);
        const int b = warp;  // this warp's column within the block
        // block j0 -> ring buffer (j0 / 4) % NBB: 4 columns of L and the 4 x 4 scalars of the finish
        auto stage_block = [&](int j0, int buf) {
            if (j0 >= 0) {
                if (st_a0 < 4) {
                    const int j = j0 + st_a0;
                    const bool ok = j < n;
                    cp8(&lring[buf][st_a0][st_d0], ok ? Lc + (size_t)j * W1 + st_d0 : Lc, ok);
                }
                if (st_a1 < 4) {
                    const int j = j0 + st_a1;
                    const bool ok = j < n;
                    cp8(&lring[buf][st_a1][st_d1], ok ? Lc + (size_t)j * W1 + st_d1 : Lc, ok);
                }
                if (tid < 16) {
                    const int bb = tid & 3, which = tid >> 2, jj = j0 + bb;
                    const double* srcv = which == 0 ? yb : (which == 1 ? y1 : (which == 2 ? zs : inv_c));
                    cp8(&sring[buf][which][bb], jj < n ? srcv + jj : srcv, jj < n);
                }
            }
            asm volatile("cp.async.commit_group;\n" ::);
        };
        const int last = n4 - 4;
        for (int pre = 0; pre < NBB - 1; pre++) stage_block(last - 4 * pre, ((last >> 2) - pre + 4 * NBB) % NBB);
        asm volatile("cp.async.wait_group %0;\n" ::"n"(NBB - 2));
        __syncthreads();
        for (int j0 = last; j0 >= 0; j0 -= 4) {
            const int buf = (j0 >> 2) % NBB;
            const int j = j0 + b;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
            // rows beyond the block (e >= 4 - b) are already solved and sit in the ring; entries past row n are
            // garbage in L's band storage, so they are masked by index rather than relied upon to be zero
            for (int e = lane + 4 - b; e <= w; e += 32) {
                if (j + e < n) {
                    const double l = lring[buf][b][e];
                    const int slot = (j + e) % W8;
                    s0 += l * xr[0][slot];
                    s1 += l * xr[1][slot];
                    s2 += l * xr[2][slot];
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            if (lane == 0) {
                S4[b][0] = s0;
                S4[b][1] = s1;
                S4[b][2] = s2;
            }
            __syncthreads();
            if (tid < 3) {  // one lane per right-hand side finishes the 4x4 triangular solve
                const int rr = tid;
                double fr[4], fi[4], xb4[4];
#pragma unroll
                for (int bb = 0; bb < 4; bb++) {
                    const double ybv = sring[buf][0][bb];
                    fr[bb] = rr == 0 ? ybv + sring[buf][2][bb] : (rr == 1 ? sring[buf][1][bb] : ybv);
                    fi[bb] = sring[buf][3][bb];  // zero beyond n => x = 0 there
                }
                const bool in1 = j0 + 1 < n, in2 = j0 + 2 < n, in3 = j0 + 3 < n;
                const double l10 = in1 ? lring[buf][0][1] : 0.0, l20 = in2 ? lring[buf][0][2] : 0.0;
                const double l30 = in3 ? lring[buf][0][3] : 0.0, l21 = in2 ? lring[buf][1][1] : 0.0;
                const double l31 = in3 ? lring[buf][1][2] : 0.0, l32 = in3 ? lring[buf][2][1] : 0.0;
                xb4[3] = (fr[3] - S4[3][rr]) * fi[3];
                xb4[2] = (fr[2] - S4[2][rr] - l32 * xb4[3]) * fi[2];
                xb4[1] = (fr[1] - S4[1][rr] - l21 * xb4[2] - l31 * xb4[3]) * fi[1];
                xb4[0] = (fr[0] - S4[0][rr] - l10 * xb4[1] - l20 * xb4[2] - l30 * xb4[3]) * fi[0];
#pragma unroll
                for (int bb = 0; bb < 4; bb++) xr[rr][(j0 + bb) % W8] = xb4[bb];
                double* dst = rr == 0 ? xs : (rr == 1 ? zs : ms);   // (zs: this block's normals are in the ring already)
#pragma unroll
                for (int bb = 0; bb < 4; bb++) dst[j0 + bb] = xb4[bb];
            }
            // block NBB-1 steps ahead -> the buffer consumed at the previous step; then wait for the next block
            stage_block(j0 - 4 * (NBB - 1), ((j0 >> 2) + 1) % NBB);
            asm volatile("cp.async.wait_group %0;\n" ::"n"(NBB - 2));
            __syncthreads();
        }
    }

    // ---- eta = x - (sum x / sum zv) zv (distributions.pyx:24-39); beta system (logit.py:226-231) ----
    double sx = 0.0, sz = 0.0;
    for (int i = tid; i < n; i += T) {
        sx += xs[i];
        sz += zs[i];
    }
    sx = block_sum_128(sx, red, tid);
    sz = block_sum_128(sz, red, tid);
    const double shift = -sx / sz;
    for (int i = tid; i < n; i += T) {
        const double e = xs[i] + shift * zs[i];
        xs[i] = e;
        v.eta_n[(size_t)i * Cp + c] = e;
        if (dbg.enabled && dbg.out.eta_mean && c < v.C) dbg.out.eta_mean[(size_t)c * n + i] = ms[i];
    }
    __syncthreads();
    for (int idx = 0; idx < p * p; idx++) {  // A = X' Om X + b_prec
        const int a = idx / p, b = idx - a * p;
        if (b > a) continue;
        double sacc = 0.0;
        for (int i = tid; i < n; i += T) sacc += om_c[i] * v.X[(size_t)i * p + a] * v.X[(size_t)i * p + b];
        sacc = block_sum_128(sacc, red, tid);
        if (tid == 0) {
            small_S[a * p + b] = sacc + hp->b_prec[a * p + b];
            small_S[b * p + a] = sacc + hp->b_prec[b * p + a];
        }
    }
    for (int a = 0; a < p; a++) {  // b = X'(k - omega * eta) + b_prec b_mu
        double sacc = 0.0;
        for (int i = tid; i < n; i += T) sacc += v.X[(size_t)i * p + a] * (kd_c[i] - om_c[i] * xs[i]);
        sacc = block_sum_128(sacc, red, tid);
        if (tid == 0) small_b[a] = sacc + hp->b_pm[a];
    }
    __syncthreads();
    if (dbg.enabled && c < v.C) {
        if (dbg.out.beta_prec)
            for (int idx = tid; idx < p * p; idx += T) dbg.out.beta_prec[(size_t)c * p * p + idx] = small_S[idx];
        if (dbg.out.beta_b)
            for (int a = tid; a < p; a += T) dbg.out.beta_b[(size_t)c * p + a] = small_b[a];
    }
    __syncthreads();
    if (tid == 0) {
        if (small_precision_mvnorm(small_S, p, small_b, nrm, small_x, small_t)) {
            for (int a = 0; a < p; a++) beta_s[a] = small_x[a];
        } else {
            ok_flag = 0;
        }
    }
    __syncthreads();
    if (!ok_flag) {
        if (tid == 0) {
            v.status[c] = 1;
            if (c < v.C) atomicAdd(&v.ctl->fails, 1);
        }
        return;
    }
    for (int a = tid; a < p; a += T) v.theta[(size_t)a * Cp + c] = beta_s[a];
    if (tid == 0) v.tau[c] = tau;
    if (rec_beta && rec_index >= 0 && rec_index < rec_capacity && c < v.C) {
        const size_t row = (size_t)c * rec_capacity + rec_index;
        for (int a = tid; a < p; a += T) rec_beta[row * p + a] = beta_s[a];
        if (tid == 0) rec_tau[row] = tau;
    }
}

// ------------------------------------------------------------------------------------------------
// host-side drivers
// ------------------------------------------------------------------------------------------------
// The RSR logit sweep computes psi inside the site pass (FUSE_PSI) in production precision; the FP64-sampler
// validation build and OCCU_FUSE_PSI=0 keep the separate GEMM.
static bool fuse_psi(const occu_engine* e) {
    static int allowed = -1;
    if (allowed < 0) {
        const char* s = getenv("OCCU_FUSE_PSI");
        allowed = s ? atoi(s) : 1;
    }
    return allowed && e->v.model == OCCU_LOGIT_RSR && !e->pg_fp64;
}

template <typename PGT, bool SLICES, bool FUSE_PSI>
static void launch_site_pass_logit(occu_engine* e, cudaStream_t st) {
    const View& v = e->v;
    dim3 grid(v.n_tiles, v.Cp / CHAIN_TILE);
    const size_t smem = (size_t)SITE_TILE * CHAIN_TILE * sizeof(PGT) + (FUSE_PSI ? PSI_SUB * CHAIN_TILE * sizeof(double) : 0);
    switch (v.r) {
#define CASE(R)                                                                                                      \
    case R:                                                                                                          \
        if (smem > 48 * 1024)                                                                                        \
            cudaFuncSetAttribute(site_pass_logit_kernel<R, PGT, SLICES, FUSE_PSI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        site_pass_logit_kernel<R, PGT, SLICES, FUSE_PSI><<<grid, CHAIN_TILE, smem, st>>>(v);                          \
        break;
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    }
    e->launches++;
}

static void launch_site_pass(occu_engine* e, const SiteDebug& dbg, cudaStream_t st) {
    const View& v = e->v;
    if (v.model == OCCU_PROBIT_RSR) {
        dim3 grid(v.n_tiles, v.Cp / CHAIN_TILE);
        switch (v.r) {
#define CASE(R)                                                                             \
    case R:                                                                                 \
        if (e->pg_fp64) site_pass_probit_kernel<R, true><<<grid, CHAIN_TILE, 0, st>>>(v, dbg);  \
        else site_pass_probit_kernel<R, false><<<grid, CHAIN_TILE, 0, st>>>(v, dbg);            \
        break;
            CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
        }
        e->launches++;
    } else if (v.int8) {
        // the site pass writes the digits of (omega, z) for the tcgen05 Gram itself
        if (e->pg_fp64) launch_site_pass_logit<double, true, false>(e, st);
        else if (fuse_psi(e)) launch_site_pass_logit<float, true, true>(e, st);
        else launch_site_pass_logit<float, true, false>(e, st);
        e->slices_valid = true;
        e->omega_in_slices = v.zprob == nullptr;
    } else if (e->pg_fp64) {
        launch_site_pass_logit<double, false, false>(e, st);
    } else if (fuse_psi(e)) {
        launch_site_pass_logit<float, false, true>(e, st);
    } else {
        launch_site_pass_logit<float, false, false>(e, st);
    }
}

static void launch_alpha_injected(occu_engine* e, const double* omega_a, cudaStream_t st) {
    const View& v = e->v;
    dim3 grid(v.n_tiles, v.Cp / CHAIN_TILE);
    switch (v.r) {
#define CASE(R) case R: alpha_partials_injected_kernel<R><<<grid, CHAIN_TILE, 0, st>>>(v, omega_a); break;
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    }
    e->launches++;
}

static int launch_psi_gemm(occu_engine* e, cudaStream_t st) {
    const View& v = e->v;
    if (v.model == OCCU_LOGIT_ICAR) {
        const size_t total = (size_t)v.n * v.Cp;
        psi_icar_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(v);
        e->launches++;
        CUDA_OK(cudaGetLastError());
        return 0;
    }
    dgemm::Args a;
    memset(&a, 0, sizeof(a));
    a.A1 = v.UT; a.B = v.theta; a.Out = v.psi;
    a.M = v.n_pad; a.N = v.Cp; a.K = v.kq_pad;
    a.lda = v.n_pad; a.ldb = v.Cp; a.ldo = v.Cp;
    CUDA_OK(dgemm::launch(a, st));
    e->launches++;
    return 0;
}

// G and H on the tcgen05 tensor cores (gram_tc.cuh).  The site pass of a sweep writes the digits of (omega, z)
// itself; when omega or z were set from outside (debug / test hooks) they are sliced here first.
static int launch_gram_tc(occu_engine* e, cudaStream_t st) {
    const View& v = e->v;
    const gtc::Geometry& g = e->geo;
    if (!e->slices_valid) {
        gtc::slice_a_kernel<<<dim3((v.tc_Kp + 31) / 32, v.Cp / 32), 256, 0, st>>>(v.omega, v.z, v.n, v.tc_Kp, v.Cp, v.Aq);
        e->launches++;
        e->slices_valid = true;
        e->omega_in_slices = false;
    }
    gtc::Params p;
    p.Cp = v.Cp; p.kblocks = g.kblocks; p.nsplit = g.nsplit; p.kb_per_chunk = g.kb_per_chunk; p.MT = g.MT;
    p.n_gt = g.n_gt; p.n_ht = g.n_ht; p.Mrows = g.Mrows; p.M = v.M; p.pq = v.pq; p.ldg = v.M_pad; p.ldh = v.ldh;
    p.scale = v.tscale; p.hshift = v.hshift; p.err = v.tc_err; p.mode = 0;
    p.order = e->tc_order;
    p.Gpart = g.nsplit > 1 ? v.Gp : v.G;
    p.Hpart = g.nsplit > 1 ? v.Hp : v.Hpart;
    const int nitems = g.MT * (g.n_gt + g.n_ht) * g.nsplit;
    const int grid = nitems < e->sm_count ? nitems : e->sm_count;
    gtc::gram_tc_kernel<TC_BN><<<grid, gtc::THREADS, gtc::Cfg<TC_BN>::SMEM_BYTES, st>>>(e->tmA, e->tmB, p);
    e->launches++;
    if (g.nsplit > 1) {
        const size_t tg = (size_t)v.Cp * v.M_pad, th = (size_t)v.Cp * v.ldh;
        gtc::reduce_parts_kernel<<<(unsigned)((tg + 255) / 256), 256, 0, st>>>(v.Gp, g.nsplit, tg, tg, v.M_pad, v.M, v.G);
        gtc::reduce_parts_kernel<<<(unsigned)((th + 255) / 256), 256, 0, st>>>(v.Hp, g.nsplit, th, th, v.ldh, v.pq, v.Hpart);
        e->launches += 2;
    }
    CUDA_OK(cudaGetLastError());
    return 0;
}

static int launch_gram_gemm(occu_engine* e, cudaStream_t st) {
    const View& v = e->v;
    if (v.int8) return launch_gram_tc(e, st);
    dgemm::Args a;
    memset(&a, 0, sizeof(a));
    a.A1 = v.omega; a.B = v.ZU; a.Out = v.G;
    a.M = v.Cp; a.N = v.model == OCCU_PROBIT_RSR ? 0 : v.M_pad; a.K = v.n_pad;  // probit: only U'd
    a.lda = v.Cp; a.ldb = v.ldz; a.ldo = v.M_pad;
    a.A2 = v.kd; a.Out2 = v.Hpart; a.N2 = v.ldh; a.col2 = v.M_pad; a.ldo2 = v.ldh; a.ksplit = v.ksplit;
    CUDA_OK(dgemm::launch(a, st));
    e->launches++;
    return 0;
}

static size_t icar_smem_bytes(const View& v) {
    const size_t W1 = v.bw + 1;
    return (W1 * W1 + 6 * W1 + 2 * MAX_SMALL + 8) * sizeof(double) + (size_t)v.bw * (v.bw + 1) + 16;
}

static size_t solve_smem_bytes(const View& v) { return solve_bytes_for(v.q, v.p, v.pq); }

// Threads per chain in the RSR solve.  Most phases of the blocked factorisation keep only a few dozen threads busy
// and the kernel is barrier bound: for small q fewer threads per chain put more chains on an SM.
static int solve_threads_for(int q) {
    static int forced = -1;
    if (forced < 0) {
        const char* s = getenv("OCCU_SOLVE_THREADS");
        forced = s ? atoi(s) : 0;
    }
    if (forced >= 32 && forced <= SOLVE_THREADS && forced % 32 == 0 && (q + 3) / 4 <= forced) return forced;
    return SOLVE_THREADS;
}

#define PROF_MARK(e, k, st)                                           \
    do {                                                              \
        if ((e)->profiling) cudaEventRecord((e)->ev[k], st);          \
    } while (0)

// profiling mode synchronises after every iteration to read the event pairs
static int prof_collect(occu_engine* e, cudaStream_t st) {
    if (!e->profiling) return 0;
    CUDA_OK(cudaEventSynchronize(e->ev[KC_COUNT]));
    for (int k = 0; k < KC_COUNT; k++) {
        float ms = 0.f;
        CUDA_OK(cudaEventElapsedTime(&ms, e->ev[k], e->ev[k + 1]));
        e->kernel_ms[k] += ms;
        e->kernel_calls[k] += 1;
    }
    return 0;
}

static int launch_gaussian_half(occu_engine* e, const SolveDebug& dbg, double* ra, double* rb, double* rt, int cap,
                                cudaStream_t st) {
    const View& v = e->v;
    PROF_MARK(e, KC_GRAM_GEMM, st);
    if (v.model != OCCU_LOGIT_ICAR && launch_gram_gemm(e, st)) return -1;
    PROF_MARK(e, KC_REDUCE, st);
    const size_t per = (size_t)v.NV * v.Cp;
    reduce_partials_kernel<<<(unsigned)((per + 31) / 32), dim3(32, RED_SPLIT), 0, st>>>(v);
    alpha_tau_kernel<<<(v.Cp + 31) / 32, 32, 0, st>>>(v, dbg, ra, cap);  // 1 warp per CTA: spread over all SMs
    e->launches += 2;
    PROF_MARK(e, KC_SOLVE, st);
    if (v.model == OCCU_LOGIT_ICAR) {
        const size_t smem_i = icar_smem_bytes(v);
        size_t& configured_i = e->smem_configured[0];   // per engine: the attribute is per device
        if (smem_i > configured_i) {
            CUDA_OK(cudaFuncSetAttribute(solve_logit_icar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_i));
            configured_i = smem_i;
        }
        const int nt = (v.bw + 4 + 7) / 8;
        switch (nt <= 8 ? nt : 0) {
#define CASE(NT) case NT: solve_logit_icar_mma_kernel<NT><<<v.Cp, 128, 0, st>>>(v, dbg, rb, rt, cap); break;
            CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
            default: solve_logit_icar_kernel<<<v.Cp, ICAR_THREADS, smem_i, st>>>(v, dbg, ra, rb, rt, cap);
        }
        e->launches++;
        PROF_MARK(e, KC_COUNT, st);
        CUDA_OK(cudaGetLastError());
        return 0;
    }
    const size_t smem = v.solve_stride ? 0 : solve_smem_bytes(v);
    size_t& configured = e->smem_configured[1];
    if (smem > configured) {
        CUDA_OK(cudaFuncSetAttribute(solve_logit_rsr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    if (v.solve_stride) {   // large q: tiled factorisation out of global scratch
        const size_t smem_t = tiled_smem_bytes(v.q, v.pq);
        size_t& configured_t = e->smem_configured[2];
        if (smem_t > configured_t) {
            CUDA_OK(cudaFuncSetAttribute(solve_logit_rsr_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
            configured_t = smem_t;
        }
        solve_logit_rsr_tiled_kernel<<<v.Cp, TILED_THREADS, smem_t, st>>>(v, dbg, rb, rt, cap);
    } else {
        solve_logit_rsr_kernel<<<v.Cp, solve_threads_for(v.q), smem, st>>>(v, dbg, ra, rb, rt, cap);
    }
    e->launches++;
    PROF_MARK(e, KC_COUNT, st);
    CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" {

const char* occu_last_error(void) { return g_err; }
int occu_abi_version(void) { return OCCU_ABI_VERSION; }

size_t occu_workspace_bytes(const occu_problem_t* prob) {
    View v;
    memset(&v, 0, sizeof(v));
    if (fill_dims(prob, v)) return 0;
    size_t total = 0;
    carve(v, nullptr, &total, prob->debug);
    return total;
}

int occu_create(const occu_problem_t* P, void* workspace, size_t workspace_bytes, void* stream, occu_engine_t** out) {
    if (!out) return set_err("null out pointer");
    cudaStream_t st = (cudaStream_t)stream;
    View v;
    memset(&v, 0, sizeof(v));
    if (fill_dims(P, v)) return -1;
    size_t need = 0;
    carve(v, nullptr, &need, P->debug);
    if (!workspace || workspace_bytes < need) return set_err("workspace too small");
    if ((uintptr_t)workspace % 256) return set_err("workspace must be 256-byte aligned");
    const bool icar = P->model == OCCU_LOGIT_ICAR;
    if (!P->X || !P->W || !P->y || !P->visit_ptr || !P->site_class) return set_err("null problem array");
    if (!icar && (!P->K || !P->Qrsr)) return set_err("RSR models need K and Qrsr");
    if (icar && !P->Qband) return set_err("ICAR model needs Qband");
    if (!P->a_mu || !P->a_prec || !P->b_mu || !P->b_prec) return set_err("null hyper-parameter array");
    carve(v, (char*)workspace, &need, P->debug);

    occu_engine* e = new (std::nothrow) occu_engine;
    if (!e) return set_err("out of host memory");
    e->v = v;
    e->prob = *P;
    e->pg_fp64 = P->pg_fp64;
    e->debug = P->debug;
    e->launches = 0;
    e->ws = (char*)workspace;
    e->ws_bytes = need;
    e->profiling = 0;
    {
        int dev = 0;
        e->sm_count = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, dev);
        const char* ord = getenv("OCCU_TC_ORDER");
        e->tc_order = ord ? atoi(ord) : 0;
        e->smem_configured[0] = e->smem_configured[1] = e->smem_configured[2] = 0;
    }
    for (int k = 0; k < KC_COUNT; k++) { e->kernel_ms[k] = 0.0; e->kernel_calls[k] = 0; }
    for (int k = 0; k <= KC_COUNT; k++) e->ev[k] = nullptr;

    // hyper-parameters: prec @ mu folded once (base.py:161-162)
    HParams h;
    memset(&h, 0, sizeof(h));
    h.tau_rate = P->tau_rate;
    h.tau_shape = P->tau_shape;
    for (int a = 0; a < v.r; a++) {
        double s = 0.0;
        for (int b = 0; b < v.r; b++) {
            h.a_prec[a * v.r + b] = P->a_prec[a * v.r + b];
            s += P->a_prec[a * v.r + b] * P->a_mu[b];
        }
        h.a_pm[a] = s;
    }
    for (int a = 0; a < v.p; a++) {
        double s = 0.0;
        for (int b = 0; b < v.p; b++) {
            h.b_prec[a * v.p + b] = P->b_prec[a * v.p + b];
            s += P->b_prec[a * v.p + b] * P->b_mu[b];
        }
        h.b_pm[a] = s;
    }
#define FAIL_IF(call)                                                        \
    do {                                                                     \
        cudaError_t _e = (call);                                             \
        if (_e != cudaSuccess) {                                             \
            delete e;                                                        \
            return set_err("CUDA error %s at %s", cudaGetErrorString(_e), #call); \
        }                                                                    \
    } while (0)
    FAIL_IF(cudaMemsetAsync(workspace, 0, need, st));
    FAIL_IF(cudaMemcpyAsync((void*)v.hp, &h, sizeof(h), cudaMemcpyHostToDevice, st));
    FAIL_IF(cudaStreamSynchronize(st));  // h is a stack object
    FAIL_IF(cudaMemcpyAsync((void*)v.W, P->W, (size_t)v.V * v.r * 8, cudaMemcpyDeviceToDevice, st));
    FAIL_IF(cudaMemcpyAsync((void*)v.y, P->y, (size_t)v.V, cudaMemcpyDeviceToDevice, st));
    FAIL_IF(cudaMemcpyAsync((void*)v.visit_ptr, P->visit_ptr, ((size_t)v.n + 1) * 4, cudaMemcpyDeviceToDevice, st));
    FAIL_IF(cudaMemcpyAsync((void*)v.site_class, P->site_class, (size_t)v.n, cudaMemcpyDeviceToDevice, st));
    if (icar) {
        FAIL_IF(cudaMemcpyAsync((void*)v.X, P->X, (size_t)v.n * v.p * 8, cudaMemcpyDeviceToDevice, st));
        FAIL_IF(cudaMemcpyAsync((void*)v.Qband, P->Qband, (size_t)v.n * (v.bw + 1) * 8, cudaMemcpyDeviceToDevice, st));
    } else {
        FAIL_IF(cudaMemcpyAsync((void*)v.Qrsr, P->Qrsr, (size_t)v.q * v.q * 8, cudaMemcpyDeviceToDevice, st));
        build_design_kernel<<<v.n, 128, v.pq * sizeof(double), st>>>(v, P->X, P->K);
        e->launches++;
    }
    if (v.model == OCCU_PROBIT_RSR) {
        gram_unit_kernel<<<(v.M + 127) / 128, 128, 0, st>>>(v);
        e->launches++;
    }
    e->slices_valid = false;
    e->omega_in_slices = false;
    if (v.int8) {
        // slice [Z | U] once (formed on the fly from U'), encode the TMA descriptors of the two operand arrays
        e->geo = gtc::make_geometry(v.n, v.Cp, v.M, v.pq, TC_BN);
        const gtc::Geometry& g = e->geo;
        gtc::slice_rows_from_ut_kernel<<<g.Mrows, 256, 0, st>>>(v.UT, v.n, v.n_pad, g.Kp, v.M, g.Mg_pad, v.pq, g.Mrows, v.tscale,
                                                                 v.Bq);
        gtc::hshift_from_ut_kernel<<<(v.ldh + 127) / 128, 128, 0, st>>>(v.UT, v.n, v.n_pad, v.pq, v.ldh, v.hshift);
        e->launches += 2;
        FAIL_IF(cudaMemsetAsync(v.Aq, 0, (size_t)(gtc::SA + 1) * v.Cp * g.Kp, st));
        FAIL_IF(cudaMemsetAsync(v.G, 0, (size_t)v.Cp * v.M_pad * 8, st));
        FAIL_IF(cudaMemsetAsync(v.tc_err, 0, sizeof(int), st));
        const char* why = gtc::make_map(&e->tmA, v.Aq, (uint64_t)(gtc::SA + 1) * v.Cp, g.Kp, gtc::BM);
        if (!why) why = gtc::make_map(&e->tmB, v.Bq, (uint64_t)gtc::SB * g.Mrows, g.Kp, TC_BN);
        if (!why && cudaFuncSetAttribute(gtc::gram_tc_kernel<TC_BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         gtc::Cfg<TC_BN>::SMEM_BYTES) != cudaSuccess)
            why = "cudaFuncSetAttribute(shared memory) failed";
        if (why) {
            occu_destroy(e);
            return set_err("INT8 Gram path: %s", why);
        }
    }
    const size_t total = (size_t)v.n * v.Cp;
    reset_z_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(v);
    set_control_kernel<<<1, 1, 0, st>>>(v.ctl, 0u, 1, -1, 1);
    e->launches += 2;
    FAIL_IF(cudaGetLastError());
#undef FAIL_IF
    *out = e;
    return 0;
}

void occu_destroy(occu_engine_t* eng) {
    if (!eng) return;
    for (int k = 0; k <= KC_COUNT; k++)
        if (eng->ev[k]) cudaEventDestroy(eng->ev[k]);
    delete eng;
}

int occu_set_profiling(occu_engine_t* e, int on) {
    if (!e) return set_err("null engine");
    if (on && !e->ev[0])
        for (int k = 0; k <= KC_COUNT; k++) CUDA_OK(cudaEventCreate(&e->ev[k]));
    e->profiling = on ? 1 : 0;
    for (int k = 0; k < KC_COUNT; k++) { e->kernel_ms[k] = 0.0; e->kernel_calls[k] = 0; }
    return 0;
}

int occu_get_kernel_ms(const occu_engine_t* e, double* ms, long long* calls, int count) {
    if (!e || !ms || !calls) return set_err("null argument");
    for (int k = 0; k < count && k < KC_COUNT; k++) { ms[k] = e->kernel_ms[k]; calls[k] = e->kernel_calls[k]; }
    return KC_COUNT;
}

long long occu_launch_count(const occu_engine_t* eng) { return eng ? eng->launches : 0; }

static int field_info(const occu_engine* e, int field, void** ptr, int* d, int* is_u8) {
    const View& v = e->v;
    *is_u8 = 0;
    switch (field) {
        case OCCU_F_ALPHA: *ptr = v.alpha; *d = v.r; return 0;
        case OCCU_F_BETA: *ptr = v.theta; *d = v.p; return 0;
        case OCCU_F_ETA:
            if (v.model == OCCU_LOGIT_ICAR) { *ptr = v.eta_n; *d = v.n; return 0; }
            *ptr = v.theta + (size_t)v.p * v.Cp; *d = v.q; return 0;
        case OCCU_F_TAU: *ptr = v.tau; *d = 1; return 0;
        case OCCU_F_Z: *ptr = v.z; *d = v.n; *is_u8 = 1; return 0;
        case OCCU_F_OMEGA: *ptr = v.omega; *d = v.n; return 0;
        case OCCU_F_PSI: *ptr = v.psi; *d = v.n; return 0;
        case OCCU_F_EPS:
            if (!v.eps) return set_err("EPS exists for the probit model only");
            *ptr = v.eps; *d = v.n; return 0;
        case OCCU_F_ZPROB:
            if (!v.zprob) return set_err("ZPROB needs a debug engine");
            *ptr = v.zprob; *d = v.n; return 0;
        default: return set_err("unknown or unsupported field");
    }
}

// kd must follow z when z is set from outside
__global__ void kd_from_z_kernel(View v) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v.kd && idx < (size_t)v.n * v.Cp) v.kd[idx] = (double)v.z[idx] - 0.5;
}

// omega from its digits (tcgen05 Gram path: a non-debug site pass writes only the digits)
__global__ void omega_from_slices_kernel(View v) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)v.n * v.Cp) return;
    const size_t i = idx / v.Cp, c = idx % v.Cp;
    const size_t plane = (size_t)v.Cp * v.tc_Kp;
    unsigned long long q = 0ull;
    for (int s = 0; s < gtc::SA; s++) q = (q << 8) | v.Aq[(size_t)s * plane + c * v.tc_Kp + i];
    v.omega[idx] = (double)q * (1.0 / 35184372088832.0);
}

static int materialise_omega(occu_engine* e, cudaStream_t st) {
    if (!(e->v.int8 && e->slices_valid && e->omega_in_slices)) return 0;
    const size_t total = (size_t)e->v.n * e->v.Cp;
    omega_from_slices_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(e->v);
    e->launches++;
    e->omega_in_slices = false;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int occu_set_field(occu_engine_t* e, int field, const void* src, void* stream) {
    if (!e || !src) return set_err("null argument");
    if (field == OCCU_F_PSI || field == OCCU_F_ZPROB) return set_err("field is read-only");
    cudaStream_t st = (cudaStream_t)stream;
    void* ptr; int d, u8;
    if (field_info(e, field, &ptr, &d, &u8)) return -1;
    if (field == OCCU_F_Z || field == OCCU_F_OMEGA) {
        // the digits no longer describe (omega, z): the next Gram re-slices from the FP64 / u8 fields
        if (field == OCCU_F_Z && materialise_omega(e, st)) return -1;
        e->slices_valid = false;
        e->omega_in_slices = false;
    }
    const View& v = e->v;
    dim3 grid((v.Cp + 31) / 32, (d + 31) / 32), block(32, 8);
    if (u8) to_chain_minor_kernel<uint8_t><<<grid, block, 0, st>>>((const uint8_t*)src, (uint8_t*)ptr, v.C, v.Cp, d);
    else to_chain_minor_kernel<double><<<grid, block, 0, st>>>((const double*)src, (double*)ptr, v.C, v.Cp, d);
    e->launches++;
    if (field == OCCU_F_Z) {
        const size_t total = (size_t)v.n * v.Cp;
        kd_from_z_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(v);
        e->launches++;
    }
    CUDA_OK(cudaGetLastError());
    return 0;
}

int occu_get_field(occu_engine_t* e, int field, void* dst, void* stream) {
    if (!e || !dst) return set_err("null argument");
    cudaStream_t st = (cudaStream_t)stream;
    void* ptr; int d, u8;
    if (field_info(e, field, &ptr, &d, &u8)) return -1;
    if (field == OCCU_F_OMEGA && materialise_omega(e, st)) return -1;
    // psi lives only inside the site pass when it is fused there (non-debug engines): recompute it from the current state
    if (field == OCCU_F_PSI && fuse_psi(e) && !e->debug && launch_psi_gemm(e, st)) return -1;
    const View& v = e->v;
    dim3 grid((v.Cp + 31) / 32, (d + 31) / 32), block(32, 8);
    if (u8) to_chain_major_kernel<uint8_t><<<grid, block, 0, st>>>((const uint8_t*)ptr, (uint8_t*)dst, v.C, v.Cp, d);
    else to_chain_major_kernel<double><<<grid, block, 0, st>>>((const double*)ptr, (double*)dst, v.C, v.Cp, d);
    e->launches++;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int occu_reset(occu_engine_t* e, void* stream) {
    if (!e) return set_err("null engine");
    cudaStream_t st = (cudaStream_t)stream;
    const View& v = e->v;
    const size_t total = (size_t)v.n * v.Cp;
    reset_z_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(v);
    set_control_kernel<<<1, 1, 0, st>>>(v.ctl, 0u, 1, -1, 1);
    e->slices_valid = false;
    e->omega_in_slices = false;
    CUDA_OK(cudaMemsetAsync(v.status, 0, (size_t)v.Cp * 4, st));
    e->launches += 2;
    CUDA_OK(cudaGetLastError());
    return 0;
}

static int read_fails(occu_engine* e, cudaStream_t st) {
    Control h;
    CUDA_OK(cudaMemcpyAsync(&h, e->v.ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    return h.fails;
}

int occu_sweep_enqueue(occu_engine_t* e, int n_iters, double* rec_alpha, double* rec_beta, double* rec_tau,
                       int rec_start, int rec_capacity, void* stream) {
    if (!e) return set_err("null engine");
    if (n_iters < 0) return set_err("n_iters < 0");
    if (rec_alpha && (!rec_beta || !rec_tau || rec_start < 0 || rec_start + n_iters > rec_capacity))
        return set_err("bad record buffers");
    cudaStream_t st = (cudaStream_t)stream;
    const View& v = e->v;
    // rec_index lives on the device so that the iteration body is launch-parameter free
    set_rec_index_kernel<<<1, 1, 0, st>>>(v.ctl, rec_alpha ? rec_start : -1);
    e->launches++;
    SolveDebug dbg;
    memset(&dbg, 0, sizeof(dbg));
    SiteDebug sdbg;
    memset(&sdbg, 0, sizeof(sdbg));
    for (int it = 0; it < n_iters; it++) {
        PROF_MARK(e, KC_PSI_GEMM, st);
        if (!fuse_psi(e) && launch_psi_gemm(e, st)) return -1;   // fused into the site pass otherwise
        PROF_MARK(e, KC_SITE_PASS, st);
        launch_site_pass(e, sdbg, st);
        if (launch_gaussian_half(e, dbg, rec_alpha, rec_beta, rec_tau, rec_capacity, st)) return -1;
        advance_kernel<<<1, 1, 0, st>>>(v.ctl);
        e->launches++;
        if (prof_collect(e, st)) return -1;
    }
    CUDA_OK(cudaGetLastError());
    return 0;
}

int occu_finish(occu_engine_t* e, void* stream) {
    if (!e) return set_err("null engine");
    return read_fails(e, (cudaStream_t)stream);
}

int occu_sweep(occu_engine_t* e, int n_iters, double* rec_alpha, double* rec_beta, double* rec_tau, int rec_start,
               int rec_capacity, void* stream) {
    if (occu_sweep_enqueue(e, n_iters, rec_alpha, rec_beta, rec_tau, rec_start, rec_capacity, stream)) return -1;
    return occu_finish(e, stream);
}

int occu_debug_gaussian_half(occu_engine_t* e, const double* normals, const double* tau_in, const double* omega_a,
                             const occu_debug_out_t* out, void* stream) {
    if (!e) return set_err("null engine");
    if (!e->debug) return set_err("engine was not created with debug = 1");
    cudaStream_t st = (cudaStream_t)stream;
    SolveDebug dbg;
    memset(&dbg, 0, sizeof(dbg));
    dbg.normals = normals;
    dbg.tau_in = tau_in;
    dbg.enabled = out ? 1 : 0;
    if (out) dbg.out = *out;
    if (omega_a) launch_alpha_injected(e, omega_a, st);
    if (launch_gaussian_half(e, dbg, nullptr, nullptr, nullptr, 0, st)) return -1;
    return read_fails(e, st);
}

int occu_debug_site_half(occu_engine_t* e, int keep_z, const double* u_site, const double* u_visit,
                         const double* nrm_eps, void* stream) {
    if (!e) return set_err("null engine");
    if (!e->debug) return set_err("engine was not created with debug = 1");
    cudaStream_t st = (cudaStream_t)stream;
    SiteDebug sdbg;
    sdbg.u_site = u_site;
    sdbg.u_visit = u_visit;
    sdbg.nrm_eps = nrm_eps;
    sdbg.keep_z = keep_z;
    // force the z-update branch even on a fresh engine
    Control h;
    CUDA_OK(cudaMemcpyAsync(&h, e->v.ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    set_control_kernel<<<1, 1, 0, st>>>(e->v.ctl, h.iter, keep_z ? 1 : 0, -1, 0);
    e->launches++;
    if (launch_psi_gemm(e, st)) return -1;
    launch_site_pass(e, sdbg, st);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int occu_pg1(const double* c, size_t count, uint64_t seed, int fp64, double* out, void* stream) {
    if (!c || !out) return set_err("null argument");
    if (!count) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)((count + 255) / 256);
    PhiloxKeys rk;
    philox_round_keys(seed, rk);
    if (fp64) pg1_kernel<double><<<blocks, 256, 0, st>>>(c, count, seed, rk, out);
    else pg1_kernel<float><<<blocks, 256, 0, st>>>(c, count, seed, rk, out);
    CUDA_OK(cudaGetLastError());
    return 0;
}

int occu_truncnorm(const double* loc, const uint8_t* positive, const double* u, size_t count, uint64_t seed, int fp64,
                   double* out, void* stream) {
    if (!loc || !positive || !out) return set_err("null argument");
    if (!count) return 0;
    truncnorm_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(loc, positive, u, count, seed,
                                                                                         fp64, out);
    CUDA_OK(cudaGetLastError());
    return 0;
}

int occu_gamma(const double* shape, const double* rate, size_t count, uint64_t seed, double* out, void* stream) {
    if (!shape || !rate || !out) return set_err("null argument");
    if (!count) return 0;
    gamma_kernel<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(shape, rate, count, seed, out);
    CUDA_OK(cudaGetLastError());
    return 0;
}

int occu_precision_mvnorm_batched(int d, int count, const double* b, double* P, const double* z, uint64_t seed, double* x,
                                  void* stream) {
    if (!b || !P || !x) return set_err("null argument");
    if (d < 1 || d > 224) return set_err("d must be in [1, 224]");
    if (count < 1) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = ((size_t)d * (d + 1) / 2 + d) * sizeof(double);
    CUDA_OK(cudaFuncSetAttribute(precision_mvnorm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int* fails = nullptr;
    CUDA_OK(cudaMalloc(&fails, sizeof(int)));  // test hook only; the sweep never allocates
    CUDA_OK(cudaMemsetAsync(fails, 0, sizeof(int), st));
    precision_mvnorm_kernel<<<count, SOLVE_THREADS, smem, st>>>(d, b, P, z, seed, x, fails);
    int h = 0;
    cudaError_t err = cudaMemcpyAsync(&h, fails, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (err == cudaSuccess) err = cudaStreamSynchronize(st);
    cudaFree(fails);
    CUDA_OK(err);
    if (h) set_err("Cholesky factorization/solver failed!");
    return h;
}

int occu_ensure_sums_to_zero_batched(int d, int count, const double* x, const double* z, double* out, void* stream) {
    if (!x || !z || !out) return set_err("null argument");
    if (count < 1 || d < 1) return 0;
    sums_to_zero_kernel<<<count, 32, 0, (cudaStream_t)stream>>>(d, x, z, out);
    CUDA_OK(cudaGetLastError());
    return 0;
}

int occu_gemm_tn(const double* A, const double* B, double* Out, int M, int N, int K, int lda, int ldb, int ldo,
                 void* stream) {
    if (!A || !B || !Out) return set_err("null argument");
    if (M % dgemm::BM || N % dgemm::BN || K % dgemm::BK)
        return set_err("M must be a multiple of 128, N of 64 and K of 16");
    dgemm::Args a;
    memset(&a, 0, sizeof(a));
    a.A1 = A; a.B = B; a.Out = Out;
    a.M = M; a.N = N; a.K = K; a.lda = lda; a.ldb = ldb; a.ldo = ldo;
    CUDA_OK(dgemm::launch(a, (cudaStream_t)stream));
    return 0;
}

int occu_chain_summaries(const double* rec, int chains, int draws, int dim, int max_lag, double* out, void* stream) {
    if (!rec || !out) return set_err("null argument");
    if (chains < 1 || draws < 2 || dim < 1 || max_lag < 0 || max_lag >= draws) return set_err("bad summary dimensions");
    chain_summaries_kernel<<<chains * dim, 128, 0, (cudaStream_t)stream>>>(rec, chains, draws, dim, max_lag, out);
    CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
