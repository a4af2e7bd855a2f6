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
// The kernels live in the headers below (one translation unit; the order matters: later parts use earlier ones).
#include "engine_types.cuh"
#include "kernels_setup.cuh"
#include "site_pass.cuh"
#include "solve_rsr.cuh"
#include "solve_tiled.cuh"
#include "primitives.cuh"
#include "solve_icar.cuh"
#include "solve_icar_warp.cuh"

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
    // beyond ~128 columns of [X K] the in-pass DMMA product (no operand reuse across the 8-site sub-tiles) costs more
    // than the tiled GEMM plus the psi round trip (measured at q = 1256: 2.7 vs 1.3 ms per 1024 chains)
    // probit debug engines keep the separate FP64 GEMM: their pass reads psi in FP64 (production forms it in FP32)
    if (e->v.model == OCCU_PROBIT_RSR && e->debug) return false;
    return allowed && (e->v.model == OCCU_LOGIT_RSR || e->v.model == OCCU_PROBIT_RSR) && !e->pg_fp64 && e->v.pq <= 128;
}

template <typename PGT, bool SLICES, int PSI_MODE>
static void launch_site_pass_logit(occu_engine* e, cudaStream_t st) {
    const View& v = e->v;
    dim3 grid(v.n_tiles, v.Cp / CHAIN_TILE);
    switch (v.r) {
#define CASE(R)                                                                                                      \
    case R: {                                                                                                        \
        const size_t smem = site_pass_logit_smem<PGT, R>(PSI_MODE);                                                  \
        if (smem > 48 * 1024)                                                                                        \
            cudaFuncSetAttribute(site_pass_logit_kernel<R, PGT, SLICES, PSI_MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        site_pass_logit_kernel<R, PGT, SLICES, PSI_MODE><<<grid, CHAIN_TILE, smem, st>>>(v);                          \
        break;                                                                                                       \
    }
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    }
    e->launches++;
}

// psi inside the logit site pass: FP32 only in production engines, with an FP64 staging tile when the engine dumps
// the (FP64) occupancy probabilities
static int psi_mode(const occu_engine* e) { return !fuse_psi(e) ? PSI_NONE : (e->v.zprob ? PSI_F64 : PSI_F32); }

static void launch_site_pass(occu_engine* e, const SiteDebug& dbg, cudaStream_t st) {
    const View& v = e->v;
    if (v.model == OCCU_PROBIT_RSR) {
        dim3 grid(v.n_tiles, v.Cp / CHAIN_TILE);
        switch (v.r) {
#define CASE(R)                                                                                             \
    case R:                                                                                                 \
        if (e->pg_fp64) site_pass_probit_kernel<R, true, false><<<grid, CHAIN_TILE, 0, st>>>(v, dbg);       \
        else if (fuse_psi(e)) site_pass_probit_kernel<R, false, true><<<grid, CHAIN_TILE, 0, st>>>(v, dbg);  \
        else site_pass_probit_kernel<R, false, false><<<grid, CHAIN_TILE, 0, st>>>(v, dbg);                  \
        break;
            CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
        }
        e->launches++;
        e->omega_in_slices = v.zprob == nullptr;   // omega_b = d + eps, materialised on request (debug engines store it)
    } else if (v.int8) {
        // the site pass writes the digits of (omega, z) for the tcgen05 Gram itself
        if (!e->pg_fp64 && e->low_planes_dirty) {   // an injected 48-bit omega was sliced earlier: clear its low planes
            const size_t plane = (size_t)v.Cp * v.tc_Kp;
            cudaMemsetAsync(v.Aq + (size_t)gtc::SA_F32 * plane, 0, (size_t)(gtc::SA - gtc::SA_F32) * plane, st);
            e->low_planes_dirty = false;
        }
        if (e->pg_fp64) launch_site_pass_logit<double, true, PSI_NONE>(e, st);
        else if (psi_mode(e) == PSI_F32) launch_site_pass_logit<float, true, PSI_F32>(e, st);
        else if (psi_mode(e) == PSI_F64) launch_site_pass_logit<float, true, PSI_F64>(e, st);
        else launch_site_pass_logit<float, true, PSI_NONE>(e, st);
        e->slices_valid = true;
        e->omega_in_slices = v.zprob == nullptr;
        e->omega_planes = e->pg_fp64 ? gtc::SA : gtc::SA_F32;
    } else if (e->pg_fp64) {
        launch_site_pass_logit<double, false, PSI_NONE>(e, st);
    } else if (psi_mode(e) == PSI_F32) {
        launch_site_pass_logit<float, false, PSI_F32>(e, st);
    } else if (psi_mode(e) == PSI_F64) {
        launch_site_pass_logit<float, false, PSI_F64>(e, st);
    } else {
        launch_site_pass_logit<float, false, PSI_NONE>(e, st);
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
        e->omega_planes = gtc::SA;
        e->low_planes_dirty = true;
    }
    gtc::Params p;
    p.Cp = v.Cp; p.kblocks = g.kblocks; p.nsplit = g.nsplit; p.kb_per_chunk = g.kb_per_chunk; p.MT = g.MT;
    p.pq = v.pq; p.ldg = v.M_pad; p.ldh = v.ldh;
    p.hshift = v.hshift; p.err = v.tc_err; p.mode = 0;
    p.order = e->tc_order;
    p.Hpart = g.nsplit > 1 ? v.Hp : v.Hpart;
    // stored planes: one launch over all column tiles.  Streamed planes (large q): groups of column tiles, each sliced
    // into the operand buffer and consumed; the last group carries the U rows of the H tiles.
    const int group = v.tc_group_tiles ? v.tc_group_tiles : g.n_gt;
    for (int nt0 = 0; nt0 < g.n_gt; nt0 += group) {
        const int ntg = g.n_gt - nt0 < group ? g.n_gt - nt0 : group;
        const bool last = nt0 + ntg == g.n_gt;
        p.n_gt = ntg;
        p.n_ht = last ? g.n_ht : 0;
        p.Mrows = v.tc_cap_rows;
        const int col0 = nt0 * TC_BN;
        p.M = v.M - col0 < ntg * TC_BN ? v.M - col0 : ntg * TC_BN;
        p.scale = v.tscale + col0;
        p.Gpart = (g.nsplit > 1 ? v.Gp : v.G) + col0;
        if (v.tc_group_tiles) {
            const int rows = (ntg + p.n_ht) * TC_BN;
            gtc::slice_group_from_ut_kernel<<<rows, 256, 0, st>>>(v.UT, v.n, v.n_pad, g.Kp, v.M, g.Mg_pad, v.pq, col0, ntg * TC_BN,
                                                                  last ? ntg * TC_BN : -1, v.tc_cap_rows, v.tscale, v.Bq);
            e->launches++;
        }
        const int nitems = g.MT * (p.n_gt + p.n_ht) * g.nsplit;
        const int grid = nitems < e->sm_count ? nitems : e->sm_count;
        if (e->omega_planes == gtc::SA_F32)
            gtc::gram_tc_kernel<TC_BN, gtc::SA_F32><<<grid, gtc::THREADS, gtc::Cfg<TC_BN, gtc::SA_F32>::SMEM_BYTES, st>>>(e->tmA, e->tmB, p);
        else
            gtc::gram_tc_kernel<TC_BN, gtc::SA><<<grid, gtc::THREADS, gtc::Cfg<TC_BN, gtc::SA>::SMEM_BYTES, st>>>(e->tmA, e->tmB, p);
        e->launches++;
    }
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
    a.n2_valid = v.pq;
    CUDA_OK(dgemm::launch(a, st));
    e->launches++;
    return 0;
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
        {
            const size_t smem_w = icarw_smem_bytes(v.bw, v.p);
            size_t& configured_w = e->smem_configured[0];   // per engine: the attribute is per device
            const int nt = icarw_nt(v.bw);
            const bool grow = smem_w > configured_w;
            switch (nt <= 8 ? nt : 0) {
#define CASE(NT)                                                                                                            \
    case NT:                                                                                                                \
        if (grow) CUDA_OK(cudaFuncSetAttribute(solve_logit_icar_warp_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w)); \
        solve_logit_icar_warp_kernel<NT><<<v.Cp, ICARW_THREADS, smem_w, st>>>(v, dbg, rb, rt, cap);                          \
        break;
                CASE(2) CASE(4) CASE(6) CASE(8)
                default:
                    if (grow) CUDA_OK(cudaFuncSetAttribute(solve_logit_icar_warp_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
                    solve_logit_icar_warp_kernel<0><<<v.Cp, ICARW_THREADS, smem_w, st>>>(v, dbg, rb, rt, cap);
#undef CASE
            }
            if (grow) configured_w = smem_w;
            e->launches++;
            PROF_MARK(e, KC_COUNT, st);
            CUDA_OK(cudaGetLastError());
            return 0;
        }
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

template <typename T>
__global__ void gather_chain_kernel(const T* __restrict__ src, T* __restrict__ dst, int chain, int Cp, int d) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < d) dst[e] = src[(size_t)e * Cp + chain];
}

// lower Cholesky factor of scale * A (d x d, row-major) on the host; false if the matrix is not positive definite
static bool host_cholesky_scaled(const double* A, int d, double scale, double* L) {
    for (int i = 0; i < d; i++)
        for (int j = 0; j <= i; j++) {
            double s = scale * A[i * d + j];
            for (int k = 0; k < j; k++) s -= L[i * d + k] * L[j * d + k];
            if (i == j) {
                if (!(s > 0.0)) return false;
                L[i * d + i] = sqrt(s);
            } else {
                L[i * d + j] = s / L[j * d + j];
            }
        }
    return true;
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
    e->start_ok = false;
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
    for (int a = 0; a < v.r; a++) h.a_mu[a] = P->a_mu[a];
    for (int a = 0; a < v.p; a++) h.b_mu[a] = P->b_mu[a];
    h.start_ok = host_cholesky_scaled(P->a_prec, v.r, 100.0, h.a_chol) && host_cholesky_scaled(P->b_prec, v.p, 100.0, h.b_chol);
    e->start_ok = h.start_ok != 0;
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
        icar_nonzeros_kernel<<<(v.n + 127) / 128, 128, 0, st>>>(v.Qband, v.n, v.bw + 1, v.qnz);
        e->launches++;
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
    e->omega_planes = gtc::SA;
    e->low_planes_dirty = false;
    if (v.int8) {
        // slice [Z | U] once (formed on the fly from U'), encode the TMA descriptors of the two operand arrays
        e->geo = gtc::make_geometry(v.n, v.Cp, v.M, v.pq, TC_BN);
        const gtc::Geometry& g = e->geo;
        if (v.tc_group_tiles)   // streamed planes: only the row scales are computed here
            gtc::row_scales_from_ut_kernel<<<g.Mrows, 256, 0, st>>>(v.UT, v.n, v.n_pad, v.M, g.Mg_pad, v.pq, v.tscale);
        else
            gtc::slice_rows_from_ut_kernel<<<g.Mrows, 256, 0, st>>>(v.UT, v.n, v.n_pad, g.Kp, v.M, g.Mg_pad, v.pq, g.Mrows, v.tscale,
                                                                     v.Bq);
        gtc::hshift_from_ut_kernel<<<v.ldh, 128, 0, st>>>(v.UT, v.n, v.n_pad, v.pq, v.ldh, v.hshift);
        e->launches += 2;
        FAIL_IF(cudaMemsetAsync(v.Aq, 0, (size_t)(gtc::SA + 1) * v.Cp * g.Kp, st));
        FAIL_IF(cudaMemsetAsync(v.G, 0, (size_t)v.Cp * v.M_pad * 8, st));
        FAIL_IF(cudaMemsetAsync(v.tc_err, 0, sizeof(int), st));
        const char* why = gtc::make_map(&e->tmA, v.Aq, (uint64_t)(gtc::SA + 1) * v.Cp, g.Kp, gtc::BM);
        if (!why) why = gtc::make_map(&e->tmB, v.Bq, (uint64_t)gtc::SB * v.tc_cap_rows, g.Kp, TC_BN);
        if (!why && (cudaFuncSetAttribute(gtc::gram_tc_kernel<TC_BN, gtc::SA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          gtc::Cfg<TC_BN>::SMEM_BYTES) != cudaSuccess ||
                     cudaFuncSetAttribute(gtc::gram_tc_kernel<TC_BN, gtc::SA_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          gtc::Cfg<TC_BN, gtc::SA_F32>::SMEM_BYTES) != cudaSuccess))
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

// probit production engines do not store omega_b: the site pass leaves d = omega_b - eps and eps
__global__ void omega_from_d_eps_kernel(View v) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (size_t)v.n * v.Cp) v.omega[idx] = v.kd[idx] + v.eps[idx];
}

static int materialise_omega(occu_engine* e, cudaStream_t st) {
    if (e->v.model == OCCU_PROBIT_RSR && e->omega_in_slices) {
        const size_t total = (size_t)e->v.n * e->v.Cp;
        omega_from_d_eps_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(e->v);
        e->launches++;
        e->omega_in_slices = false;
        CUDA_OK(cudaGetLastError());
        return 0;
    }
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

int occu_get_field_chain(occu_engine_t* e, int field, int chain, void* dst, void* stream) {
    if (!e || !dst) return set_err("null argument");
    if (chain < 0 || chain >= e->v.C) return set_err("chain index out of range");
    cudaStream_t st = (cudaStream_t)stream;
    void* ptr; int d, u8;
    if (field_info(e, field, &ptr, &d, &u8)) return -1;
    if (field == OCCU_F_OMEGA && materialise_omega(e, st)) return -1;
    if (field == OCCU_F_PSI && fuse_psi(e) && !e->debug && launch_psi_gemm(e, st)) return -1;
    const View& v = e->v;
    if (u8) gather_chain_kernel<uint8_t><<<(d + 255) / 256, 256, 0, st>>>((const uint8_t*)ptr, (uint8_t*)dst, chain, v.Cp, d);
    else gather_chain_kernel<double><<<(d + 255) / 256, 256, 0, st>>>((const double*)ptr, (double*)dst, chain, v.Cp, d);
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

int occu_draw_start(occu_engine_t* e, uint64_t seed, double eta_scale, int centre_eta, void* stream) {
    if (!e) return set_err("null engine");
    if (!e->start_ok) return set_err("100 * a_prec or 100 * b_prec is not positive definite: no default start values");
    if (occu_reset(e, stream)) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    const View& v = e->v;
    void* eta; int d, u8;
    if (field_info(e, OCCU_F_ETA, &eta, &d, &u8)) return -1;
    draw_start_small_kernel<<<(v.Cp + 127) / 128, 128, 0, st>>>(v, seed);
    const size_t total = (size_t)d * v.Cp;
    draw_start_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(v, seed, (double*)eta, d, eta_scale, START_ETA);
    e->launches += 2;
    if (centre_eta) {
        centre_rows_kernel<<<(v.Cp + 31) / 32, dim3(32, 8), 0, st>>>((double*)eta, d, v.Cp);
        e->launches++;
    }
    if (v.eps) {
        const size_t tot = (size_t)v.n * v.Cp;
        draw_start_rows_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(v, seed, v.eps, v.n, 1.0, START_EPS);
        e->launches++;
    }
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
