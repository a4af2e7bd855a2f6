// engine_types.cuh -- constants, error plumbing, the device view of an engine and the workspace layout.
// Part of the single translation unit engine.cu (included there, in order; not a stand-alone header).
#pragma once

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
constexpr int ICAR_NZ = 15;            // compact list of a band row's non-zero offsets (queen lattice: 4)
constexpr int OCCU_ICAR_MAX_BW = 200;  // 26 x 27 / 2 tiles of 512 B = 180 KB of shared memory
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
    // default start values (occu_draw_start): prior means and the lower Cholesky factors of 100 a_prec / 100 b_prec
    // (base.py:199-211 draws alpha, beta ~ N(mu, 100 prec): the precision used as a covariance), row-major
    double a_mu[MAX_SMALL], b_mu[MAX_SMALL];
    double a_chol[MAX_SMALL * MAX_SMALL], b_chol[MAX_SMALL * MAX_SMALL];
    int start_ok;   // both factorisations succeeded
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
    int tc_group_tiles;     // 0: the Z planes are stored; > 0: streamed, this many 64-column tiles per group (large q)
    int tc_cap_rows;        // rows per plane of the B operand buffer (= tc_Mrows when stored)
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
    uint8_t* qnz;         // [n][ICAR_NZ + 1]  count, then the offsets d >= 1 of the non-zeros of row i of the band (the
                          //                   quadratic form eta' Q eta walks ~4 of them instead of bw); count 255: more than ICAR_NZ
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
    int omega_planes;       // omega byte planes that can be non-zero: SA_F32 after a production (FP32) site pass, else SA
    bool low_planes_dirty;  // planes SA_F32 .. SA - 1 hold digits of an older 48-bit omega
    bool omega_in_slices;   // ... and v.omega was NOT written by the last site pass (non-debug engines; probit: omega_b = d + eps)
    int profiling;
    bool start_ok;          // the Cholesky factors of the default start values exist (HParams::a_chol / b_chol)
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
constexpr int TILED_QMAX = 3072;     // the per-chain vectors (5 q doubles) and 9 panel slots stay within the 227 KB of shared memory
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
    // the window of the banded factorisation (8 x 8 tiles over bw + 8 rows, solve_icar_warp.cuh) lives in shared memory
    if (icar && P->icar_bw > OCCU_ICAR_MAX_BW)
        return set_err("ICAR half bandwidth too large for the shared-memory window (200 in this build; reorder the sites)");
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
        // Z planes: SB x Mrows x Kp bytes.  Kept when they fit OCCU_TC_STORE_GB (default 16 GiB: q <= ~700 at n = 10^4);
        // beyond that they are re-formed every iteration in groups of column tiles through a buffer of OCCU_TC_GROUP_GB
        // (default 2 GiB) -- the operand of the reference's default call at n = 10^4 (q = 1256) would be 48 GB
        const char* sg = getenv("OCCU_TC_STORE_GB");
        const char* gg = getenv("OCCU_TC_GROUP_GB");
        const double store_max = (sg ? atof(sg) : 16.0) * 1073741824.0, group_max = (gg ? atof(gg) : 2.0) * 1073741824.0;
        const double plane_bytes = (double)gtc::SB * g.Mrows * g.Kp;
        v.tc_group_tiles = 0;
        v.tc_cap_rows = g.Mrows;
        if (plane_bytes > store_max) {
            long long tiles = (long long)(group_max / ((double)gtc::SB * g.Kp * TC_BN)) - g.n_ht;
            if (tiles < 8) tiles = 8;
            if (tiles > g.n_gt) tiles = g.n_gt;
            v.tc_group_tiles = (int)tiles;
            v.tc_cap_rows = ((int)tiles + g.n_ht) * TC_BN;
        }
    }
    if (P->model == OCCU_PROBIT_RSR) {
        // only the K-split product U'd runs.  The split depends on n alone (never on the number of chains): it fixes the
        // FP64 summation order, and a chain's draws must not change with the shard it runs in.  >= 32 K tiles per split,
        // at most 16 splits: n = 10^4 gives 16, so that even one 128-chain tile spreads over 32 CTAs
        const int kt = v.n_pad / dgemm::BK;
        const int ks = kt / 32;
        v.ksplit = ks < 1 ? 1 : (ks > 16 ? 16 : ks);
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
            return set_err("RSR rank q too large (q <= 3072 in this build)");
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
    v.qnz = icar ? c.take<uint8_t>((size_t)v.n * (ICAR_NZ + 1)) : nullptr;
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
        v.Bq = c.take<int8_t>((size_t)gtc::SB * v.tc_cap_rows * v.tc_Kp);
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
