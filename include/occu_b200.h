/*
 * occu_b200.h -- C ABI of the B200-native Gibbs-sweep engine for OccuSpytial's samplers.
 *
 * Drop-in boundary: the reference has no FFI of its own; its native layer is two
 * Cython modules (occuspytial/distributions.pyx, occuspytial/data.pyx) called from the pure-Python
 * sweep (occuspytial/gibbs/{base,logit,probit}.py).  This library replaces that whole L2+L1 layer:
 * the Python host (occuspytial_b200/gibbs/) keeps the reference's sampler API and calls the entry
 * points below through ctypes.  Each entry point cites the reference interface it replaces.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer (e.g. torch.Tensor.data_ptr()) unless the
 *     parameter is documented as HOST;
 *   - "chain-major" arrays are [chains][dim] row-major float64 (what NumPy/torch hand over);
 *     the engine keeps its own chain-minor copies inside the caller-provided workspace;
 *   - return value: 0 = OK, < 0 = argument / CUDA error (see occu_last_error()),
 *     > 0 = number of chains that hit a numerical failure (non-positive-definite pivot, NaN);
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are asynchronous
 *     unless stated;
 *   - the library owns nothing but the small host-side handle: all device memory is the caller's.
 */
#ifndef OCCU_B200_H
#define OCCU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OCCU_ABI_VERSION 2

/* model ids */
#define OCCU_LOGIT_RSR 0  /* LogitRSRGibbs   occuspytial/gibbs/logit.py:340-485 */
#define OCCU_PROBIT_RSR 1 /* ProbitRSRGibbs  occuspytial/gibbs/probit.py:27-270 */
#define OCCU_LOGIT_ICAR 2 /* LogitICARGibbs  occuspytial/gibbs/logit.py:102-266 ("PGGibbs") */

/* site classes (partition built once by GibbsBase._configure, base.py:125-137) */
#define OCCU_SITE_NOT_SURVEYED 0
#define OCCU_SITE_NOT_OBS 1 /* surveyed, species never detected */
#define OCCU_SITE_OBS 2     /* surveyed, detected at least once => z = 1 */

/* per-chain state fields for occu_set_field / occu_get_field (chain-major on the caller's side) */
#define OCCU_F_ALPHA 0 /* [chains][r]  f64 */
#define OCCU_F_BETA 1  /* [chains][p]  f64 */
#define OCCU_F_ETA 2   /* [chains][q]  f64 (RSR) or [chains][n] (ICAR) */
#define OCCU_F_TAU 3   /* [chains]     f64 */
#define OCCU_F_Z 4     /* [chains][n]  u8  */
#define OCCU_F_OMEGA 5 /* [chains][n]  f64  omega_b */
#define OCCU_F_PSI 6   /* [chains][n]  f64  X beta + spatial (read-only) */
#define OCCU_F_EPS 7   /* [chains][n]  f64  probit only */
#define OCCU_F_ZPROB 8 /* [chains][n]  f64  last occupancy probabilities (debug engines only, read-only) */

typedef struct occu_engine occu_engine_t;

/*
 * Problem description = everything GibbsBase.__init__/_configure (base.py:84-164),
 * LogitRSRGibbs._configure_rsr (logit.py:415-460) and ProbitRSRGibbs._configure (probit.py:109-153)
 * leave in `self.fixed`, with the ragged `Data` dictionaries (data.pyx:34-147) flattened to CSR.
 */
typedef struct {
    int32_t abi_version;  /* OCCU_ABI_VERSION */
    int32_t model;        /* OCCU_LOGIT_RSR | OCCU_PROBIT_RSR | OCCU_LOGIT_ICAR */
    int32_t n;            /* sites */
    int32_t p;            /* occupancy covariates (columns of X) */
    int32_t r;            /* detection covariates (columns of W) */
    int32_t q;            /* RSR rank (columns of K); 0 for ICAR */
    int32_t V;            /* total visit rows */
    int32_t chains;       /* chains advanced in lock-step on this device */
    int32_t chain_offset; /* global id of this device's first chain (Philox stream id) */
    int32_t debug;        /* 1: keep occupancy probabilities / allow occu_debug_* */
    int32_t pg_fp64;      /* 1: PG / truncated-normal sampler transcendental work in FP64 (validation), 0: FP32 */
    int32_t icar_bw;      /* ICAR: half bandwidth of Q in the given site order; else 0 */
    int32_t gram_int8;    /* LogitRSR: 1 = Omega-weighted Gram on tcgen05 kind::i8 by exact integer slicing (gram_tc.cuh), 0 = FP64 DMMA */
    uint64_t seed;        /* Philox key (random_state of the sampler) */
    const double* X;      /* [n][p] */
    const double* K;      /* [n][q] RSR basis (fixed.K), NULL for ICAR */
    const double* Qrsr;   /* [q][q] K'QK (fixed.Q after _configure_rsr), NULL for ICAR */
    const double* Qband;  /* ICAR: [n][icar_bw + 1] lower band of Q, Qband[i][d] = Q[i][i - d]; else NULL */
    const double* W;      /* [V][r] visit rows grouped by site, sites in id order */
    const uint8_t* y;     /* [V] detections */
    const int32_t* visit_ptr;  /* [n + 1] CSR offsets into W / y (empty range for unsurveyed sites) */
    const uint8_t* site_class; /* [n] OCCU_SITE_* */
    double tau_rate;      /* fixed.tau_rate */
    double tau_shape;     /* fixed.tau_shape */
    const double* a_mu;   /* HOST [r] */
    const double* a_prec; /* HOST [r][r] */
    const double* b_mu;   /* HOST [p] */
    const double* b_prec; /* HOST [p][p] */
} occu_problem_t;

/* Debug outputs of one conditional-Gaussian stage, all chain-major device pointers, any may be NULL. */
typedef struct {
    double* eta_prec;  /* [chains][q][q] Lambda of eta | .  (logit.py:335 / probit.py:226) */
    double* eta_b;     /* [chains][q]    b */
    double* eta_chol;  /* [chains][q][q] lower Cholesky factor of Lambda (upper part zero) */
    double* eta_mean;  /* [chains][q]    Lambda^-1 b */
    double* beta_prec; /* [chains][p][p] (logit.py:230 / probit.py:112) */
    double* beta_b;    /* [chains][p]    (logit.py:231 / probit.py:238-240) */
    double* alpha_prec;/* [chains][r][r] (logit.py:222 / probit.py:233) */
    double* alpha_b;   /* [chains][r]    (logit.py:223 / probit.py:234) */
    double* tau_rate;  /* [chains]       rate of the Gamma conditional (logit.py:208) */
} occu_debug_out_t;

const char* occu_last_error(void);
int occu_abi_version(void);

/* Bytes of device workspace the engine needs for this problem (HOST call, no CUDA work). */
size_t occu_workspace_bytes(const occu_problem_t* prob);

/*
 * Build an engine inside `workspace` (device memory, >= occu_workspace_bytes, 256-byte aligned).
 * Copies the problem arrays into the workspace, builds the face-splitting design Z and the
 * initial z (base.py:113-118).  Replaces GibbsBase._configure's per-sampler state setup.
 */
int occu_create(const occu_problem_t* prob, void* workspace, size_t workspace_bytes, void* stream,
                occu_engine_t** out);
void occu_destroy(occu_engine_t* eng);

/* Start values (GibbsBase._initialize_posterior_state, base.py:188-197; logit.py:468-476; probit.py:162-171). */
int occu_set_field(occu_engine_t* eng, int field, const void* src_chain_major, void* stream);
int occu_get_field(occu_engine_t* eng, int field, void* dst_chain_major, void* stream);
/* One chain's copy of a field, [dim]: what the reference exposes as `sampler.state.<name>` after a run (base.py:15-27,
 * 236-239) without gathering the other chains. */
int occu_get_field_chain(occu_engine_t* eng, int field, int chain, void* dst, void* stream);

/* Re-arm the engine at iteration 0 with z reset from the data (start of GibbsBase._run, base.py:226). */
int occu_reset(occu_engine_t* eng, void* stream);

/*
 * Default start values of every chain, drawn on the device (GibbsBase._initialize_default_start, base.py:199-211, with the
 * overrides of logit.py:462-466 and probit.py:155-160), followed by occu_reset:
 *   tau ~ Gamma(1/2, scale 1 / tau_rate); alpha ~ N(a_mu, 100 a_prec); beta ~ N(b_mu, 100 b_prec) (Cholesky method);
 *   eta ~ eta_scale * N(0, I) (RSR samplers: 5), minus its mean over the sites when centre_eta != 0 (ICAR sampler);
 *   probit: eps ~ N(0, I_n).
 * The draws are Philox blocks addressed by (seed; global chain id, element, field): a chain's start does not depend on
 * how the job is sharded.  Fails if 100 a_prec or 100 b_prec is not positive definite.
 */
int occu_draw_start(occu_engine_t* eng, uint64_t seed, double eta_scale, int centre_eta, void* stream);

/*
 * Advance all chains by `n_iters` full Gibbs iterations (the loop of GibbsBase._run, base.py:236-239,
 * with <Sampler>.step() = logit.py:254-266 / probit.py:262-270 as the body).  When `rec_alpha` is not
 * NULL, iteration j of this call is stored at draw index rec_start + j (Chain.append, chain.py:56-74):
 *   rec_alpha [chains][rec_capacity][r], rec_beta [chains][rec_capacity][p], rec_tau [chains][rec_capacity].
 * Synchronises the stream before returning; returns the number of failed chains.
 */
int occu_sweep(occu_engine_t* eng, int n_iters, double* rec_alpha, double* rec_beta, double* rec_tau,
               int rec_start, int rec_capacity, void* stream);

/*
 * The two halves of occu_sweep: occu_sweep_enqueue only enqueues the kernels of `n_iters` iterations on `stream`
 * (no host synchronisation), occu_finish synchronises the stream and returns the number of failed chains.  This is
 * what sample() uses: every block of iterations is enqueued without waiting, the record block that has just filled
 * crosses PCIe on a second stream meanwhile, and engines on several devices (one per GPU, driven from host threads)
 * run concurrently.
 */
int occu_sweep_enqueue(occu_engine_t* eng, int n_iters, double* rec_alpha, double* rec_beta, double* rec_tau,
                       int rec_start, int rec_capacity, void* stream);
int occu_finish(occu_engine_t* eng, void* stream);

/* Number of kernels launched by this engine so far (bench.py's gpu_launches). */
long long occu_launch_count(const occu_engine_t* eng);

/*
 * Optional per-kernel-class timing with CUDA events on the engine's stream (bench.py's roofline figures).
 * Classes: 0 psi GEMM, 1 site pass, 2 Gram GEMM, 3 partial-sum reduction, 4 per-chain solve.
 * Profiling mode synchronises once per iteration; leave it off for throughput runs.
 * occu_get_kernel_ms fills accumulated milliseconds and launch counts, returns the number of classes.
 */
int occu_set_profiling(occu_engine_t* eng, int on);
int occu_get_kernel_ms(const occu_engine_t* eng, double* ms, long long* calls, int count);

/* ---- deterministic-piece hooks (parity tests; need prob.debug = 1) ---------------------------- */

/*
 * Run the conditional-Gaussian half of one iteration (tau, eta, beta, alpha updates) from the engine's
 * current omega_b / z / theta with injected randomness:
 *   normals  [chains][q + p + r]  N(0,1) draws used for eta, beta, alpha in that order (NULL: Philox)
 *   tau_in   [chains]             tau to use instead of a Gamma draw (NULL: draw)
 * and dump every intermediate system.  Replaces for the tests: _update_tau/_update_eta/_update_beta/
 * _update_alpha (logit.py:206-232, 478-485; probit.py:211-243) + precision_mvnorm (distributions.pyx:42-110).
 * alpha's system needs omega_a: `omega_a` [chains][V] (NULL: draw PG from the current alpha; rows of sites
 * with z = 0 are ignored).
 */
int occu_debug_gaussian_half(occu_engine_t* eng, const double* normals, const double* tau_in,
                             const double* omega_a, const occu_debug_out_t* out, void* stream);

/*
 * Run psi = X beta + K eta and the site half of one iteration: the z-update (logit.py:234-252 /
 * probit.py:245-260) unless keep_z != 0, the omega_a / omega_b draws and -- probit -- the eps update.
 * Afterwards OCCU_F_ZPROB holds the occupancy probabilities (keep_z == 0), OCCU_F_Z / OCCU_F_OMEGA /
 * OCCU_F_EPS the new draws.  Probit only, chain-major, NULL = Philox: u_site [chains][n] and u_visit
 * [chains][V] inject the uniforms of the reference's inverse-CDF truncated normals (probit.py:190-194,
 * 205-209), nrm_eps [chains][n] the normals of _update_eps (probit.py:220).
 */
int occu_debug_site_half(occu_engine_t* eng, int keep_z, const double* u_site, const double* u_visit,
                         const double* nrm_eps, void* stream);

/* ---- stand-alone batched primitives ----------------------------------------------------------- */

/* out[i] ~ PG(1, c[i]); element i uses Philox stream (seed, i).  Replaces polyagamma.random_polyagamma(1, c)
 * (logit.py:191-193, 202-204). */
int occu_pg1(const double* c, size_t count, uint64_t seed, int fp64, double* out, void* stream);

/* out[i] ~ N(loc[i], 1) truncated to (0, inf) if positive[i] else (-inf, 0).  `u` (may be NULL) injects the
 * uniforms of the reference's inverse-CDF formula (always FP64); otherwise Philox draws with the transcendental
 * work in FP64 (fp64 != 0, validation) or FP32 (production).  Replaces truncnorm_inf_ppf / truncnorm_neginf_ppf
 * (probit.py:11-24). */
int occu_truncnorm(const double* loc, const uint8_t* positive, const double* u, size_t count, uint64_t seed, int fp64,
                   double* out, void* stream);

/* out[i] ~ Gamma(shape[i], scale = 1 / rate[i])  (rng.gamma at logit.py:209 / probit.py:214). */
int occu_gamma(const double* shape, const double* rate, size_t count, uint64_t seed, double* out, void* stream);

/*
 * Batched precision_mvnorm (distributions.pyx:42-110): for each of `count` systems, x = P^-1 b + L^-T z with
 * P = L L'.  b [count][d], P [count][d][d] (lower triangle overwritten with L, as the reference overwrites
 * `prec`), z [count][d] injected normals (NULL: Philox(seed, system)), x [count][d].  d <= 224.
 */
int occu_precision_mvnorm_batched(int d, int count, const double* b, double* P, const double* z, uint64_t seed,
                                  double* x, void* stream);

/* ensure_sums_to_zero (distributions.pyx:24-39), batched: out = x - (sum x / sum z) z per row of [count][d]. */
int occu_ensure_sums_to_zero_batched(int d, int count, const double* x, const double* z, double* out,
                                     void* stream);

/* Out[m][n] = sum_k A[k][m] B[k][n] on the FP64 tensor pipe; M multiple of 128, N of 64, K of 16. */
int occu_gemm_tn(const double* A, const double* B, double* Out, int M, int N, int K, int lda, int ldb, int ldo,
                 void* stream);

/* ---- chain summaries (R-hat / ESS inputs; the only multi-GPU exchange is an all-gather of these) ---- */

/* From recorded draws rec [chains][draws][dim] compute per chain and component: mean, M2 (sum of squared
 * deviations) of the whole chain and of its two halves, and autocovariances at lags 0..max_lag:
 *   out [chains][dim][6 + max_lag + 1] = {mean, M2, mean_h1, M2_h1, mean_h2, M2_h2, acov_0..acov_max_lag}. */
int occu_chain_summaries(const double* rec, int chains, int draws, int dim, int max_lag, double* out,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OCCU_B200_H */
