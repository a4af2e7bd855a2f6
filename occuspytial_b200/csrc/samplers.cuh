// samplers.cuh -- device samplers for the augmentation variables and tau.
//
//  pg1_draw<T>      PG(1, c)  -- replaces polyagamma.random_polyagamma(1, c) called at
//                   occuspytial/gibbs/logit.py:191-193 and 202-204 (third-party, restated from
//                   Polson, Scott & Windle 2013, Alg. 1: Devroye's alternating-series sampler of J*(1, c/2)).
//  truncnorm_draw   one-sided unit-variance truncated normal, replaces truncnorm_inf_ppf /
//                   truncnorm_neginf_ppf (occuspytial/gibbs/probit.py:11-24) with a tail-safe branch.
//  gamma_draw       Gamma(shape, 1) by Marsaglia-Tsang, replaces rng.gamma (logit.py:209, probit.py:214).
#pragma once
#include "common.cuh"

#define PG_TRUNC 0.64  // truncation point t between the two series representations

// ---------------------------------------------------------------------------------------------
// PG(1, c) = J*(1, c/2) / 4 by Devroye's alternating-series method, arranged for SIMT execution.
//
// The textbook algorithm (Polson-Scott-Windle Alg. 1, as in polyagamma) nests three data-dependent
// loops: the outer accept loop, the truncated inverse-Gaussian sub-sampler and its own rejection loops.
// On a GPU that serialises lanes.  Here the sampler is flattened into i.i.d. ROUNDS with no inner loop:
// every round proposes X from an envelope g >= f and accepts or rejects; a rejected lane simply runs
// another round, an accepted lane is refilled with its next draw by the caller (site pass), so the
// lanes of a warp stay busy ("compaction by refill").  With z = |c|/2, K = pi^2/8 + z^2/2, t = 0.64:
//   right piece  x > t : g = (pi/2) exp(-K x),                    mass p  = (pi / 2K) exp(-K t),  X = t + Exp(1)/K
//   left, z t < 1      : g = a_0(x) (Levy(1) density x 2),          mass q' = 4 Phi(-1/sqrt t),     X = 1 / Y^2, Y ~ N(0,1) | |Y| > 1/sqrt t,
//                        pre-accept with probability exp(-z^2 X / 2)
//   left, z t >= 1     : g = a_0(x) exp(-z^2 x / 2) on (0, inf),    mass q''= 2 exp(-z),            X ~ IG(1/z, 1) (Michael-Schucany-Haas),
//                        pre-accept iff X <= t
// followed by the alternating-series squeeze on the ratios r_n = a_n(x)/a_0(x) = (2n+1) exp(-n(n+1) e),
// e = 2/x (x <= t) or pi^2 x/2 (x > t).  Restarting from the mixture choice after any rejection keeps the
// accepted draw exactly J*(1, z) distributed because each round is an independent rejection trial against
// the same envelope.  One round consumes exactly one Philox block (4 uniforms).
// ---------------------------------------------------------------------------------------------
template <typename T>
struct PGTask {
    T K, p_right, mu, hz2;
    bool small_z;
};

template <typename T>
__device__ __forceinline__ void pg1_setup(T c, PGTask<T>& s) {
    typedef RT<T> R;
    const T t = (T)PG_TRUNC;
    const T z = (T)0.5 * fabs(c);
    s.hz2 = (T)0.5 * z * z;
    s.K = (T)(OCCU_PI * OCCU_PI / 8.0) + s.hz2;
    s.small_z = z * t < (T)1;
    s.mu = s.small_z ? (T)1 : (T)1 / z;
    // q/p with q = q' (small z) or q'' (large z); exp may overflow to inf for huge z => p_right = 0
    const T two_k_over_pi = s.K * (T)(2.0 / OCCU_PI);
    const T qdivp = s.small_z ? (T)0.42259909466742100 * two_k_over_pi * R::exp_(s.K * t)  // 4 Phi(-1.25)
                              : (T)2 * two_k_over_pi * R::exp_(s.K * t - z);
    s.p_right = (T)1 / ((T)1 + qdivp);
}

// One rejection round.  u[0..3]: uniforms on (0,1).  Returns true and the J*(1,z) variate in x on acceptance.
template <typename T>
__device__ __forceinline__ bool pg1_round(const PGTask<T>& s, T ub, T ux, T u3, T us, T& x) {
    typedef RT<T> R;
    const T t = (T)PG_TRUNC;
    bool pre;
    if (ub < s.p_right) {
        x = t - R::log_(ux) / s.K;
        pre = true;
    } else {
        const T y = R::normcdfinv_(s.small_z ? ux * (T)0.10564977366685525 : ux);  // Phi(-1.25)
        const T y2 = y * y;
        if (s.small_z) {
            x = (T)1 / y2;
            pre = (x <= t) && (u3 <= R::exp_(-s.hz2 * x));
        } else {
            const T w = s.mu * y2;
            x = s.mu / ((T)1 + (T)0.5 * w + R::sqrt_(w + (T)0.25 * w * w));  // cancellation-free small root
            if (u3 * (s.mu + x) > s.mu) x = s.mu * s.mu / x;
            pre = x <= t;
        }
    }
    const T e = (x > t) ? (T)(OCCU_PI * OCCU_PI * 0.5) * x : (T)2 / x;
    T sum = (T)1;
    bool acc = false;
    for (int n = 1;; n++) {
        const T rn = (T)(2 * n + 1) * R::exp_(-(T)(n * (n + 1)) * e);
        if (n & 1) {
            sum -= rn;
            if (us <= sum) {
                acc = true;
                break;
            }
        } else {
            sum += rn;
            if (us > sum) break;
        }
    }
    return pre && acc;
}

// ---------------------------------------------------------------------------------------------
// FP32 production round.  Same envelope, same acceptance tests and the same use of the four uniforms as the
// generic round above, arranged for issue rate (the site pass is instruction bound):
//   * one log serves all three proposals: L = -log(ux) is the Exp(1) variate of the right piece, the tail
//     quantile argument of the left piece for small z (p = ux Phi(-1.25) => -log p = L + 2.2476) and the
//     Box-Muller radius of the chi-square(1) variate of the inverse-Gaussian proposal (with u3 as the angle; the
//     root choice then takes the residual of ub, which is U(0,1) given ub >= p_right);
//   * the tail quantile y = -Phi^-1(p), p < Phi(-1.25), is a (3,2) rational in sqrt(-2 log p) (max relative error
//     4.4e-7 in FP32 over p down to 2^-25 Phi(-1.25)) instead of normcdfinvf;
//   * MUFU-based exp / log / rcp / rsqrt (2 ulp) instead of IEEE division and sqrt sequences;
//   * the alternating series is tested at n = 1 inline (it decides > 99.4 % of the proposals); only the rest
//     enters the loop.
// The KS / moment tests against the oracle sampler arbitrate (tests/test_gpu_primitives.py).
// ---------------------------------------------------------------------------------------------
// MUFU approximations with flush-to-zero (one instruction each, <= 2 ulp; no IEEE fix-up or denormal paths)
__device__ __forceinline__ float f_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float f_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float f_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float f_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float f_cos(float x) { float r; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float f_exp(float x) { return f_ex2(x * 1.4426950408889634f); }

struct PGTaskF {
    float invK, p_right, z, hz2;
    bool small_z;
};

__device__ __forceinline__ void pg1_setup_fast(float c, PGTaskF& s) {
    const float t = (float)PG_TRUNC;
    const float z = 0.5f * fabsf(c);
    s.hz2 = 0.5f * z * z;
    const float K = (float)(OCCU_PI * OCCU_PI / 8.0) + s.hz2;
    s.invK = f_rcp(K);
    s.small_z = z * t < 1.0f;
    s.z = z;
    // q / p = (2K / pi) exp(K t) x { 4 Phi(-1.25)  |  2 exp(-z) }; overflow => p_right = 0
    const float qdivp = K * (s.small_z ? (float)(0.42259909466742100 * 2.0 / OCCU_PI) : (float)(4.0 / OCCU_PI)) *
                        f_exp(fmaf(K, t, s.small_z ? 0.0f : -z));
    s.p_right = f_rcp(1.0f + qdivp);  // 0 when qdivp overflows; never 1 (qdivp >= 0.33)
}

__device__ __forceinline__ bool pg1_round_fast(const PGTaskF& s, float ub, float ux, float u3, float us, float& x) {
    const float t = (float)PG_TRUNC;
    const float L = -0.6931471805599453f * f_lg2(ux);
    bool pre = true;
    float m2e;  // -2 e log2(e), e = pi^2 x / 2 (x > t) or 2 / x (x <= t): r_1 = 3 exp(-2 e)
    if (ub < s.p_right) {
        x = fmaf(L, s.invK, t);
        m2e = -(float)(OCCU_PI * OCCU_PI * 1.4426950408889634) * x;
    } else {
        float rx;  // 1 / x
        if (s.small_z) {
            const float q2 = 2.0f * (L + 2.2476256772143173f);
            const float q = q2 * f_rsqrt(q2);
            const float num = fmaf(fmaf(fmaf(0.5547605086386899f, q, 2.046679917526049f), q, -0.8955289493983294f), q,
                                   -2.686210871246895f);
            const float den = fmaf(fmaf(0.55384050985106f, q, 2.090629666186503f), q, 1.0f);
            const float y = num * f_rcp(den);
            rx = y * y;
            x = f_rcp(rx);
            pre = (x <= t) && (u3 <= f_exp(-s.hz2 * x));
        } else {
            const float mu = f_rcp(s.z);  // z t >= 1 on this path
            const float cs = f_cos(6.283185307179586f * u3);
            const float w = mu * (2.0f * L * cs * cs);  // mu Y^2, Y ~ N(0,1)
            const float disc = fmaf(0.25f * w, w, w);
            const float root = fmaf(disc, f_rsqrt(fmaxf(disc, 1e-30f)), fmaf(0.5f, w, 1.0f));  // 1 + w/2 + sqrt(w + w^2/4)
            // the two roots of Michael-Schucany-Haas are mu / root and mu root
            const float ur = (ub - s.p_right) * f_rcp(1.0f - s.p_right);
            const bool small_root = ur * (root + 1.0f) <= root;  // with probability mu / (mu + x), x = mu / root
            x = small_root ? mu * f_rcp(root) : mu * root;
            rx = f_rcp(x);
            pre = x <= t;
        }
        m2e = -4.0f * 1.4426950408889634f * rx;
    }
    float sum = fmaf(-3.0f, f_ex2(m2e), 1.0f);
    if (us <= sum) return pre;
    // rare: the first term did not decide
    const float e = -0.5f * 0.6931471805599453f * m2e;
    for (int n = 2;; n++) {
        const float rn = (float)(2 * n + 1) * f_exp(-(float)(n * (n + 1)) * e);
        if (n & 1) {
            sum -= rn;
            if (us <= sum) return pre;
        } else {
            sum += rn;
            if (us > sum) return false;
        }
    }
}

// Precision dispatch.  float: the production round above (the draw is a random variate, not a deterministic
// quantity; the KS / moment tests in tests/test_gpu_primitives.py arbitrate).  double: the generic round, FP64
// transcendental work -- the validation build (pg_fp64).
template <typename T>
struct PGSel {
    typedef PGTask<T> Task;
    static __device__ __forceinline__ void setup(T c, Task& s) { pg1_setup<T>(c, s); }
    static __device__ __forceinline__ bool round(const Task& s, T ub, T ux, T u3, T us, T& x) {
        return pg1_round<T>(s, ub, ux, u3, us, x);
    }
};
template <>
struct PGSel<float> {
    typedef PGTaskF Task;
    static __device__ __forceinline__ void setup(float c, Task& s) { pg1_setup_fast(c, s); }
    static __device__ __forceinline__ bool round(const Task& s, float ub, float ux, float u3, float us, float& x) {
        return pg1_round_fast(s, ub, ux, u3, us, x);
    }
};

// One PG(1, c) variate: the production loop of the site pass (setup once, one Philox block per rejection round,
// round keys from the kernel parameters) as a stand-alone function for the test hook occu_pg1.
template <typename T>
__device__ __forceinline__ T pg1_draw(T c, Philox& g, const PhiloxKeys& rk) {
    typedef RT<T> R;
    typename PGSel<T>::Task s;
    PGSel<T>::setup(c, s);
    for (;;) {
        T ub, ux, u3, us, x;
        R::unif4_rk(g, rk, ub, ux, u3, us);
        if (PGSel<T>::round(s, ub, ux, u3, us, x)) return (T)0.25 * x;
    }
}

// Standard-normal draw restricted to (0, inf) (positive) or (-inf, 0) after shifting by loc:
// returns w ~ N(loc, 1) truncated to w > 0 (positive) or w < 0 (!positive).
// Reference: inverse-CDF (probit.py:11-24), which underflows for |loc| >~ 8; here the far tail
// (truncation point more than 5 sd out) uses Robert's (1995) exponential-rejection sampler.
// Exact reference formula with the uniform injected (probit.py:16 / probit.py:24):
//   positive:  loc - ndtri(ndtr(loc) (1 - u))        !positive:  loc + ndtri(ndtr(-loc) u)
__device__ __forceinline__ double truncnorm_icdf(double loc, bool positive, double u) {
    const double a = positive ? -loc : loc;  // draw x ~ N(0,1) | x > a, then reflect
    const double pa = ndtr_d(-a);
    double x = -normcdfinv(pa * (positive ? 1.0 - u : u));
    if (x < a) x = a;  // guard the last-ulp rounding of the inverse CDF
    return positive ? loc + x : loc - x;
}

__device__ __forceinline__ double truncnorm_draw(double loc, bool positive, Philox& g) {
    const double a = positive ? -loc : loc;
    if (a < 5.0) return truncnorm_icdf(loc, positive, g.unif_d());
    const double lam = 0.5 * (a + sqrt(a * a + 4.0));
    double x;
    for (;;) {
        x = a - log(g.unif_d()) / lam;
        const double d = x - lam;
        if (g.unif_d() <= exp(-0.5 * d * d)) break;
    }
    return positive ? loc + x : loc - x;
}

// Production-precision variant: the same draw with the transcendental work in FP32 (one 32-bit uniform for the
// inverse-CDF branch).  Like the PG draws these are latent random variates, not deterministic quantities; the
// KS tests against scipy.stats.truncnorm arbitrate (tests/test_gpu_primitives.py).
// ---------------------------------------------------------------------------------------------
// FP32 building blocks of the production probit pass (all transcendental work is single MUFU instructions; the
// coefficients are least-squares fits made by benchmarks/fits/fit_probit_f32.py, which also prints the error bounds;
// the KS tests against scipy.stats.truncnorm arbitrate, tests/test_gpu_primitives.py).
//   erfc_pos_fast(z), z >= 0:  t exp(-z^2 + P(t)), t = 1 / (1 + z / 2)  (the form of Numerical Recipes' erfcc, refit to
//                              degree 9: 3.3e-7 in exact arithmetic, 4e-6 relative as evaluated in FP32 up to z = 6)
//   ndtri_fast(p), 0 < p < 1:  |p - 1/2| <= 0.425: (p - 1/2) A(s) / B(s), s = 0.180625 - (p - 1/2)^2; tails: rational in
//                              r = sqrt(-ln min(p, 1 - p)) (the split of Wichura's AS241, refit: 3.1e-7 / 4.6e-7 relative
//                              in FP32 down to p = 1e-19)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float erfc_pos_fast(float z) {
    const float t = f_rcp(fmaf(0.5f, z, 1.0f));
    float p = 1.879873355e-01f;
    p = fmaf(p, t, -9.019287893e-01f);
    p = fmaf(p, t, 1.645356232e+00f);
    p = fmaf(p, t, -1.304344333e+00f);
    p = fmaf(p, t, 3.878528898e-01f);
    p = fmaf(p, t, -2.291440062e-01f);
    p = fmaf(p, t, 1.068652698e-01f);
    p = fmaf(p, t, 3.727594862e-01f);
    p = fmaf(p, t, 1.000110509e+00f);
    p = fmaf(p, t, -1.265514260e+00f);
    return t * f_ex2(1.4426950408889634f * fmaf(-z, z, p));
}

// Phi(x) and Phi(-x) from one erfc evaluation
__device__ __forceinline__ void phi_pair_fast(float x, float& phi, float& phi_neg) {
    const float e = 0.5f * erfc_pos_fast(fabsf(x) * 0.70710678118654752f);   // Phi(-|x|)
    const float big = 1.0f - e;
    phi = x >= 0.0f ? big : e;
    phi_neg = x >= 0.0f ? e : big;
}

__device__ __forceinline__ float ndtri_fast(float p) {
    const float d = p - 0.5f;
    if (fabsf(d) <= 0.425f) {
        const float s = fmaf(-d, d, 0.180625f);
        const float num = fmaf(fmaf(fmaf(5.505418006e+01f, s, 1.531967914e+02f), s, 4.966260758e+01f), s, 3.387132222e+00f);
        const float den = fmaf(fmaf(fmaf(6.338554050e+01f, s, 7.627822074e+01f), s, 1.766729577e+01f), s, 1.0f);
        return d * num * f_rcp(den);
    }
    const float q = d < 0.0f ? p : 1.0f - p;
    const float L = -0.6931471805599453f * f_lg2(q);   // >= 2.59
    const float r = L * f_rsqrt(L);
    const float num = fmaf(fmaf(fmaf(1.710931907e+00f, r, 4.151230509e+00f), r, -1.414777384e+00f), r, -2.684062117e+00f);
    const float den = fmaf(fmaf(fmaf(1.019275364e-04f, r, 1.206112088e+00f), r, 3.015726461e+00f), r, 1.0f);
    const float y = num * f_rcp(den);                   // -ndtri(q) > 0
    return d < 0.0f ? -y : y;
}

// `word`: the 32 random bits of the inverse-CDF branch; `g`: generator for the (rare) far-tail rejection branch
__device__ __forceinline__ double truncnorm_word_f32(double loc, bool positive, uint32_t word, Philox& g) {
    const float a = positive ? -(float)loc : (float)loc;
    float x;
    if (a < 5.0f) {
        const float u = fmaf(__uint2float_rz(word), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
        float pa_neg, pa;
        phi_pair_fast(a, pa_neg, pa);   // pa = Phi(-a): mass of the kept side
        x = -ndtri_fast(pa * u);        // u and 1 - u are exchangeable
        // the approximations can land a few 1e-6 on the wrong side of the truncation point: reflect back inside, and stay
        // strictly inside (the support is open)
        x = a + fmaxf(fabsf(x - a), 1e-7f * (1.0f + fabsf(a)));
    } else {
        const float lam = 0.5f * (a + sqrtf(a * a + 4.0f));
        for (;;) {
            x = a - __logf(g.unif_f()) / lam;
            const float d = x - lam;
            if (g.unif_f() <= __expf(-0.5f * d * d)) break;
        }
    }
    const double xd = (double)x;
    return positive ? loc + xd : loc - xd;
}

__device__ __forceinline__ double truncnorm_draw_f32(double loc, bool positive, Philox& g) {
    const uint32_t word = g.next_u32();
    return truncnorm_word_f32(loc, positive, word, g);
}

// N(0,1) in FP32 (Box-Muller, 32-bit uniform for the radius: tail to 6.6 sigma)
__device__ __forceinline__ double normal_words_f32(uint32_t w1, uint32_t w2) {
    const float u1 = fmaf(__uint2float_rz(w1), 2.3283064365386963e-10f, 1.1641532182693481e-10f);   // < 1
    const float ang = ((float)(w2 >> 8) - 8388608.0f) * (6.283185307179586f / 16777216.0f);          // [-pi, pi)
    const float r2 = -1.3862943611198906f * f_lg2(u1);                                               // -2 ln u1 > 0
    return (double)(r2 * f_rsqrt(r2) * f_cos(ang));
}
__device__ __forceinline__ double normal_f32(Philox& g) {
    const uint32_t w1 = g.next_u32(), w2 = g.next_u32();
    return normal_words_f32(w1, w2);
}

// Gamma(shape, 1): Marsaglia & Tsang (2000); shape < 1 through the U^(1/shape) boost.
__device__ __forceinline__ double gamma_draw(double shape, Philox& g) {
    double boost = 1.0;
    if (shape < 1.0) {
        boost = pow(g.unif_d(), 1.0 / shape);
        shape += 1.0;
    }
    const double d = shape - 1.0 / 3.0;
    const double cc = 1.0 / sqrt(9.0 * d);
    for (;;) {
        double x, v;
        do {
            x = g.normal_d();
            v = 1.0 + cc * x;
        } while (v <= 0.0);
        v = v * v * v;
        const double u = g.unif_d();
        const double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2) return boost * d * v;
        if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) return boost * d * v;
    }
}
