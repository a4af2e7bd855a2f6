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

// T = float: transcendental work in FP32 (the draw is a random variate, not a deterministic quantity;
// the KS / moment tests in tests/test_gpu_primitives.py arbitrate).  T = double: validation build.
template <typename T>
__device__ __forceinline__ T pg1_draw(T c, Philox& g) {
    typedef RT<T> R;
    PGTask<T> s;
    pg1_setup<T>(c, s);
    for (;;) {
        T ub, ux, u3, us, x;
        R::unif4(g, ub, ux, u3, us);
        if (pg1_round<T>(s, ub, ux, u3, us, x)) return (T)0.25 * x;
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
__device__ __forceinline__ double truncnorm_draw_f32(double loc, bool positive, Philox& g) {
    const float a = positive ? -(float)loc : (float)loc;
    float x;
    if (a < 5.0f) {
        const float u = ((float)(g.next_u32() >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float pa = normcdff(-a);
        x = -normcdfinvf(pa * u);   // u and 1 - u are exchangeable
        x = fmaxf(x, a);
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

// N(0,1) in FP32 (Box-Muller, 32-bit uniform for the radius: tail to 6.6 sigma)
__device__ __forceinline__ double normal_f32(Philox& g) {
    const float u1 = ((float)g.next_u32() + 0.5f) * (1.0f / 4294967296.0f);
    const float u2 = g.unif_f();
    return (double)(sqrtf(-2.0f * __logf(u1)) * cospif(2.0f * u2));
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
