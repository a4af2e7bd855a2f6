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

// T = float: transcendental work in FP32 (the draw is a random variate, not a deterministic quantity;
// the KS / moment tests in tests/test_gpu_samplers.py arbitrate).  T = double: validation build.
template <typename T>
__device__ __forceinline__ T pg1_draw(T c, Philox& g) {
    typedef RT<T> R;
    const T t = (T)PG_TRUNC;
    const T z = (T)0.5 * fabs(c);
    const T K = (T)(OCCU_PI * OCCU_PI / 8.0) + (T)0.5 * z * z;
    // mixture weight of the right (exponential) piece: p / (p + q), evaluated through q / p in log space
    const T x0 = R::log_(K) + K * t;
    const T rs = (T)1.25;  // sqrt(1 / t)
    const T la = R::log_((T)0.5 * R::erfc_(-rs * (t * z - (T)1) * (T)0.70710678118654752)) ;
    const T lb = R::log_((T)0.5 * R::erfc_(rs * (t * z + (T)1) * (T)0.70710678118654752));
    const T qdivp = (T)(4.0 / OCCU_PI) * (R::exp_(x0 - z + la) + R::exp_(x0 + z + lb));
    const T p_right = (T)1 / ((T)1 + qdivp);
    const bool small_z = z * t < (T)1;  // 1/z > t: truncated IG by the exponential-pair rejection
    const T mu = small_z ? (T)0 : (T)1 / z;
    const T hz2 = (T)0.5 * z * z;

    for (;;) {
        T x;
        if (R::unif(g) < p_right) {
            x = t - R::log_(R::unif(g)) / K;
        } else if (small_z) {
            for (;;) {
                T e1, e2;
                do {
                    e1 = -R::log_(R::unif(g));
                    e2 = -R::log_(R::unif(g));
                } while (e1 * e1 * t > (T)2 * e2);
                x = (T)1 + t * e1;
                x = t / (x * x);
                if (R::unif(g) <= R::exp_(-hz2 * x)) break;
            }
        } else {
            do {
                // IG(mu, 1) by Michael-Schucany-Haas; the small root is written in the cancellation-free form
                const T u1 = R::unif(g), u2 = R::unif(g);
                const T nrm = R::sqrt_((T)-2 * R::log_(u1)) * R::cospi_((T)2 * u2);
                const T w = mu * nrm * nrm;
                x = mu / ((T)1 + (T)0.5 * w + R::sqrt_(w + (T)0.25 * w * w));
                if (R::unif(g) * (mu + x) > mu) x = mu * mu / x;
            } while (x > t);
        }
        // alternating-series squeeze on the ratios r_n = a_n(x) / a_0(x) = (2n+1) exp(-n(n+1) e), with
        // e = 2/x (x <= t) or pi^2 x / 2 (x > t); accept when U <= partial sum after an odd term.
        const T e = (x > t) ? (T)(OCCU_PI * OCCU_PI * 0.5) * x : (T)2 / x;
        const T u = R::unif(g);
        T s = (T)1;
        bool accepted = false;
        for (int n = 1;; n++) {
            const T rn = (T)(2 * n + 1) * R::exp_(-(T)(n * (n + 1)) * e);
            if (n & 1) {
                s -= rn;
                if (u <= s) {
                    accepted = true;
                    break;
                }
            } else {
                s += rn;
                if (u > s) break;
            }
        }
        if (accepted) return (T)0.25 * x;
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
