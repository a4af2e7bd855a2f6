/*
 * oracle/pg_devroye.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the PG(1, c) sampler that the reference obtains from the
 * third-party package polyagamma==1.2.0 (poetry.lock:865-866), called at
 *   occuspytial/gibbs/logit.py:191-193  (omega_a ~ PG(1, W_e alpha))
 *   occuspytial/gibbs/logit.py:202-204  (omega_b ~ PG(1, X beta + spatial))
 * The package is not vendored under /root/reference and is not installable
 * here, so the algorithm is restated from its published description:
 * Polson, Scott & Windle (2013), "Bayesian inference for logistic models using
 * Polya-Gamma latent variables", Algorithm 1 / supplementary S1 (Devroye's
 * alternating-series method for J*(1, z)), which is the method polyagamma's
 * hybrid sampler selects for h = 1.
 *
 * PARITY UNPINNED at this boundary: the reference has no test that pins PG
 * draws.  The restatement is pinned instead against the
 * analytic moments / Laplace transform and an independent sum-of-exponentials
 * representation (tests/test_oracle_pg.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * call this file.  It also backs the `polyagamma` shim (oracle/shims) used to
 * import the real reference in the build container.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#define PG_T 0.64 /* truncation point t of the two-piece density */
#define PG_PI 3.141592653589793238462643383279502884
#define PG_PI2_8 1.233700550136169827354311374984518891 /* pi^2 / 8 */

/* ---------- private RNG: xoshiro256++ seeded through splitmix64 ---------- */
typedef struct {
    uint64_t s[4];
} rng_t;

static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

static uint64_t splitmix64(uint64_t* x) {
    uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

static void rng_seed(rng_t* r, uint64_t seed) {
    for (int i = 0; i < 4; i++) r->s[i] = splitmix64(&seed);
}

static inline uint64_t rng_next(rng_t* r) {
    uint64_t* s = r->s;
    const uint64_t result = rotl(s[0] + s[3], 23) + s[0];
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return result;
}

/* uniform on (0, 1): 53 random bits, never exactly 0 */
static inline double rng_unif(rng_t* r) {
    return ((double)(rng_next(r) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

static inline double rng_exp(rng_t* r) { return -log(rng_unif(r)); }

static inline double rng_norm(rng_t* r) {
    /* Marsaglia polar method, one value per call (the spare is discarded) */
    for (;;) {
        double u = 2.0 * rng_unif(r) - 1.0, v = 2.0 * rng_unif(r) - 1.0;
        double s = u * u + v * v;
        if (s > 0.0 && s < 1.0) return u * sqrt(-2.0 * log(s) / s);
    }
}

/* log Phi(x), stable in the lower tail */
static inline double log_norm_cdf(double x) { return log(0.5 * erfc(-x * M_SQRT1_2)); }

/* a_n(x): n-th coefficient of the alternating series for the J*(1,0) density */
static inline double a_coef(int n, double x) {
    const double h = (double)n + 0.5;
    if (x > PG_T) return PG_PI * h * exp(-0.5 * h * h * PG_PI * PG_PI * x);
    const double c = 2.0 / (PG_PI * x);
    return PG_PI * h * c * sqrt(c) * exp(-2.0 * h * h / x);
}

/* IG(mu = 1/z, lambda = 1) restricted to (0, t) */
static double trunc_invgauss(rng_t* r, double z) {
    double x;
    if (z * PG_T < 1.0) { /* 1/z > t, includes z == 0 */
        for (;;) {
            double e1, e2;
            do {
                e1 = rng_exp(r);
                e2 = rng_exp(r);
            } while (e1 * e1 > 2.0 * e2 / PG_T);
            x = 1.0 + PG_T * e1;
            x = PG_T / (x * x);
            if (rng_unif(r) <= exp(-0.5 * z * z * x)) return x;
        }
    }
    const double mu = 1.0 / z;
    do {
        double y = rng_norm(r);
        y *= y;
        const double muy = mu * y;
        x = mu + 0.5 * mu * muy - 0.5 * mu * sqrt(4.0 * muy + muy * muy);
        if (rng_unif(r) > mu / (mu + x)) x = mu * mu / x;
    } while (x > PG_T);
    return x;
}

/* one PG(1, c) draw */
static double pg1(rng_t* r, double c) {
    const double z = 0.5 * fabs(c);
    const double K = PG_PI2_8 + 0.5 * z * z;
    const double x0 = log(K) + K * PG_T;
    const double st = sqrt(1.0 / PG_T);
    const double qdivp = (4.0 / PG_PI) * (exp(x0 - z + log_norm_cdf(st * (PG_T * z - 1.0))) +
                                          exp(x0 + z + log_norm_cdf(-st * (PG_T * z + 1.0))));
    const double p_right = 1.0 / (1.0 + qdivp);
    for (;;) {
        double x;
        if (rng_unif(r) < p_right)
            x = PG_T + rng_exp(r) / K;
        else
            x = trunc_invgauss(r, z);
        double s = a_coef(0, x);
        const double y = rng_unif(r) * s;
        for (int n = 1;; n++) {
            if (n & 1) {
                s -= a_coef(n, x);
                if (y <= s) return 0.25 * x;
            } else {
                s += a_coef(n, x);
                if (y > s) break;
            }
        }
    }
}

/* out[i] ~ PG(1, c[i]), i < n; stream fully determined by `seed` */
void occu_oracle_pg1_fill(uint64_t seed, size_t n, const double* c, double* out) {
    rng_t r;
    rng_seed(&r, seed);
    for (size_t i = 0; i < n; i++) out[i] = pg1(&r, c[i]);
}

/*
 * Independent representation used only to cross-check the sampler above:
 * PG(1, c) = (1 / 2 pi^2) sum_{k>=1} g_k / ((k - 1/2)^2 + c^2 / (4 pi^2)),
 * g_k ~ Exp(1), truncated at `terms` with the tail replaced by its mean.
 */
void occu_oracle_pg1_series_fill(uint64_t seed, size_t n, const double* c, int terms, double* out) {
    rng_t r;
    rng_seed(&r, seed);
    for (size_t i = 0; i < n; i++) {
        const double c2 = c[i] * c[i] / (4.0 * PG_PI * PG_PI);
        double acc = 0.0;
        for (int k = 1; k <= terms; k++) {
            const double d = ((double)k - 0.5) * ((double)k - 0.5) + c2;
            acc += rng_exp(&r) / d;
        }
        /* tail mean: sum_{k>terms} 1/((k-1/2)^2 + c2) ~ integral bound, midpoint rule */
        double tail = 0.0;
        if (c2 > 0.0) {
            const double sc = sqrt(c2);
            tail = (0.5 * PG_PI - atan((double)terms / sc)) / sc;
        } else {
            tail = 1.0 / (double)terms;
        }
        out[i] = (acc + tail) / (2.0 * PG_PI * PG_PI);
    }
}
