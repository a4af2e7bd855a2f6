// common.cuh -- shared device helpers for the occuspytial_b200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define OCCU_PI 3.141592653589793238462643383279502884
#define OCCU_PI_F 3.14159265358979f

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator (Salmon et al. 2011).  A draw is addressed by logical
// coordinates only -- (seed; chain, element, iteration, stage/block) -- so that the same seed gives
// the same chains whatever the grid shape or the number of GPUs (reference contract: same
// random_state => same chains, gibbs/tests/test_samplers.py:63-68).
// ---------------------------------------------------------------------------------------------
// The ten round keys (key + i * Weyl constants) depend on the seed only.  The host computes them once; kernels
// that take them inside their parameter block use them as constant-bank operands, which removes the 18 key-update
// instructions from every block (refill_rk below; same stream as refill()).
struct PhiloxKeys {
    uint32_t k[20];
};
static inline void philox_round_keys(uint64_t seed, PhiloxKeys& rk) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int i = 0; i < 10; i++) {
        rk.k[2 * i] = k0;
        rk.k[2 * i + 1] = k1;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

constexpr int PHILOX_FAST_ROUNDS = 7;

struct Philox {
    uint32_t ctr[4];
    uint32_t key[2];
    uint32_t buf[4];
    int have;

    // chain: global chain id, elem: site / visit row / vector slot, iter: sweep number,
    // stage: which conditional (8 bits); the low 24 bits of ctr[3] count 128-bit blocks.
    __device__ __forceinline__ void init(uint64_t seed, uint32_t chain, uint32_t elem, uint32_t iter, uint32_t stage) {
        key[0] = (uint32_t)seed;
        key[1] = (uint32_t)(seed >> 32);
        ctr[0] = chain;
        ctr[1] = elem;
        ctr[2] = iter;
        ctr[3] = stage << 24;
        have = 0;
    }

    __device__ __forceinline__ void refill() {
        uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
        uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
        for (int i = 0; i < 10; i++) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            c0 = hi1 ^ c1 ^ k0;
            c1 = lo1;
            c2 = hi0 ^ c3 ^ k1;
            c3 = lo0;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        buf[0] = c0;
        buf[1] = c1;
        buf[2] = c2;
        buf[3] = c3;
        ctr[3] += 1u;
        have = 4;
    }

    // ROUNDS = 7 (Philox4x32-7, the smallest round count of the family that passes BigCrush, Salmon et al. 2011 table 2)
    // is used by the production PG / truncated-normal draws of the site pass, whose cost is dominated by this function;
    // everything else keeps the conventional 10.
    template <int ROUNDS = 10>
    __device__ __forceinline__ void refill_rk(const PhiloxKeys& rk) {
        uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
#pragma unroll
        for (int i = 0; i < ROUNDS; i++) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            c0 = hi1 ^ c1 ^ rk.k[2 * i];
            c1 = lo1;
            c2 = hi0 ^ c3 ^ rk.k[2 * i + 1];
            c3 = lo0;
        }
        buf[0] = c0;
        buf[1] = c1;
        buf[2] = c2;
        buf[3] = c3;
        ctr[3] += 1u;
        have = 4;
    }

    __device__ __forceinline__ uint32_t next_u32() {
        if (have == 0) refill();
        have--;
        // static indexing keeps buf[] in registers
        return have == 3 ? buf[0] : have == 2 ? buf[1] : have == 1 ? buf[2] : buf[3];
    }

    // uniform on (0,1), 24 random bits, never 0 or 1
    __device__ __forceinline__ float unif_f() { return ((float)(next_u32() >> 8) + 0.5f) * (1.0f / 16777216.0f); }

    // uniform on (0,1), 53 random bits, never 0 or 1
    __device__ __forceinline__ double unif_d() {
        const uint32_t a = next_u32(), b = next_u32();
        const uint64_t v = ((uint64_t)a << 21) ^ (uint64_t)(b >> 11);  // 53 bits
        return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
    }

    // standard normal, Box-Muller in FP64 (one value per call; only O(q) of these per chain-iteration)
    __device__ __forceinline__ double normal_d() {
        const double u1 = unif_d(), u2 = unif_d();
        return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
};

template <typename T>
struct RT;  // real-type dispatch for samplers that run in either precision
template <>
struct RT<float> {
    static __device__ __forceinline__ float unif(Philox& g) { return g.unif_f(); }
    static __device__ __forceinline__ float exp_(float x) { return __expf(x); }
    static __device__ __forceinline__ float log_(float x) { return __logf(x); }
    static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float erfc_(float x) { return erfcf(x); }
    static __device__ __forceinline__ float cospi_(float x) { return cospif(x); }
    static __device__ __forceinline__ float normcdfinv_(float x) { return normcdfinvf(x); }
    // four uniforms from exactly one Philox block, round keys from the kernel parameters; (0, 1): the word is
    // truncated to 24 significant bits (round toward zero) and centred by 2^-33
    static __device__ __forceinline__ void unif4_rk(Philox& g, const PhiloxKeys& rk, float& a, float& b, float& c, float& d) {
        g.refill_rk<PHILOX_FAST_ROUNDS>(rk);
        g.have = 0;
        a = fmaf(__uint2float_rz(g.buf[0]), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
        b = fmaf(__uint2float_rz(g.buf[1]), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
        c = fmaf(__uint2float_rz(g.buf[2]), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
        d = fmaf(__uint2float_rz(g.buf[3]), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    }
    // four uniforms from exactly one Philox block
    static __device__ __forceinline__ void unif4(Philox& g, float& a, float& b, float& c, float& d) {
        g.refill();
        g.have = 0;
        a = ((float)(g.buf[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        b = ((float)(g.buf[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        c = ((float)(g.buf[2] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        d = ((float)(g.buf[3] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    }
};
template <>
struct RT<double> {
    static __device__ __forceinline__ double unif(Philox& g) { return g.unif_d(); }
    static __device__ __forceinline__ double exp_(double x) { return exp(x); }
    static __device__ __forceinline__ double log_(double x) { return log(x); }
    static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
    static __device__ __forceinline__ double erfc_(double x) { return erfc(x); }
    static __device__ __forceinline__ double cospi_(double x) { return cospi(x); }
    static __device__ __forceinline__ double normcdfinv_(double x) { return normcdfinv(x); }
    static __device__ __forceinline__ void unif4_rk(Philox& g, const PhiloxKeys&, double& a, double& b, double& c, double& d) {
        unif4(g, a, b, c, d);
    }
    // four 53-bit uniforms from two Philox blocks
    static __device__ __forceinline__ void unif4(Philox& g, double& a, double& b, double& c, double& d) {
        g.have = 0;
        a = g.unif_d();
        b = g.unif_d();
        c = g.unif_d();
        d = g.unif_d();
    }
};

// logistic function in FP64, safe for large |x| (scipy.special.expit semantics)
__device__ __forceinline__ double expit_d(double x) {
    if (x >= 0.0) return 1.0 / (1.0 + exp(-x));
    const double e = exp(x);
    return e / (1.0 + e);
}

// Phi(x) in FP64 (scipy.special.ndtr semantics)
__device__ __forceinline__ double ndtr_d(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }
