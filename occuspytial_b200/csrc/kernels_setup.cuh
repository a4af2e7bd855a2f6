// kernels_setup.cuh -- one-off set-up kernels and chain-major <-> chain-minor transposes.
// Part of the single translation unit engine.cu (included there, in order; not a stand-alone header).
#pragma once

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

// ------------------------------------------------------------------------------------------------
// default start values, drawn on the device (base.py:199-211, logit.py:462-466, probit.py:155-160): a chain's values are
// addressed by (seed; global chain id, element, START_ITER, field) only -- independent of how the chains are sharded
// ------------------------------------------------------------------------------------------------
constexpr uint32_t START_ITER = 0xFFFFFFFFu;   // Philox iteration coordinate no sweep reaches
enum { START_TAU = 0, START_ALPHA = 1, START_BETA = 2, START_ETA = 3, START_EPS = 4 };

// dst [d][Cp] <- scale * N(0, 1); padded chains replicate chain (c mod C) like to_chain_minor_kernel
__global__ void draw_start_rows_kernel(View v, uint64_t seed, double* __restrict__ dst, int d, double scale, uint32_t field) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)d * v.Cp) return;
    const uint32_t e = (uint32_t)(idx / v.Cp), c = (uint32_t)(idx % v.Cp);
    Philox g;
    g.init(seed, (uint32_t)v.chain_offset + c % (uint32_t)v.C, e, START_ITER, field);
    dst[idx] = scale * g.normal_d();
}

// x [d][Cp] <- x - mean over the d rows, per chain (base.py:202-203); block (32, 8), one CTA per 32 chains
__global__ void centre_rows_kernel(double* __restrict__ x, int d, int Cp) {
    __shared__ double part[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    double s = 0.0;
    if (c < Cp)
        for (int e = threadIdx.y; e < d; e += 8) s += x[(size_t)e * Cp + c];
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    double m = 0.0;
    for (int k = 0; k < 8; k++) m += part[k][threadIdx.x];
    m /= (double)d;
    if (c < Cp)
        for (int e = threadIdx.y; e < d; e += 8) x[(size_t)e * Cp + c] -= m;
}

// tau ~ Gamma(1/2, scale 1 / tau_rate) = N(0,1)^2 / (2 tau_rate); alpha = a_mu + L_a z; beta = b_mu + L_b z.  One thread per chain.
__global__ void draw_start_small_kernel(View v, uint64_t seed) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= v.Cp) return;
    const uint32_t gid = (uint32_t)v.chain_offset + (uint32_t)(c % v.C);
    const HParams* hp = v.hp;
    Philox g;
    g.init(seed, gid, 0u, START_ITER, START_TAU);
    const double zt = g.normal_d();
    v.tau[c] = 0.5 * zt * zt / hp->tau_rate;
    double z[MAX_SMALL];
    g.init(seed, gid, 0u, START_ITER, START_ALPHA);
    for (int j = 0; j < v.r; j++) z[j] = g.normal_d();
    for (int a = 0; a < v.r; a++) {
        double s = hp->a_mu[a];
        for (int b = 0; b <= a; b++) s += hp->a_chol[a * v.r + b] * z[b];
        v.alpha[(size_t)a * v.Cp + c] = s;
    }
    g.init(seed, gid, 0u, START_ITER, START_BETA);
    for (int j = 0; j < v.p; j++) z[j] = g.normal_d();
    for (int a = 0; a < v.p; a++) {
        double s = hp->b_mu[a];
        for (int b = 0; b <= a; b++) s += hp->b_chol[a * v.p + b] * z[b];
        v.theta[(size_t)a * v.Cp + c] = s;
    }
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
