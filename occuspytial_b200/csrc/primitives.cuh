// primitives.cuh -- stand-alone primitives behind the C-ABI test hooks, and the chain-summary kernel.
// Part of the single translation unit engine.cu (included there, in order; not a stand-alone header).
#pragma once

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
