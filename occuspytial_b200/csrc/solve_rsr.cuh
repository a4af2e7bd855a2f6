// solve_rsr.cuh -- alpha / tau kernel, partial-sum reduction and the shared-memory RSR solve (q <~ 200).
// Part of the single translation unit engine.cu (included there, in order; not a stand-alone header).
#pragma once

// ------------------------------------------------------------------------------------------------
// per-chain solve: tau, eta, beta, alpha
// ------------------------------------------------------------------------------------------------
struct SolveDebug {
    const double* normals;  // [C][q+p+r] or NULL
    const double* tau_in;   // [C] or NULL
    occu_debug_out_t out;
    int enabled;
};

__device__ __forceinline__ int tri(int a, int b) { return a * (a + 1) / 2 + b; }  // b <= a

// In-place Cholesky of the d x d block whose (i, j) entry (j <= i) lives at A[tri(o + i, o + j)].
// Returns false (uniformly) on a non-positive pivot.
__device__ bool block_cholesky_packed(double* A, int o, int d, int tid, int nthreads) {
    for (int j = 0; j < d; j++) {
        __syncthreads();
        const double piv = A[tri(o + j, o + j)];
        if (!(piv > 0.0)) return false;
        const double dj = sqrt(piv);
        const double inv = 1.0 / dj;
        __syncthreads();
        for (int i = j + tid; i < d; i += nthreads) {
            if (i == j) A[tri(o + j, o + j)] = dj;
            else A[tri(o + i, o + j)] *= inv;
        }
        __syncthreads();
        const int m = d - 1 - j;
        const int cnt = m * (m + 1) / 2;
        for (int t = tid; t < cnt; t += nthreads) {
            int ii = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
            while (ii * (ii + 1) / 2 > t) ii--;
            while ((ii + 1) * (ii + 2) / 2 <= t) ii++;
            const int kk = t - ii * (ii + 1) / 2;
            const int i = j + 1 + ii, k = j + 1 + kk;
            A[tri(o + i, o + k)] -= A[tri(o + i, o + j)] * A[tri(o + k, o + j)];
        }
    }
    __syncthreads();
    return true;
}

// Sequential small dense solve by one thread: x = P^-1 b + L^-T z, P = L L' (full d x d row-major in S,
// lower triangle overwritten by L).  Returns false on a non-positive pivot.  distributions.pyx:88-105.
__device__ bool small_precision_mvnorm(double* S, int d, const double* b, const double* zn, double* x, double* tmp) {
    // one rsqrt per pivot and multiplications by the reciprocal pivot instead of a square root and 2 d divisions (each
    // a few dozen dependent FP64 instructions in a thread that a whole CTA may be waiting for); the diagonal of S
    // receives 1 / L_jj
    for (int j = 0; j < d; j++) {
        double s = S[j * d + j];
        for (int k = 0; k < j; k++) s -= S[j * d + k] * S[j * d + k];
        if (!(s > 0.0)) return false;
        const double inv = rsqrt(s);
        S[j * d + j] = inv;
        for (int i = j + 1; i < d; i++) {
            double t = S[i * d + j];
            for (int k = 0; k < j; k++) t -= S[i * d + k] * S[j * d + k];
            S[i * d + j] = t * inv;
        }
    }
    for (int i = 0; i < d; i++) {  // L y = b
        double t = b[i];
        for (int k = 0; k < i; k++) t -= S[i * d + k] * tmp[k];
        tmp[i] = t * S[i * d + i];
    }
    for (int i = d - 1; i >= 0; i--) {  // L' x = y + z
        double t = tmp[i] + zn[i];
        for (int k = i + 1; k < d; k++) t -= S[k * d + i] * x[k];
        x[i] = t * S[i * d + i];
    }
    return true;
}

// Per-chain scalar work that does not depend on the Gram, one THREAD per chain (chain-minor, coalesced):
//  * alpha | omega_a, z  (logit.py:219-224 / probit.py:231-235), committed and recorded here,
//  * the unit-rate Gamma variate of tau | eta (tau = gam / rate, logit.py:206-209).
// Keeping these serial FP64 chains (sqrt / divide / log) out of the one-CTA-per-chain solve kernels matters:
// there a single lane would hold 127 others at a barrier for ~10^4 cycles per chain.
__global__ void __launch_bounds__(32) alpha_tau_kernel(const View v, const SolveDebug dbg, double* rec_alpha,
                                                        int rec_capacity) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= v.Cp) return;
    const int r = v.r, NA = r * (r + 1) / 2, NV = v.NV;
    const size_t Cp = v.Cp;
    const HParams* hp = v.hp;
    const uint32_t iter = v.ctl->iter;
    const int rec_index = v.ctl->rec_index;
    const uint32_t gc = (uint32_t)(v.chain_offset + c);
    const int cr = c % v.C;
    const int d_eta = v.model == OCCU_LOGIT_ICAR ? v.n : v.q;   // normals stream: eta, beta, alpha
    double sums[8 * 9 / 2 + 8];
    for (int k = 0; k < NV; k++) sums[k] = v.asum[(size_t)k * Cp + c];  // reduced by reduce_partials_kernel
    double S[64], b[8], x[8], tmp[8], zn[8];
    for (int a = 0; a < r; a++) {
        for (int bb = 0; bb < r; bb++) S[a * r + bb] = sums[bb <= a ? tri(a, bb) : tri(bb, a)] + hp->a_prec[a * r + bb];
        b[a] = sums[NA + a] + hp->a_pm[a];
        if (dbg.normals) {
            zn[a] = dbg.normals[(size_t)cr * (d_eta + v.p + r) + d_eta + v.p + a];
        } else {
            Philox g;
            g.init(v.seed, gc, (uint32_t)(d_eta + v.p + a), iter, ST_NORMALS);
            zn[a] = g.normal_d();
        }
    }
    if (dbg.enabled && c < v.C) {
        if (dbg.out.alpha_prec)
            for (int idx = 0; idx < r * r; idx++) dbg.out.alpha_prec[(size_t)c * r * r + idx] = S[idx];
        if (dbg.out.alpha_b)
            for (int a = 0; a < r; a++) dbg.out.alpha_b[(size_t)c * r + a] = b[a];
    }
    if (!small_precision_mvnorm(S, r, b, zn, x, tmp)) {
        v.status[c] = 1;
        if (c < v.C) atomicAdd(&v.ctl->fails, 1);
    } else if (!v.status[c]) {
        for (int a = 0; a < r; a++) v.alpha[(size_t)a * Cp + c] = x[a];
        if (rec_alpha && rec_index >= 0 && rec_index < rec_capacity && c < v.C) {
            const size_t row = (size_t)c * rec_capacity + rec_index;
            for (int a = 0; a < r; a++) rec_alpha[row * r + a] = x[a];
        }
    }
    Philox g;
    g.init(v.seed, gc, 0u, iter, ST_TAU);
    v.gam[c] = gamma_draw(hp->tau_shape, g);
}

// asum[k][c] = sum_tile apart[tile][k][c], fixed order
// Block = 32 sums x 8 interleaved slices of the site tiles (8 x the loads in flight of a one-thread-per-sum loop),
// combined in slice order: the summation order is fixed, so the result is reproducible.
constexpr int RED_SPLIT = 8;
__global__ void __launch_bounds__(32 * RED_SPLIT) reduce_partials_kernel(const View v) {
    __shared__ double part[RED_SPLIT][33];
    const size_t idx = (size_t)blockIdx.x * 32 + threadIdx.x;
    const size_t per = (size_t)v.NV * v.Cp;
    double s = 0.0;
    if (idx < per) {
#pragma unroll 4
        for (int t = threadIdx.y; t < v.n_tiles; t += RED_SPLIT) s += v.apart[(size_t)t * per + idx];
    }
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && idx < per) {
        double tot = 0.0;
#pragma unroll
        for (int k = 0; k < RED_SPLIT; k++) tot += part[k][threadIdx.x];
        v.asum[idx] = tot;
    }
}

// ---- 4x4-blocked Cholesky machinery of the RSR solve ------------------------------------------------
// Lambda (q x q, padded to qb = ceil(q/4) blocks with an identity tail) plus ONE extra block row holding the
// right-hand side b (row 0 of block row qb) are kept in shared memory as packed lower-triangular 4x4 blocks,
// block (I, K) at BLK * (I (I + 1) / 2 + K), row-major inside, BLK = 17 doubles (odd stride: threads working on
// consecutive blocks hit different banks).  Factorising the augmented matrix yields y = L^-1 b in the extra
// row for free.  Per block column J: (1) every thread that owns a panel block (I, J) factorises the 4x4
// diagonal block redundantly in registers and solves its own block against it -- no serial section, no extra
// barrier; (2) the trailing blocks (I, K), J < K <= I, are updated by A_IK -= P_I P_K' (64 FMAs per block).
// Two barriers per block column instead of three per column, and no per-element index decoding.
constexpr int BLK = 17;
__device__ __forceinline__ int bidx(int I, int K) { return (I * (I + 1) / 2 + K) * BLK; }

// lower Cholesky of a 4x4 block held in registers (a[r][c], c <= r used); returns false on a bad pivot
// (inv[j] = 1 / L_jj; one rsqrt per pivot instead of a sqrt and a divide on the critical path)
__device__ __forceinline__ bool chol4(double (&a)[4][4], double (&invd)[4]) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
        double d = a[j][j];
#pragma unroll
        for (int k = 0; k < j; k++) d -= a[j][k] * a[j][k];
        if (!(d > 0.0)) return false;
        const double inv = rsqrt(d), dj = d * inv;
        invd[j] = inv;
        a[j][j] = dj;
#pragma unroll
        for (int i = j + 1; i < 4; i++) {
            double t = a[i][j];
#pragma unroll
            for (int k = 0; k < j; k++) t -= a[i][k] * a[j][k];
            a[i][j] = t * inv;
        }
    }
    return true;
}

__global__ void __launch_bounds__(SOLVE_THREADS, 6) solve_logit_rsr_kernel(const View v, const SolveDebug dbg,
                                                                         double* rec_alpha, double* rec_beta,
                                                                         double* rec_tau, int rec_capacity) {
#ifdef RSRS_PROF   // measurement builds: cycles per phase of chain 0, thread 0
#define RSRS_MARK(k) do { if (tid == 0) { const long long t_ = clock64(); rs_t[k] += t_ - rs_last; rs_last = t_; } } while (0)
    long long rs_t[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, rs_last = clock64();
#else
#define RSRS_MARK(k)
#endif
    extern __shared__ double sm_dyn[];
    const int c = blockIdx.x, tid = threadIdx.x, T = blockDim.x;  // 64 (q <= 128) or SOLVE_THREADS
    double* sm = v.solve_stride ? v.solve_scratch + (size_t)c * v.solve_stride : sm_dyn;
    const int p = v.p, q = v.q, r = v.r, pq = v.pq;
    const int qb = (q + 3) / 4, q4 = qb * 4;
    const int nblk = (qb + 1) * (qb + 2) / 2;  // lower-triangular blocks incl. the right-hand-side block row
    const size_t Cp = v.Cp;
    const HParams* hp = v.hp;
    double* Bs = sm;                          // [nblk][BLK]
    double* gx = Bs + (size_t)nblk * BLK;     // [q4][p]  K' Om X  (rows of the joint Gram below the XX block)
    double* gxx = gx + (size_t)q4 * p;        // [p][p]   X' Om X (symmetric, full)
    double* h = gxx + p * p;                  // [pq]     U'(z - 1/2)
    double* th = h + pq;                      // [pq + 4] current beta, eta (eta padded)
    double* xm = th + pq + 4;                 // [q4] backward-solve result: mean
    double* xe = xm + q4;                     // [q4] backward-solve result: L^-T eps
    double* nrm = xe + q4;                    // [q4 + p + r] normals (eta part padded to q4)
    double* red = nrm + q4 + 2 * MAX_SMALL;   // [8]
    double* Ld = red + 8;                     // [qb][16] factored diagonal blocks (kept apart: the panel threads
                                              //          read the unfactored block while its owner factorises it)
    double* Li = Ld + qb * 16;                // [qb][4] reciprocals of the diagonal of L
    uint8_t* pair_i = reinterpret_cast<uint8_t*>(Li + qb * 4);
    uint8_t* pair_k = pair_i + nblk;
    __shared__ double small_S[MAX_SMALL * MAX_SMALL];
    __shared__ double small_b[MAX_SMALL], small_x[MAX_SMALL], small_t[MAX_SMALL];
    __shared__ int ok_flag;

    const uint32_t iter = v.ctl->iter;
    const int rec_index = v.ctl->rec_index;
    const uint32_t gc = (uint32_t)(v.chain_offset + c);
    const int cr = c % v.C;  // chain whose injected values a padded chain replicates

    // logit: per-chain omega-weighted Gram; probit: the shared unit-weight Gram (A_eta = K'K + tau Q, probit.py:226)
    const double* Grow = v.model == OCCU_PROBIT_RSR ? v.G0 : v.G + (size_t)c * v.M_pad;

    for (int a = tid; a < pq; a += T) {
        double hs = 0.0;  // K-split partials, fixed order
        for (int s = 0; s < v.ksplit; s++) hs += v.Hpart[((size_t)s * Cp + c) * v.ldh + a];
        h[a] = hs;
        th[a] = v.theta[(size_t)a * Cp + c];
    }
    // normals: eta (q), beta (p), alpha (r); stored as eta[0..q4) (zero padded), then beta, alpha
    for (int k = tid; k < q4 + p + r; k += T) {
        const int src = k < q4 ? k : k - q4 + q;  // index in the (q + p + r) stream
        double val = 0.0;
        if (k >= q4 || k < q) {
            if (dbg.normals) {
                val = dbg.normals[(size_t)cr * (pq + r) + src];
            } else {
                Philox g;
                g.init(v.seed, gc, (uint32_t)src, iter, ST_NORMALS);
                val = g.normal_d();
            }
        }
        nrm[k] = val;
    }
    {   // block-pair table: t -> (ii, kk), kk <= ii, row-major over the lower triangle
        int t = 0;
        for (int ii = 0; ii <= qb; ii++) {
            for (int kk = tid; kk <= ii; kk += T) {
                pair_i[t + kk] = (uint8_t)ii;
                pair_k[t + kk] = (uint8_t)kk;
            }
            t += ii + 1;
        }
    }
    if (tid == 0) ok_flag = 1;
    __syncthreads();

    RSRS_MARK(0);
    // ---- tau | eta   (logit.py:206-209) ----
    {
        double part = 0.0;
        // one warp per row of Qrsr, lanes along the row (coalesced; no division to decode a flat index), eight rows in
        // flight per warp: all loads of a batch are issued before any is consumed
        const int warp = tid >> 5, lane = tid & 31, nwarp = T >> 5;
        const double t0 = lane < q ? th[p + lane] : 0.0, t1 = lane + 32 < q ? th[p + lane + 32] : 0.0;
        for (int i0 = warp; i0 < q; i0 += 8 * nwarp) {
            double qv[8][2];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int i = i0 + u * nwarp;
                const double* qrow = v.Qrsr + (size_t)(i < q ? i : 0) * q;
                qv[u][0] = (i < q && lane < q) ? qrow[lane] : 0.0;
                qv[u][1] = (i < q && lane + 32 < q) ? qrow[lane + 32] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int i = i0 + u * nwarp;
                if (i < q) part += th[p + i] * (qv[u][0] * t0 + qv[u][1] * t1);
            }
            for (int i = i0; i < q && i < i0 + 8 * nwarp; i += nwarp) {   // columns beyond 64 (q > 64)
                double rowsum = 0.0;
                for (int j = 64 + lane; j < q; j += 32) rowsum += v.Qrsr[(size_t)i * q + j] * th[p + j];
                part += th[p + i] * rowsum;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if ((tid & 31) == 0) red[tid >> 5] = part;
        __syncthreads();
        if (tid == 0) {
            double quad = 0.0;
            for (int w = 0; w < T / 32; w++) quad += red[w];
            const double rate = 0.5 * quad + hp->tau_rate;
            const double tau = dbg.tau_in ? dbg.tau_in[cr] : v.gam[c] / rate;
            red[T / 32] = tau;
            if (dbg.enabled && dbg.out.tau_rate && c < v.C) dbg.out.tau_rate[c] = rate;
        }
        __syncthreads();
    }
    const double tau = red[T / 32];

    RSRS_MARK(1);
    // ---- stage the joint Gram: XX block, KX rows, and Lambda = K' Om K + tau Qrsr in blocked form ----
    for (int idx = tid; idx < nblk * BLK; idx += T) Bs[idx] = 0.0;
    for (int idx = tid; idx < p * p; idx += T) {
        const int a = idx / p, b = idx - a * p;
        gxx[idx] = (b <= a) ? Grow[tri(a, b)] : Grow[tri(b, a)];
    }
    for (int idx = tid; idx < q4 * p; idx += T) {
        const int i = idx / p, a = idx - i * p;
        gx[idx] = i < q ? Grow[tri(p + i, a)] : 0.0;
    }
    __syncthreads();
    {   // Lambda = K' Om K + tau Qrsr.  One warp per row (contiguous in the packed Gram), four rows in flight per
        // warp: the loads of a batch are issued before any is consumed (a plain row-by-row loop serialises q
        // global-load latencies; decoding (i, j) from a flat triangle index costs ~30 instructions per element)
        const int warp = tid >> 5, lane = tid & 31, nwarp = T >> 5;
        for (int i0 = warp; i0 < q; i0 += 4 * nwarp) {
            double gv[4][2], qv[4][2];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + u * nwarp;
#pragma unroll
                for (int w = 0; w < 2; w++) {
                    const int j = lane + 32 * w;
                    const bool on = i < q && j <= i;
                    gv[u][w] = on ? Grow[tri(p + i, p) + j] : 0.0;
                    qv[u][w] = on ? v.Qrsr[(size_t)i * q + j] : 0.0;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int i = i0 + u * nwarp;
#pragma unroll
                for (int w = 0; w < 2; w++) {
                    const int j = lane + 32 * w;
                    if (i < q && j <= i) Bs[bidx(i >> 2, j >> 2) + (i & 3) * 4 + (j & 3)] = gv[u][w] + tau * qv[u][w];
                }
            }
            for (int i = i0; i < q && i < i0 + 4 * nwarp; i += nwarp)   // columns beyond 64 (q > 64)
                for (int j = 64 + lane; j <= i; j += 32)
                    Bs[bidx(i >> 2, j >> 2) + (i & 3) * 4 + (j & 3)] = Grow[tri(p + i, p) + j] + tau * v.Qrsr[(size_t)i * q + j];
        }
    }
    for (int i = q + tid; i < q4; i += T) Bs[bidx(i >> 2, i >> 2) + (i & 3) * 5] = 1.0;  // identity padding
    __syncthreads();
    // right-hand side b = K'k - (K' Om X) beta  (logit.py:481) into row 0 of block row qb
    for (int i = tid; i < q; i += T) {
        double b = h[p + i];
        for (int a = 0; a < p; a++) b -= gx[i * p + a] * th[a];
        Bs[bidx(qb, i >> 2) + (i & 3)] = b;
        if (dbg.enabled && dbg.out.eta_b && c < v.C) dbg.out.eta_b[(size_t)c * q + i] = b;
    }
    __syncthreads();
    if (dbg.enabled && dbg.out.eta_prec && c < v.C)
        for (int idx = tid; idx < q * q; idx += T) {
            const int i = idx / q, j = idx - i * q;
            const int ii = i >= j ? i : j, jj = i >= j ? j : i;
            dbg.out.eta_prec[(size_t)c * q * q + idx] = Bs[bidx(ii >> 2, jj >> 2) + (ii & 3) * 4 + (jj & 3)];
        }

    RSRS_MARK(2);
    // ---- blocked right-looking Cholesky of the augmented matrix ----
    for (int J = 0; J < qb; J++) {
        // (1) panel: blocks (I, J), I = J .. qb
        for (int I = J + tid; I <= qb; I += T) {
            double d[4][4];
            const double* dj = Bs + bidx(J, J);
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) d[a][b] = dj[a * 4 + b];
            double invd[4];
            if (!chol4(d, invd)) {
                ok_flag = 0;
                continue;
            }
            double* blk = Bs + bidx(I, J);
            if (I == J) {
#pragma unroll
                for (int a = 0; a < 4; a++) {
#pragma unroll
                    for (int b = 0; b < 4; b++) Ld[J * 16 + a * 4 + b] = b <= a ? d[a][b] : 0.0;
                    Li[J * 4 + a] = invd[a];
                }
            } else {  // P = A_IJ L_JJ^-T, row by row
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    double x[4];
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        double t = blk[a * 4 + b];
#pragma unroll
                        for (int k = 0; k < b; k++) t -= x[k] * d[b][k];
                        x[b] = t * invd[b];
                    }
#pragma unroll
                    for (int b = 0; b < 4; b++) blk[a * 4 + b] = x[b];
                }
            }
        }
        RSRS_MARK(3);
        __syncthreads();
        if (!ok_flag) break;
        // (2) trailing update: blocks (J+1+ii, J+1+kk), kk <= ii <= qb-J-1
        const int m1 = qb - J;  // block rows below J, including the right-hand-side row
        const int cnt = m1 * (m1 + 1) / 2;
        for (int t = tid; t < cnt; t += T) {
            const int I = J + 1 + pair_i[t], K = J + 1 + pair_k[t];
            if (K > qb - 1 + (I == qb ? 0 : 1)) continue;  // (qb, qb) does not exist
            const double* PI = Bs + bidx(I, J);
            const double* PK = Bs + bidx(K, J);
            double* A = Bs + bidx(I, K);
            double pk[4][4];
#pragma unroll
            for (int e = 0; e < 16; e++) pk[e >> 2][e & 3] = PK[e];
#pragma unroll
            for (int a = 0; a < 4; a++) {
                const double p0 = PI[a * 4], p1 = PI[a * 4 + 1], p2 = PI[a * 4 + 2], p3 = PI[a * 4 + 3];
#pragma unroll
                for (int b = 0; b < 4; b++)
                    A[a * 4 + b] -= p0 * pk[b][0] + p1 * pk[b][1] + p2 * pk[b][2] + p3 * pk[b][3];
            }
        }
        RSRS_MARK(4);
        __syncthreads();
        RSRS_MARK(5);
    }
    if (!ok_flag) {
        if (tid == 0) {
            v.status[c] = 1;
            if (c < v.C) atomicAdd(&v.ctl->fails, 1);
        }
        return;  // chain is frozen at its previous state; the failure is reported to the host
    }
    if (dbg.enabled && dbg.out.eta_chol && c < v.C)
        for (int idx = tid; idx < q * q; idx += T) {
            const int i = idx / q, j = idx - i * q;
            const int I = i >> 2, K = j >> 2, e = (i & 3) * 4 + (j & 3);
            dbg.out.eta_chol[(size_t)c * q * q + idx] = (j <= i) ? (I == K ? Ld[I * 16 + e] : Bs[bidx(I, K) + e]) : 0.0;
        }

    RSRS_MARK(6);
    // ---- backward, right-looking over block rows: L' m = y (mean) and L' e = eps ----
    // thread J owns the two 4-vectors of block row J
    double rm[4] = {0, 0, 0, 0}, re[4] = {0, 0, 0, 0};
    if (tid < qb) {
#pragma unroll
        for (int b = 0; b < 4; b++) {
            rm[b] = Bs[bidx(qb, tid) + b];  // y = L^-1 b (row 0 of the extra block row)
            re[b] = nrm[tid * 4 + b];
        }
    }
    for (int I = qb - 1; I >= 0; I--) {
        if (tid == I) {  // x_I = L_II^-T rhs_I
            const double* L = Ld + I * 16;
#pragma unroll
            for (int a = 3; a >= 0; a--) {
                double tm = rm[a], te = re[a];
#pragma unroll
                for (int k = a + 1; k < 4; k++) {
                    tm -= L[k * 4 + a] * rm[k];
                    te -= L[k * 4 + a] * re[k];
                }
                rm[a] = tm * Li[I * 4 + a];
                re[a] = te * Li[I * 4 + a];
            }
#pragma unroll
            for (int a = 0; a < 4; a++) {
                xm[I * 4 + a] = rm[a];
                xe[I * 4 + a] = re[a];
            }
        }
        __syncthreads();
        if (tid < I) {  // rhs_J -= L_IJ' x_I
            const double* L = Bs + bidx(I, tid);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const double m = xm[I * 4 + k], e = xe[I * 4 + k];
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    rm[b] -= L[k * 4 + b] * m;
                    re[b] -= L[k * 4 + b] * e;
                }
            }
        }
    }
    __syncthreads();
    if (dbg.enabled && dbg.out.eta_mean && c < v.C)
        for (int i = tid; i < q; i += T) dbg.out.eta_mean[(size_t)c * q + i] = xm[i];
    for (int i = tid; i < q; i += T) th[p + i] = xm[i] + xe[i];  // eta = mean + L^-T eps
    __syncthreads();

    RSRS_MARK(7);
    // ---- beta | .   A = X' Om X + b_prec,  b = X'k - (X' Om K) eta + b_prec b_mu   (logit.py:226-232) ----
    for (int idx = tid; idx < p * p; idx += T) small_S[idx] = gxx[idx] + hp->b_prec[idx];
    for (int a = tid >> 5; a < p; a += T >> 5) {   // one warp per component: lanes along eta, fixed-order reduction
        double acc = 0.0;
        for (int j = tid & 31; j < q; j += 32) acc += gx[j * p + a] * th[p + j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((tid & 31) == 0) small_b[a] = h[a] + hp->b_pm[a] - acc;
    }
    __syncthreads();
    if (dbg.enabled && c < v.C) {
        if (dbg.out.beta_prec)
            for (int idx = tid; idx < p * p; idx += T) dbg.out.beta_prec[(size_t)c * p * p + idx] = small_S[idx];
        if (dbg.out.beta_b)
            for (int a = tid; a < p; a += T) dbg.out.beta_b[(size_t)c * p + a] = small_b[a];
    }
    __syncthreads();
    if (tid == 0) {
        if (small_precision_mvnorm(small_S, p, small_b, nrm + q4, small_x, small_t)) {
            for (int a = 0; a < p; a++) th[a] = small_x[a];
        } else {
            ok_flag = 0;
        }
    }
    __syncthreads();

    if (!ok_flag) {
        if (tid == 0) {
            v.status[c] = 1;
            if (c < v.C) atomicAdd(&v.ctl->fails, 1);
        }
        return;
    }

    RSRS_MARK(8);
    // ---- commit state, record the draw (Chain.append, chain.py:56-74) ----
    for (int a = tid; a < pq; a += T) v.theta[(size_t)a * Cp + c] = th[a];
    if (tid == 0) v.tau[c] = tau;
    if (rec_beta && rec_index >= 0 && rec_index < rec_capacity && c < v.C) {
        const size_t row = (size_t)c * rec_capacity + rec_index;
        for (int a = tid; a < p; a += T) rec_beta[row * p + a] = th[a];
        if (tid == 0) rec_tau[row] = tau;
    }
#ifdef RSRS_PROF
    RSRS_MARK(9);
    if (c == 0 && tid == 0 && v.ctl->iter == 3)
        printf("rsr solve chain 0 (thread 0): prologue %lld | tau %lld | stage-in %lld | chol: panel %lld, barrier + update %lld, "
               "barrier %lld | post %lld | backward %lld | beta %lld | commit %lld cycles\n",
               rs_t[0], rs_t[1], rs_t[2], rs_t[3], rs_t[4], rs_t[5], rs_t[6], rs_t[7], rs_t[8], rs_t[9]);
#endif
}
