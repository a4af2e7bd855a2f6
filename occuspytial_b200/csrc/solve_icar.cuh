// solve_icar.cuh -- LogitICARGibbs ("PGGibbs"): fixed-pattern banded Cholesky solves.
// Part of the single translation unit engine.cu (included there, in order; not a stand-alone header).
#pragma once

// ------------------------------------------------------------------------------------------------
// LogitICARGibbs ("PGGibbs"): eta | . ~ N(Lambda^-1 b, Lambda^-1) on {1'eta = 0}, Lambda = tau Q + diag(omega_b)
//
// The reference perturbs b through a dense eigen-factor of Q (2n - 1 normals) and solves two systems with
// MINRES at rtol 1e-5 (logit.py:73-99).  Here Lambda -- whose sparsity pattern is fixed by Q -- is factorised
// exactly per chain by a banded Cholesky over the fixed symbolic pattern (half bandwidth bw after the host's
// reordering), the draw is x = Lambda^-1 b + L^-T eps (n normals, same law), zv = Lambda^-1 1, and
// eta = x - (sum x / sum zv) zv (ensure_sums_to_zero, distributions.pyx:24-39).
//
// One CTA per chain.  Right-looking column Cholesky on a (bw+1) x (bw+1) ring-indexed window in shared
// memory (no data movement when the window slides), forward substitution of both right-hand sides fused
// into the factorisation, L streamed to HBM once in column-band form and read back once by the backward pass.
// ------------------------------------------------------------------------------------------------
constexpr int ICAR_THREADS = 128;

// psi = X beta + eta   (logit.py:201 with spatial = eta)
__global__ void psi_icar_kernel(const View v) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)v.n * v.Cp;
    if (idx >= total) return;
    const int i = (int)(idx / v.Cp), c = (int)(idx - (size_t)i * v.Cp);
    double s = v.eta_n[idx];
    for (int a = 0; a < v.p; a++) s += v.X[(size_t)i * v.p + a] * v.theta[(size_t)a * v.Cp + c];
    v.psi[idx] = s;
}

__device__ __forceinline__ double block_sum_128(double val, double* red, int tid) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = val;
    __syncthreads();
    return red[0] + red[1] + red[2] + red[3];
}

__global__ void __launch_bounds__(ICAR_THREADS) solve_logit_icar_kernel(const View v, const SolveDebug dbg,
                                                                         double* rec_alpha, double* rec_beta,
                                                                         double* rec_tau, int rec_capacity) {
    extern __shared__ double sm[];
    const int c = blockIdx.x, tid = threadIdx.x, T = ICAR_THREADS;
    const int n = v.n, p = v.p, r = v.r, w = v.bw, W1 = v.bw + 1;
    const size_t Cp = v.Cp;
    const HParams* hp = v.hp;
    double* Wn = sm;                     // [W1][W1] ring window, element (i, k) at Wn[(i % W1) * W1 + (k % W1)]
    double* col = Wn + W1 * W1;          // [W1] current column of L
    double* r1 = col + W1;               // [W1] ring: forward residual of b
    double* r2 = r1 + W1;                // [W1] ring: forward residual of 1
    double* xr = r2 + W1;                // [3][W1] rings of the backward solves
    double* nrm = xr + 3 * W1;           // [p + r] normals of beta, alpha
    double* red = nrm + 2 * MAX_SMALL;   // [8]
    uint8_t* pair_i = reinterpret_cast<uint8_t*>(red + 8);  // [w (w+1) / 2] trailing-update pair table
    uint8_t* pair_k = pair_i + (w * (w + 1)) / 2;
    __shared__ double small_S[MAX_SMALL * MAX_SMALL];
    __shared__ double small_b[MAX_SMALL], small_x[MAX_SMALL], small_t[MAX_SMALL], beta_s[MAX_SMALL];
    __shared__ double ybc[2];
    __shared__ int ok_flag;

    const uint32_t iter = v.ctl->iter;
    const int rec_index = v.ctl->rec_index;
    const uint32_t gc = (uint32_t)(v.chain_offset + c);
    const int cr = c % v.C;
    double* Lc = v.Lc + (size_t)c * n * W1;
    double* yb = v.scratch + (size_t)c * 8 * round_up_dev(n, 4);  // forward results / backward inputs
    double* y1 = yb + n;
    double* xs = y1 + n;
    double* zs = xs + n;
    double* ms = zs + n;

    for (int a = tid; a < p; a += T) beta_s[a] = v.theta[(size_t)a * Cp + c];
    for (int k = tid; k < p + r; k += T) {
        double val;
        if (dbg.normals) {
            val = dbg.normals[(size_t)cr * (n + p + r) + n + k];
        } else {
            Philox g;
            g.init(v.seed, gc, (uint32_t)(n + k), iter, ST_NORMALS);
            val = g.normal_d();
        }
        nrm[k] = val;
    }
    {
        int t = 0;  // pair table (ii, kk), kk <= ii < w, built redundantly but cheaply by thread 0..: parallel fill
        for (int ii = 0; ii < w; ii++) {
            for (int kk = tid; kk <= ii; kk += T) {
                pair_i[t + kk] = (uint8_t)ii;
                pair_k[t + kk] = (uint8_t)kk;
            }
            t += ii + 1;
        }
    }
    if (tid == 0) ok_flag = 1;
    __syncthreads();

    // ---- tau | eta: rate = eta' Q eta / 2 + tau_rate (logit.py:206-209), Q in lower-band form ----
    double part = 0.0;
    for (int i = tid; i < n; i += T) {
        const double ei = v.eta_n[(size_t)i * Cp + c];
        double s = 0.5 * v.Qband[(size_t)i * W1] * ei;
        const int dmax = min(w, i);
        for (int d = 1; d <= dmax; d++) s += v.Qband[(size_t)i * W1 + d] * v.eta_n[(size_t)(i - d) * Cp + c];
        part += ei * s;  // eta' Q eta = 2 * sum_i eta_i (Q_ii eta_i / 2 + sum_{k<i} Q_ik eta_k)
    }
    const double quad = 2.0 * block_sum_128(part, red, tid);
    __syncthreads();
    if (tid == 0) {
        const double rate = 0.5 * quad + hp->tau_rate;
        const double tau = dbg.tau_in ? dbg.tau_in[cr] : v.gam[c] / rate;
        red[4] = tau;
        if (dbg.enabled && dbg.out.tau_rate && c < v.C) dbg.out.tau_rate[c] = rate;
    }
    __syncthreads();
    const double tau = red[4];

    // Stage this chain's per-site inputs CONTIGUOUSLY (chain-major scratch) before the sequential loops: the
    // chain-minor state arrays would cost one uncoalesced DRAM round trip per row inside the dependency chain,
    // and the Box-Muller normals of the backward solve must not sit on its critical path either.
    double* om_c = ms + n;  // [n] omega of this chain ("spare" slot of the scratch block)
    for (int i = tid; i < n; i += T) {
        const double om = v.omega[(size_t)i * Cp + c];
        double xb = 0.0;
        for (int a = 0; a < p; a++) xb += v.X[(size_t)i * p + a] * beta_s[a];
        const double bi = v.kd[(size_t)i * Cp + c] - om * xb;  // b_i = k_i - omega_i x_i'beta   (logit.py:213)
        om_c[i] = om;
        xs[i] = bi;  // xs / zs are free until the backward pass: b and the normals live there meanwhile
        if (dbg.enabled && dbg.out.eta_b && c < v.C) dbg.out.eta_b[(size_t)c * n + i] = bi;
        double ez;
        if (dbg.normals) {
            ez = dbg.normals[(size_t)cr * (n + p + r) + i];
        } else {
            Philox g;
            g.init(v.seed, gc, (uint32_t)i, iter, ST_NORMALS);
            ez = g.normal_d();
        }
        zs[i] = ez;
    }
    __syncthreads();

    // row i of Lambda (band part) and of the right-hand sides enters ring slot i % W1
    auto load_row = [&](int i) {
        const int slot = i % W1;
        for (int d = tid; d < W1; d += T) {  // k = i - d
            double val = 0.0;
            if (i < n) {
                if (d <= i) val = tau * v.Qband[(size_t)i * W1 + d];
                if (d == 0) val += om_c[i];
            } else if (d == 0) {
                val = 1.0;  // identity padding past the last row
            }
            const int k = i - d;
            const int kslot = ((k % W1) + W1) % W1;
            Wn[slot * W1 + kslot] = val;
        }
        if (tid == 0) {
            // b_i = k_i - omega_i x_i'beta   (logit.py:213)
            double bi = 0.0, one = 0.0;
            if (i < n) {
                bi = xs[i];
                one = 1.0;
            }
            r1[slot] = bi;
            r2[slot] = one;
        }
    };
    for (int i = 0; i <= w; i++) load_row(i);
    __syncthreads();

    // ---- factorisation + fused forward substitution ----
    const int npairs = w * (w + 1) / 2;
    for (int j = 0; j < n; j++) {
        const int jm = j % W1;
        const double piv = Wn[jm * W1 + jm];
        if (!(piv > 0.0)) {
            if (tid == 0) {
                v.status[c] = 1;
                if (c < v.C) atomicAdd(&v.ctl->fails, 1);
            }
            return;
        }
        const double inv = rsqrt(piv), dj = piv * inv;  // one rsqrt on the critical path instead of sqrt + divide
        if (tid == 0) om_c[j] = inv;                    // om_c[j] is dead (row j entered long ago): keep 1 / L_jj there
        for (int e = tid; e < W1; e += T) {
            int slot = jm + e;
            if (slot >= W1) slot -= W1;
            const double le = (e == 0) ? dj : Wn[slot * W1 + jm] * inv;
            col[e] = le;
            Lc[(size_t)j * W1 + e] = le;
        }
        if (tid == T - 1) {
            const double a = r1[jm] * inv, b = r2[jm] * inv;
            ybc[0] = a;
            ybc[1] = b;
            yb[j] = a;
            y1[j] = b;
        }
        __syncthreads();
        const int j1 = (jm + 1 == W1) ? 0 : jm + 1;
        for (int t = tid; t < npairs; t += T) {
            int si = j1 + pair_i[t], sk = j1 + pair_k[t];
            if (si >= W1) si -= W1;
            if (sk >= W1) sk -= W1;
            Wn[si * W1 + sk] -= col[1 + pair_i[t]] * col[1 + pair_k[t]];
        }
        for (int e = 1 + tid; e <= w; e += T) {  // forward residuals of rows j+1 .. j+w
            int slot = jm + e;
            if (slot >= W1) slot -= W1;
            r1[slot] -= col[e] * ybc[0];
            r2[slot] -= col[e] * ybc[1];
        }
        load_row(j + W1);  // reuses slot jm: row / column j are no longer referenced
        __syncthreads();
    }
    if (dbg.enabled && dbg.out.eta_chol && c < v.C)  // column-band factor [n][bw + 1]
        for (size_t idx = tid; idx < (size_t)n * W1; idx += T) dbg.out.eta_chol[(size_t)c * n * W1 + idx] = Lc[idx];
    __threadfence_block();
    __syncthreads();

    // ---- backward: L' x = y_b + eps, L' zv = y_1, (debug) L' m = y_b; one warp, lanes over the band ----
    if (tid < 32) {
        for (int e = tid; e < 3 * W1; e += 32) xr[e] = 0.0;
        __syncwarp();
        constexpr int NL = 5;  // band entries per lane (bw <= 160)
        double lnext[NL], sc[4] = {0.0, 0.0, 0.0, 0.0};
        auto fetch = [&](int j) {  // loads of column j: independent of the recurrence, issued one step ahead
#pragma unroll
            for (int u = 0; u < NL; u++) {
                const int e = 1 + tid + 32 * u;
                lnext[u] = (e <= w && j + e < n) ? Lc[(size_t)j * W1 + e] : 0.0;
            }
            if (tid == 0) {
                sc[0] = yb[j];
                sc[1] = y1[j];
                sc[2] = zs[j];    // staged normal
                sc[3] = om_c[j];  // 1 / L_jj
            }
        };
        fetch(n - 1);
        for (int j = n - 1; j >= 0; j--) {
            const int jm = j % W1;
            double lcur[NL];
#pragma unroll
            for (int u = 0; u < NL; u++) lcur[u] = lnext[u];
            const double ybj = sc[0], y1j = sc[1], ez = sc[2], inv = sc[3];
            if (j > 0) fetch(j - 1);
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
            for (int u = 0; u < NL; u++) {
                const int e = 1 + tid + 32 * u;
                if (e <= w) {
                    int slot = jm + e;
                    if (slot >= W1) slot -= W1;
                    s0 += lcur[u] * xr[slot];
                    s1 += lcur[u] * xr[W1 + slot];
                    s2 += lcur[u] * xr[2 * W1 + slot];
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            if (tid == 0) {
                const double x = (ybj + ez - s0) * inv, z = (y1j - s1) * inv, m = (ybj - s2) * inv;
                xr[jm] = x;
                xr[W1 + jm] = z;
                xr[2 * W1 + jm] = m;
                xs[j] = x;
                zs[j] = z;
                ms[j] = m;
            }
            __syncwarp();
        }
    }
    __threadfence_block();
    __syncthreads();

    // ---- eta = x - (sum x / sum zv) zv (distributions.pyx:24-39); beta system (logit.py:226-231) ----
    double sx = 0.0, sz = 0.0;
    for (int i = tid; i < n; i += T) {
        sx += xs[i];
        sz += zs[i];
    }
    sx = block_sum_128(sx, red, tid);
    sz = block_sum_128(sz, red, tid);
    const double shift = -sx / sz;
    for (int i = tid; i < n; i += T) {
        const double e = xs[i] + shift * zs[i];
        xs[i] = e;
        v.eta_n[(size_t)i * Cp + c] = e;
        if (dbg.enabled && dbg.out.eta_mean && c < v.C) dbg.out.eta_mean[(size_t)c * n + i] = ms[i];
    }
    __syncthreads();
    for (int idx = 0; idx < p * p; idx++) {  // A = X' Om X + b_prec
        const int a = idx / p, b = idx - a * p;
        if (b > a) continue;
        double s = 0.0;
        for (int i = tid; i < n; i += T) s += v.omega[(size_t)i * Cp + c] * v.X[(size_t)i * p + a] * v.X[(size_t)i * p + b];
        s = block_sum_128(s, red, tid);
        if (tid == 0) {
            small_S[a * p + b] = s + hp->b_prec[a * p + b];
            small_S[b * p + a] = s + hp->b_prec[b * p + a];
        }
    }
    for (int a = 0; a < p; a++) {  // b = X'(k - omega * eta) + b_prec b_mu
        double s = 0.0;
        for (int i = tid; i < n; i += T)
            s += v.X[(size_t)i * p + a] * (v.kd[(size_t)i * Cp + c] - v.omega[(size_t)i * Cp + c] * xs[i]);
        s = block_sum_128(s, red, tid);
        if (tid == 0) small_b[a] = s + hp->b_pm[a];
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
        if (small_precision_mvnorm(small_S, p, small_b, nrm, small_x, small_t)) {
            for (int a = 0; a < p; a++) beta_s[a] = small_x[a];
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
    for (int a = tid; a < p; a += T) v.theta[(size_t)a * Cp + c] = beta_s[a];
    if (tid == 0) v.tau[c] = tau;
    if (rec_beta && rec_index >= 0 && rec_index < rec_capacity && c < v.C) {
        const size_t row = (size_t)c * rec_capacity + rec_index;
        for (int a = tid; a < p; a += T) rec_beta[row * p + a] = beta_s[a];
        if (tid == 0) rec_tau[row] = tau;
    }
}

// ------------------------------------------------------------------------------------------------
// ICAR solve on the FP64 tensor pipe (half bandwidth bw <= 60): 4-column blocked banded Cholesky whose
// rank-4 trailing update A -= P P' is exactly the k = 4 of DMMA.8x8x4.
//
// The sliding window of the factorisation covers matrix indices [4s, 4s + W8), W8 = 8 NT >= bw + 4, and is
// held ENTIRELY IN REGISTERS as m8n8 accumulator fragments, addressed by ring slot (index mod W8) so that
// sliding the window moves no data: 4 slots leave, 4 enter, per block step.  Two extra rows (tile row NT)
// carry the right-hand sides b and 1, so the forward substitutions ride along in the same MMAs.
// Per block step (2 barriers): (1) the owners of block-column J publish their 4 raw columns to shared
// memory; (2) one thread per window row factorises the 4x4 diagonal block redundantly (rsqrt, no divide),
// solves its row against it, emits L (column-band layout) and the panel P; (3) every warp updates its tiles
// with DMMA and overwrites the 4 entering rows / columns with tau Q + diag(omega).
// The backward pass is blocked the same way: 4 warps reduce the 4 columns of a block against the ring of
// already-known x in parallel, one lane finishes the 4x4 triangular solve; L is prefetched one block ahead.
// ------------------------------------------------------------------------------------------------
template <int NT>
// 7 resident CTAs per SM (72 registers, < 32 KB of shared memory): 1036 chain slots on 148 SMs, so that the 1024 chains
// of BASELINE config 4 are one wave instead of 1.4 (measured 7.7 -> 6.6 ms per iteration); the widest bands keep 5
__global__ void __launch_bounds__(128, NT <= 7 ? 7 : 5) solve_logit_icar_mma_kernel(const View v, const SolveDebug dbg, double* rec_beta,
                                                                    double* rec_tau, int rec_capacity) {
    constexpr int W8 = 8 * NT, T = 128, RPW = (NT + 1 + 3) / 4;  // tile rows per warp (incl. the RHS tile row NT)
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int n = v.n, n4 = round_up_dev(v.n, 4), p = v.p, w = v.bw, W1 = v.bw + 1;
    const size_t Cp = v.Cp;
    const HParams* hp = v.hp;
    __shared__ double raw[W8 + 8][4];
    __shared__ double P[W8 + 8][4];
    __shared__ double xr[3][W8];          // rings of the backward solves
    __shared__ double S4[4][3];           // backward: the four column sums of a block
    constexpr int NBF = 2, NBB = 4;       // prefetch depth (block steps) of the forward / backward pass
    __shared__ double qin[NBF][4][64];    // ring: lower-band rows of Q for the 4 matrix rows entering at a block step
    __shared__ double oin[NBF][4], bin[NBF][4];  // their omega and b
    __shared__ double lring[NBB][4][64];  // backward ring: the 4 columns of L of a block (column-band layout)
    __shared__ double sring[NBB][4][4];   // backward ring: y_b, y_1, normals, 1 / L_jj of the block
    __shared__ double red[8];
    __shared__ double small_S[MAX_SMALL * MAX_SMALL];
    __shared__ double small_b[MAX_SMALL], small_x[MAX_SMALL], small_t[MAX_SMALL], beta_s[MAX_SMALL], nrm[MAX_SMALL];
    __shared__ int ok_flag;

    const uint32_t iter = v.ctl->iter;
    const int rec_index = v.ctl->rec_index;
    const uint32_t gc = (uint32_t)(v.chain_offset + c);
    const int cr = c % v.C;
    double* Lc = v.Lc + (size_t)c * n * W1;
    double* yb = v.scratch + (size_t)c * 8 * n4;
    double* y1 = yb + n4;
    double* xs = y1 + n4;    // b (forward) then x (backward)
    double* zs = xs + n4;    // normals (forward) then zv (backward)
    double* ms = zs + n4;    // L^-T L^-1 b (debug mean)
    double* om_c = ms + n4;  // omega of this chain
    double* inv_c = om_c + n4;  // 1 / L_jj
    double* kd_c = inv_c + n4;  // z - 1/2 of this chain

    for (int a = tid; a < p; a += T) beta_s[a] = v.theta[(size_t)a * Cp + c];
    for (int k = tid; k < p; k += T) {
        double val;
        if (dbg.normals) {
            val = dbg.normals[(size_t)cr * (n + p + v.r) + n + k];
        } else {
            Philox gph;
            gph.init(v.seed, gc, (uint32_t)(n + k), iter, ST_NORMALS);
            val = gph.normal_d();
        }
        nrm[k] = val;
    }
    if (tid == 0) ok_flag = 1;
    __syncthreads();

    // ---- tau | eta (logit.py:206-209); eta is first copied chain-major so that the band products hit L1 ----
    for (int i = tid; i < n; i += T) ms[i] = v.eta_n[(size_t)i * Cp + c];
    __syncthreads();
    double part = 0.0;
    for (int i = tid; i < n; i += T) {
        const double ei = ms[i];
        const double* qrow = v.Qband + (size_t)i * W1;
        double sacc = 0.5 * qrow[0] * ei;
        const int dmax = min(w, i);
        for (int d = 1; d <= dmax; d++) sacc += qrow[d] * ms[i - d];
        part += ei * sacc;
    }
    const double quad = 2.0 * block_sum_128(part, red, tid);
    __syncthreads();
    if (tid == 0) {
        const double rate = 0.5 * quad + hp->tau_rate;
        red[4] = dbg.tau_in ? dbg.tau_in[cr] : v.gam[c] / rate;
        if (dbg.enabled && dbg.out.tau_rate && c < v.C) dbg.out.tau_rate[c] = rate;
    }
    __syncthreads();
    const double tau = red[4];

    // ---- stage this chain's per-site inputs contiguously (see solve_logit_icar_kernel) ----
    for (int i = tid; i < n4; i += T) {
        double om = 0.0, bi = 0.0, ez = 0.0, kk = 0.0;
        if (i < n) {
            om = v.omega[(size_t)i * Cp + c];
            kk = v.kd[(size_t)i * Cp + c];
            double xb = 0.0;
            for (int a = 0; a < p; a++) xb += v.X[(size_t)i * p + a] * beta_s[a];
            bi = kk - om * xb;  // logit.py:213
            if (dbg.enabled && dbg.out.eta_b && c < v.C) dbg.out.eta_b[(size_t)c * n + i] = bi;
            if (dbg.normals) {
                ez = dbg.normals[(size_t)cr * (n + p + v.r) + i];
            } else {
                Philox gph;
                gph.init(v.seed, gc, (uint32_t)i, iter, ST_NORMALS);
                ez = gph.normal_d();
            }
        }
        om_c[i] = om;
        kd_c[i] = kk;
        xs[i] = bi;
        zs[i] = ez;
    }
    __syncthreads();

    // entering rows i0 .. i0 + 3 -> staging buffer `buf` (plain loads issued a whole block step before use)
    // (each thread stages at most two band entries; their (row, offset) coordinates never change)
    const int st_a0 = tid / W1, st_d0 = tid - st_a0 * W1;
    const int st_a1 = (tid + T) / W1, st_d1 = (tid + T) - st_a1 * W1;
    // 8-byte cp.async with zero fill when `ok` is false: the copies land while the block steps in between run
    auto cp8 = [&](double* dst, const double* src, bool ok) {
        const uint32_t sdst = (uint32_t)__cvta_generic_to_shared(dst);
        const int bytes = ok ? 8 : 0;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sdst), "l"(src), "r"(bytes));
    };
    auto stage_rows = [&](int i0, int buf) {
        {
            const int i = i0 + st_a0;
            const bool ok = st_a0 < 4 && i < n && st_d0 <= i;
            if (st_a0 < 4) cp8(&qin[buf][st_a0][st_d0], ok ? v.Qband + (size_t)i * W1 + st_d0 : v.Qband, ok);
        }
        if (st_a1 < 4) {
            const int i = i0 + st_a1;
            const bool ok = i < n && st_d1 <= i;
            cp8(&qin[buf][st_a1][st_d1], ok ? v.Qband + (size_t)i * W1 + st_d1 : v.Qband, ok);
        }
        if (tid < 8) {
            const int a = tid & 3, i = i0 + a;
            const double* srcv = (tid < 4) ? om_c : xs;
            cp8(tid < 4 ? &oin[buf][a] : &bin[buf][a], i < n ? srcv + i : srcv, i < n);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    // Lambda[i][k] for an entering row or column (max(i, k) in [i0, i0 + 4)), from the staging buffer
    auto lam_in = [&](int i, int k, int i0, int buf) -> double {
        if (i < k) {
            const int t2 = i;
            i = k;
            k = t2;
        }
        const int d = i - k;
        if (i >= n) return d == 0 ? 1.0 : 0.0;
        if (d > w) return 0.0;
        double val = tau * qin[buf][i - i0][d];
        if (d == 0) val += oin[buf][i - i0];
        return val;
    };
    // Lambda[i][k] (identity beyond n) and the right-hand-side rows
    auto lam = [&](int i, int k) -> double {
        if (i < k) {
            const int t2 = i;
            i = k;
            k = t2;
        }
        const int d = i - k;
        if (i >= n) return d == 0 ? 1.0 : 0.0;
        if (d > w) return 0.0;
        double val = tau * v.Qband[(size_t)i * W1 + d];
        if (d == 0) val += om_c[i];
        return val;
    };
    auto rhs = [&](int row, int k) -> double { return (k < n) ? (row == 0 ? xs[k] : (row == 1 ? 1.0 : 0.0)) : 0.0; };

    // ---- window in registers ----
    double acc[RPW][NT][2];
#pragma unroll
    for (int ri = 0; ri < RPW; ri++) {
        const int I = warp + 4 * ri;
#pragma unroll
        for (int K = 0; K < NT; K++) {
            const int sc = 8 * K + 2 * tq;
            if (I < NT) {
                acc[ri][K][0] = lam(8 * I + g, sc);
                acc[ri][K][1] = lam(8 * I + g, sc + 1);
            } else if (I == NT) {
                acc[ri][K][0] = rhs(g, sc);
                acc[ri][K][1] = rhs(g, sc + 1);
            } else {
                acc[ri][K][0] = acc[ri][K][1] = 0.0;
            }
        }
    }

    // run f(K) with K a compile-time constant equal to the (warp-uniform) run-time tile column kc
    auto with_tile_col = [&](int kc, auto f) {
        switch (kc) {
            case 0: f(std::integral_constant<int, 0>{}); break;
            case 1: if constexpr (NT > 1) f(std::integral_constant<int, 1>{}); break;
            case 2: if constexpr (NT > 2) f(std::integral_constant<int, 2>{}); break;
            case 3: if constexpr (NT > 3) f(std::integral_constant<int, 3>{}); break;
            case 4: if constexpr (NT > 4) f(std::integral_constant<int, 4>{}); break;
            case 5: if constexpr (NT > 5) f(std::integral_constant<int, 5>{}); break;
            case 6: if constexpr (NT > 6) f(std::integral_constant<int, 6>{}); break;
            case 7: if constexpr (NT > 7) f(std::integral_constant<int, 7>{}); break;
        }
    };

    // ---- blocked factorisation + fused forward substitutions ----
    int jm4 = 0;
    for (int pre = 0; pre < NBF - 1; pre++) stage_rows(W8 + 4 * pre, pre);  // rows entering at steps 0 .. NBF-2
    for (int j0 = 0; j0 < n4; j0 += 4) {
        const int Kc = jm4 >> 3, h = (jm4 >> 2) & 1;
        const int buf = (j0 >> 2) % NBF, inew = j0 + W8;  // rows inew .. inew + 3 enter at the end of this step
        // (1) publish the raw block column
        with_tile_col(Kc, [&](auto kc) {
            constexpr int K = decltype(kc)::value;
            if ((tq >> 1) == h) {
#pragma unroll
                for (int ri = 0; ri < RPW; ri++) {
                    const int I = warp + 4 * ri;
                    if (I <= NT) {
                        const int row = (I < NT ? 8 * I : W8) + g;
                        raw[row][2 * (tq & 1)] = acc[ri][K][0];
                        raw[row][2 * (tq & 1) + 1] = acc[ri][K][1];
                    }
                }
            }
        });
        __syncthreads();
        // rows entering NBF-1 steps from now -> the ring buffer read last at the previous step (everyone is past
        // it); then wait for THIS step's rows (issued NBF-1 steps ago), published by the barrier below
        stage_rows(inew + 4 * (NBF - 1), ((j0 >> 2) + NBF - 1) % NBF);
        asm volatile("cp.async.wait_group %0;\n" ::"n"(NBF - 1));
        // (2) one thread per window row: factorise the diagonal block, solve the row, emit L and P
        if (tid < W8 + 8) {
            double d[4][4], invd[4];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) d[a][b] = raw[jm4 + a][b];
            if (!chol4(d, invd)) {
                ok_flag = 0;
            } else {
                double x[4];
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    double t = raw[tid][b];
#pragma unroll
                    for (int k = 0; k < b; k++) t -= x[k] * d[b][k];
                    x[b] = t * invd[b];
                }
                if (tid < W8) {
                    int dt = tid - jm4;
                    if (dt < 0) dt += W8;
                    const bool leaving = dt < 4;  // rows of the diagonal block: x_b = L_JJ[dt][b] for b <= dt
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const int e = dt - b;
                        if (e >= 0 && e <= w && j0 + b < n) Lc[(size_t)(j0 + b) * W1 + e] = x[b];
                        P[tid][b] = leaving ? 0.0 : x[b];
                    }
                    if (dt == 0) {
#pragma unroll
                        for (int b = 0; b < 4; b++) inv_c[j0 + b] = invd[b];
                    }
                } else {
                    const int rr = tid - W8;
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        P[tid][b] = rr < 2 ? x[b] : 0.0;
                        if (rr == 0) yb[j0 + b] = x[b];  // y_b = L^-1 b
                        if (rr == 1) y1[j0 + b] = x[b];  // y_1 = L^-1 1
                    }
                }
            }
        }
        __syncthreads();
        if (!ok_flag) break;
        // (3) rank-4 trailing update on the tensor pipe
        const int base = jm4 + 4;  // ring slot of matrix index j0 + 4 after the slide
#pragma unroll
        for (int ri = 0; ri < RPW; ri++) {
            const int I = warp + 4 * ri;
            if (I <= NT) {
                const int prow = (I < NT ? 8 * I : W8) + g;
                const double aF = -P[prow][tq];
#pragma unroll
                for (int K = 0; K < NT; K++) dgemm::dmma884(acc[ri][K][0], acc[ri][K][1], aF, P[8 * K + g][tq]);
            }
        }
        // (4a) the 4 entering ROWS live in tile row Kc: only its owner warp, only the lanes on those rows
#pragma unroll
        for (int ri = 0; ri < RPW; ri++) {
            if (warp + 4 * ri == Kc && (g >> 2) == h) {
                const int ir = inew + (g & 3);
#pragma unroll
                for (int K = 0; K < NT; K++) {
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        int dts = 8 * K + 2 * tq + u - base;
                        if (dts < 0) dts += W8;
                        acc[ri][K][u] = lam_in(ir, j0 + 4 + dts, inew, buf);
                    }
                }
            }
        }
        // (4b) the 4 entering COLUMNS live in tile column Kc: the lanes holding those columns, every tile row
        if ((tq >> 1) == h) {
            const int ic = inew + 2 * (tq & 1);
#pragma unroll
            for (int ri = 0; ri < RPW; ri++) {
                const int I = warp + 4 * ri;
                if (I <= NT) {
                    double v0, v1;
                    if (I < NT) {
                        int dts = 8 * I + g - base;
                        if (dts < 0) dts += W8;
                        const int ir = j0 + 4 + dts;
                        v0 = lam_in(ir, ic, inew, buf);
                        v1 = lam_in(ir, ic + 1, inew, buf);
                    } else {
                        v0 = (ic < n) ? (g == 0 ? bin[buf][ic - inew] : (g == 1 ? 1.0 : 0.0)) : 0.0;
                        v1 = (ic + 1 < n) ? (g == 0 ? bin[buf][ic + 1 - inew] : (g == 1 ? 1.0 : 0.0)) : 0.0;
                    }
                    with_tile_col(Kc, [&](auto kc) {
                        constexpr int K = decltype(kc)::value;
                        acc[ri][K][0] = v0;
                        acc[ri][K][1] = v1;
                    });
                }
            }
        }
        jm4 = base == W8 ? 0 : base;
    }
    if (!ok_flag) {
        if (tid == 0) {
            v.status[c] = 1;
            if (c < v.C) atomicAdd(&v.ctl->fails, 1);
        }
        return;
    }
    if (dbg.enabled && dbg.out.eta_chol && c < v.C)  // column-band factor [n][bw + 1]
        for (size_t idx = tid; idx < (size_t)n * W1; idx += T) {
            const int j = (int)(idx / W1), e = (int)(idx - (size_t)j * W1);
            dbg.out.eta_chol[(size_t)c * n * W1 + idx] = (j + e < n) ? Lc[idx] : 0.0;
        }
    for (int e = tid; e < 3 * W8; e += T) (&xr[0][0])[e] = 0.0;
    __syncthreads();

    // ---- blocked backward: L' x = y_b + eps, L' zv = y_1, L' m = y_b ----
    {
        asm volatile("cp.async.wait_group 0;\n" ::);
        const int b = warp;  // this warp's column within the block
        // block j0 -> ring buffer (j0 / 4) % NBB: 4 columns of L and the 4 x 4 scalars of the finish
        auto stage_block = [&](int j0, int buf) {
            if (j0 >= 0) {
                if (st_a0 < 4) {
                    const int j = j0 + st_a0;
                    const bool ok = j < n;
                    cp8(&lring[buf][st_a0][st_d0], ok ? Lc + (size_t)j * W1 + st_d0 : Lc, ok);
                }
                if (st_a1 < 4) {
                    const int j = j0 + st_a1;
                    const bool ok = j < n;
                    cp8(&lring[buf][st_a1][st_d1], ok ? Lc + (size_t)j * W1 + st_d1 : Lc, ok);
                }
                if (tid < 16) {
                    const int bb = tid & 3, which = tid >> 2, jj = j0 + bb;
                    const double* srcv = which == 0 ? yb : (which == 1 ? y1 : (which == 2 ? zs : inv_c));
                    cp8(&sring[buf][which][bb], jj < n ? srcv + jj : srcv, jj < n);
                }
            }
            asm volatile("cp.async.commit_group;\n" ::);
        };
        const int last = n4 - 4;
        for (int pre = 0; pre < NBB - 1; pre++) stage_block(last - 4 * pre, ((last >> 2) - pre + 4 * NBB) % NBB);
        asm volatile("cp.async.wait_group %0;\n" ::"n"(NBB - 2));
        __syncthreads();
        for (int j0 = last; j0 >= 0; j0 -= 4) {
            const int buf = (j0 >> 2) % NBB;
            const int j = j0 + b;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
            // rows beyond the block (e >= 4 - b) are already solved and sit in the ring; entries past row n are
            // garbage in L's band storage, so they are masked by index rather than relied upon to be zero
            for (int e = lane + 4 - b; e <= w; e += 32) {
                if (j + e < n) {
                    const double l = lring[buf][b][e];
                    const int slot = (j + e) % W8;
                    s0 += l * xr[0][slot];
                    s1 += l * xr[1][slot];
                    s2 += l * xr[2][slot];
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            if (lane == 0) {
                S4[b][0] = s0;
                S4[b][1] = s1;
                S4[b][2] = s2;
            }
            __syncthreads();
            if (tid < 3) {  // one lane per right-hand side finishes the 4x4 triangular solve
                const int rr = tid;
                double fr[4], fi[4], xb4[4];
#pragma unroll
                for (int bb = 0; bb < 4; bb++) {
                    const double ybv = sring[buf][0][bb];
                    fr[bb] = rr == 0 ? ybv + sring[buf][2][bb] : (rr == 1 ? sring[buf][1][bb] : ybv);
                    fi[bb] = sring[buf][3][bb];  // zero beyond n => x = 0 there
                }
                const bool in1 = j0 + 1 < n, in2 = j0 + 2 < n, in3 = j0 + 3 < n;
                const double l10 = in1 ? lring[buf][0][1] : 0.0, l20 = in2 ? lring[buf][0][2] : 0.0;
                const double l30 = in3 ? lring[buf][0][3] : 0.0, l21 = in2 ? lring[buf][1][1] : 0.0;
                const double l31 = in3 ? lring[buf][1][2] : 0.0, l32 = in3 ? lring[buf][2][1] : 0.0;
                xb4[3] = (fr[3] - S4[3][rr]) * fi[3];
                xb4[2] = (fr[2] - S4[2][rr] - l32 * xb4[3]) * fi[2];
                xb4[1] = (fr[1] - S4[1][rr] - l21 * xb4[2] - l31 * xb4[3]) * fi[1];
                xb4[0] = (fr[0] - S4[0][rr] - l10 * xb4[1] - l20 * xb4[2] - l30 * xb4[3]) * fi[0];
#pragma unroll
                for (int bb = 0; bb < 4; bb++) xr[rr][(j0 + bb) % W8] = xb4[bb];
                double* dst = rr == 0 ? xs : (rr == 1 ? zs : ms);   // (zs: this block's normals are in the ring already)
#pragma unroll
                for (int bb = 0; bb < 4; bb++) dst[j0 + bb] = xb4[bb];
            }
            // block NBB-1 steps ahead -> the buffer consumed at the previous step; then wait for the next block
            stage_block(j0 - 4 * (NBB - 1), ((j0 >> 2) + 1) % NBB);
            asm volatile("cp.async.wait_group %0;\n" ::"n"(NBB - 2));
            __syncthreads();
        }
    }

    // ---- eta = x - (sum x / sum zv) zv (distributions.pyx:24-39); beta system (logit.py:226-231) ----
    double sx = 0.0, sz = 0.0;
    for (int i = tid; i < n; i += T) {
        sx += xs[i];
        sz += zs[i];
    }
    sx = block_sum_128(sx, red, tid);
    sz = block_sum_128(sz, red, tid);
    const double shift = -sx / sz;
    for (int i = tid; i < n; i += T) {
        const double e = xs[i] + shift * zs[i];
        xs[i] = e;
        v.eta_n[(size_t)i * Cp + c] = e;
        if (dbg.enabled && dbg.out.eta_mean && c < v.C) dbg.out.eta_mean[(size_t)c * n + i] = ms[i];
    }
    __syncthreads();
    for (int idx = 0; idx < p * p; idx++) {  // A = X' Om X + b_prec
        const int a = idx / p, b = idx - a * p;
        if (b > a) continue;
        double sacc = 0.0;
        for (int i = tid; i < n; i += T) sacc += om_c[i] * v.X[(size_t)i * p + a] * v.X[(size_t)i * p + b];
        sacc = block_sum_128(sacc, red, tid);
        if (tid == 0) {
            small_S[a * p + b] = sacc + hp->b_prec[a * p + b];
            small_S[b * p + a] = sacc + hp->b_prec[b * p + a];
        }
    }
    for (int a = 0; a < p; a++) {  // b = X'(k - omega * eta) + b_prec b_mu
        double sacc = 0.0;
        for (int i = tid; i < n; i += T) sacc += v.X[(size_t)i * p + a] * (kd_c[i] - om_c[i] * xs[i]);
        sacc = block_sum_128(sacc, red, tid);
        if (tid == 0) small_b[a] = sacc + hp->b_pm[a];
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
        if (small_precision_mvnorm(small_S, p, small_b, nrm, small_x, small_t)) {
            for (int a = 0; a < p; a++) beta_s[a] = small_x[a];
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
    for (int a = tid; a < p; a += T) v.theta[(size_t)a * Cp + c] = beta_s[a];
    if (tid == 0) v.tau[c] = tau;
    if (rec_beta && rec_index >= 0 && rec_index < rec_capacity && c < v.C) {
        const size_t row = (size_t)c * rec_capacity + rec_index;
        for (int a = tid; a < p; a += T) rec_beta[row * p + a] = beta_s[a];
        if (tid == 0) rec_tau[row] = tau;
    }
}
