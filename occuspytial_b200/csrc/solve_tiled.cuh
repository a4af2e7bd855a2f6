// solve_tiled.cuh -- the tiled large-q RSR solve (200 < q <= 1312).
// Part of the single translation unit engine.cu (included there, in order; not a stand-alone header).
#pragma once

// ------------------------------------------------------------------------------------------------
// large-q RSR solve: when Lambda no longer fits in shared memory (q > ~200) it is kept in
// per-chain global scratch as 32 x 32 tiles and factorised by a left-looking tiled Cholesky: for block column k
//   (A) every warp takes tiles (I, k), I >= k, and applies all updates  A_Ik -= sum_{j<k} L_Ij L_kj'  on the FP64
//       tensor pipe (DMMA m8n8k4, accumulators in registers, operands straight from L1/L2),
//   (B) warp 0 factorises the diagonal tile in shared memory,
//   (C) the rows of the tiles below are solved against it (one thread per row, row in registers),
//   (D) the finished panel is written back -- every tile is written once.
// The right-hand side b rides along as an extra tile row, so y = L^-1 b comes out of the factorisation; the two
// backward solves (mean and noise) are tiled the same way.  Everything around the factorisation (tau, beta, alpha,
// commit, record) is the code of solve_logit_rsr_kernel.  Reference: logit.py:325-337, 478-485 (dgemm + dgesv).
// ------------------------------------------------------------------------------------------------
constexpr int TB = 32;               // tile size
constexpr int TP = TB + 1;           // padded row length of a tile in shared memory
#ifndef TILED_THREADS_N
#define TILED_THREADS_N 256
#endif
constexpr int TILED_THREADS = TILED_THREADS_N;   // 8 warps
constexpr int TILED_WARPS = TILED_THREADS / 32;
constexpr int PANEL_SLOTS = 1 + TILED_THREADS / 32;   // diagonal tile + one tile per warp
__device__ __forceinline__ size_t tile_off(int I, int J) { return ((size_t)I * (I + 1) / 2 + J) * (TB * TB); }  // J <= I

static size_t tiled_scratch_doubles(int q, int p) {
    const size_t nbk = (q + TB - 1) / TB;
    return (nbk * (nbk + 1) / 2 + nbk) * TB * TB + nbk * TB * (size_t)p + 64;
}
// shared memory: panel slots, then h[pq], th[pq + 4], xm[q32], xe[q32], nrm[q32 + 2 MAX_SMALL]
static size_t tiled_smem_bytes(int q, int pq) {
    const size_t q32 = (size_t)(q + TB - 1) / TB * TB;
    return ((size_t)PANEL_SLOTS * TB * TP + 2 * (size_t)pq + 4 + 3 * q32 + 2 * MAX_SMALL + 8) * sizeof(double);
}

__global__ void __launch_bounds__(TILED_THREADS, 512 / TILED_THREADS) solve_logit_rsr_tiled_kernel(const View v, const SolveDebug dbg,
                                                                               double* rec_beta, double* rec_tau,
                                                                               int rec_capacity) {
    extern __shared__ double panel[];                 // [PANEL_SLOTS][TB][TP], then the per-chain vectors
    __shared__ double partial[TILED_THREADS / 32][2][TB];
    __shared__ double invd[TB], red[TILED_THREADS / 32 + 1];
    __shared__ double small_S[MAX_SMALL * MAX_SMALL];
    __shared__ double small_b[MAX_SMALL], small_x[MAX_SMALL], small_t[MAX_SMALL];
    __shared__ int ok_flag;
    const int c = blockIdx.x, tid = threadIdx.x, T = TILED_THREADS;
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int p = v.p, q = v.q, r = v.r, pq = v.pq;
    const int nbk = (q + TB - 1) / TB, q32 = nbk * TB;
    const size_t Cp = v.Cp;
    const HParams* hp = v.hp;
    double* h = panel + (size_t)PANEL_SLOTS * TB * TP;   // [pq]      U'(z - 1/2)
    double* th = h + pq;                                 // [pq + 4]  beta, eta
    double* xm = th + pq + 4;                            // [q32]     backward solve: mean
    double* xe = xm + q32;                               // [q32]     backward solve: L^-T eps
    double* nrm = xe + q32;                              // [q32 + 2 MAX_SMALL] normals
    double* tiles = v.solve_scratch + (size_t)c * v.solve_stride;
    double* gx = tiles + ((size_t)nbk * (nbk + 1) / 2 + nbk) * TB * TB;   // [q32][p]  K' Om X
    const uint32_t iter = v.ctl->iter;
    const int rec_index = v.ctl->rec_index;
    const uint32_t gc = (uint32_t)(v.chain_offset + c);
    const int cr = c % v.C;
    const double* Grow = v.model == OCCU_PROBIT_RSR ? v.G0 : v.G + (size_t)c * v.M_pad;

    for (int a = tid; a < pq; a += T) {
        double hs = 0.0;
        for (int s = 0; s < v.ksplit; s++) hs += v.Hpart[((size_t)s * Cp + c) * v.ldh + a];
        h[a] = hs;
        th[a] = v.theta[(size_t)a * Cp + c];
    }
    for (int k = tid; k < q32 + p + r; k += T) {   // normals: eta (zero padded to q32), then beta, alpha
        const int src = k < q32 ? k : k - q32 + q;
        double val = 0.0;
        if (k >= q32 || k < q) {
            if (dbg.normals) {
                val = dbg.normals[(size_t)cr * (pq + r) + src];
            } else {
                Philox gph;
                gph.init(v.seed, gc, (uint32_t)src, iter, ST_NORMALS);
                val = gph.normal_d();
            }
        }
        nrm[k] = val;
    }
    if (tid == 0) ok_flag = 1;
    __syncthreads();

    // ---- tau | eta   (logit.py:206-209) ----
    {
        double part = 0.0;
#pragma unroll 8
        for (int idx = tid; idx < q * q; idx += T) {
            const int i = idx / q, j = idx - i * q;
            part += th[p + i] * v.Qrsr[idx] * th[p + j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) red[warp] = part;
        __syncthreads();
        if (tid == 0) {
            double quad = 0.0;
            for (int w = 0; w < T / 32; w++) quad += red[w];
            const double rate = 0.5 * quad + hp->tau_rate;
            red[T / 32] = dbg.tau_in ? dbg.tau_in[cr] : v.gam[c] / rate;
            if (dbg.enabled && dbg.out.tau_rate && c < v.C) dbg.out.tau_rate[c] = rate;
        }
        __syncthreads();
    }
    const double tau = red[T / 32];

    // ---- K' Om X rows, then Lambda = K' Om K + tau Qrsr and b = K'k - (K' Om X) beta as tiles ----
    for (int idx = tid; idx < q32 * p; idx += T) {
        const int i = idx / p, a = idx - i * p;
        gx[idx] = i < q ? Grow[tri(p + i, a)] : 0.0;
    }
    __syncthreads();
    for (int I = 0; I < nbk; I++) {   // Lambda tiles: four elements per thread in flight (read-only loads first)
        const int count = (I + 1) * TB * TB;
        for (int e0 = tid; e0 < count; e0 += 4 * T) {
            double gv[4], qv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int e = e0 + u * T;
                gv[u] = 0.0;
                qv[u] = 0.0;
                if (e < count) {
                    const int J = e / (TB * TB), rc = e - J * TB * TB, rr = rc / TB, cc = rc - rr * TB;
                    const int i = I * TB + rr, j = J * TB + cc;
                    if (i < q && j < q) {
                        const int hi = i >= j ? i : j, lo = i >= j ? j : i;
                        gv[u] = __ldg(Grow + tri(p + hi, p + lo));
                        qv[u] = __ldg(v.Qrsr + (size_t)hi * q + lo);
                    } else {
                        gv[u] = i == j ? 1.0 : 0.0;   // identity padding
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int e = e0 + u * T;
                if (e < count) {
                    const int J = e / (TB * TB), rc = e - J * TB * TB;
                    tiles[tile_off(I, J) + rc] = gv[u] + tau * qv[u];
                }
            }
        }
    }
    for (int e = tid; e < nbk * TB * TB; e += T) {   // right-hand side row (logit.py:481): row 0 of tile row nbk
        const int J = e / (TB * TB), rc = e - J * TB * TB, rr = rc / TB, cc = rc - rr * TB;
        const int j = J * TB + cc;
        double val = 0.0;
        if (rr == 0 && j < q) {
            double b = h[p + j];
            for (int a = 0; a < p; a++) b -= gx[j * p + a] * th[a];
            val = b;
            if (dbg.enabled && dbg.out.eta_b && c < v.C) dbg.out.eta_b[(size_t)c * q + j] = b;
        }
        tiles[tile_off(nbk, J) + rc] = val;
    }
    __syncthreads();
    if (dbg.enabled && dbg.out.eta_prec && c < v.C)
        for (int idx = tid; idx < q * q; idx += T) {
            const int i = idx / q, j = idx - i * q;
            const int hi = i >= j ? i : j, lo = i >= j ? j : i;
            dbg.out.eta_prec[(size_t)c * q * q + idx] = tiles[tile_off(hi / TB, lo / TB) + (hi % TB) * TB + (lo % TB)];
        }

    // ---- left-looking tiled Cholesky of the augmented matrix ----
    for (int k = 0; k < nbk && ok_flag; k++) {
        // tiles of block column k are processed in groups: the first group holds the diagonal tile (warp 0, slot 0)
        // and the next TILED_WARPS - 1 tiles below it, every later group TILED_WARPS tiles; slots 1.. hold the group's
        // off-diagonal tiles
        const int nbelow = nbk - k;                       // tiles below the diagonal, incl. the right-hand-side row
        bool first = true;
        for (int g0 = 0; first || g0 < nbelow; g0 = first ? TILED_WARPS - 1 : g0 + TILED_WARPS, first = false) {
            // tile handled by this warp: first group -> warp 0: diagonal, warps 1..7: below[0..6];
            // later groups -> below[g0 + warp]
            const int below_idx = first ? warp - 1 : g0 + warp;
            const int gcount = first ? min(TILED_WARPS - 1, nbelow) : min(TILED_WARPS, nbelow - g0);   // off-diagonal tiles
            const bool has_tile = first ? (warp == 0 || below_idx < nbelow) : (below_idx < nbelow);
            const int I = (first && warp == 0) ? k : k + 1 + below_idx;
            const int slot = (first && warp == 0) ? 0 : (first ? warp : 1 + warp);
            // (A) A_Ik -= sum_{j<k} L_Ij L_kj'
            if (has_tile) {
                double acc[4][4][2];
                const double* src = tiles + tile_off(I, k);
#pragma unroll
                for (int mb = 0; mb < 4; mb++)
#pragma unroll
                    for (int nb = 0; nb < 4; nb++) {
                        const double2 v2 = *reinterpret_cast<const double2*>(src + (mb * 8 + g) * TB + nb * 8 + 2 * t4);
                        acc[mb][nb][0] = v2.x;
                        acc[mb][nb][1] = v2.y;
                    }
                for (int j = 0; j < k; j++) {
                    const double* A = tiles + tile_off(I, j);
                    const double* B = tiles + tile_off(k, j);
#pragma unroll 2
                    for (int ks = 0; ks < TB / 4; ks++) {
                        double a[4], b[4];
#pragma unroll
                        for (int mb = 0; mb < 4; mb++) a[mb] = -A[(mb * 8 + g) * TB + ks * 4 + t4];
#pragma unroll
                        for (int nb = 0; nb < 4; nb++) b[nb] = B[(nb * 8 + g) * TB + ks * 4 + t4];
#pragma unroll
                        for (int mb = 0; mb < 4; mb++)
#pragma unroll
                            for (int nb = 0; nb < 4; nb++) dgemm::dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
                    }
                }
                double* dst = panel + (size_t)slot * TB * TP;
#pragma unroll
                for (int mb = 0; mb < 4; mb++)
#pragma unroll
                    for (int nb = 0; nb < 4; nb++) {
                        dst[(mb * 8 + g) * TP + nb * 8 + 2 * t4] = acc[mb][nb][0];
                        dst[(mb * 8 + g) * TP + nb * 8 + 2 * t4 + 1] = acc[mb][nb][1];
                    }
            }
            __syncthreads();
            if (first) {
                // (B) diagonal tile, whole CTA, one barrier per column: the trailing update of column j uses the
                //     UNSCALED column (a_ic -= a_ij a_cj / a_jj, j < c <= i); column j is scaled to
                //     l_ij = a_ij / sqrt(a_jj) during step j + 1, when nobody reads it any more
                double* D = panel;
                for (int j = 0; j <= TB; j++) {
                    double piv = 1.0;
                    if (j < TB) {
                        piv = D[j * TP + j];
                        if (!(piv > 0.0)) {   // uniform: every thread reads the same pivot
                            if (tid == 0) ok_flag = 0;
                            break;
                        }
                        if (lane > j) {   // lane = row, warp = column group (columns j + 1 + warp, + 8, ... <= row)
                            const double lij = D[lane * TP + j] / piv;
                            for (int cc = j + 1 + warp; cc <= lane; cc += T / 32) D[lane * TP + cc] -= lij * D[cc * TP + j];
                        }
                    }
                    if (j > 0 && tid < TB && tid >= j - 1) D[tid * TP + j - 1] *= invd[j - 1];   // column j - 1
                    if (j < TB && tid == 0) invd[j] = rsqrt(piv);
                    __syncthreads();
                }
                __syncthreads();
                if (!ok_flag) break;
            }
            // (C) rows of the group's off-diagonal tiles: x L_kk' = a, one thread per row
            if (tid < gcount * TB) {
                double* R = panel + (size_t)(1 + tid / TB) * TB * TP + (tid % TB) * TP;   // row stride TP: conflict-free
                for (int cc = 0; cc < TB; cc++) {
                    double s0 = R[cc], s1 = 0.0;          // two accumulators: the dependent chain is cc / 2 long
                    int m = 0;
                    for (; m + 1 < cc; m += 2) {
                        s0 -= R[m] * panel[cc * TP + m];
                        s1 -= R[m + 1] * panel[cc * TP + m + 1];
                    }
                    if (m < cc) s0 -= R[m] * panel[cc * TP + m];
                    R[cc] = (s0 + s1) * invd[cc];
                }
            }
            __syncthreads();
            // (D) write the group back (the strictly upper part of the diagonal tile is never read)
            {
                const int s0 = first ? 0 : 1, ns = gcount + (first ? 1 : 0);
                for (int e = tid; e < ns * TB * TB; e += T) {
                    const int s = s0 + e / (TB * TB), rc = e % (TB * TB);
                    const int It = s == 0 ? k : k + 1 + (first ? 0 : g0) + (s - 1);
                    tiles[tile_off(It, k) + rc] = panel[(size_t)s * TB * TP + (rc / TB) * TP + (rc % TB)];
                }
            }
            __syncthreads();
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
    if (dbg.enabled && dbg.out.eta_chol && c < v.C)
        for (int idx = tid; idx < q * q; idx += T) {
            const int i = idx / q, j = idx - i * q;
            dbg.out.eta_chol[(size_t)c * q * q + idx] = j <= i ? tiles[tile_off(i / TB, j / TB) + (i % TB) * TB + (j % TB)] : 0.0;
        }

    // ---- backward: L' m = y (mean) and L' e = eps, block rows from the bottom ----
    for (int i = tid; i < q32; i += T) {
        xm[i] = tiles[tile_off(nbk, i / TB) + (i % TB)];   // y = L^-1 b: row 0 of the extra tile row
        xe[i] = nrm[i];
    }
    __syncthreads();
    for (int k = nbk - 1; k >= 0; k--) {
        // rhs_k -= sum_{I > k} L_Ik' x_I : warps split the tiles below, lane = column
        double sm_ = 0.0, se_ = 0.0;
        for (int I = k + 1 + warp; I < nbk; I += T / 32) {
            const double* L = tiles + tile_off(I, k);
#pragma unroll 8
            for (int rr = 0; rr < TB; rr++) {
                const double l = L[rr * TB + lane];
                sm_ += l * xm[I * TB + rr];
                se_ += l * xe[I * TB + rr];
            }
        }
        partial[warp][0][lane] = sm_;
        partial[warp][1][lane] = se_;
        // diagonal tile into shared memory for the triangular solve
        for (int e = tid; e < TB * TB; e += T) panel[(e / TB) * TP + (e % TB)] = tiles[tile_off(k, k) + e];
        __syncthreads();
        if (warp < 2) {   // warp 0: mean, warp 1: noise
            double* x = warp == 0 ? xm : xe;
            double rhs = x[k * TB + lane];
            for (int w = 0; w < T / 32; w++) rhs -= partial[w][warp][lane];
            // L_kk' x = rhs, backward: x_c = (rhs_c - sum_{m > c} L[m][c] x_m) / L[c][c]
            for (int cc = TB - 1; cc >= 0; cc--) {
                const double xc = __shfl_sync(0xffffffffu, rhs, cc) / panel[cc * TP + cc];
                if (lane == cc) rhs = xc;
                if (lane < cc) rhs -= panel[cc * TP + lane] * xc;
            }
            x[k * TB + lane] = rhs;
        }
        __syncthreads();
    }
    if (dbg.enabled && dbg.out.eta_mean && c < v.C)
        for (int i = tid; i < q; i += T) dbg.out.eta_mean[(size_t)c * q + i] = xm[i];
    for (int i = tid; i < q; i += T) th[p + i] = xm[i] + xe[i];  // eta = mean + L^-T eps
    __syncthreads();

    // ---- beta | .   A = X' Om X + b_prec,  b = X'k - (X' Om K) eta + b_prec b_mu   (logit.py:226-232) ----
    for (int idx = tid; idx < p * p; idx += T) {
        const int a = idx / p, b = idx - a * p;
        small_S[idx] = ((b <= a) ? Grow[tri(a, b)] : Grow[tri(b, a)]) + hp->b_prec[idx];
    }
    for (int a = warp; a < p; a += T / 32) {   // one warp per row of the system: q-long dot product
        double b = 0.0;
        for (int j = lane; j < q; j += 32) b -= gx[j * p + a] * th[p + j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
        if (lane == 0) small_b[a] = b + h[a] + hp->b_pm[a];
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
        if (small_precision_mvnorm(small_S, p, small_b, nrm + q32, small_x, small_t)) {
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
    // ---- commit state, record the draw (Chain.append, chain.py:56-74) ----
    for (int a = tid; a < pq; a += T) v.theta[(size_t)a * Cp + c] = th[a];
    if (tid == 0) v.tau[c] = tau;
    if (rec_beta && rec_index >= 0 && rec_index < rec_capacity && c < v.C) {
        const size_t row = (size_t)c * rec_capacity + rec_index;
        for (int a = tid; a < p; a += T) rec_beta[row * p + a] = th[a];
        if (tid == 0) rec_tau[row] = tau;
    }
}
