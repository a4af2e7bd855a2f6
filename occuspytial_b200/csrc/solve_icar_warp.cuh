// solve_icar_warp.cuh -- LogitICARGibbs ("PGGibbs"): fixed-pattern banded Cholesky solve, four warps per chain.
// Part of the single translation unit engine.cu (included there, in order; not a stand-alone header).
#pragma once

// ------------------------------------------------------------------------------------------------
// eta | . ~ N(Lambda^-1 b, Lambda^-1) on {1'eta = 0}, Lambda = tau Q + diag(omega_b)   (logit.py:73-99, 211-217)
//
// The reference perturbs b through a dense eigen-factor of Q (2n - 1 normals) and solves two systems with MINRES at
// rtol 1e-5.  Here Lambda -- whose sparsity pattern is fixed by Q -- is factorised exactly per chain by a banded
// Cholesky over the fixed symbolic pattern (half bandwidth bw after the host's reordering), the draw is
// x = Lambda^-1 b + L^-T eps (n normals, same law), zv = Lambda^-1 1, and eta = x - (sum x / sum zv) zv
// (ensure_sums_to_zero, distributions.pyx:24-39).
//
// Mapping.  FOUR WARPS own a chain (a 128-thread CTA at <= 72 registers, so that 7 CTAs -- the 1024 chains of BASELINE
// config 4 in one wave -- stay resident per SM), each with its own job, two barriers per block step:
//   warp 0      the serial part: the 8 x 8 pivot block of the NEXT block step -- its rank-8 update, then its Cholesky
//               factorisation, lanes 0..7 holding one row each and exchanging columns by shuffles (8 dependent rsqrt:
//               the irreducible chain) -- runs while the other warps do the trailing update of the CURRENT step
//               (look-ahead); during the row phase it stages the next entering rows (cp.async) and solves the two
//               right-hand sides;
//   warps 1..3  row phase: every thread solves one window row against the factor (read as shared-memory broadcasts),
//               writes that row of L to HBM (column-band layout) and into the panel P; update phase: forward residuals,
//               rank-8 DMMA update of the tiles that stay (dealt round-robin), refill of the leaving slot.
// One warp per chain (7 warps per SM) was bound by one chain's dependency latency at 25 % issue utilisation
// (profiles/r02_ncu_icar_warp_v3.txt); two warps with a redundant pivot factorisation gained nothing.
// The factorisation is right-looking in blocks of 8 columns over a sliding window of W = 8 NT >= bw + 8 matrix rows, kept
// in shared memory as 8 x 8 tiles addressed by RING SLOT (8-row group index mod NT), so sliding the window moves no data:
// per block step one slot leaves (the pivot block) and is refilled with the 8 entering rows.  Only slot pairs a >= b are
// stored: tile (a, b) holds Lambda[rows of slot a][columns of slot b], whichever of the two is the lower triangle in
// matrix order -- the matrix is symmetric and the rank-8 update tile(a, b) -= P_a P_b' has the same form for every pair,
// two DMMA.8x8x4 per tile with the A and B fragments read from the same column-major panel.  The right-hand sides b and
// 1 ride along (forward substitution fused into the factorisation).  The entering band rows of Q (forward) and the eight
// column bands of the factor (backward: one contiguous piece) are staged one step ahead by cp.async into rows padded with
// zeros, so neither pass has bounds checks.  The backward pass is blocked the same way: the 8 x bw product
// L_below' [x zv m] is a chain of DMMAs split over the four warps, three lanes finish the 8 x 8 triangular solves.
// ------------------------------------------------------------------------------------------------
constexpr int ICARW_THREADS = 128;

// window tiles per side: the smallest compiled instance (2, 4, 6, 8: fully unrolled tile loops, constant shared-memory
// offsets) that covers bw + 8 rows, else the exact count with run-time loops (template argument 0)
__host__ __device__ inline int icarw_nt_needed(int bw) { return (bw + 8 + 7) / 8; }
__host__ __device__ inline int icarw_nt(int bw) {
    const int need = icarw_nt_needed(bw);
    return need <= 8 ? (need + 1) / 2 * 2 : need;
}
__host__ __device__ inline int icarw_wp(int nt) { return (nt & 1) ? 8 * nt : 8 * nt + 8; }   // panel pitch = 8 mod 16 doubles
// staged band rows: W + 8 doubles, zero beyond the band, so that both passes read them without bounds checks (forward:
// offsets up to W - 1; backward: up to bw + 10 <= W + 2)
__host__ __device__ inline int icarw_pw(int nt) { return 8 * nt + 8; }
// shared-memory layout (doubles): region A = window tiles (forward) / x ring + block sums + two staged factor blocks
// (backward) | panel | forward residuals | one staged group of entering rows | beta system
__host__ __device__ inline size_t icarw_region_a(int nt, int bw) {
    const size_t tiles = (size_t)nt * (nt + 1) / 2 * 64, back = 8 * 8 * (size_t)nt + 128 + 2 * (8 * (size_t)icarw_pw(nt) + 32);
    return tiles > back ? tiles : back;
}
static size_t icarw_smem_bytes(int bw, int p) {
    const size_t nt = icarw_nt(bw), W = 8 * nt, Wp = icarw_wp((int)nt);
    return (icarw_region_a((int)nt, bw) + 8 * Wp + 2 * W + (8 * (size_t)icarw_pw((int)nt) + 32) + 96 + (size_t)p * p +
            5 * (size_t)p) * sizeof(double);
}

// 8-byte cp.async with zero fill when `ok` is false: lands while the block step in between runs
__device__ __forceinline__ void icarw_cp8(double* dst, const double* src, bool ok) {
    const uint32_t sdst = (uint32_t)__cvta_generic_to_shared(dst);
    const int bytes = ok ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sdst), "l"(src), "r"(bytes));
}

// compact non-zero offsets of every band row of Q (set-up, once): qnz[i][0] = count (255: too many), then the offsets
__global__ void icar_nonzeros_kernel(const double* __restrict__ Qband, int n, int W1, uint8_t* __restrict__ qnz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cnt = 0;
    for (int d = 1; d < W1 && d <= i; d++) {
        if (Qband[(size_t)i * W1 + d] != 0.0) {
            if (cnt < ICAR_NZ) qnz[(size_t)i * (ICAR_NZ + 1) + 1 + cnt] = (uint8_t)d;
            cnt++;
        }
    }
    qnz[(size_t)i * (ICAR_NZ + 1)] = cnt <= ICAR_NZ ? (uint8_t)cnt : (uint8_t)255;
}

__device__ __forceinline__ double warp_sum(double val) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    return val;
}

// sum over the CTA's four warps (red: four doubles of shared memory); every thread gets the total
__device__ __forceinline__ double cta4_sum(double val, double* red) {
    val = warp_sum(val);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = val;
    __syncthreads();
    return (red[0] + red[1]) + (red[2] + red[3]);
}

// Cholesky factorisation of an 8 x 8 block by ONE WARP: lane a < 8 holds row a of the lower triangle (row[b], b <= a);
// column j is scaled by 1 / sqrt(pivot) (rsqrt, no divide) and exchanged by shuffles for the rank-1 update.  Writes the
// factor (row-major 8 x 8, lower) and the reciprocal pivots to shared memory.  Returns false on a non-positive pivot.
__device__ __forceinline__ bool icarw_chol8(const double* __restrict__ D, double* __restrict__ Lds, double* __restrict__ invs,
                                            int lane) {
    const int a = lane & 7;
    double row[8];
#pragma unroll
    for (int b = 0; b < 8; b++) row[b] = D[a * 8 + b];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const double piv = __shfl_sync(0xffffffffu, row[j], j);
        if (!(piv > 0.0)) ok = false;
        const double ri = rsqrt(piv);
        const double laj = (a == j) ? piv * ri : row[j] * ri;
        row[j] = laj;
        if (lane == j) invs[j] = ri;
#pragma unroll
        for (int b = j + 1; b < 8; b++) {
            const double lbj = __shfl_sync(0xffffffffu, laj, b);
            row[b] = fma(-laj, lbj, row[b]);   // only b <= a is ever read back
        }
    }
    if (lane < 8) {
#pragma unroll
        for (int b = 0; b < 8; b++)
            if (b <= a) Lds[a * 8 + b] = row[b];
    }
    return ok;
}

template <int NT_T>
__global__ void __launch_bounds__(ICARW_THREADS, NT_T > 0 ? 7 : 1) solve_logit_icar_warp_kernel(const View v, const SolveDebug dbg,
                                                                                              double* rec_beta, double* rec_tau,
                                                                                              int rec_capacity) {
    extern __shared__ __align__(16) double icw_sm[];
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, T = ICARW_THREADS;
    const int g = lane >> 2, t4 = lane & 3;
    const int n = v.n, n4 = round_up_dev(v.n, 4), p = v.p, w = v.bw, W1 = v.bw + 1;
    const int NT = NT_T > 0 ? NT_T : icarw_nt(w), W = 8 * NT, Wp = icarw_wp(NT);
    const int W1p = icarw_pw(NT);
    const int n8 = round_up_dev(n, 8);
    const size_t Cp = v.Cp;
    const HParams* hp = v.hp;
    const int stg_sz = 8 * W1p + 32;              // 8 staged band rows + 4 x 8 scalars
    double* Ts = icw_sm;                          // [NT (NT + 1) / 2][8][8]  window tiles, pair (a, b), a >= b   (forward)
    double* xr = icw_sm;                          // [8][W]   backward: ring of solved x, zv, m by (matrix row mod W); rows 3..7 zero
    double* Ss = xr + 8 * W;                      // [4][8][4]  backward: block sums, one set per warp
    double* stgB = Ss + 128;                      // [2][stg_sz]  backward: staged factor blocks
    double* Ps = icw_sm + icarw_region_a(NT, w);  // [8][Wp]  panel, column-major: Ps[col][window row]
    double* rhs = Ps + 8 * Wp;                    // [2][W]   forward residuals of b and 1, by window row
    double* stg = rhs + 2 * W;                    // [stg_sz] forward: staged entering rows
    double* Lds = stg + stg_sz;                   // [8][8]   factor of the pivot block (lower, row-major)
    double* invs = Lds + 64;                      // [8]      1 / L_jj of the pivot block
    double* yvs = invs + 8;                       // [2][8]   y_b, y_1 of the pivot block
    double* red = yvs + 16;                       // [8]      CTA reductions
    double* small_S = Lds + 96;                   // [p][p]   beta system
    double* small_b = small_S + p * p;
    double* small_x = small_b + p;
    double* small_t = small_x + p;
    double* beta_s = small_t + p;
    double* nrm = beta_s + p;
    __shared__ int ok_flag, piv_ok;

    const uint32_t iter = v.ctl->iter;
    const int rec_index = v.ctl->rec_index;
    const uint32_t gc = (uint32_t)(v.chain_offset + c);
    const int cr = c % v.C;
    double* Lc = v.Lc + (size_t)c * n * W1;
    double* yb = v.scratch + (size_t)c * 8 * n4;
    double* y1 = yb + n4;
    double* xs = y1 + n4;    // b (forward) then x (backward)
    double* zs = xs + n4;    // normals (forward) then zv (backward)
    double* ms = zs + n4;    // L^-T L^-1 b (debug mean); eta during the tau quadratic form
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
    if (tid == 0) {
        ok_flag = 1;
        piv_ok = 1;
    }
    __syncthreads();

    // ---- tau | eta (logit.py:206-209); eta is first copied chain-major so that the band products hit L1 ----
    for (int i = tid; i < n; i += T) ms[i] = v.eta_n[(size_t)i * Cp + c];
    __syncthreads();
    double part = 0.0;
    for (int i = tid; i < n; i += T) {
        const double ei = ms[i];
        const double* qrow = v.Qband + (size_t)i * W1;
        double sacc = 0.5 * qrow[0] * ei;
        const uint8_t* nz = v.qnz + (size_t)i * (ICAR_NZ + 1);
        const int cnt = nz[0];
        if (cnt != 255) {   // the few non-zeros of the row, ascending offset
            for (int t = 1; t <= cnt; t++) {
                const int d = nz[t];
                sacc += qrow[d] * ms[i - d];
            }
        } else {
            const int dmax = min(w, i);
            for (int d = 1; d <= dmax; d++) sacc += qrow[d] * ms[i - d];
        }
        part += ei * sacc;
    }
    const double quad = 2.0 * cta4_sum(part, red);
    const double rate = 0.5 * quad + hp->tau_rate;
    const double tau = dbg.tau_in ? dbg.tau_in[cr] : v.gam[c] / rate;
    if (tid == 0 && dbg.enabled && dbg.out.tau_rate && c < v.C) dbg.out.tau_rate[c] = rate;

    // ---- stage this chain's per-site inputs contiguously (the chain-minor state arrays would cost one uncoalesced
    //      access per row inside the sequential loops; the Box-Muller normals must not sit on the critical path) ----
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

    // ---- entering rows: the 8 band rows of Q, omega and b of the group starting at `base` -> the staging buffer
    //      (cp.async by the calling warp(s): `first`, `step` give the rows each warp takes) ----
    auto stage_rows = [&](int base, int first, int step) {
        double* dst = stg;
#pragma unroll 1
        for (int rr = first; rr < 8; rr += step) {
            const int i = base + rr;
            const double* srow = v.Qband + (size_t)(i < n ? i : 0) * W1;
            for (int d = lane; d < W1; d += 32) icarw_cp8(dst + rr * W1p + d, srow + d, i < n && d <= i);
        }
        if (first == 0 && lane < 16) {
            const int rr = lane & 7, i = base + rr;
            const double* srcv = lane < 8 ? om_c : xs;
            icarw_cp8(dst + 8 * W1p + (lane < 8 ? 0 : 8) + rr, i < n ? srcv + i : srcv, i < n);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    // tile (a, b), a >= b, with the staged group (matrix rows base .. base + 7) on one side and an OLDER group on the other:
    // rows = matrix rows ra .. ra + 7, columns = rb .. rb + 7.  Lambda[hi][lo] = tau Q[hi][hi - lo]; the staged rows are
    // zero beyond the band and for rows past n, so the read needs no bounds checks.  Two adjacent elements per lane.
    auto fill_tile_off = [&](const double* sb, int base, int a, int b, int ra, int rb) {
        const int i = ra + g, k = rb + 2 * t4;
        double v0, v1;
        if (ra == base) {   // rows are the staged group: hi = i
            const double* row = sb + g * W1p + (i - k);
            v0 = tau * row[0];
            v1 = tau * row[-1];
        } else {            // columns are the staged group: hi = k
            const double* row = sb + (k - base) * W1p + (k - i);
            v0 = tau * row[0];
            v1 = tau * row[W1p + 1];
        }
        *reinterpret_cast<double2*>(Ts + (size_t)tri(a, b) * 64 + g * 8 + 2 * t4) = make_double2(v0, v1);
    };
    // the diagonal tile of the staged group (symmetric; identity beyond n)
    auto fill_tile_diag = [&](const double* sb, int base, int a) {
        double vv[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int ri = g, ci = 2 * t4 + u;
            const int hi = max(ri, ci), d = hi - min(ri, ci);
            double val = tau * sb[hi * W1p + d];
            if (d == 0) val += base + hi < n ? sb[8 * W1p + hi] : 1.0;
            vv[u] = val;
        }
        *reinterpret_cast<double2*>(Ts + (size_t)tri(a, a) * 64 + g * 8 + 2 * t4) = make_double2(vv[0], vv[1]);
    };
    auto fill_rhs = [&](const double* sb, int slot, int base) {   // right-hand sides b and 1 of the 8 rows (lanes 0..15 of the caller)
        if (lane < 16) {
            const int rr = lane & 7, r = lane >> 3, i = base + rr;
            rhs[r * W + 8 * slot + rr] = i < n ? (r == 0 ? sb[8 * W1p + 8 + rr] : 1.0) : 0.0;
        }
    };

    // ---- initial window: matrix rows 0 .. W - 1, group s in slot s; group a is staged, its tiles against the groups
    //      b <= a are filled (every stored pair has its later group as the row side here) ----
    for (int e = tid; e < stg_sz; e += T) stg[e] = 0.0;   // the pads beyond the band stay zero for good
    __syncthreads();
    for (int a = 0; a < NT; a++) {
        stage_rows(8 * a, warp, 4);
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();
        for (int b = warp; b < a; b += 4) fill_tile_off(stg, 8 * a, a, b, 8 * a, 8 * b);
        if (warp == (a & 3)) fill_tile_diag(stg, 8 * a, a);
        if (warp == ((a + 1) & 3)) fill_rhs(stg, a, 8 * a);
        __syncthreads();
    }
    // factor of the first pivot block
    if (warp == 0) {
        if (!icarw_chol8(Ts, Lds, invs, lane)) piv_ok = 0;
    }
    __syncthreads();

    // ---- blocked factorisation + fused forward substitutions ----
    const int nsteps = n8 / 8;
    for (int m = 0; m < nsteps; m++) {
        if (!piv_ok) break;   // uniform (read after a barrier)
        const int cs = m % NT, j0 = 8 * m;
        const int cs1 = cs + 1 == NT ? 0 : cs + 1;   // the next pivot slot
        // ===== row phase =====
        if (warp == 0) {
            // entering rows of this step (group m + NT): staged now, consumed by the refill after the barrier
            stage_rows(8 * (m + NT), 0, 1);
            // the two right-hand sides of the pivot rows: y_r = L_JJ^-1 rhs_r (lane r)
            if (lane < 2) {
                double yv[8];
#pragma unroll
                for (int b = 0; b < 8; b++) {
                    double t = rhs[lane * W + 8 * cs + b];
#pragma unroll
                    for (int k = 0; k < b; k++) t = fma(-yv[k], Lds[b * 8 + k], t);
                    yv[b] = t * invs[b];
                }
                double* dstg = lane == 0 ? yb : y1;
#pragma unroll
                for (int b = 0; b < 8; b++) {
                    yvs[lane * 8 + b] = yv[b];
                    if (j0 + b < n) dstg[j0 + b] = yv[b];
                }
            } else if (lane >= 8 && lane < 16) {
                if (j0 + lane - 8 < n) inv_c[j0 + lane - 8] = invs[lane - 8];
            }
            asm volatile("cp.async.wait_group 0;\n" ::);
        } else {
            // every window row against the factor: L (to HBM) and the panel P
#pragma unroll 1
            for (int wrow = tid - 32; wrow < W; wrow += T - 32) {
                const int a = wrow >> 3, rr = wrow & 7;
                const int grp = m + (a >= cs ? a - cs : a - cs + NT);   // 8-row group held by slot a
                const int dt = 8 * (grp - m) + rr;                       // matrix row - j0
                double x[8];
                if (a >= cs) {
                    const double* src = Ts + (size_t)tri(a, cs) * 64 + rr * 8;
#pragma unroll
                    for (int b = 0; b < 8; b += 2) {
                        const double2 t2 = *reinterpret_cast<const double2*>(src + b);
                        x[b] = t2.x;
                        x[b + 1] = t2.y;
                    }
                } else {
                    const double* src = Ts + (size_t)tri(cs, a) * 64 + rr;
#pragma unroll
                    for (int b = 0; b < 8; b++) x[b] = src[b * 8];
                }
#pragma unroll
                for (int b = 0; b < 8; b++) {
                    double t = x[b];
#pragma unroll
                    for (int k = 0; k < b; k++) t = fma(-x[k], Lds[b * 8 + k], t);
                    x[b] = t * invs[b];
                }
                const bool pivot_row = a == cs;
                const int i = j0 + dt;
                if (i < n) {   // L[i][j0 + b] -> column-band entry e = dt - b of column j0 + b (consecutive b: stride W1 - 1)
                    double* lp = Lc + (size_t)j0 * W1 + dt;
                    const int blo = dt - w, bhi = min(dt, n - 1 - j0);   // 0 <= e <= w and j0 + b < n  <=>  blo <= b <= bhi
#pragma unroll
                    for (int b = 0; b < 8; b++)
                        if (b >= blo && b <= bhi) lp[b * (W1 - 1)] = x[b];
                }
#pragma unroll
                for (int b = 0; b < 8; b++) Ps[b * Wp + wrow] = pivot_row ? 0.0 : x[b];
            }
        }
        __syncthreads();   // panel, y and the staged entering rows are visible
        // ===== update phase =====
        if (warp == 0) {
            // look-ahead: the next pivot block gets its rank-8 update first and is factorised right away, while the
            // other warps update the rest of the window
            if (m + 1 < nsteps) {
                double* Dn = Ts + (size_t)tri(cs1, cs1) * 64;
                const double a0 = -Ps[t4 * Wp + 8 * cs1 + g], a1 = -Ps[(4 + t4) * Wp + 8 * cs1 + g];
                const double b0 = -a0, b1 = -a1;
                double2* cp = reinterpret_cast<double2*>(Dn + g * 8 + 2 * t4);
                double2 cv = *cp;
                dgemm::dmma884(cv.x, cv.y, a0, b0);
                dgemm::dmma884(cv.x, cv.y, a1, b1);
                *cp = cv;
                __syncwarp();
                if (!icarw_chol8(Dn, Lds, invs, lane)) piv_ok = 0;
            }
        } else {
            const int hw = warp - 1;   // 0..2
            // forward residuals of the rows that stay: rhs_r[row] -= y_r . P[row]
#pragma unroll 1
            for (int wrow = tid - 32; wrow < W; wrow += T - 32) {
                if ((wrow >> 3) != cs) {
                    double r0 = rhs[wrow], r1 = rhs[W + wrow];
#pragma unroll
                    for (int b = 0; b < 8; b++) {
                        const double pv = Ps[b * Wp + wrow];
                        r0 = fma(-yvs[b], pv, r0);
                        r1 = fma(-yvs[8 + b], pv, r1);
                    }
                    rhs[wrow] = r0;
                    rhs[W + wrow] = r1;
                }
            }
            // rank-8 trailing update on the tensor pipe: tile (a, b) -= P_a P_b' for the slots that stay, dealt round-robin
            // to the three warps; the pivot slot's tiles are being refilled and the next pivot block belongs to warp 0
            if (NT_T > 0) {
                // this warp's tiles of the step as a bit mask over the NT (NT + 1) / 2 stored pairs (<= 36): every third
                // tile, without the pivot slot's row and column and without the next pivot block
                unsigned long long mine = 0ull;
                {
                    constexpr int NTILES = (NT_T > 0 ? NT_T : 1) * ((NT_T > 0 ? NT_T : 1) + 1) / 2;
                    unsigned long long third = 0ull;
#pragma unroll
                    for (int t = 0; t < NTILES; t++)
                        if (t % 3 == 0) third |= 1ull << t;
                    mine = third << hw;                               // tiles t = hw (mod 3)
                    unsigned long long dead = 1ull << tri(cs1, cs1);
                    for (int b = 0; b <= cs; b++) dead |= 1ull << tri(cs, b);        // row of the pivot slot
                    for (int a = cs + 1; a < NT; a++) dead |= 1ull << tri(a, cs);    // its column
                    mine &= ~dead & ((1ull << NTILES) - 1ull);
                }
#pragma unroll
                for (int a = 0; a < (NT_T > 0 ? NT_T : 1); a++) {
                    const double a0 = -Ps[t4 * Wp + 8 * a + g], a1 = -Ps[(4 + t4) * Wp + 8 * a + g];
#pragma unroll
                    for (int b = 0; b <= a; b++) {
                        if ((mine >> (a * (a + 1) / 2 + b)) & 1ull) {
                            const double b0 = Ps[t4 * Wp + 8 * b + g], b1 = Ps[(4 + t4) * Wp + 8 * b + g];
                            double2* cp = reinterpret_cast<double2*>(Ts + (a * (a + 1) / 2 + b) * 64 + g * 8 + 2 * t4);
                            double2 cv = *cp;
                            dgemm::dmma884(cv.x, cv.y, a0, b0);
                            dgemm::dmma884(cv.x, cv.y, a1, b1);
                            *cp = cv;
                        }
                    }
                }
            } else {
                int tix = 0;
                for (int a = 0; a < NT; a++) {
                    const double a0 = -Ps[t4 * Wp + 8 * a + g], a1 = -Ps[(4 + t4) * Wp + 8 * a + g];
                    for (int b = 0; b <= a; b++, tix++) {
                        if ((tix % 3) != hw || a == cs || b == cs || (a == cs1 && b == cs1)) continue;
                        const double b0 = Ps[t4 * Wp + 8 * b + g], b1 = Ps[(4 + t4) * Wp + 8 * b + g];
                        double2* cp = reinterpret_cast<double2*>(Ts + (size_t)tix * 64 + g * 8 + 2 * t4);
                        double2 cv = *cp;
                        dgemm::dmma884(cv.x, cv.y, a0, b0);
                        dgemm::dmma884(cv.x, cv.y, a1, b1);
                        *cp = cv;
                    }
                }
            }
            // the pivot slot leaves; it takes the 8 entering rows (group m + NT)
            const int rn = 8 * (m + NT);
            for (int a = hw; a < NT; a += 3) {
                if (a == cs) {
                    fill_tile_diag(stg, rn, cs);
                } else {
                    const int grp = m + (a >= cs ? a - cs : a - cs + NT);
                    if (a > cs) fill_tile_off(stg, rn, a, cs, 8 * grp, rn);
                    else fill_tile_off(stg, rn, cs, a, rn, 8 * grp);
                }
            }
            if (hw == 0) fill_rhs(stg, cs, rn);
        }
        __syncthreads();   // window updated and refilled; next factor ready; staged group consumed
    }
    if (!piv_ok) {
        if (tid == 0) {
            v.status[c] = 1;
            if (c < v.C) atomicAdd(&v.ctl->fails, 1);
        }
        return;
    }
    __threadfence_block();
    __syncthreads();
    if (dbg.enabled && dbg.out.eta_chol && c < v.C)  // column-band factor [n][bw + 1]
        for (size_t idx = tid; idx < (size_t)n * W1; idx += T) {
            const int j = (int)(idx / W1), e = (int)(idx - (size_t)j * W1);
            dbg.out.eta_chol[(size_t)c * n * W1 + idx] = (j + e < n) ? Lc[idx] : 0.0;
        }

    // ---- blocked backward: L' x = y_b + eps, L' zv = y_1, L' m = y_b (its arrays take over the window's memory) ----
    __syncthreads();
    for (int e = tid; e < 8 * W + 128 + 2 * stg_sz; e += T) xr[e] = 0.0;   // ring, sums, staging pads
    __syncthreads();
    {
        // the 8 column bands of a block are one contiguous piece of the factor: staged one block ahead (cp.async, two
        // columns per warp), with y_b, y_1, the normals and 1 / L_jj of the block.  Entries of L below row n are never
        // written: they keep the zero of the workspace, like the staged rows' pads, so nothing below needs a bounds check.
        auto stage_block = [&](int j0, int buf) {
            double* dst = stgB + buf * stg_sz;
            if (j0 >= 0) {
#pragma unroll 1
                for (int bb = warp; bb < 8; bb += 4) {
                    const bool okk = j0 + bb < n;
                    const double* scol = Lc + (size_t)(okk ? j0 + bb : 0) * W1;
                    for (int e = lane; e < W1; e += 32) icarw_cp8(dst + bb * W1p + e, scol + e, okk);
                }
                if (tid < 32) {
                    const int which = tid >> 3, bb = tid & 7, j = j0 + bb;
                    const double* srcv = which == 0 ? yb : (which == 1 ? y1 : (which == 2 ? zs : inv_c));
                    icarw_cp8(dst + 8 * W1p + which * 8 + bb, j < n ? srcv + j : srcv, j < n);
                }
            }
            asm volatile("cp.async.commit_group;\n" ::);
        };
        const int ksteps = (w + 3) >> 2;   // rows j0 + 8 .. j0 + 7 + bw below the block
        const int mlast = n8 / 8 - 1;
        stage_block(8 * mlast, mlast & 1);
        int ring0 = (8 * mlast + 8 + t4 + 4 * warp) % W;   // ring slot of row j0 + 8 + 4 warp + t4
        for (int m = mlast; m >= 0; m--) {
            const int j0 = 8 * m;
            stage_block(j0 - 8, (m + 1) & 1);                 // next block; its buffer was consumed one block ago
            asm volatile("cp.async.wait_group 1;\n" ::);     // this block's copies have landed
            __syncthreads();
            const double* sb = stgB + (m & 1) * stg_sz;
            // S[b][r] = sum_i L[i][j0 + b] x_r[i]:  D (8 columns b x 8 right-hand sides) += A (8 x 4: L[i][j0 + g]) B (4 x 8: x_g[i]);
            // warp w takes the k steps w, w + 4, ...
            double s0 = 0.0, s1 = 0.0;
            const double* lcol = sb + g * W1p + (8 + t4 - g) + 4 * warp;   // entry e = 8 + 4 ks + t4 - g of column j0 + g
            const double* xg = xr + g * W;
            int ring = ring0;
            for (int ks = warp; ks < ksteps; ks += 4) {
                dgemm::dmma884(s0, s1, lcol[4 * (ks - warp)], xg[ring]);
                ring += 16;
                if (ring >= W) ring -= W;
                if (ring >= W) ring -= W;
            }
            if (t4 < 2) {   // right-hand sides 0..3 live in the lanes t4 = 0, 1
                Ss[warp * 32 + g * 4 + 2 * t4] = s0;
                Ss[warp * 32 + g * 4 + 2 * t4 + 1] = s1;
            }
            __syncthreads();
            if (tid < 3) {  // one lane per right-hand side finishes the 8 x 8 triangular solve
                const int r = tid;
                double xb8[8];
#pragma unroll
                for (int b = 7; b >= 0; b--) {
                    const double ybv = sb[8 * W1p + b];
                    double t = (r == 0 ? ybv + sb[8 * W1p + 16 + b] : (r == 1 ? sb[8 * W1p + 8 + b] : ybv)) -
                               ((Ss[b * 4 + r] + Ss[32 + b * 4 + r]) + (Ss[64 + b * 4 + r] + Ss[96 + b * 4 + r]));
#pragma unroll
                    for (int bb = b + 1; bb < 8; bb++) t = fma(-sb[b * W1p + (bb - b)], xb8[bb], t);
                    xb8[b] = t * sb[8 * W1p + 24 + b];   // 1 / L_jj, staged as zero for the padded columns j >= n
                }
                double* dst = r == 0 ? xs : (r == 1 ? zs : ms);
                const int slot = j0 % W;
#pragma unroll
                for (int b = 0; b < 8; b++) {
                    xr[r * W + slot + b] = xb8[b];       // blocks are 8-aligned: no wrap inside a block
                    if (j0 + b < n) dst[j0 + b] = xb8[b];
                }
            }
            ring0 -= 8;
            if (ring0 < 0) ring0 += W;
            __syncthreads();
        }
        asm volatile("cp.async.wait_group 0;\n" ::);
    }
    __threadfence_block();
    __syncthreads();

    // ---- eta = x - (sum x / sum zv) zv (distributions.pyx:24-39); beta system (logit.py:226-231) ----
    double sx = 0.0, sz = 0.0;
    for (int i = tid; i < n; i += T) {
        sx += xs[i];
        sz += zs[i];
    }
    sx = cta4_sum(sx, red);
    sz = cta4_sum(sz, red + 4);
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
        sacc = cta4_sum(sacc, red);
        if (tid == 0) {
            small_S[a * p + b] = sacc + hp->b_prec[a * p + b];
            small_S[b * p + a] = sacc + hp->b_prec[b * p + a];
        }
    }
    for (int a = 0; a < p; a++) {  // b = X'(k - omega * eta) + b_prec b_mu
        double sacc = 0.0;
        for (int i = tid; i < n; i += T) sacc += v.X[(size_t)i * p + a] * (kd_c[i] - om_c[i] * xs[i]);
        sacc = cta4_sum(sacc, red + 4);
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
