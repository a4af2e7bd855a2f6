// solve_icar_warp.cuh -- LogitICARGibbs ("PGGibbs"): fixed-pattern banded Cholesky solve, one warp per chain.
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
// Mapping.  ONE WARP owns a chain: no CTA barriers anywhere, only __syncwarp.  The factorisation is right-looking in
// blocks of 8 columns over a sliding window of W = 8 NT >= bw + 8 matrix rows, kept in shared memory as 8 x 8 tiles
// addressed by RING SLOT (8-row group index mod NT), so sliding the window moves no data: per block step one slot
// leaves (the pivot block) and is refilled with the 8 entering rows.  Only slot pairs a >= b are stored: tile (a, b)
// holds Lambda[rows of slot a][columns of slot b], whichever of the two is the lower triangle in matrix order -- the
// matrix is symmetric and the rank-8 update tile(a, b) -= P_a P_b' has the same form for every pair, so the update is
// NT (NT - 1) / 2 tiles x two DMMA.8x8x4 with the A and B fragments read from the same panel array.
// Per block step:
//   1. every lane owns window rows: it reads its row of the pivot block column (a tile row, or a tile column when the
//      pair is stored the other way round), all lanes read the 8 x 8 pivot block (shared-memory broadcast) and factorise
//      it redundantly in registers (a warp instruction costs the same for 1 or 32 lanes; rsqrt, no divide);
//   2. the lane solves its row against the factor: that row of L goes to HBM (column-band layout, read back once by the
//      backward pass) and into the panel P (column-major in shared memory: conflict-free fragment reads); the two
//      right-hand sides b and 1 ride along (forward substitution fused into the factorisation);
//   3. rank-8 trailing update of the window on the FP64 tensor pipe (DMMA), accumulators through shared memory;
//   4. the leaving slot is refilled with tau Q + diag(omega) of the entering rows.
// The backward pass is blocked the same way: the 8 x bw product L_below' [x zv m] is a chain of DMMAs (A fragments
// straight from the column-band factor in HBM/L2, prefetched one block ahead), three lanes finish the 8 x 8 triangular
// solves.
// ------------------------------------------------------------------------------------------------
constexpr int ICARW_THREADS = 32;

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
    const size_t tiles = (size_t)nt * (nt + 1) / 2 * 64, back = 8 * 8 * (size_t)nt + 32 + 2 * (8 * (size_t)icarw_pw(nt) + 32);
    return tiles > back ? tiles : back;
}
static size_t icarw_smem_bytes(int bw, int p) {
    const size_t nt = icarw_nt(bw), W = 8 * nt, Wp = icarw_wp((int)nt);
    return (icarw_region_a((int)nt, bw) + 8 * Wp + 2 * W + (8 * (size_t)icarw_pw((int)nt) + 32) + (size_t)p * p + 5 * (size_t)p) *
           sizeof(double);
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

template <int NT_T>
__global__ void __launch_bounds__(ICARW_THREADS) solve_logit_icar_warp_kernel(const View v, const SolveDebug dbg, double* rec_beta,
                                                                              double* rec_tau, int rec_capacity) {
    extern __shared__ __align__(16) double icw_sm[];
    const int c = blockIdx.x, lane = threadIdx.x, T = ICARW_THREADS;
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
    double* Ss = xr + 8 * W;                      // [8][4]   backward: block sums
    double* stgB = Ss + 32;                       // [2][stg_sz]  backward: staged factor blocks
    double* Ps = icw_sm + icarw_region_a(NT, w);  // [8][Wp]  panel, column-major: Ps[col][window row]
    double* rhs = Ps + 8 * Wp;                    // [2][W]   forward residuals of b and 1, by window row
    double* stg = rhs + 2 * W;                    // [stg_sz] forward: staged entering rows
    double* small_S = stg + stg_sz;               // [p][p]   beta system
    double* small_b = small_S + p * p;
    double* small_x = small_b + p;
    double* small_t = small_x + p;
    double* beta_s = small_t + p;
    double* nrm = beta_s + p;
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
    double* ms = zs + n4;    // L^-T L^-1 b (debug mean); eta during the tau quadratic form
    double* om_c = ms + n4;  // omega of this chain
    double* inv_c = om_c + n4;  // 1 / L_jj
    double* kd_c = inv_c + n4;  // z - 1/2 of this chain

    for (int a = lane; a < p; a += T) beta_s[a] = v.theta[(size_t)a * Cp + c];
    for (int k = lane; k < p; k += T) {
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
    if (lane == 0) ok_flag = 1;
    __syncwarp();

    // ---- tau | eta (logit.py:206-209); eta is first copied chain-major so that the band products hit L1 ----
    for (int i = lane; i < n; i += T) ms[i] = v.eta_n[(size_t)i * Cp + c];
    __syncwarp();
    double part = 0.0;
    for (int i = lane; i < n; i += T) {
        const double ei = ms[i];
        const double* qrow = v.Qband + (size_t)i * W1;
        double sacc = 0.5 * qrow[0] * ei;
        const uint8_t* nz = v.qnz + (size_t)i * (ICAR_NZ + 1);
        const int cnt = nz[0];
        if (cnt != 255) {   // the few non-zeros of the row (same order of summation as the dense walk: ascending offset)
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
    const double quad = 2.0 * warp_sum(part);
    const double rate = 0.5 * quad + hp->tau_rate;
    const double tau = dbg.tau_in ? dbg.tau_in[cr] : v.gam[c] / rate;
    if (lane == 0 && dbg.enabled && dbg.out.tau_rate && c < v.C) dbg.out.tau_rate[c] = rate;

    // ---- stage this chain's per-site inputs contiguously (the chain-minor state arrays would cost one uncoalesced
    //      access per row inside the sequential loops; the Box-Muller normals must not sit on the critical path) ----
    for (int i = lane; i < n4; i += T) {
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
    __syncwarp();

    // ---- entering rows: the 8 band rows of Q, omega and b of group `grp8 / 8` -> staging buffer `buf` (cp.async; the
    //      copies land while the block step in between runs) ----
    auto stage_rows = [&](int base) {
        double* dst = stg;
#pragma unroll 1
        for (int rr = 0; rr < 8; rr++) {
            const int i = base + rr;
            const double* srow = v.Qband + (size_t)(i < n ? i : 0) * W1;
            for (int d = lane; d < W1; d += T) icarw_cp8(dst + rr * W1p + d, srow + d, i < n && d <= i);
        }
        if (lane < 16) {
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
    auto fill_rhs = [&](const double* sb, int slot, int base) {   // right-hand sides b and 1 of the 8 rows base .. base + 7
        if (lane < 16) {
            const int rr = lane & 7, r = lane >> 3, i = base + rr;
            rhs[r * W + 8 * slot + rr] = i < n ? (r == 0 ? sb[8 * W1p + 8 + rr] : 1.0) : 0.0;
        }
    };

    // ---- initial window: matrix rows 0 .. W - 1, group s in slot s; group a is staged, its tiles against the groups
    //      b <= a are filled (every stored pair has its later group as the row side here) ----
    for (int e = lane; e < stg_sz; e += T) stg[e] = 0.0;   // the pads beyond the band stay zero for good
    __syncwarp();
    for (int a = 0; a < NT; a++) {
        stage_rows(8 * a);
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncwarp();
        for (int b = 0; b < a; b++) fill_tile_off(stg, 8 * a, a, b, 8 * a, 8 * b);
        fill_tile_diag(stg, 8 * a, a);
        fill_rhs(stg, a, 8 * a);
        __syncwarp();
    }
    stage_rows(8 * NT);   // the group entering at the end of block step 0

    // ---- blocked factorisation + fused forward substitutions ----
    const int passes = (W + 31) >> 5;
    bool ok = true;
    for (int m = 0; m < n8 / 8; m++) {
        const int cs = m % NT, j0 = 8 * m;
        // (1) pivot block: lower triangle of tile (cs, cs), factorised redundantly by every lane
        double Ld[8][8], inv[8];
        {
            const double* D = Ts + (size_t)tri(cs, cs) * 64;
#pragma unroll
            for (int a = 0; a < 8; a++)
#pragma unroll
                for (int b = 0; b <= a; b++) Ld[a][b] = D[a * 8 + b];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const double piv = Ld[j][j];
                if (!(piv > 0.0)) ok = false;
                const double ri = rsqrt(piv);
                inv[j] = ri;
                Ld[j][j] = piv * ri;
#pragma unroll
                for (int a = j + 1; a < 8; a++) Ld[a][j] *= ri;
#pragma unroll
                for (int a = j + 1; a < 8; a++)
#pragma unroll
                    for (int b = j + 1; b <= a; b++) Ld[a][b] = fma(-Ld[a][j], Ld[b][j], Ld[a][b]);
            }
        }
        if (!ok) break;   // uniform: every lane factorised the same block
        // (2) the two right-hand sides of the pivot rows: y = L_JJ^-1 rhs, computed redundantly (no broadcast needed)
        double yv[2][8];
#pragma unroll
        for (int r = 0; r < 2; r++) {
#pragma unroll
            for (int b = 0; b < 8; b++) {
                double t = rhs[r * W + 8 * cs + b];
#pragma unroll
                for (int k = 0; k < b; k++) t = fma(-yv[r][k], Ld[b][k], t);
                yv[r][b] = t * inv[b];
            }
        }
        if (lane < 16) {   // y_b = L^-1 b, y_1 = L^-1 1 and 1 / L_jj of the block
            const int b = lane & 7, r = lane >> 3;
            if (j0 + b < n) {
                double val = 0.0, iv = 0.0;
#pragma unroll
                for (int bb = 0; bb < 8; bb++)
                    if (bb == b) {
                        val = r == 0 ? yv[0][bb] : yv[1][bb];
                        iv = inv[bb];
                    }
                (r == 0 ? yb : y1)[j0 + b] = val;
                if (r == 0) inv_c[j0 + b] = iv;
            }
        }
        // (3) every window row against the factor: L (to HBM), the panel P and the forward residuals
        for (int ps = 0; ps < passes; ps++) {
            const int wrow = ps * 32 + lane;
            if (wrow < W) {
                const int a = wrow >> 3, rr = wrow & 7;
                const int grp = m + (a >= cs ? a - cs : a - cs + NT);   // 8-row group held by slot a
                const int dt = 8 * (grp - m) + rr;                       // matrix row - j0
                double raw[8], x[8];
                if (a >= cs) {
                    const double* src = Ts + (size_t)tri(a, cs) * 64 + rr * 8;
#pragma unroll
                    for (int b = 0; b < 8; b += 2) {
                        const double2 t2 = *reinterpret_cast<const double2*>(src + b);
                        raw[b] = t2.x;
                        raw[b + 1] = t2.y;
                    }
                } else {
                    const double* src = Ts + (size_t)tri(cs, a) * 64 + rr;
#pragma unroll
                    for (int b = 0; b < 8; b++) raw[b] = src[b * 8];
                }
#pragma unroll
                for (int b = 0; b < 8; b++) {
                    double t = raw[b];
#pragma unroll
                    for (int k = 0; k < b; k++) t = fma(-x[k], Ld[b][k], t);
                    x[b] = t * inv[b];
                }
                const bool pivot_row = a == cs;
                const int i = j0 + dt;
                if (i < n) {
#pragma unroll
                    for (int b = 0; b < 8; b++) {
                        const int e = dt - b;
                        if (e >= 0 && e <= w && j0 + b < n) Lc[(size_t)(j0 + b) * W1 + e] = x[b];
                    }
                }
#pragma unroll
                for (int b = 0; b < 8; b++) Ps[b * Wp + wrow] = pivot_row ? 0.0 : x[b];
                if (!pivot_row) {
                    double r0 = rhs[wrow], r1 = rhs[W + wrow];
#pragma unroll
                    for (int b = 0; b < 8; b++) {
                        r0 = fma(-yv[0][b], x[b], r0);
                        r1 = fma(-yv[1][b], x[b], r1);
                    }
                    rhs[wrow] = r0;
                    rhs[W + wrow] = r1;
                }
            }
        }
        __syncwarp();
        // (4) rank-8 trailing update on the tensor pipe: tile (a, b) -= P_a P_b' for EVERY stored pair -- the pivot
        //     slot's panel rows are zero, so its (dead) tiles are left unchanged and the loops carry no conditions
        if (NT_T > 0) {
#pragma unroll
            for (int a = 0; a < (NT_T > 0 ? NT_T : 1); a++) {
                const double a0 = -Ps[t4 * Wp + 8 * a + g], a1 = -Ps[(4 + t4) * Wp + 8 * a + g];
#pragma unroll
                for (int b = 0; b <= a; b++) {
                    const double b0 = Ps[t4 * Wp + 8 * b + g], b1 = Ps[(4 + t4) * Wp + 8 * b + g];
                    double2* cp = reinterpret_cast<double2*>(Ts + (a * (a + 1) / 2 + b) * 64 + g * 8 + 2 * t4);
                    double2 cv = *cp;
                    dgemm::dmma884(cv.x, cv.y, a0, b0);
                    dgemm::dmma884(cv.x, cv.y, a1, b1);
                    *cp = cv;
                }
            }
        } else {
            double2* cp = reinterpret_cast<double2*>(Ts + g * 8 + 2 * t4);
            for (int a = 0; a < NT; a++) {
                const double a0 = -Ps[t4 * Wp + 8 * a + g], a1 = -Ps[(4 + t4) * Wp + 8 * a + g];
                for (int b = 0; b <= a; b++, cp += 32) {
                    const double b0 = Ps[t4 * Wp + 8 * b + g], b1 = Ps[(4 + t4) * Wp + 8 * b + g];
                    double2 cv = *cp;
                    dgemm::dmma884(cv.x, cv.y, a0, b0);
                    dgemm::dmma884(cv.x, cv.y, a1, b1);
                    *cp = cv;
                }
            }
        }
        // (5) the pivot slot leaves; it takes the 8 entering rows (group m + NT), staged during this block step
        {
            const int rn = 8 * (m + NT);
            asm volatile("cp.async.wait_group 0;\n" ::);
            __syncwarp();
            const double* sb = stg;
            for (int a = 0; a < NT; a++) {
                if (a == cs) {
                    fill_tile_diag(sb, rn, cs);
                } else {
                    const int grp = m + (a >= cs ? a - cs : a - cs + NT);
                    if (a > cs) fill_tile_off(sb, rn, a, cs, 8 * grp, rn);
                    else fill_tile_off(sb, rn, cs, a, rn, 8 * grp);
                }
            }
            fill_rhs(sb, cs, rn);
            __syncwarp();             // the staged group is consumed: the buffer takes the next one
            stage_rows(rn + 8);       // lands during the next block step
        }
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
    if (!ok) {
        if (lane == 0) {
            v.status[c] = 1;
            if (c < v.C) atomicAdd(&v.ctl->fails, 1);
        }
        return;
    }
    __threadfence_block();
    __syncwarp();
    if (dbg.enabled && dbg.out.eta_chol && c < v.C)  // column-band factor [n][bw + 1]
        for (size_t idx = lane; idx < (size_t)n * W1; idx += T) {
            const int j = (int)(idx / W1), e = (int)(idx - (size_t)j * W1);
            dbg.out.eta_chol[(size_t)c * n * W1 + idx] = (j + e < n) ? Lc[idx] : 0.0;
        }

    // ---- blocked backward: L' x = y_b + eps, L' zv = y_1, L' m = y_b (its arrays take over the window's memory) ----
    __syncwarp();
    for (int e = lane; e < 8 * W + 32 + 2 * stg_sz; e += T) xr[e] = 0.0;   // ring, sums, staging pads
    __syncwarp();
    {
        // the 8 column bands of a block are one contiguous piece of the factor: staged one block ahead (cp.async), with
        // y_b, y_1, the normals and 1 / L_jj of the block.  Entries of L below row n are never written: they keep the
        // zero of the workspace, like the staged rows' pads, so nothing below needs a bounds check.
        auto stage_block = [&](int j0, int buf) {
            double* dst = stgB + buf * stg_sz;
            if (j0 >= 0) {
#pragma unroll 1
                for (int bb = 0; bb < 8; bb++) {
                    const bool okk = j0 + bb < n;
                    const double* scol = Lc + (size_t)(okk ? j0 + bb : 0) * W1;
                    for (int e = lane; e < W1; e += T) icarw_cp8(dst + bb * W1p + e, scol + e, okk);
                }
                {
                    const int which = lane >> 3, bb = lane & 7, j = j0 + bb;
                    const double* srcv = which == 0 ? yb : (which == 1 ? y1 : (which == 2 ? zs : inv_c));
                    icarw_cp8(dst + 8 * W1p + which * 8 + bb, j < n ? srcv + j : srcv, j < n);
                }
            }
            asm volatile("cp.async.commit_group;\n" ::);
        };
        const int ksteps = (w + 3) >> 2;   // rows j0 + 8 .. j0 + 7 + bw below the block
        const int mlast = n8 / 8 - 1;
        stage_block(8 * mlast, mlast & 1);
        int ring0 = (8 * mlast + 8 + t4) % W;   // ring slot of row j0 + 8 + t4
        for (int m = mlast; m >= 0; m--) {
            const int j0 = 8 * m;
            stage_block(j0 - 8, (m + 1) & 1);                 // next block; its buffer was consumed one block ago
            asm volatile("cp.async.wait_group 1;\n" ::);     // this block's copies have landed
            __syncwarp();
            const double* sb = stgB + (m & 1) * stg_sz;
            // S[b][r] = sum_i L[i][j0 + b] x_r[i]:  D (8 columns b x 8 right-hand sides) += A (8 x 4: L[i][j0 + g]) B (4 x 8: x_g[i])
            double s0a = 0.0, s1a = 0.0, s0b = 0.0, s1b = 0.0;
            const double* lcol = sb + g * W1p + (8 + t4 - g);   // entry e = 8 + 4 ks + t4 - g of column j0 + g
            const double* xg = xr + g * W;
            int ring = ring0;
#pragma unroll 2
            for (int ks = 0; ks < ksteps; ks += 2) {
                int ring1 = ring + 4;
                if (ring1 >= W) ring1 -= W;
                dgemm::dmma884(s0a, s1a, lcol[4 * ks], xg[ring]);
                dgemm::dmma884(s0b, s1b, lcol[4 * ks + 4], xg[ring1]);   // (an odd count runs one step into the zero pad)
                ring = ring1 + 4;
                if (ring >= W) ring -= W;
            }
            if (t4 < 2) {   // right-hand sides 0..3 live in the lanes t4 = 0, 1
                Ss[g * 4 + 2 * t4] = s0a + s0b;
                Ss[g * 4 + 2 * t4 + 1] = s1a + s1b;
            }
            __syncwarp();
            if (lane < 3) {  // one lane per right-hand side finishes the 8 x 8 triangular solve
                const int r = lane;
                double xb8[8];
#pragma unroll
                for (int b = 7; b >= 0; b--) {
                    const double ybv = sb[8 * W1p + b];
                    double t = (r == 0 ? ybv + sb[8 * W1p + 16 + b] : (r == 1 ? sb[8 * W1p + 8 + b] : ybv)) - Ss[b * 4 + r];
#pragma unroll
                    for (int bb = b + 1; bb < 8; bb++) t = fma(-sb[b * W1p + (bb - b)], xb8[bb], t);
                    xb8[b] = t * sb[8 * W1p + 24 + b];   // 1 / L_jj, staged as zero for the padded columns j >= n
                }
                double* dst = r == 0 ? xs : (r == 1 ? zs : ms);
                int slot = j0 % W;
#pragma unroll
                for (int b = 0; b < 8; b++) {
                    xr[r * W + slot + b] = xb8[b];       // blocks are 8-aligned: no wrap inside a block
                    if (j0 + b < n) dst[j0 + b] = xb8[b];
                }
            }
            ring0 -= 8;
            if (ring0 < 0) ring0 += W;
            __syncwarp();
        }
        asm volatile("cp.async.wait_group 0;\n" ::);
    }
    __threadfence_block();
    __syncwarp();

    // ---- eta = x - (sum x / sum zv) zv (distributions.pyx:24-39); beta system (logit.py:226-231) ----
    double sx = 0.0, sz = 0.0;
    for (int i = lane; i < n; i += T) {
        sx += xs[i];
        sz += zs[i];
    }
    sx = warp_sum(sx);
    sz = warp_sum(sz);
    const double shift = -sx / sz;
    for (int i = lane; i < n; i += T) {
        const double e = xs[i] + shift * zs[i];
        xs[i] = e;
        v.eta_n[(size_t)i * Cp + c] = e;
        if (dbg.enabled && dbg.out.eta_mean && c < v.C) dbg.out.eta_mean[(size_t)c * n + i] = ms[i];
    }
    __syncwarp();
    for (int idx = 0; idx < p * p; idx++) {  // A = X' Om X + b_prec
        const int a = idx / p, b = idx - a * p;
        if (b > a) continue;
        double sacc = 0.0;
        for (int i = lane; i < n; i += T) sacc += om_c[i] * v.X[(size_t)i * p + a] * v.X[(size_t)i * p + b];
        sacc = warp_sum(sacc);
        if (lane == 0) {
            small_S[a * p + b] = sacc + hp->b_prec[a * p + b];
            small_S[b * p + a] = sacc + hp->b_prec[b * p + a];
        }
    }
    for (int a = 0; a < p; a++) {  // b = X'(k - omega * eta) + b_prec b_mu
        double sacc = 0.0;
        for (int i = lane; i < n; i += T) sacc += v.X[(size_t)i * p + a] * (kd_c[i] - om_c[i] * xs[i]);
        sacc = warp_sum(sacc);
        if (lane == 0) small_b[a] = sacc + hp->b_pm[a];
    }
    __syncwarp();
    if (dbg.enabled && c < v.C) {
        if (dbg.out.beta_prec)
            for (int idx = lane; idx < p * p; idx += T) dbg.out.beta_prec[(size_t)c * p * p + idx] = small_S[idx];
        if (dbg.out.beta_b)
            for (int a = lane; a < p; a += T) dbg.out.beta_b[(size_t)c * p + a] = small_b[a];
    }
    __syncwarp();
    if (lane == 0) {
        if (small_precision_mvnorm(small_S, p, small_b, nrm, small_x, small_t)) {
            for (int a = 0; a < p; a++) beta_s[a] = small_x[a];
        } else {
            ok_flag = 0;
        }
    }
    __syncwarp();
    if (!ok_flag) {
        if (lane == 0) {
            v.status[c] = 1;
            if (c < v.C) atomicAdd(&v.ctl->fails, 1);
        }
        return;
    }
    for (int a = lane; a < p; a += T) v.theta[(size_t)a * Cp + c] = beta_s[a];
    if (lane == 0) v.tau[c] = tau;
    if (rec_beta && rec_index >= 0 && rec_index < rec_capacity && c < v.C) {
        const size_t row = (size_t)c * rec_capacity + rec_index;
        for (int a = lane; a < p; a += T) rec_beta[row * p + a] = beta_s[a];
        if (lane == 0) rec_tau[row] = tau;
    }
}
