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
// config 4 in one wave -- stay resident per SM), each with its own job, two CTA barriers per block step:
//   warp 0      the serial part: the 8 x 8 pivot block of the NEXT block step -- its rank-8 update, then its Cholesky
//               factorisation, lanes 0..7 holding one row each and exchanging columns by shuffles (8 dependent rsqrt:
//               the irreducible chain) -- runs while the other warps do the trailing update of the CURRENT step
//               (look-ahead); during the row phase it solves the two right-hand sides (eight lanes each, shuffles);
//   warps 1..3  row phase: every thread solves one window row against the factor (read as shared-memory broadcasts),
//               writes that row of L to HBM (column-band layout) and into the panel P; update phase: refill of the
//               leaving slot from the staged rows, then -- after a named barrier of these three warps -- the last warp
//               requests the rows that enter in the NEXT step (cp.async, consumed a whole phase later), forward
//               residuals, rank-8 DMMA update of the tiles that stay (compile-time list per warp, dealt by window row).
// A lone warp issues an instruction every ~6 cycles, so the instruction count of whichever warp the others wait for IS
// the critical path (-DICARW_PROF prints cycle counts per phase and warp; DESIGN.md section 4 has the history).
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
// -DICARW_PROF: cycle counts per phase and warp of chain 0, printed at the end of the kernel (measurement builds only)
#ifdef ICARW_PROF
#define ICARW_T(var) const long long var = clock64()
#define ICARW_SET(var) var = clock64()
#define ICARW_ACC(acc, t1, t0) acc += (t1) - (t0)
#else
#define ICARW_T(var)
#define ICARW_SET(var)
#define ICARW_ACC(acc, t1, t0)
#endif

// window tiles per side: the smallest compiled instance (2, 4, 6, 8: fully unrolled tile loops, constant shared-memory
// offsets) that covers bw + 8 rows, else the exact count with run-time loops (template argument 0)
__host__ __device__ inline int icarw_nt_needed(int bw) { return (bw + 8 + 7) / 8; }
__host__ __device__ inline int icarw_nt(int bw) {
    const int need = icarw_nt_needed(bw);
    return need <= 8 ? (need + 1) / 2 * 2 : need;
}
// panel pitch = 4 mod 16 doubles: the DMMA fragment loads Ps[t4 * Wp + 8 slot + g] of a half warp (g = 0..3, t4 = 0..3)
// then touch 16 distinct bank pairs (with 8 mod 16 the rows t4 and t4 + 2 collided: 4 wavefronts per load instead of 2,
// and the shared-memory pipe is what seven resident chains per SM saturate)
__host__ __device__ inline int icarw_wp(int nt) { return (nt & 1) ? 8 * nt + 12 : 8 * nt + 4; }
// pitch of the backward pass's ring of solved right-hand sides, same rule (B fragments xr[g * Wx + ring])
__host__ __device__ inline int icarw_wx(int nt) { return 8 * nt + 4; }
// staged band rows: W + 8 doubles, zero beyond the band, so that both passes read them without bounds checks (forward:
// offsets up to W - 1; backward: up to bw + 10 <= W + 2)
__host__ __device__ inline int icarw_pw(int nt) { return 8 * nt + 8; }
// shared-memory layout (doubles): region A = window tiles (forward) / x ring + block sums + two staged factor blocks
// (backward) | panel | forward residuals | one staged group of entering rows | beta system
__host__ __device__ inline size_t icarw_region_a(int nt, int bw) {
    const size_t tiles = (size_t)nt * (nt + 1) / 2 * 64, back = 8 * (size_t)icarw_wx(nt) + 128 + 2 * (8 * (size_t)icarw_pw(nt) + 32);
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

// 16-byte variant (two doubles; both addresses 16-byte aligned)
__device__ __forceinline__ void icarw_cp16(double* dst, const double* src, bool ok) {
    const uint32_t sdst = (uint32_t)__cvta_generic_to_shared(dst);
    const int bytes = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sdst), "l"(src), "r"(bytes));
}
// one band row / factor column of `count` doubles by a warp, all of it valid or none (zero fill): pairs when the rows
// are 16-byte aligned (count even: every row of a [.][count] array then starts on a 16-byte boundary)
__device__ __forceinline__ void icarw_stage_line(double* dst, const double* src, int count, bool ok, int lane) {
    if ((count & 1) == 0) {
        for (int e = 2 * lane; e < count; e += 64) icarw_cp16(dst + e, src + e, ok);
    } else {
        for (int e = lane; e < count; e += 32) icarw_cp8(dst + e, src + e, ok);
    }
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

// 1 / sqrt(a) for a pivot of Lambda (positive, far inside the FP32 range: tau Q_ii + omega_i): the FP32 hardware
// approximation (relative error < 2^-22) refined by two Newton steps y <- y (3/2 - a/2 y^2) in FP64 (2^-44, then FP64
// rounding).  Ten dependent instructions on the one chain of the factorisation that cannot be overlapped.
__device__ __forceinline__ double icarw_rsqrt(double a) {
    float yf;
    const float af = (float)a;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"(af));
    double y = (double)yf;
    const double h = 0.5 * a;
    y = y * fma(-h * y, y, 1.5);
    y = y * fma(-h * y, y, 1.5);
    return y;
}

// Cholesky factorisation of an 8 x 8 block by ONE WARP: lane a < 8 holds row a of the lower triangle (row[b], b <= a);
// column j is scaled by 1 / sqrt(pivot) (rsqrt, no divide) and exchanged by shuffles for the rank-1 update.  Writes the
// factor (row-major 8 x 8, lower) and the reciprocal pivots to shared memory.  Returns false on a non-positive pivot.
__device__ __forceinline__ bool icarw_chol8(const double* __restrict__ D, double* __restrict__ Lds, double* __restrict__ invs,
                                            int lane) {
    const int a = lane & 7;
    double row[8];
#pragma unroll
    for (int b = 0; b < 8; b++) row[b] = lane < 8 ? D[a * 8 + b] : 1.0;   // lanes 8..31 only take part in the shuffles
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const double piv = __shfl_sync(0xffffffffu, row[j], j);
        if (!(piv > 0.0)) ok = false;
        const double ri = icarw_rsqrt(piv);
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

// Rank-8 trailing update of one of the three update warps (HW = 0..2) for a compiled window size.  The tiles that stay
// are named by their slot distances (ra, rb) from the pivot slot, 1 <= rb <= ra <= NT - 1, without (1, 1) (the next
// pivot block, which warp 0 updates and factorises).  Whole ROWS ra are dealt to the warps, from the longest down in the
// order 2 0 1 1 0 2 (NT = 8: 9 / 9 / 9 tiles: the last warp has no window rows to solve up to W = 64 but issues the
// copies of the entering rows, which costs as much under load), so a warp reads
// the panel fragment of its row once and executes exactly its own tiles -- the list is a compile-time constant.  (A lone
// warp issues an instruction every ~6 cycles: the earlier version, which walked all NT (NT + 1) / 2 stored pairs under
// predicates, spent 342 issue slots per step on the critical path of the factorisation.)
// tile (hi, lo) -= P_hi P_lo' with hi / lo the larger / smaller SLOT of the pair (only slot pairs a >= b are stored);
// the A and B fragments of a slot are the same two panel entries, the minus sign is an operand modifier of the DMMA.
template <int NT, int HW>
__device__ __forceinline__ void icarw_update_tiles(double* __restrict__ Ts, const double* __restrict__ Ps, int Wp, int cs,
                                                   int g, int t4) {
    int slot[NT > 0 ? NT : 1];
#pragma unroll
    for (int r = 1; r < NT; r++) {
        const int sl = cs + r;
        slot[r] = sl >= NT ? sl - NT : sl;
    }
    const double* p0 = Ps + t4 * Wp + g;
    const double* p1 = Ps + (4 + t4) * Wp + g;
#pragma unroll
    for (int ra = NT - 1; ra >= 2; ra--) {
        const int k = (NT - 1 - ra) % 6;
        const int owner = k == 0 ? 2 : (k == 1 ? 0 : (k == 2 ? 1 : (k == 3 ? 1 : (k == 4 ? 0 : 2))));
        if (owner != HW) continue;
        const int sa = slot[ra];
        const double fa0 = p0[8 * sa], fa1 = p1[8 * sa];
#pragma unroll
        for (int rb = 1; rb <= ra; rb++) {
            const int sb = slot[rb];
            const double fb0 = p0[8 * sb], fb1 = p1[8 * sb];
            const int hi = max(sa, sb), lo = min(sa, sb);
            double2* cp = reinterpret_cast<double2*>(Ts + (hi * (hi + 1) / 2 + lo) * 64 + g * 8 + 2 * t4);
            double2 cv = *cp;
            if (sa >= sb) {   // uniform: the row slot is the tile's row side
                dgemm::dmma884(cv.x, cv.y, -fa0, fb0);
                dgemm::dmma884(cv.x, cv.y, -fa1, fb1);
            } else {
                dgemm::dmma884(cv.x, cv.y, -fb0, fa0);
                dgemm::dmma884(cv.x, cv.y, -fb1, fa1);
            }
            *cp = cv;
        }
    }
}

template <int NT_T>
__global__ void __launch_bounds__(ICARW_THREADS, NT_T > 0 ? 7 : 1) solve_logit_icar_warp_kernel(const View v, const SolveDebug dbg,
                                                                                              double* rec_beta, double* rec_tau,
                                                                                              int rec_capacity) {
    extern __shared__ __align__(16) double icw_sm[];
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, T = ICARW_THREADS;
    const int g = lane >> 2, t4 = lane & 3;
#ifdef ICARW_PROF
    const long long pr_t_start = clock64();
    long long pr_t_tau = 0, pr_t_inputs = 0, pr_t_bwd_end = 0;
#endif
    const int n = v.n, n4 = round_up_dev(v.n, 4), p = v.p, w = v.bw, W1 = v.bw + 1;
    const int NT = NT_T > 0 ? NT_T : icarw_nt(w), W = 8 * NT, Wp = icarw_wp(NT);
    const int W1p = icarw_pw(NT);
    const int n8 = round_up_dev(n, 8);
    const size_t Cp = v.Cp;
    const HParams* hp = v.hp;
    const int stg_sz = 8 * W1p + 32;              // 8 staged band rows + 4 x 8 scalars
    double* Ts = icw_sm;                          // [NT (NT + 1) / 2][8][8]  window tiles, pair (a, b), a >= b   (forward)
    const int Wx = icarw_wx(NT);
    double* xr = icw_sm;                          // [8][Wx]  backward: ring of solved x, zv, m by (matrix row mod W); rows 3..7 zero
    double* Ss = xr + 8 * Wx;                     // [4][8][4]  backward: block sums, one set per warp
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

#ifdef ICARW_PROF
    pr_t_tau = clock64();
#endif
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

#ifdef ICARW_PROF
    pr_t_inputs = clock64();
#endif
    // ---- entering rows: the 8 band rows of Q, omega and b of the group starting at `base` -> the staging buffer
    //      (cp.async by the calling warp(s): `first`, `step` give the rows each warp takes) ----
    auto stage_rows = [&](int base, int first, int step) {
        double* dst = stg;
#pragma unroll 1
        for (int rr = first; rr < 8; rr += step) {
            const int i = base + rr;
            const double* srow = v.Qband + (size_t)(i < n ? i : 0) * W1;
            if (i >= w) {   // the whole band row exists (all rows but the first bw)
                icarw_stage_line(dst + rr * W1p, srow, W1, i < n, lane);
            } else {
                for (int d = lane; d < W1; d += 32) icarw_cp8(dst + rr * W1p + d, srow + d, d <= i);
            }
        }
        if (first == 0 && lane < 16) {
            const int rr = lane & 7, i = base + rr;
            const double* srcv = lane < 8 ? om_c : xs;
            icarw_cp8(dst + 8 * W1p + (lane < 8 ? 0 : 8) + rr, i < n ? srcv + i : srcv, i < n);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    // the same for a group of complete band rows (base >= bw; rows past n are zero-filled) when the rows are 16-byte
    // aligned: four lanes per row, 16-byte pieces -- ~40 instructions instead of ~400 for the warp that issues them
    auto stage_rows_fast = [&](int base) {
        const int rr = lane >> 2, i = base + rr, q4 = 2 * (lane & 3);
        const bool ok = i < n;
        const double* src = v.Qband + (size_t)(ok ? i : 0) * W1 + q4;
        double* dst = stg + rr * W1p + q4;
        if (NT_T > 0) {   // W1 <= 8 NT - 7: at most NT pieces per lane, constant offsets
#pragma unroll
            for (int it = 0; it < (NT_T > 0 ? NT_T : 1); it++)
                if (q4 + 8 * it < W1) icarw_cp16(dst + 8 * it, src + 8 * it, ok);
        } else {
            for (int e = q4; e < W1; e += 8, dst += 8, src += 8) icarw_cp16(dst, src, ok);
        }
        if (lane < 16) {
            const int r2 = lane & 7, i2 = base + r2;
            const double* srcv = lane < 8 ? om_c : xs;
            icarw_cp8(stg + 8 * W1p + (lane < 8 ? 0 : 8) + r2, i2 < n ? srcv + i2 : srcv, i2 < n);
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
    // factor of the first pivot block; the rows that enter in step 0 (group NT) start their way to shared memory
    if (warp == 0) {
        if (!icarw_chol8(Ts, Lds, invs, lane)) piv_ok = 0;
    } else if (warp == 3) {
        stage_rows(8 * NT, 0, 1);
    }
    __syncthreads();

    // ---- blocked factorisation + fused forward substitutions ----
    const int nsteps = n8 / 8;
    const int rf_rows = g * W1p + g - 2 * t4, rf_cols = 2 * t4 * W1p + 2 * t4 - g;   // refill: lane offsets into the staged rows
#ifdef ICARW_PROF
    long long pr_row = 0, pr_b1 = 0, pr_upd = 0, pr_b2 = 0, pr_chol = 0, pr_refill = 0, pr_nb = 0, pr_resid = 0, pr_dmma = 0;
    const long long pr_t_fwd0 = clock64();
#endif
    for (int m = 0; m < nsteps; m++) {
        if (!piv_ok) break;   // uniform (read after a barrier)
        const int cs = m % NT, j0 = 8 * m;
        const int cs1 = cs + 1 == NT ? 0 : cs + 1;   // the next pivot slot
        // ===== row phase =====
        ICARW_T(pt0);
        if (warp == 0) {
            // the two right-hand sides of the pivot rows, y_r = L_JJ^-1 rhs_r: lane 8 r + b holds entry b of right-hand
            // side r; column by column the solved entry is broadcast by a shuffle and eliminated from the entries below it
            // (8 dependent shuffle + multiply-add steps instead of 36 in a lane that solves a whole column alone)
            {
                const int b = lane & 7, r = (lane >> 3) & 1;
                double t = rhs[r * W + 8 * cs + b];
                double yv = 0.0;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const double yj = __shfl_sync(0xffffffffu, t, (lane & 24) | j) * invs[j];
                    if (b == j) yv = yj;
                    if (b > j) t = fma(-yj, Lds[b * 8 + j], t);
                }
                if (lane < 16) {
                    yvs[r * 8 + b] = yv;
                    if (j0 + b < n) (r == 0 ? yb : y1)[j0 + b] = yv;
                } else if (lane < 24) {
                    if (j0 + b < n) inv_c[j0 + b] = invs[b];
                }
            }
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
#ifndef ICARW_NOLSTORE
                        if (b >= blo && b <= bhi) lp[b * (W1 - 1)] = x[b];
#else
                        if (b >= blo && b <= bhi && x[b] == 1.2345e300) lp[b * (W1 - 1)] = x[b];
#endif
                }
#pragma unroll
                for (int b = 0; b < 8; b++) Ps[b * Wp + wrow] = pivot_row ? 0.0 : x[b];
            }
            // the entering rows of this step (group m + NT) were requested one update phase ago by the last warp (which
            // has the fewest window rows, none at all up to W = 64); they are consumed right after the barrier
            if (warp == 3) asm volatile("cp.async.wait_group 0;\n" ::);
        }
        ICARW_T(pt1);
        __syncthreads();   // panel, y and the staged entering rows are visible
        // ===== update phase =====
        ICARW_T(pt2);
#ifdef ICARW_PROF
        long long ps0 = 0, ps1 = 0, ps2 = 0, ps3 = 0;
#endif
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
                ICARW_T(pc0);
                if (!icarw_chol8(Dn, Lds, invs, lane)) piv_ok = 0;
                ICARW_T(pc1);
                ICARW_ACC(pr_chol, pc1, pc0);
            }
        } else {
            const int hw = warp - 1;   // 0..2
            // the pivot slot leaves -- the row phase was the last reader of its tiles -- and takes the 8 entering rows
            // (group m + NT) FIRST, so that the staging buffer is free again for the rest of the phase
            ICARW_SET(ps0);
            if (NT_T > 0) {
                // the slot at distance r from the pivot slot holds group m + r, D = 8 (NT - r) matrix rows above the
                // entering group: Lambda[hi][lo] = tau Q[hi][hi - lo] with hi always in the entering (staged) group, so a
                // tile differs from the next only by D -- the lane's two offsets into the staged rows are loop constants
                for (int r = 1 + hw; r < NT; r += 3) {
                    const int a = cs + r >= NT ? cs + r - NT : cs + r;
                    const int D = 8 * (NT - r);
                    double v0, v1;
                    double* tile;
                    if (a > cs) {   // tile (a, cs): the columns are the entering group
                        const double* row = stg + rf_cols + D;
                        v0 = tau * row[0];
                        v1 = tau * row[W1p + 1];
                        tile = Ts + (a * (a + 1) / 2 + cs) * 64;
                    } else {        // tile (cs, a): the rows are the entering group
                        const double* row = stg + rf_rows + D;
                        v0 = tau * row[0];
                        v1 = tau * row[-1];
                        tile = Ts + (cs * (cs + 1) / 2 + a) * 64;
                    }
                    *reinterpret_cast<double2*>(tile + g * 8 + 2 * t4) = make_double2(v0, v1);
                }
                if (hw == 1) fill_tile_diag(stg, 8 * (m + NT), cs);
                if (hw == 0) fill_rhs(stg, cs, 8 * (m + NT));
            } else {
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
            ICARW_SET(ps1);
            asm volatile("bar.sync 1, 96;\n" ::: "memory");   // warps 1..3: everybody has read the staged group
            ICARW_SET(ps2);
            // the rows of the NEXT step travel (L2 / HBM latency: ~3000 cycles when exposed) while the window is updated
#ifndef ICARW_NOSTAGE   // (measurement builds: -DICARW_NOSTAGE / -DICARW_NOLSTORE drop a piece to time the rest; wrong results)
            if (warp == 3 && m + 1 < nsteps) {
                const int nb = 8 * (m + 1 + NT);
                if ((W1 & 1) == 0 && nb >= w) stage_rows_fast(nb);
                else stage_rows(nb, 0, 1);
            }
#endif
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
            ICARW_SET(ps3);
            // rank-8 trailing update on the tensor pipe: tile (a, b) -= P_a P_b' for the slots that stay, dealt round-robin
            // to the three warps; the pivot slot's tiles are being refilled and the next pivot block belongs to warp 0
            if (NT_T > 0) {
                if (hw == 0) icarw_update_tiles<NT_T, 0>(Ts, Ps, Wp, cs, g, t4);
                else if (hw == 1) icarw_update_tiles<NT_T, 1>(Ts, Ps, Wp, cs, g, t4);
                else icarw_update_tiles<NT_T, 2>(Ts, Ps, Wp, cs, g, t4);
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
        }
#ifdef ICARW_PROF
        if (warp != 0) {
            const long long ps4 = clock64();
            pr_refill += ps1 - ps0; pr_nb += ps2 - ps1; pr_resid += ps3 - ps2; pr_dmma += ps4 - ps3;
        }
#endif
        ICARW_T(pt3);
        __syncthreads();   // window updated and refilled; next factor ready; staged group consumed
        ICARW_T(pt4);
        ICARW_ACC(pr_row, pt1, pt0);
        ICARW_ACC(pr_b1, pt2, pt1);
        ICARW_ACC(pr_upd, pt3, pt2);
        ICARW_ACC(pr_b2, pt4, pt3);
    }
#ifdef ICARW_PROF
    const long long pr_t_fwd1 = clock64();
#endif
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
    for (int e = tid; e < 8 * Wx + 128 + 2 * stg_sz; e += T) xr[e] = 0.0;   // ring, sums, staging pads
    __syncthreads();
    {
        // the 8 column bands of a block are one contiguous piece of the factor: staged one block ahead (cp.async, two
        // columns per warp), with y_b, y_1, the normals and 1 / L_jj of the block.  Entries of L below row n are never
        // written: they keep the zero of the workspace, like the staged rows' pads, so nothing below needs a bounds check.
        auto stage_block = [&](int j0, int buf) {
            double* dst = stgB + buf * stg_sz;
            if (j0 >= 0) {
                if ((W1 & 1) == 0) {
                    // 16-byte pieces, eight lanes per column: warps 0 and 1 take four columns each
                    if (warp < 2) {
                        const int bb = 4 * warp + (lane >> 3), e0 = 2 * (lane & 7);
                        const bool okk = j0 + bb < n;
                        const double* scol = Lc + (size_t)(okk ? j0 + bb : 0) * W1 + e0;
                        double* d = dst + bb * W1p + e0;
                        for (int e = e0; e < W1; e += 16, d += 16, scol += 16) icarw_cp16(d, scol, okk);
                    }
                } else {
#pragma unroll 1
                    for (int bb = warp; bb < 8; bb += 4) {
                        const bool okk = j0 + bb < n;
                        const double* scol = Lc + (size_t)(okk ? j0 + bb : 0) * W1;
                        icarw_stage_line(dst + bb * W1p, scol, W1, okk, lane);
                    }
                }
                if (warp == 2) {
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
            const double* xg = xr + g * Wx;
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
            if (warp == 0) {
                // the 8 x 8 triangular solves L_JJ' x = t of the three right-hand sides: lane 8 r + b holds entry b of
                // right-hand side r; from the last column up, the solved entry is broadcast by a shuffle and eliminated
                // from the entries above it (8 dependent steps; one lane per right-hand side needed 36)
                const int b = lane & 7, r = lane >> 3;   // r = 3: idle copy of r = 2's arithmetic, never stored
                const int rr = r < 3 ? r : 2;
                const double ybv = sb[8 * W1p + b];
                double t = (rr == 0 ? ybv + sb[8 * W1p + 16 + b] : (rr == 1 ? sb[8 * W1p + 8 + b] : ybv)) -
                           ((Ss[b * 4 + rr] + Ss[32 + b * 4 + rr]) + (Ss[64 + b * 4 + rr] + Ss[96 + b * 4 + rr]));
                const double ib = sb[8 * W1p + 24 + b];   // 1 / L_jj, staged as zero for the padded columns j >= n
                double xv = 0.0;
#pragma unroll
                for (int j = 7; j >= 0; j--) {
                    const double xj = __shfl_sync(0xffffffffu, t * ib, (lane & 24) | j);
                    if (b == j) xv = xj;
                    if (b < j) t = fma(-sb[b * W1p + (j - b)], xj, t);
                }
                if (r < 3) {
                    double* dst = r == 0 ? xs : (r == 1 ? zs : ms);
                    xr[r * Wx + (j0 % W) + b] = xv;      // blocks are 8-aligned: no wrap inside a block
                    if (j0 + b < n) dst[j0 + b] = xv;
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
#ifdef ICARW_PROF
    pr_t_bwd_end = clock64();
    if (c == 0 && tid == 0 && v.ctl->iter == 3)
        printf("icarw chain 0: tau %lld | inputs + normals %lld | initial window %lld | forward %lld | backward %lld cycles\n",
               pr_t_tau - pr_t_start, pr_t_inputs - pr_t_tau, pr_t_fwd0 - pr_t_inputs, pr_t_fwd1 - pr_t_fwd0, pr_t_bwd_end - pr_t_fwd1);
    if ((c == 0 || c == 500) && lane == 0 && v.ctl->iter == 3)
        printf("icarw chain %d warp %d: steps %d  forward %lld cyc (row %lld, barrier %lld, update %lld [chol8 %lld | refill %lld, named barrier %lld, residuals %lld, dmma %lld], barrier %lld)  backward %lld cyc\n",
               c, warp, nsteps, pr_t_fwd1 - pr_t_fwd0, pr_row, pr_b1, pr_upd, pr_chol, pr_refill, pr_nb, pr_resid, pr_dmma, pr_b2, clock64() - pr_t_fwd1);
#endif

    // ---- eta = x - (sum x / sum zv) zv (distributions.pyx:24-39); beta system (logit.py:226-231) ----
    // ONE pass over the sites gathers every sum the rest needs -- sum x, sum zv, the p (p + 1) / 2 entries of X' Om X and,
    // because eta = x + shift zv is linear in the not yet known shift, the two halves X'(k - Om x) and X' Om zv of the
    // right-hand side -- and one fixed-order reduction combines them (eleven separate passes with two barriers each, every
    // one a chain of dependent global loads, cost 0.3 M cycles per chain).  p <= 4 keeps the sums in registers; larger p
    // falls back to a pass per sum.
    double shift;
    if (p <= 4) {
        constexpr int PM = 4, NS = 2 + PM * (PM + 1) / 2 + 2 * PM;   // 20 sums
        double acc[NS];
#pragma unroll
        for (int k = 0; k < NS; k++) acc[k] = 0.0;
        for (int i = tid; i < n; i += T) {
            const double xi = xs[i], zi = zs[i], om = om_c[i], kk = kd_c[i];
            double xr_[PM];
#pragma unroll
            for (int a = 0; a < PM; a++) xr_[a] = a < p ? v.X[(size_t)i * p + a] : 0.0;
            acc[0] += xi;
            acc[1] += zi;
            int k = 2;
#pragma unroll
            for (int a = 0; a < PM; a++)
#pragma unroll
                for (int bb = 0; bb <= a; bb++, k++) acc[k] += om * xr_[a] * xr_[bb];
#pragma unroll
            for (int a = 0; a < PM; a++) {
                acc[k + a] += xr_[a] * (kk - om * xi);
                acc[k + PM + a] += xr_[a] * (om * zi);
            }
        }
#pragma unroll
        for (int k = 0; k < NS; k++) acc[k] = warp_sum(acc[k]);
        double* redv = Ps;   // [4][NS]: the panel is free after the factorisation
        __syncthreads();
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < NS; k++) redv[warp * NS + k] = acc[k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < NS; k++) acc[k] = (redv[k] + redv[NS + k]) + (redv[2 * NS + k] + redv[3 * NS + k]);
        shift = -acc[0] / acc[1];
        if (tid == 0) {
            int k = 2;
#pragma unroll
            for (int a = 0; a < PM; a++)
#pragma unroll
                for (int bb = 0; bb <= a; bb++, k++)
                    if (a < p) {
                        small_S[a * p + bb] = acc[k] + hp->b_prec[a * p + bb];
                        small_S[bb * p + a] = acc[k] + hp->b_prec[bb * p + a];
                    }
#pragma unroll
            for (int a = 0; a < PM; a++)   // X'(k - Om eta) + b_prec b_mu
                if (a < p) small_b[a] = (acc[k + a] - shift * acc[k + PM + a]) + hp->b_pm[a];
        }
        for (int i = tid; i < n; i += T) {
            const double e = xs[i] + shift * zs[i];
            xs[i] = e;
            v.eta_n[(size_t)i * Cp + c] = e;
            if (dbg.enabled && dbg.out.eta_mean && c < v.C) dbg.out.eta_mean[(size_t)c * n + i] = ms[i];
        }
    } else {
        double sx = 0.0, sz = 0.0;
        for (int i = tid; i < n; i += T) {
            sx += xs[i];
            sz += zs[i];
        }
        sx = cta4_sum(sx, red);
        sz = cta4_sum(sz, red + 4);
        shift = -sx / sz;
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
#ifdef ICARW_PROF
    if (c == 0 && tid == 0 && v.ctl->iter == 3) printf("icarw chain 0: epilogue (projection, beta system, commit) %lld cycles\n", clock64() - pr_t_bwd_end);
#endif
}
