// site_pass.cuh -- the per-site half of the sweep: z, omega_a, omega_b (logit: PG draws, psi product; probit: truncated normals).
// Part of the single translation unit engine.cu (included there, in order; not a stand-alone header).
#pragma once

// ------------------------------------------------------------------------------------------------
// site pass (logit): z-update, omega_a draws + alpha partial sums, omega_b draw
//
// CTA = 128 chains x SITE_TILE sites, thread = chain.
//  phase 1 (lock-step over the tile's sites, coalesced): z | psi, alpha; z written; psi kept in shared memory as
//          the PG argument; the z bits and the set of sites whose visits need omega_a draws kept in registers.
//          The occupancy probability is evaluated in FP32 log2 space by production engines (it only feeds a
//          Bernoulli draw) and in FP64 by debug / validation engines, which expose it for the 1e-10 parity tests.
//  phase 2 (lanes desynchronise): one rejection round of the PG sampler per loop trip (pg1_round).  Two kinds of work:
//          * omega_a: every lane walks ITS OWN list of visit rows (the visits of the sites its chain occupies); the
//            accepted draw goes into the lane's FP64 accumulators of W_e' Om_a W_e and W_e'(y_e - 1/2), so this work is
//            tied to the lane;
//          * omega_b: the SITE_TILE x 32 draws PG(1, psi) of a warp are a POOL: arguments and results live in shared
//            memory, so any lane can serve any (site, chain) pair.  A lane without a visit draw takes the next pool
//            task (warp ballot + prefix count: no atomics).  The pool evens out the lanes' different list lengths and
//            rejection counts -- all 32 lanes stay busy until both kinds of work run out together.
//          An accepted lane commits and fetches its next draw in the same trip, a rejected lane retries: no lane waits
//          for another lane's rejection loop ("compaction by refill").
//  phase 3: omega_b (shared memory) is written out coalesced.
// Philox streams are keyed by (chain, visit row | site, iteration, stage), so the result does not depend on which
// lane served a draw or on how lanes interleave.
// ------------------------------------------------------------------------------------------------
//  SLICES (tcgen05 Gram path): phase 3 writes omega as SA base-256 digits of its 48-bit fixed-point value and z as one
//          more byte plane, site-contiguous per chain (gram_tc.cuh: Aq[s][chain][site]) -- the A operand of the Gram
//          kernel, 7 bytes per site instead of the two FP64 arrays omega and kd; a debug engine also writes omega.
// resident CTAs per SM the register allocation aims at: 8 (64 registers) for the usual r <= 4, fewer when the
// r (r + 1) / 2 + r FP64 accumulators no longer fit
constexpr int site_min_blocks(int r) { return r <= 4 ? 8 : (r <= 6 ? 5 : 3); }
//  FUSE_PSI: psi = [X K] theta of the tile is computed here, on the FP64 tensor pipe (DMMA m8n8k4: 8 sites x 32 chains
//          per warp and step, operands straight from L1/L2), instead of by a separate GEMM that writes psi [n x chains]
//          to HBM for this kernel to read back: the DMMA work hides under the ALU-bound PG sampling of the other warps.
// psi[SITE_TILE sites][32 chains of this warp] = U theta on the FP64 tensor pipe:  D (8 sites x 8 chains) += A (8 x 4:
// UT'[site][k]) B (4 x 8: theta[k][chain]); rows of UT and theta beyond pq are zero up to kq_pad.  The loop nest keeps one
// block of 8 chains fixed and sweeps k once for all the tile's sites, so every element of theta is read ONCE per tile (it
// comes from L2: 0.5 MB per 128 chains) and the small UT tile (SITE_TILE x pq, shared by the CTA's warps) is what gets
// re-read, from L1.  Result into dst[site][chain column of the CTA]; a warp only touches its own 32 columns.
template <typename OutT>
__device__ __forceinline__ void psi_tile_dmma(const View& v, int i0, int tid, int c, OutT (*dst)[CHAIN_TILE]) {
    constexpr int SB = SITE_TILE / 8;
    const int lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int cw = c - lane;   // first chain of this warp
    const size_t Cp = v.Cp;
    const int ksteps = (v.pq + 3) >> 2;
    const double* ua = v.UT + (size_t)t4 * v.n_pad + i0 + g;       // n_pad >= i0 + SITE_TILE: no bounds check
#pragma unroll 1
    for (int nb = 0; nb < 4; nb++) {
        double acc[SB][2];
#pragma unroll
        for (int sb = 0; sb < SB; sb++) acc[sb][0] = acc[sb][1] = 0.0;
        const double* tb = v.theta + (size_t)t4 * Cp + cw + nb * 8 + g;
#pragma unroll 2
        for (int ks = 0; ks < ksteps; ks++) {
            const double b = tb[(size_t)ks * 4 * Cp];
            double a[SB];
#pragma unroll
            for (int sb = 0; sb < SB; sb++) a[sb] = ua[(size_t)ks * 4 * v.n_pad + sb * 8];
#pragma unroll
            for (int sb = 0; sb < SB; sb++) dgemm::dmma884(acc[sb][0], acc[sb][1], a[sb], b);
        }
#pragma unroll
        for (int sb = 0; sb < SB; sb++) {
            dst[sb * 8 + g][(tid & ~31) + nb * 8 + 2 * t4] = (OutT)acc[sb][0];
            dst[sb * 8 + g][(tid & ~31) + nb * 8 + 2 * t4 + 1] = (OutT)acc[sb][1];
        }
    }
    __syncwarp();
}
// PSI_MODE: 0 = psi read from v.psi (separate GEMM), 1 = formed here with an FP64 staging tile (debug / validation engines:
// the occupancy probability is FP64), 2 = formed here and kept in FP32 only (production: two DMMA passes of 16 sites
// write straight into the PG argument tile).
constexpr int PSI_NONE = 0, PSI_F64 = 1, PSI_F32 = 2;
// shared-memory copy of the tile's visit rows (design row + y - 1/2): rows that fit 7 KB
constexpr int site_stage_rows(int r) { return 7168 / (8 * (r + 1)); }
template <typename PGT, int R>
constexpr size_t site_pass_logit_smem(int psi_mode) {
    return sizeof(PGT) * SITE_TILE * CHAIN_TILE + (psi_mode == PSI_F64 ? SITE_TILE * CHAIN_TILE * sizeof(double) : 0) +
           (size_t)site_stage_rows(R) * (R + 1) * sizeof(double) + (SITE_TILE + 4) * sizeof(int) + SITE_TILE;
}
template <int R, typename PGT, bool SLICES, int PSI_MODE>
__global__ void __launch_bounds__(CHAIN_TILE, site_min_blocks(R)) site_pass_logit_kernel(const View v) {
    constexpr int NA = R * (R + 1) / 2, NV = NA + R;
    constexpr int ROWS_MAX = site_stage_rows(R);
    static_assert(SITE_TILE <= 64, "per-lane site masks are 64-bit");
    static_assert(PSI_MODE != PSI_F32 || std::is_same<PGT, float>::value, "FP32 psi feeds the FP32 sampler");
    extern __shared__ __align__(16) unsigned char site_smem[];
    PGT (*arg_s)[CHAIN_TILE] = reinterpret_cast<PGT(*)[CHAIN_TILE]>(site_smem);  // psi in, omega_b out
    // PSI_F64: FP64 psi of the tile, [SITE_TILE][CHAIN_TILE]; a warp only touches its own 32 columns
    unsigned char* sp = site_smem + sizeof(PGT) * SITE_TILE * CHAIN_TILE;
    double (*psi_d)[CHAIN_TILE] = reinterpret_cast<double(*)[CHAIN_TILE]>(sp);
    if (PSI_MODE == PSI_F64) sp += SITE_TILE * CHAIN_TILE * sizeof(double);
    double* wt_s = reinterpret_cast<double*>(sp);                 // [ROWS_MAX][R + 1]: w_ij, y_ij - 1/2
    int* vp_s = reinterpret_cast<int*>(wt_s + ROWS_MAX * (R + 1));   // visit_ptr[i0 .. i0 + SITE_TILE]
    uint8_t* cls_s = reinterpret_cast<uint8_t*>(vp_s + SITE_TILE + 4);
    const int tid = threadIdx.x;
    const uint32_t iter = v.ctl->iter;
    const bool first = v.ctl->first != 0;
    // FP64 occupancy probabilities: validation build and debug engines (which dump them, OCCU_F_ZPROB)
    const bool zf64 = PSI_MODE == PSI_F32 ? false : (std::is_same<PGT, double>::value || v.zprob != nullptr);
    const size_t Cp = v.Cp;
    const int tile = blockIdx.x;
    const int c = blockIdx.y * CHAIN_TILE + tid;
    const int i0 = tile * SITE_TILE;
    const int i1 = min(i0 + SITE_TILE, v.n);
    const uint32_t gc = (uint32_t)(v.chain_offset + c);

    // ---- stage the tile's share of the CSR visit layout: offsets, site classes, and (when they fit) the visit rows ----
    if (tid <= SITE_TILE) vp_s[tid] = v.visit_ptr[min(i0 + tid, v.n)];
    if (tid < SITE_TILE) cls_s[tid] = i0 + tid < v.n ? v.site_class[i0 + tid] : (uint8_t)OCCU_SITE_NOT_SURVEYED;
    const int row0 = v.visit_ptr[i0];
    const int nrows = v.visit_ptr[i1] - row0;
    const bool staged = nrows <= ROWS_MAX;
    if (staged) {
        for (int idx = tid; idx < nrows * (R + 1); idx += CHAIN_TILE) {
            const int row = idx / (R + 1), k = idx - row * (R + 1);
            wt_s[idx] = k < R ? v.W[(size_t)(row0 + row) * R + k] : (double)v.y[row0 + row] - 0.5;
        }
    }
    __syncthreads();
    // design row of visit vv and y - 1/2 (logit.py:189, 221)
    auto load_row = [&](int vv, double (&w)[R], double& yk) {
        if (staged) {
            const double* rowp = wt_s + (vv - row0) * (R + 1);
#pragma unroll
            for (int k = 0; k < R; k++) w[k] = rowp[k];
            yk = rowp[R];
        } else {
#pragma unroll
            for (int k = 0; k < R; k++) w[k] = v.W[(size_t)vv * R + k];
            yk = (double)v.y[vv] - 0.5;
        }
    };

    double al[R];
#pragma unroll
    for (int k = 0; k < R; k++) al[k] = v.alpha[(size_t)k * Cp + c];

    // ---- phase 1: z | . ----
    uint64_t zbits = 0ull;   // z of the tile's sites
    uint64_t pend = 0ull;    // sites whose visits get omega_a draws: surveyed and occupied
    // psi of the whole tile first: straight into the PG argument tile as FP32 (the only precision production engines use
    // psi in), or into the FP64 staging tile of a debug engine
    if (PSI_MODE == PSI_F32) psi_tile_dmma(v, i0, tid, c, reinterpret_cast<float(*)[CHAIN_TILE]>(site_smem));
    if (PSI_MODE == PSI_F64) psi_tile_dmma(v, i0, tid, c, psi_d);
    for (int i = i0; i < i1; i++) {
        const int cls = cls_s[i - i0];
        const size_t off = (size_t)i * Cp + c;
        const double psi = PSI_MODE == PSI_F32 ? (double)arg_s[i - i0][tid] : (PSI_MODE == PSI_F64 ? psi_d[i - i0][tid] : v.psi[off]);
        if (PSI_MODE == PSI_F64 && v.zprob) v.psi[off] = psi;   // debug engines keep psi readable
        int zc;
        if (first) {
            zc = v.z[off];
        } else {
            if (cls == OCCU_SITE_OBS) {
                zc = 1;
            } else {
                // logit.py:241-245 / 249:  pr = expit(psi) prod / (1 - expit(psi) + expit(psi) prod) with
                // prod = prod_j expit(-w_ij alpha)  ==  1 / (1 + exp(-psi) prod_j (1 + exp(w_ij alpha))): one exp per
                // visit and one division per site; not-surveyed sites have the empty product
                Philox g;
                g.init(v.seed, gc, (uint32_t)i, iter, ST_Z);
                const int v0 = vp_s[i - i0], v1 = vp_s[i - i0 + 1];
                if (zf64) {
                    g.refill_rk(v.rk);
                    double grow = exp(-psi);
                    if (cls == OCCU_SITE_NOT_OBS) {
                        for (int vv = v0; vv < v1; vv++) {
                            double w[R], yk;
                            load_row(vv, w, yk);
                            double lin = 0.0;
#pragma unroll
                            for (int k = 0; k < R; k++) lin += w[k] * al[k];
                            grow *= 1.0 + exp(lin);
                        }
                    }
                    const double pr = 1.0 / (1.0 + grow);
                    if (v.zprob) v.zprob[off] = pr;
                    const uint64_t bits = ((uint64_t)g.buf[0] << 21) ^ (uint64_t)(g.buf[1] >> 11);  // == g.unif_d()
                    zc = ((double)bits + 0.5) * (1.0 / 9007199254740992.0) < pr ? 1 : 0;  // logit.py:246 / 250
                } else {
                    // same probability in FP32 log2 space (MUFU ex2 / lg2): lg = log2(exp(-psi) prod (1 + exp(w alpha)));
                    // an overflowing factor gives lg = +inf and probability 0, its limit
                    g.refill_rk<PHILOX_FAST_ROUNDS>(v.rk);
                    float lg = -1.4426950408889634f * (float)psi;
                    if (cls == OCCU_SITE_NOT_OBS) {
                        for (int vv = v0; vv < v1; vv++) {
                            double w[R], yk;
                            load_row(vv, w, yk);
                            double lin = 0.0;
#pragma unroll
                            for (int k = 0; k < R; k++) lin += w[k] * al[k];
                            lg += f_lg2(1.0f + f_ex2(1.4426950408889634f * (float)lin));
                        }
                    }
                    // z = 1 with probability 1 / (1 + grow):  U < 1 / (1 + grow)  <=>  U (1 + grow) < 1; a small U
                    // keeps its full relative precision (the integer is converted with a floating exponent)
                    const float u = fmaf(__uint2float_rz(g.buf[0]), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
                    zc = u * (1.0f + f_ex2(lg)) < 1.0f ? 1 : 0;
                }
            }
            v.z[off] = (uint8_t)zc;
            if (!SLICES) v.kd[off] = (double)zc - 0.5;  // logit.py:252
        }
        zbits |= (uint64_t)zc << (i - i0);
        if (cls != OCCU_SITE_NOT_SURVEYED && zc) pend |= 1ull << (i - i0);
        if (PSI_MODE != PSI_F32) arg_s[i - i0][tid] = (PGT)psi;
    }
    __syncwarp();   // phase 2 reads the warp's 32 columns of arg_s across lanes

    // ---- phase 2: PG draws ----
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) acc[k] = 0.0;
    const int lane = tid & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int wbase = tid - lane;              // first column of this warp within the tile
    const uint32_t gcw = gc - (uint32_t)lane;  // global chain id of the warp's first lane
    const int pool_total = (i1 - i0) * 32;     // omega_b draws of this warp, task = site-in-tile * 32 + chain-in-warp
    int pool_next = 0;                         // warp-uniform: next task not handed out yet
    int vv = 0, vend = 0, cur = 0;
    bool have = false, is_visit = false;
    typename PGSel<PGT>::Task task;
    Philox g;
    double wv[R], wy = 0.0;  // design row and y - 1/2 of the visit being drawn (kept from the fetch to the commit)
#pragma unroll
    for (int k = 0; k < R; k++) wv[k] = 0.0;
    for (;;) {
        // -- fetch: the lane's next visit draw (omega_a ~ PG(1, w_ij alpha), logit.py:187-193), else a pool task
        //    (omega_b ~ PG(1, psi), logit.py:201-204) --
        bool fetched = false, want_pool = false;
        PGT carg = (PGT)0;
        uint32_t pchain = gc, pelem = 0u, pstage = ST_OMEGA_A;
        if (!have) {
            if (vv >= vend && pend) {   // next occupied surveyed site of this lane's chain
                const int s = __ffsll((long long)pend) - 1;
                pend &= pend - 1ull;
                vv = vp_s[s];
                vend = vp_s[s + 1];
            }
            if (vv < vend) {
                cur = vv++;
                load_row(cur, wv, wy);
                double lin = 0.0;
#pragma unroll
                for (int k = 0; k < R; k++) lin += wv[k] * al[k];
                carg = (PGT)lin;
                pelem = (uint32_t)cur;
                is_visit = true;
                fetched = true;
            } else {
                want_pool = true;
            }
        }
        const unsigned wants = __ballot_sync(0xffffffffu, want_pool);
        if (wants) {
            const int mine = pool_next + __popc(wants & lt_mask);
            pool_next += __popc(wants);
            if (want_pool && mine < pool_total) {
                cur = mine;
                const int s = mine >> 5, ch = mine & 31;
                carg = arg_s[s][wbase + ch];
                pchain = gcw + (uint32_t)ch;
                pelem = (uint32_t)(i0 + s);
                pstage = ST_OMEGA_B;
                is_visit = false;
                fetched = true;
            }
        }
        if (fetched) {
            g.init(v.seed, pchain, pelem, iter, pstage);
            PGSel<PGT>::setup(carg, task);  // one copy of the setup code for both kinds of draw
            have = true;
        }
        if (!__any_sync(0xffffffffu, have)) break;
        // -- one rejection round --
        if (have) {
            PGT ub, ux, u3, us, x;
            RT<PGT>::unif4_rk(g, v.rk, ub, ux, u3, us);
            if (PGSel<PGT>::round(task, ub, ux, u3, us, x)) {
                have = false;
                if (is_visit) {
                    const double om = 0.25 * (double)x;
                    const double yk = wy;
                    int idx = 0;
#pragma unroll
                    for (int a = 0; a < R; a++) {
                        const double owa = om * wv[a];
#pragma unroll
                        for (int b = 0; b <= a; b++, idx++) acc[idx] = fma(owa, wv[b], acc[idx]);  // logit.py:222
                    }
#pragma unroll
                    for (int a = 0; a < R; a++) acc[NA + a] = fma(wv[a], yk, acc[NA + a]);  // logit.py:223
                } else {
                    arg_s[cur >> 5][wbase + (cur & 31)] = (PGT)0.25 * x;
                }
            }
        }
    }
    __syncwarp();   // phase 3 reads the lane's own column, written by whichever lane served the task

    // ---- phase 3: write-out ----
    if (!SLICES)
        for (int ii = i0; ii < i1; ii++) v.omega[(size_t)ii * Cp + c] = (double)arg_s[ii - i0][tid];
    if (SLICES) {
        // omega as a fixed-point number of omega / 8.  FP32 draws (production): 32 bits -- exact for every omega >=
        // 2^-6 (a float has 24 significant bits), rounded to 2^-29 below that, which is the resolution the FP32
        // arithmetic of the draw has anyway; only the first NPL = 4 byte planes are written and the Gram kernel skips
        // the pairs of the two empty ones.  FP64 draws (validation build): all 48 bits.  omega IS that fixed-point
        // value: a debug engine's FP64 copy is written from it.
        constexpr int NPL = std::is_same<PGT, float>::value ? gtc::SA_F32 : gtc::SA;
        static_assert(SITE_TILE % 16 == 0, "16 sites per 128-bit store");
        const size_t plane = (size_t)Cp * v.tc_Kp;
        uint8_t* row = v.Aq + (size_t)c * v.tc_Kp + i0;
#pragma unroll 1
        for (int g = 0; g < SITE_TILE / 16; g++) {
            uint32_t w[NPL + 1][4];
#pragma unroll
            for (int s = 0; s <= NPL; s++)
#pragma unroll
                for (int k = 0; k < 4; k++) w[s][k] = 0u;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int ii = g * 16 + j;
                unsigned long long q = 0ull;   // (omega / 8) 2^48
                if (i0 + ii < i1) {
                    if (NPL == gtc::SA) {
                        const double x = (double)arg_s[ii][tid] * 35184372088832.0;   // 2^45
                        q = x >= 281474976710655.0 ? 281474976710655ull : (x > 0.0 ? (unsigned long long)x : 0ull);
                    } else {
                        const double x = rint((double)arg_s[ii][tid] * 536870912.0);   // 2^29, round to nearest
                        q = (x >= 4294967295.0 ? 4294967295ull : (x > 0.0 ? (unsigned long long)x : 0ull)) << 16;
                    }
                    if (v.zprob) v.omega[(size_t)(i0 + ii) * Cp + c] = (double)q * (1.0 / 35184372088832.0);
                }
                const int sh = 8 * (j & 3);
#pragma unroll
                for (int s = 0; s < NPL; s++) w[s][j >> 2] |= (uint32_t)((q >> (8 * (gtc::SA - 1 - s))) & 0xFFull) << sh;
                w[NPL][j >> 2] |= (uint32_t)((zbits >> ii) & 1ull) << sh;
            }
#pragma unroll
            for (int s = 0; s < NPL; s++)
                *reinterpret_cast<uint4*>(row + (size_t)s * plane + g * 16) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
            *reinterpret_cast<uint4*>(row + (size_t)gtc::Z_SLICE * plane + g * 16) =
                make_uint4(w[NPL][0], w[NPL][1], w[NPL][2], w[NPL][3]);
        }
    }
#pragma unroll
    for (int k = 0; k < NV; k++) v.apart[((size_t)tile * NV + k) * Cp + c] = acc[k];
}

// Variant used by occu_debug_gaussian_half when omega_a is injected: alpha partial sums only.
template <int R>
__global__ void __launch_bounds__(CHAIN_TILE) alpha_partials_injected_kernel(const View v, const double* __restrict__ omega_a_cm) {
    const bool probit = v.model == OCCU_PROBIT_RSR;
    constexpr int NA = R * (R + 1) / 2, NV = NA + R;
    const int c = blockIdx.y * CHAIN_TILE + threadIdx.x;
    const int tile = blockIdx.x;
    const int i0 = tile * SITE_TILE;
    const int i1 = min(i0 + SITE_TILE, v.n);
    const size_t Cp = v.Cp;
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) acc[k] = 0.0;
    const double* oma = omega_a_cm + (size_t)(c % v.C) * v.V;
    for (int i = i0; i < i1; i++) {
        const int cls = v.site_class[i];
        if (cls == OCCU_SITE_NOT_SURVEYED || !v.z[(size_t)i * Cp + c]) continue;
        for (int vv = v.visit_ptr[i]; vv < v.visit_ptr[i + 1]; vv++) {
            double w[R];
#pragma unroll
            for (int k = 0; k < R; k++) w[k] = v.W[(size_t)vv * R + k];
            const double om = probit ? 1.0 : oma[vv];                               // probit.py:233 vs logit.py:222
            const double yk = probit ? oma[vv] : (double)v.y[vv] - 0.5;             // probit.py:234 vs logit.py:223
            int idx = 0;
#pragma unroll
            for (int a = 0; a < R; a++) {
#pragma unroll
                for (int b = 0; b <= a; b++) acc[idx++] += om * w[a] * w[b];
            }
#pragma unroll
            for (int a = 0; a < R; a++) acc[NA + a] += w[a] * yk;
        }
    }
#pragma unroll
    for (int k = 0; k < NV; k++) v.apart[((size_t)tile * NV + k) * Cp + c] = acc[k];
}

// G0[m] = sum_i Z[i][m]: the joint Gram of [X K] with unit weights (probit: K'K, K'X, X'X; probit.py:112, 144)
__global__ void gram_unit_kernel(View v) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= v.M) return;
    double s = 0.0;
    for (int i = 0; i < v.n; i++) s += v.ZU[(size_t)i * v.ldz + m];
    v.G0[m] = s;
}

// injected randomness for the parity tests (all chain-major, NULL = Philox)
struct SiteDebug {
    const double* u_site;   // [C][n]  uniforms of the omega_b truncated normals (probit.py:203-209)
    const double* u_visit;  // [C][V]  uniforms of the omega_a truncated normals (probit.py:191-194)
    const double* nrm_eps;  // [C][n]  normals of the eps update (probit.py:220)
    int keep_z;             // 1: do not redraw z (the reference draws omega_b from the CURRENT z)
};

// ------------------------------------------------------------------------------------------------
// site pass (probit): z | ., omega_a (truncated normal) + alpha partial sums, omega_b, eps, d = omega_b - eps
// No rejection loops (inverse-CDF draws), so the lanes stay in lock-step over the sites of the tile.
//
// Production precision (F64 = false, nothing injected): the latent draws and the occupancy probability run in FP32
// (erfc_pos_fast / ndtri_fast; KS tests arbitrate, tests/test_gpu_primitives.py), ONE Philox4x32-7 block serves a site
// (words: z | omega_b | two for the eps normal) and one block serves four visits of a site; psi = [X K] theta is formed
// here on the FP64 tensor pipe (FUSE_PSI, as in the logit pass) and only eps, d and z cross HBM (25 B per site and
// chain; omega_b = d + eps is materialised on request).  Debug / validation engines keep FP64 everywhere, one
// Philox4x32-10 stream per draw, and also store omega_b and the occupancy probability.
// ------------------------------------------------------------------------------------------------
template <int R, bool F64, bool FUSE_PSI>
// seven CTAs per SM for r <= 3 (72 registers: 1.11 ms at 4096 chains against 1.15 with six at 80 and 1.12 with eight at 64 and
// spills); the alpha partial sums alone are r (r + 3) / 2 doubles: fewer CTAs for larger r
__global__ void __launch_bounds__(CHAIN_TILE, F64 ? 1 : (R <= 3 ? 7 : (R <= 5 ? 4 : 3))) site_pass_probit_kernel(const View v, const SiteDebug dbg) {
    constexpr int NA = R * (R + 1) / 2, NV = NA + R;
    __shared__ __align__(16) float psi_s[FUSE_PSI ? SITE_TILE : 1][CHAIN_TILE];   // production engines only (FP32 psi)
    const int tid = threadIdx.x;
    const int c = blockIdx.y * CHAIN_TILE + tid;
    const int tile = blockIdx.x;
    const int i0 = tile * SITE_TILE;
    const int i1 = min(i0 + SITE_TILE, v.n);
    const uint32_t iter = v.ctl->iter;
    const bool keep_z = v.ctl->first != 0 || dbg.keep_z;
    const uint32_t gc = (uint32_t)(v.chain_offset + c);
    const size_t Cp = v.Cp;
    const int cr = c % v.C;
    // FP64 draws / probabilities: validation build, debug engines (they dump the probabilities) and injected randomness
    const bool slow = F64 || v.zprob != nullptr || dbg.u_site != nullptr || dbg.u_visit != nullptr || dbg.nrm_eps != nullptr;

    double al[R];
#pragma unroll
    for (int k = 0; k < R; k++) al[k] = v.alpha[(size_t)k * Cp + c];
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) acc[k] = 0.0;

    if (FUSE_PSI) psi_tile_dmma(v, i0, tid, c, psi_s);
    // eps of the next site is requested one site ahead: it comes from HBM (the array does not fit L2) and is needed at the
    // very top of a site's work, where its latency would be exposed once per site
    double ep_next = i0 < i1 ? v.eps[(size_t)i0 * Cp + c] : 0.0;
    for (int i = i0; i < i1; i++) {
        const int cls = v.site_class[i];
        const size_t off = (size_t)i * Cp + c;
        const double lp = FUSE_PSI ? (double)psi_s[i - i0][tid] : v.psi[off];   // x_i beta + K_i eta
        const double ep = ep_next;
        if (i + 1 < i1) ep_next = v.eps[off + Cp];
        const double loc = lp + ep;
        const int v0 = v.visit_ptr[i], v1 = v.visit_ptr[i + 1];
        int zc;
        double omb, nz;
        if (slow) {
            // ---- FP64 path: one Philox stream per draw ----
            if (keep_z) {
                zc = v.z[off];
            } else {
                if (cls == OCCU_SITE_OBS) {
                    zc = 1;
                } else {
                    double pr = ndtr_d(loc);  // probit.py:251 / 258
                    if (cls == OCCU_SITE_NOT_OBS) {
                        double prod = 1.0;  // probit.py:252-253: prod_j (1 - Phi(w_ij alpha))
                        for (int vv = v0; vv < v1; vv++) {
                            double lin = 0.0;
#pragma unroll
                            for (int k = 0; k < R; k++) lin += v.W[(size_t)vv * R + k] * al[k];
                            prod *= 1.0 - ndtr_d(lin);
                        }
                        const double num = pr * prod;
                        pr = num / ((1.0 - pr) + num);  // probit.py:254-255
                    }
                    if (v.zprob) v.zprob[off] = pr;
                    Philox g;
                    g.init(v.seed, gc, (uint32_t)i, iter, ST_Z);
                    zc = g.unif_d() < pr ? 1 : 0;
                }
                v.z[off] = (uint8_t)zc;
            }
            if (cls != OCCU_SITE_NOT_SURVEYED && zc) {
                // omega_a ~ TN(w_ij alpha, 1) on (0, inf) if y_ij = 1 else (-inf, 0)   (probit.py:180-194)
                for (int vv = v0; vv < v1; vv++) {
                    double w[R];
                    double lin = 0.0;
#pragma unroll
                    for (int k = 0; k < R; k++) {
                        w[k] = v.W[(size_t)vv * R + k];
                        lin += w[k] * al[k];
                    }
                    const bool positive = v.y[vv] == 1;
                    double oma;
                    if (dbg.u_visit) {
                        oma = truncnorm_icdf(lin, positive, dbg.u_visit[(size_t)cr * v.V + vv]);
                    } else {
                        Philox g;
                        g.init(v.seed, gc, (uint32_t)vv, iter, ST_OMEGA_A);
                        oma = F64 ? truncnorm_draw(lin, positive, g) : truncnorm_draw_f32(lin, positive, g);
                    }
                    int idx = 0;
#pragma unroll
                    for (int a = 0; a < R; a++) {
#pragma unroll
                        for (int b = 0; b <= a; b++) acc[idx++] += w[a] * w[b];  // probit.py:233
                    }
#pragma unroll
                    for (int a = 0; a < R; a++) acc[NA + a] += w[a] * oma;  // probit.py:234
                }
            }
            // omega_b ~ TN(lp + eps, 1), sign by z (probit.py:196-209); eps | . (probit.py:216-221)
            Philox g;
            g.init(v.seed, gc, (uint32_t)i, iter, ST_OMEGA_B);
            omb = dbg.u_site ? truncnorm_icdf(loc, zc == 1, dbg.u_site[(size_t)cr * v.n + i])
                             : (F64 ? truncnorm_draw(loc, zc == 1, g) : truncnorm_draw_f32(loc, zc == 1, g));
            nz = dbg.nrm_eps ? dbg.nrm_eps[(size_t)cr * v.n + i] : (F64 ? g.normal_d() : normal_f32(g));
        } else {
            // ---- production path: FP32 latent draws, one Philox block per site and per four visits ----
            Philox g;
            g.init(v.seed, gc, (uint32_t)i, iter, ST_OMEGA_B);
            g.refill_rk<PHILOX_FAST_ROUNDS>(v.rk);
            g.have = 0;
            const uint32_t wz = g.buf[0], wb = g.buf[1], wn1 = g.buf[2], wn2 = g.buf[3];
            if (keep_z) {
                zc = v.z[off];
            } else {
                if (cls == OCCU_SITE_OBS) {
                    zc = 1;
                } else {
                    // z = 1 with probability num / (den + num), num = Phi(loc) prod_j Phi(-w_ij alpha), den = Phi(-loc)
                    float num, den;
                    phi_pair_fast((float)loc, num, den);
                    if (cls == OCCU_SITE_NOT_OBS) {
                        for (int vv = v0; vv < v1; vv++) {
                            double lin = 0.0;
#pragma unroll
                            for (int k = 0; k < R; k++) lin += v.W[(size_t)vv * R + k] * al[k];
                            float ph, phn;
                            phi_pair_fast((float)lin, ph, phn);
                            num *= phn;
                        }
                    }
                    const float u = fmaf(__uint2float_rz(wz), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
                    zc = u * (den + num) < num ? 1 : 0;   // probit.py:254-260
                }
                v.z[off] = (uint8_t)zc;
            }
            if (cls != OCCU_SITE_NOT_SURVEYED && zc) {
                Philox ga;
                uint32_t wd0 = 0u, wd1 = 0u, wd2 = 0u, wd3 = 0u;
#pragma unroll 1
                for (int vv = v0; vv < v1; vv++) {
                    const int j = (vv - v0) & 3;
                    if (j == 0) {   // one Philox block per four visits of the site
                        ga.init(v.seed, gc, (uint32_t)vv, iter, ST_OMEGA_A);
                        ga.refill_rk<PHILOX_FAST_ROUNDS>(v.rk);
                        ga.have = 0;
                        wd0 = ga.buf[0]; wd1 = ga.buf[1]; wd2 = ga.buf[2]; wd3 = ga.buf[3];
                    }
                    const uint32_t word = j == 0 ? wd0 : (j == 1 ? wd1 : (j == 2 ? wd2 : wd3));
                    double w[R];
                    double lin = 0.0;
#pragma unroll
                    for (int k = 0; k < R; k++) {
                        w[k] = v.W[(size_t)vv * R + k];
                        lin += w[k] * al[k];
                    }
                    const double oma = truncnorm_word_f32(lin, v.y[vv] == 1, word, ga);
                    int idx = 0;
#pragma unroll
                    for (int a = 0; a < R; a++) {
#pragma unroll
                        for (int b = 0; b <= a; b++) acc[idx++] += w[a] * w[b];  // probit.py:233
                    }
#pragma unroll
                    for (int a = 0; a < R; a++) acc[NA + a] += w[a] * oma;  // probit.py:234
                }
            }
            omb = truncnorm_word_f32(loc, zc == 1, wb, g);   // probit.py:196-209
            nz = normal_words_f32(wn1, wn2);
        }
        const double ep_new = 0.5 * (omb - lp) + nz * 0.70710678118654752440;   // probit.py:216-221
        if (v.zprob) v.omega[off] = omb;   // debug engines keep omega_b readable as drawn; else omega_b = d + eps on request
        v.eps[off] = ep_new;
        v.kd[off] = omb - ep_new;  // d: right-hand-side vector of both the eta and the beta systems
    }
#pragma unroll
    for (int k = 0; k < NV; k++) v.apart[((size_t)tile * NV + k) * Cp + c] = acc[k];
}
