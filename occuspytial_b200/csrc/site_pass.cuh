// site_pass.cuh -- the per-site half of the sweep: z, omega_a, omega_b (logit: PG draws, psi product; probit: truncated normals).
// Part of the single translation unit engine.cu (included there, in order; not a stand-alone header).
#pragma once

// ------------------------------------------------------------------------------------------------
// site pass (logit): z-update, omega_a draws + alpha partial sums, omega_b draw
//
// CTA = 128 chains x SITE_TILE sites, thread = chain.
//  phase 1 (lock-step over the tile's sites, coalesced): z | psi, alpha in FP64; z, kd written; psi kept
//          in shared memory as the PG argument; z bits kept in a register.
//  phase 2 (lanes desynchronise): every lane walks its OWN stream of PG draws -- for each site the visits
//          of an occupied surveyed site (omega_a) and then the site itself (omega_b) -- running one
//          rejection round per loop trip (pg1_round).  An accepted lane commits and immediately fetches
//          its next draw, a rejected lane retries: no lane waits for another lane's rejection loop.
//  phase 3: omega_b (staged in the lane's own shared-memory column) is written out coalesced.
// Philox streams are keyed by (chain, visit row | site, iteration, stage), so the result does not
// depend on how lanes interleave.
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
constexpr int PSI_SUB = 8;   // sites per DMMA sub-tile
template <int R, typename PGT, bool SLICES, bool FUSE_PSI>
__global__ void __launch_bounds__(CHAIN_TILE, site_min_blocks(R)) site_pass_logit_kernel(const View v) {
    constexpr int NA = R * (R + 1) / 2, NV = NA + R;
    extern __shared__ __align__(16) unsigned char site_smem[];
    PGT (*arg_s)[CHAIN_TILE] = reinterpret_cast<PGT(*)[CHAIN_TILE]>(site_smem);  // column tid is private to thread tid
    // FUSE_PSI: FP64 psi of the current sub-tile, [PSI_SUB][CHAIN_TILE]; a warp only touches its own 32 columns
    double (*psi_d)[CHAIN_TILE] = reinterpret_cast<double(*)[CHAIN_TILE]>(site_smem + sizeof(PGT) * SITE_TILE * CHAIN_TILE);
    const int tid = threadIdx.x;
    const uint32_t iter = v.ctl->iter;
    const bool first = v.ctl->first != 0;
    const size_t Cp = v.Cp;
    const int tile = blockIdx.x;
    const int c = blockIdx.y * CHAIN_TILE + tid;
    const int i0 = tile * SITE_TILE;
    const int i1 = min(i0 + SITE_TILE, v.n);
    const uint32_t gc = (uint32_t)(v.chain_offset + c);

    double al[R];
#pragma unroll
    for (int k = 0; k < R; k++) al[k] = v.alpha[(size_t)k * Cp + c];

    // ---- phase 1: z | . ----
    uint64_t zbits = 0ull;
    static_assert(SITE_TILE % PSI_SUB == 0, "whole DMMA sub-tiles");
    for (int ib = i0; ib < i1; ib += PSI_SUB) {
    if (FUSE_PSI) {
        // psi[ib .. ib + 8)[32 chains of this warp] = U theta:  D (8 sites x 8 chains) += A (8 x 4: UT'[site][k]) B (4 x 8:
        // theta[k][chain]); rows of UT and theta beyond pq are zero up to kq_pad
        const int lane = tid & 31, g = lane >> 2, t4 = lane & 3;
        const int cw = c - lane;   // first chain of this warp
        double acc[4][2];
#pragma unroll
        for (int nb = 0; nb < 4; nb++) acc[nb][0] = acc[nb][1] = 0.0;
        const double* ua = v.UT + (size_t)t4 * v.n_pad + ib + g;       // n_pad >= i0 + SITE_TILE: no bounds check
        const double* tb = v.theta + (size_t)t4 * Cp + cw + g;
        const int ksteps = (v.pq + 3) >> 2;
#pragma unroll 4
        for (int ks = 0; ks < ksteps; ks++) {
            const double a = ua[(size_t)ks * 4 * v.n_pad];
            double b[4];
#pragma unroll
            for (int nb = 0; nb < 4; nb++) b[nb] = tb[(size_t)ks * 4 * Cp + nb * 8];
#pragma unroll
            for (int nb = 0; nb < 4; nb++) dgemm::dmma884(acc[nb][0], acc[nb][1], a, b[nb]);
        }
        __syncwarp();   // the previous sub-tile's readers are done
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
            psi_d[g][(tid & ~31) + nb * 8 + 2 * t4] = acc[nb][0];
            psi_d[g][(tid & ~31) + nb * 8 + 2 * t4 + 1] = acc[nb][1];
        }
        __syncwarp();
    }
    for (int i = ib; i < min(ib + PSI_SUB, i1); i++) {
        const int cls = v.site_class[i];
        const size_t off = (size_t)i * Cp + c;
        const double psi = FUSE_PSI ? psi_d[i - ib][tid] : v.psi[off];
        if (FUSE_PSI && v.zprob) v.psi[off] = psi;   // debug engines keep psi readable
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
                double grow = exp(-psi);
                if (cls == OCCU_SITE_NOT_OBS) {
                    const int v1 = v.visit_ptr[i + 1];
                    for (int vv = v.visit_ptr[i]; vv < v1; vv++) {
                        double lin = 0.0;
#pragma unroll
                        for (int k = 0; k < R; k++) lin += v.W[(size_t)vv * R + k] * al[k];
                        grow *= 1.0 + exp(lin);
                    }
                }
                const double pr = 1.0 / (1.0 + grow);
                if (v.zprob) v.zprob[off] = pr;
                Philox g;
                g.init(v.seed, gc, (uint32_t)i, iter, ST_Z);
                g.refill_rk(v.rk);
                const uint64_t bits = ((uint64_t)g.buf[0] << 21) ^ (uint64_t)(g.buf[1] >> 11);  // == g.unif_d()
                zc = ((double)bits + 0.5) * (1.0 / 9007199254740992.0) < pr ? 1 : 0;  // logit.py:246 / 250
            }
            v.z[off] = (uint8_t)zc;
            if (!SLICES) v.kd[off] = (double)zc - 0.5;  // logit.py:252
        }
        zbits |= (uint64_t)zc << (i - i0);
        arg_s[i - i0][tid] = (PGT)psi;
    }
    }  // sub-tiles

    // ---- phase 2: per-lane stream of PG draws ----
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) acc[k] = 0.0;
    int i = i0 - 1, vv = 0, vend = 0, cur = 0;
    bool site_done = true, active = true, have = false, is_visit = false;
    typename PGSel<PGT>::Task task;
    Philox g;
    double wv[R];  // design row of the visit being drawn (kept from the fetch to the commit)
#pragma unroll
    for (int k = 0; k < R; k++) wv[k] = 0.0;
    while (__any_sync(0xffffffffu, active || have)) {
        if (!have && active) {
            // next draw of this lane's stream, loop-free: the visits of an occupied surveyed site
            // (omega_a ~ PG(1, w_ij alpha), logit.py:187-193), then the site itself (omega_b ~ PG(1, psi),
            // logit.py:201-204), then the next site
            if (vv >= vend && site_done) {
                if (++i >= i1) {
                    active = false;
                } else {
                    site_done = false;
                    if (v.site_class[i] != OCCU_SITE_NOT_SURVEYED && ((zbits >> (i - i0)) & 1ull)) {
                        vv = v.visit_ptr[i];
                        vend = v.visit_ptr[i + 1];
                    }
                }
            }
            if (active) {
                PGT carg;
                is_visit = vv < vend;
                if (is_visit) {
                    cur = vv++;
                    double lin = 0.0;
#pragma unroll
                    for (int k = 0; k < R; k++) {
                        wv[k] = v.W[(size_t)cur * R + k];
                        lin += wv[k] * al[k];
                    }
                    carg = (PGT)lin;
                } else {
                    site_done = true;
                    cur = i;
                    carg = arg_s[i - i0][tid];
                }
                g.init(v.seed, gc, (uint32_t)cur, iter, is_visit ? ST_OMEGA_A : ST_OMEGA_B);
                PGSel<PGT>::setup(carg, task);  // one shared copy of the setup code for both kinds of draw
                have = true;
            }
        }
        if (have) {
            PGT ub, ux, u3, us, x;
            RT<PGT>::unif4_rk(g, v.rk, ub, ux, u3, us);
            if (PGSel<PGT>::round(task, ub, ux, u3, us, x)) {
                have = false;
                const double om = 0.25 * (double)x;
                if (is_visit) {
                    const double yk = (double)v.y[cur] - 0.5;
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
                    arg_s[cur - i0][tid] = (PGT)om;
                }
            }
        }
    }

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
// ------------------------------------------------------------------------------------------------
template <int R, bool F64>
__global__ void __launch_bounds__(CHAIN_TILE) site_pass_probit_kernel(const View v, const SiteDebug dbg) {
    constexpr int NA = R * (R + 1) / 2, NV = NA + R;
    const int c = blockIdx.y * CHAIN_TILE + threadIdx.x;
    const int tile = blockIdx.x;
    const int i0 = tile * SITE_TILE;
    const int i1 = min(i0 + SITE_TILE, v.n);
    const uint32_t iter = v.ctl->iter;
    const bool keep_z = v.ctl->first != 0 || dbg.keep_z;
    const uint32_t gc = (uint32_t)(v.chain_offset + c);
    const size_t Cp = v.Cp;
    const int cr = c % v.C;

    double al[R];
#pragma unroll
    for (int k = 0; k < R; k++) al[k] = v.alpha[(size_t)k * Cp + c];
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) acc[k] = 0.0;

    for (int i = i0; i < i1; i++) {
        const int cls = v.site_class[i];
        const size_t off = (size_t)i * Cp + c;
        const double lp = v.psi[off];   // x_i beta + K_i eta
        const double ep = v.eps[off];
        const int v0 = v.visit_ptr[i], v1 = v.visit_ptr[i + 1];
        int zc;
        if (keep_z) {
            zc = v.z[off];
        } else {
            if (cls == OCCU_SITE_OBS) {
                zc = 1;
            } else {
                double pr = ndtr_d(lp + ep);  // probit.py:251 / 258
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
        const double loc = lp + ep;
        const double omb = dbg.u_site ? truncnorm_icdf(loc, zc == 1, dbg.u_site[(size_t)cr * v.n + i])
                                      : (F64 ? truncnorm_draw(loc, zc == 1, g) : truncnorm_draw_f32(loc, zc == 1, g));
        const double nz = dbg.nrm_eps ? dbg.nrm_eps[(size_t)cr * v.n + i] : (F64 ? g.normal_d() : normal_f32(g));
        const double ep_new = 0.5 * (omb - lp) + nz * 0.70710678118654752440;
        v.omega[off] = omb;
        v.eps[off] = ep_new;
        v.kd[off] = omb - ep_new;  // d: right-hand-side vector of both the eta and the beta systems
    }
#pragma unroll
    for (int k = 0; k < NV; k++) v.apart[((size_t)tile * NV + k) * Cp + c] = acc[k];
}
