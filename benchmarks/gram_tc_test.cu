// gram_tc_test.cu -- standalone check + timing of the tcgen05 Gram kernel (occuspytial_b200/csrc/gram_tc.cuh).
//
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -o gram_tc_test gram_tc_test.cu
//   ./gram_tc_test [n] [chains] [pq] [BN]
//
// 1. exactness: for sampled chains, every output column is recomputed from the SAME digits with INT64
//    accumulation and the epilogue's FP64 recombination -> must agree bit for bit;
// 2. accuracy: against the plain FP64 dot product of the un-sliced inputs;
// 3. time: CUDA events around the Gram kernel alone and around slice + Gram + reduce.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../occuspytial_b200/csrc/gram_tc.cuh"

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
            return 2;                                                                           \
        }                                                                                       \
    } while (0)

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ double u01(uint64_t x) { return ((mix(x) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

__global__ void fill_zu(double* ZU, int n, int ld, int M, int src_u0, int pq) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * ld) return;
    const int col = (int)(idx % ld);
    double v = 0.0;
    if (col < M || (col >= src_u0 && col < src_u0 + pq)) {
        const double colscale = exp(6.0 * (u01(0x1234567ull + col) - 0.5));  // column magnitudes over e^+-3
        v = colscale * (4.0 * u01(idx * 2 + 1) - 2.0) * (u01(idx * 2 + 2) < 0.1 ? 0.0 : 1.0);
    }
    ZU[idx] = v;
}
__global__ void fill_omega(double* om, uint8_t* z, size_t count, int planes) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    const double u = u01(idx + 0xABCDEFull);
    double w = 0.25 * exp(3.0 * (u01(idx * 3 + 7) - 0.7));  // ~ PG(1, c) magnitudes, 0.03 .. 0.6
    if (u < 1e-4) w = 7.9999;                               // near the saturation bound
    if (u > 1.0 - 1e-4) w = 1e-9;
    if (planes == gtc::SA_F32) w = rint(w * 536870912.0) * (1.0 / 536870912.0);   // 32-bit fixed point of omega / 8
    om[idx] = w;
    z[idx] = u01(idx * 5 + 11) < 0.4 ? 1 : 0;
}

// One thread per (sampled chain, output row of the sliced Z): INT64 re-evaluation from the digits.
__global__ void ref_kernel(const uint8_t* Aq, const int8_t* Bq, gtc::Params p, int Kp, const int* chains, int nsamp,
                           double* out /* [nsamp][Mrows] */) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const int si = blockIdx.y;
    if (row >= p.Mrows || si >= nsamp) return;
    const int c = chains[si];
    const bool is_h = row >= p.n_gt * (p.Mrows / (p.n_gt + p.n_ht));
    double total = 0.0;
    for (int chunk = 0; chunk < p.nsplit; chunk++) {
        long long lev[gtc::NLEV];
        for (int l = 0; l < gtc::NLEV; l++) lev[l] = 0;
        const int i0 = chunk * p.kb_per_chunk * gtc::BK;
        const int i1 = min((chunk + 1) * p.kb_per_chunk, p.kblocks) * gtc::BK;
        for (int i = i0; i < i1; i++) {
            int b[gtc::SB];
            for (int t = 0; t < gtc::SB; t++) {
                const int8_t raw = Bq[((size_t)t * p.Mrows + row) * Kp + i];
                b[t] = (int)raw;
            }
            if (is_h) {
                const int a = Aq[((size_t)gtc::Z_SLICE * p.Cp + c) * Kp + i];
                for (int t = 0; t < gtc::SB; t++) lev[t] += (long long)a * b[t];
            } else {
                for (int s = 1; s <= gtc::SA; s++) {
                    const int a = Aq[((size_t)(s - 1) * p.Cp + c) * Kp + i];
                    for (int t = 1; t <= gtc::SB; t++)
                        if (s + t <= gtc::LMAX) lev[s + t - 2] += (long long)a * b[t - 1];
                }
            }
        }
        const int nlev = is_h ? gtc::SB : gtc::NLEV, w0 = is_h ? 1 : 2;
        double acc = 0.0;
        for (int l = nlev - 1; l >= 0; l--) acc = fma((double)lev[l], ldexp(1.0, -8 * (l + w0)), acc);
        total += acc * p.scale[row];
    }
    out[(size_t)si * p.Mrows + row] = total;
}

// plain FP64 dot products of the un-sliced inputs, plus sum |omega Z| for normalisation
__global__ void dot_kernel(const double* omega, const uint8_t* z, const double* ZU, int n, int Cp, int ld, int M, int src_u0,
                           int pq, const int* chains, int nsamp, double* out, double* outabs, int ncols_out) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const int si = blockIdx.y;
    if (col >= ncols_out || si >= nsamp) return;
    const int c = chains[si];
    const bool is_h = col >= M;
    const int src = is_h ? src_u0 + (col - M) : col;
    double acc = 0.0, aabs = 0.0;
    for (int i = 0; i < n; i++) {
        const double w = is_h ? (double)z[(size_t)i * Cp + c] : omega[(size_t)i * Cp + c];
        const double t = w * ZU[(size_t)i * ld + src];
        acc += t;
        aabs += fabs(t);
    }
    out[(size_t)si * ncols_out + col] = acc;
    outabs[(size_t)si * ncols_out + col] = aabs;
}

template <int BN>
int run(int n, int C, int pq, int want_split) {
    const int M = pq * (pq + 1) / 2;
    const int Cp = (C + 127) / 128 * 128;
    const int M64 = (M + 63) / 64 * 64, pq64 = (pq + 63) / 64 * 64;
    const int ld = M64 + pq64, src_u0 = M64;
    const gtc::Geometry g = gtc::make_geometry(n, Cp, M, pq, BN, want_split);
    printf("n=%d chains=%d (Cp=%d) pq=%d M=%d BN=%d  Kp=%d kblocks=%d nsplit=%d kb/chunk=%d  tiles=%d x (%d+%d)  NPAIR=%d smem=%d\n", n,
           C, Cp, pq, M, BN, g.Kp, g.kblocks, g.nsplit, g.kb_per_chunk, g.MT, g.n_gt, g.n_ht, gtc::NPAIR, gtc::Cfg<BN>::SMEM_BYTES);

    double *ZU, *omega, *cscale, *scale, *Gpart, *Hpart, *G, *H;
    uint8_t *z, *Aq;
    int8_t* Bq;
    int* err;
    CK(cudaMalloc(&ZU, (size_t)n * ld * 8));
    CK(cudaMalloc(&omega, (size_t)n * Cp * 8));
    CK(cudaMalloc(&z, (size_t)n * Cp));
    CK(cudaMalloc(&cscale, (size_t)ld * 8));
    CK(cudaMalloc(&scale, (size_t)g.Mrows * 8));
    CK(cudaMalloc(&Aq, (size_t)(gtc::SA + 1) * Cp * g.Kp));
    CK(cudaMalloc(&Bq, (size_t)gtc::SB * g.Mrows * g.Kp));
    const size_t gstride = (size_t)Cp * M64, hstride = (size_t)Cp * pq64;
    CK(cudaMalloc(&Gpart, g.nsplit * gstride * 8));
    CK(cudaMalloc(&Hpart, g.nsplit * hstride * 8));
    CK(cudaMalloc(&G, gstride * 8));
    CK(cudaMalloc(&H, hstride * 8));
    CK(cudaMalloc(&err, 4));
    CK(cudaMemset(err, 0, 4));
    CK(cudaMemset(Aq, 0, (size_t)(gtc::SA + 1) * Cp * g.Kp));
    CK(cudaMemset(Gpart, 0xFF, g.nsplit * gstride * 8));

    fill_zu<<<(unsigned)(((size_t)n * ld + 255) / 256), 256>>>(ZU, n, ld, M, src_u0, pq);
    const int planes = getenv("SAV") ? atoi(getenv("SAV")) : gtc::SA;   // 4: omega on the 32-bit grid, kernel skips planes 5, 6
    fill_omega<<<(unsigned)(((size_t)n * Cp + 255) / 256), 256>>>(omega, z, (size_t)n * Cp, planes);
    gtc::column_scale_kernel<<<(ld + 127) / 128, 128>>>(ZU, n, ld, ld, cscale);
    gtc::slice_b_kernel<<<dim3((g.Kp + 255) / 256, g.Mrows), 256>>>(ZU, n, g.Kp, ld, M, g.Mg_pad, pq, src_u0, g.Mrows, cscale, scale,
                                                                    Bq);
    CK(cudaDeviceSynchronize());

    CUtensorMap tmA, tmB;
    const char* e1 = gtc::make_map(&tmA, Aq, (uint64_t)(gtc::SA + 1) * Cp, g.Kp, gtc::BM);
    const char* e2 = gtc::make_map(&tmB, Bq, (uint64_t)gtc::SB * g.Mrows, g.Kp, BN);
    if (e1 || e2) {
        printf("tensor map: %s\n", e1 ? e1 : e2);
        return 2;
    }
    gtc::Params p;
    p.Cp = Cp; p.kblocks = g.kblocks; p.nsplit = g.nsplit; p.kb_per_chunk = g.kb_per_chunk; p.MT = g.MT;
    p.n_gt = g.n_gt; p.n_ht = g.n_ht; p.Mrows = g.Mrows; p.M = M; p.pq = pq; p.ldg = M64; p.ldh = pq64;
    p.scale = scale; p.Gpart = Gpart; p.Hpart = Hpart; p.err = err; p.mode = 0; p.hshift = nullptr;
    p.order = getenv("TC_ORDER") ? atoi(getenv("TC_ORDER")) : 0;

    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaFuncSetAttribute(gtc::gram_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, gtc::Cfg<BN>::SMEM_BYTES));
    CK(cudaFuncSetAttribute(gtc::gram_tc_kernel<BN, gtc::SA_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, gtc::Cfg<BN, gtc::SA_F32>::SMEM_BYTES));
    const int nitems = g.MT * (g.n_gt + g.n_ht) * g.nsplit;
    const int grid = nitems < sms ? nitems : sms;

    auto step = [&]() {
        gtc::slice_a_kernel<<<dim3((g.Kp + 31) / 32, Cp / 32), 256>>>(omega, z, n, g.Kp, Cp, Aq);
        if (planes == gtc::SA_F32) gtc::gram_tc_kernel<BN, gtc::SA_F32><<<grid, gtc::THREADS, gtc::Cfg<BN, gtc::SA_F32>::SMEM_BYTES>>>(tmA, tmB, p);
        else gtc::gram_tc_kernel<BN><<<grid, gtc::THREADS, gtc::Cfg<BN>::SMEM_BYTES>>>(tmA, tmB, p);
        gtc::reduce_parts_kernel<<<(unsigned)((gstride + 255) / 256), 256>>>(Gpart, g.nsplit, gstride, gstride, M64, M, G);
        gtc::reduce_parts_kernel<<<(unsigned)((hstride + 255) / 256), 256>>>(Hpart, g.nsplit, hstride, hstride, pq64, pq, H);
    };
    step();
    cudaError_t se = cudaDeviceSynchronize();
    int herr = -1;
    cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost);
    if (se != cudaSuccess) {
        printf("kernel failed: %s (barrier-timeout flag %d)\n", cudaGetErrorString(se), herr);
        return 3;
    }

    // ---- checks on sampled chains: 2 per chain tile, one in each half of the TMEM lanes ----
    std::vector<int> hch;
    for (int mt = 0; mt < g.MT; mt++) {
        hch.push_back(mt * 128 + (mt * 37) % 64);
        hch.push_back(mt * 128 + 64 + (mt * 53) % 64);
    }
    const int nsamp = (int)hch.size();
    int* dch;
    CK(cudaMalloc(&dch, nsamp * 4));
    CK(cudaMemcpy(dch, hch.data(), nsamp * 4, cudaMemcpyHostToDevice));
    double *dref, *ddot, *dabs;
    const int ncols_out = M + pq;
    CK(cudaMalloc(&dref, (size_t)nsamp * g.Mrows * 8));
    CK(cudaMalloc(&ddot, (size_t)nsamp * ncols_out * 8));
    CK(cudaMalloc(&dabs, (size_t)nsamp * ncols_out * 8));
    ref_kernel<<<dim3((g.Mrows + 63) / 64, nsamp), 64>>>(Aq, Bq, p, g.Kp, dch, nsamp, dref);
    dot_kernel<<<dim3((ncols_out + 63) / 64, nsamp), 64>>>(omega, z, ZU, n, Cp, ld, M, src_u0, pq, dch, nsamp, ddot, dabs, ncols_out);
    CK(cudaDeviceSynchronize());
    std::vector<double> href((size_t)nsamp * g.Mrows), hdot((size_t)nsamp * ncols_out), habs((size_t)nsamp * ncols_out);
    std::vector<double> hG(gstride), hH(hstride);
    CK(cudaMemcpy(href.data(), dref, href.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hdot.data(), ddot, hdot.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(habs.data(), dabs, habs.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hG.data(), G, gstride * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hH.data(), H, hstride * 8, cudaMemcpyDeviceToHost));
    long long mism = 0, checked = 0;
    double max_rel = 0.0;
    int shown = 0;
    for (int si = 0; si < nsamp; si++) {
        const int c = hch[si];
        for (int col = 0; col < ncols_out; col++) {
            const bool is_h = col >= M;
            const int row = is_h ? g.Mg_pad + (col - M) : col;
            const double got = is_h ? hH[(size_t)c * pq64 + (col - M)] : hG[(size_t)c * M64 + col];
            const double want = href[(size_t)si * g.Mrows + row];
            checked++;
            if (!(got == want)) {
                mism++;
                if (shown++ < 12) printf("  MISMATCH chain %d col %d (%s): got %.17g want %.17g\n", c, col, is_h ? "H" : "G", got, want);
            }
            const double rel = fabs(got - hdot[(size_t)si * ncols_out + col]) / (habs[(size_t)si * ncols_out + col] + 1e-300);
            if (rel > max_rel) max_rel = rel;
        }
    }
    printf("exactness: %lld of %lld sampled outputs differ from the INT64 digit evaluation\n", mism, checked);
    printf("accuracy : max |G - fp64 dot| / sum|omega Z| = %.3e\n", max_rel);

    // ---- timing ----
    cudaEvent_t e0, e1e;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1e);
    const int reps = 20;
    for (int w = 0; w < 3; w++) step();
    cudaEventRecord(e0);
    for (int r = 0; r < reps; r++) {
        if (planes == gtc::SA_F32) gtc::gram_tc_kernel<BN, gtc::SA_F32><<<grid, gtc::THREADS, gtc::Cfg<BN, gtc::SA_F32>::SMEM_BYTES>>>(tmA, tmB, p);
        else gtc::gram_tc_kernel<BN><<<grid, gtc::THREADS, gtc::Cfg<BN>::SMEM_BYTES>>>(tmA, tmB, p);
    }
    cudaEventRecord(e1e);
    CK(cudaDeviceSynchronize());
    float ms_k = 0, ms_s = 0;
    cudaEventElapsedTime(&ms_k, e0, e1e);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; r++) step();
    cudaEventRecord(e1e);
    CK(cudaDeviceSynchronize());
    cudaEventElapsedTime(&ms_s, e0, e1e);
    int pairs = 0;   // slice pairs issued: s <= planes, t <= SB, s + t <= LMAX
    for (int s_ = 1; s_ <= planes; s_++)
        for (int t_ = 1; t_ <= gtc::SB; t_++) pairs += (s_ + t_ <= gtc::LMAX);
    const double macs = (double)Cp * g.Kp * ((double)g.n_gt * BN * pairs + (double)g.n_ht * BN * gtc::SB);
    printf("time     : gram kernel %.3f ms (%.2f int8 POP/s issued), slice + gram + reduce %.3f ms; grid %d, items %d\n", ms_k / reps,
           2.0 * macs / (ms_k / reps * 1e-3) / 1e15, ms_s / reps, grid, nitems);
    for (int mode = 1; mode <= 2; mode++) {  // bottleneck probes
        gtc::Params pm = p;
        pm.mode = mode;
        cudaEventRecord(e0);
        for (int r = 0; r < reps; r++) gtc::gram_tc_kernel<BN><<<grid, gtc::THREADS, gtc::Cfg<BN>::SMEM_BYTES>>>(tmA, tmB, pm);
        cudaEventRecord(e1e);
        CK(cudaDeviceSynchronize());
        float ms_m = 0;
        cudaEventElapsedTime(&ms_m, e0, e1e);
        printf("probe    : %s %.3f ms\n", mode == 1 ? "MMA + epilogue only (no TMA loads)" : "TMA loads + epilogue only (no MMA) ", ms_m / reps);
    }
    printf("RESULT %s\n", mism == 0 && max_rel < 1e-10 ? "PASS" : "FAIL");
    return mism == 0 && max_rel < 1e-10 ? 0 : 1;
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 10000;
    const int C = argc > 2 ? atoi(argv[2]) : 4096;
    const int pq = argc > 3 ? atoi(argv[3]) : 67;
    const int BN = argc > 4 ? atoi(argv[4]) : 80;
    const int ws = argc > 5 ? atoi(argv[5]) : 1;
    if (BN == 64) return run<64>(n, C, pq, ws);
#if GTC_SB == 5
    if (BN == 80) return run<80>(n, C, pq, ws);
#endif
    printf("BN must be 64 (or 80 when built with -DGTC_SB=5)\n");
    return 2;
}
