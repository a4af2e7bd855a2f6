// solve_icar.cuh -- LogitICARGibbs ("PGGibbs"): psi for the full ICAR model (the solve is in solve_icar_warp.cuh).
// Part of the single translation unit engine.cu (included there, in order; not a stand-alone header).
#pragma once

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
