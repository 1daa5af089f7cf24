// pm_copy_ops.cuh — per-mode arithmetic of pm_fourier_copy_modes (pm_copymodes.cu): moving a Fourier slab
// between grids of different size (reference copy_modes, mesh.py:980-1322).  __host__ __device__ like
// pm_ic_ops.cuh, so that tests/ic_host_harness.cu runs the same code on the CPU.
#pragma once

#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>

#ifndef PM_HD
#ifdef __CUDACC__
#define PM_HD __host__ __device__ __forceinline__
#else
#define PM_HD inline
#endif
#endif

namespace pm {
namespace copyops {

struct CopyParams {
    int Gs, Gd;          // source / destination grid size (one rank each: slabs [G][G][G/2+1])
    int deconv_order;    // power of Π x_l/sin x_l evaluated for the SOURCE grid (fourier_loop(gridsize_small, gridsize_from, …))
    int rotate;          // lattice shift != 0
    double th[3];        // −2π/Gs·shift[d]  (θ of fourier_loop for gridsize_corrections = Gs, mesh.py:2873-2888)
    double cell_phase;   // π/Gd − π/Gs: half-cell offset between two cell-centred grids (mesh.py:1302)
    double scale;        // 1/n_lattices
};

PM_HD double ipow(double f, int n) {
    double r = 1.0;
    while (n > 0) {
        if (n & 1) r *= f;
        f *= f;
        n >>= 1;
    }
    return r;
}

// The value a shared mode (ki, kj, kk) of the source grid — v at source slab indices (is, js, kk) — takes in the destination:
// source-grid deconvolution, 1/n_lattices, interlacing phase and the half-cell phase between the two grids.
PM_HD double2 mode_value(const double2 v, int ki, int kj, int kk, int is, int js, const CopyParams& p, const double* tab_x,
                         const double* tab_sin) {
    double factor = 1;
    if (p.deconv_order) {
        // ((xi·xj)·xk)/((si·sj)·sk), then **D  (mesh.py:2795-2856)
        factor = ((tab_x[is] * tab_x[js]) * tab_x[kk]) / ((tab_sin[is] * tab_sin[js]) * tab_sin[kk]);
        factor = ipow(factor, p.deconv_order);
    }
    factor *= p.scale;
    double theta = p.cell_phase * ((ki + kj) + kk);
    if (p.rotate) theta += (ki * p.th[0] + kj * p.th[1]) + kk * p.th[2];
    double sn, cs;
    sincos(theta, &sn, &cs);
    double2 o;
    o.x = factor * (v.x * cs - v.y * sn);
    o.y = factor * (v.x * sn + v.y * cs);
    return o;
}

// Mode idx of the destination slab.  Returns false when the mode lies outside the cube |k| < min(Gs, Gd)/2 that
// the two grids share (Nyquist planes of the smaller grid excluded, mesh.py:1134-1135): '=' writes zero there,
// '+=' leaves it alone.  tab_x / tab_sin: the SOURCE context's x_l = k_l·π/Gs + ε and sin x_l by slab index.
PM_HD bool copy_mode(int64_t idx, const double2* src, const CopyParams& p, const double* tab_x, const double* tab_sin,
                     double2* out) {
    const int Gcd = p.Gd / 2 + 1, Gcs = p.Gs / 2 + 1;
    const int n = (p.Gs < p.Gd ? p.Gs : p.Gd) / 2;
    const int nyqd = p.Gd / 2;
    const int64_t row = idx / Gcd;
    const int kk = (int)(idx - row * Gcd);
    const int i = (int)(row / p.Gd);
    const int j = (int)(row - (int64_t)i * p.Gd);
    const int ki = i - (i >= nyqd ? p.Gd : 0);
    const int kj = j - (j >= nyqd ? p.Gd : 0);
    out->x = 0.0; out->y = 0.0;
    if (!(ki > -n && ki < n && kj > -n && kj < n && kk < n)) return false;
    const int is = ki < 0 ? ki + p.Gs : ki;
    const int js = kj < 0 ? kj + p.Gs : kj;
    *out = mode_value(src[((int64_t)is * p.Gs + js) * Gcs + kk], ki, kj, kk, is, js, p, tab_x, tab_sin);
    return true;
}

}  // namespace copyops
}  // namespace pm
