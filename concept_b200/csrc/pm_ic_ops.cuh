// pm_ic_ops.cuh — per-element arithmetic of the initial-condition kernels (pm_ic.cu).
//
// Everything here is __host__ __device__ and free of CUDA-only types beyond double2, so that
// tests/ic_host_harness.cu can compile the very same code with g++ and run it on the CPU against the
// reference's golden vectors (tests/test_widen_ic.py) — like pm_fftcore.cuh does for the transform stages.
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
namespace icops {

// local Fourier slab [G (i)][njl (j_local)][G/2+1 (kk)] of complex values, j_global = j0 + j_local
struct Slab {
    int G, njl, j0;
};

// preinitialize_particles (ic.py:2197-2227): x = (ℝ[domain_bgn + ½ + shift] + i)·ℝ[boxsize/gridsize]
PM_HD void lattice_point(int64_t p, int n, double bx, double by, double bz, double cell, double* xyz) {
    const int k = (int)(p % n);
    const int64_t r = p / n;
    const int j = (int)(r % n);
    const int i = (int)(r / n);
    xyz[0] = (bx + i) * cell;
    xyz[1] = (by + j) * cell;
    xyz[2] = (bz + k) * cell;
}

// realize_grid (ic.py:711-764, scalar) followed by laplacian_inverse (mesh.py:3432-3436) for one mode:
//   amplitude·(re, im)·e^{iθ} · lap/k²,  lap = −factor/k_f²;  origin and Nyquist planes → 0
PM_HD double2 potential_mode(int64_t idx, Slab s, const double2* noise, const double* amplitudes, int k2_max,
                             double th0, double th1, double th2, int rotate, double lap) {
    const int Gc = s.G / 2 + 1;
    const int nyq = s.G / 2;
    const int64_t row = idx / Gc;
    const int kk = (int)(idx - row * Gc);
    const int i = (int)(row / s.njl);
    const int j = s.j0 + (int)(row - (int64_t)i * s.njl);
    double2 out;
    out.x = 0.0; out.y = 0.0;
    if (i == nyq || j == nyq || kk == nyq) return out;
    const int ki = i - (i >= nyq ? s.G : 0);
    const int kj = j - (j >= nyq ? s.G : 0);
    const int k2 = (kj * kj + ki * ki) + kk * kk;
    if (k2 == 0 || k2 > k2_max) return out;
    const double2 v = noise[idx];
    double re = v.x, im = v.y;
    if (rotate) {
        const double theta = (ki * th0 + kj * th1) + kk * th2;
        double sn, cs;
        sincos(theta, &sn, &cs);
        const double r2 = re * cs - im * sn;
        const double i2 = re * sn + im * cs;
        re = r2; im = i2;
    }
    const double amplitude = amplitudes[k2];
    if (lap == 0.0) {          // realize_grid on its own (the non-Gaussian path transforms before the inverse Laplacian)
        out.x = amplitude * re;
        out.y = amplitude * im;
        return out;
    }
    const double inv = lap / k2;
    out.x = (amplitude * re) * inv;
    out.y = (amplitude * im) * inv;
    return out;
}

// local non-Gaussianity (realize_grid, ic.py:766-771): x += f·x²
PM_HD double nongaussian_point(double x, double f) { return x + f * (x * x); }

// lattice particle p = (i·G + j)·G + k  →  element of the padded real grid [nxl][G][Gp]
PM_HD int64_t real_index(int64_t p, int G, int Gp) { return (p / G) * Gp + (p % G); }

// carryout_2lpt (ic.py:1570-1575), accumulated in the reference's order
PM_HD double lpt2_source(double d00, double d11, double d22, double d01, double d12, double d02) {
    double v = -(d00 * d11);
    v -= d11 * d22;
    v -= d22 * d00;
    v += d01 * d01;
    v += d12 * d12;
    v += d02 * d02;
    return v;
}

// one product of handle_lpt_term (ic.py:2003-2021): the grids are multiplied into the base grid one after the other
PM_HD double lpt_product(double a, double b, double c, int has_c) {
    double v = a * b;
    if (has_c) v *= c;
    return v;
}

// resize_grid(…, 'fourier') as the LPT code uses it, one rank: mode idx of the destination slab [Gd][Gd][Gd/2+1]
PM_HD double2 resize_mode(int64_t idx, const double2* src, int Gs, int Gd) {
    const int Gcd = Gd / 2 + 1, Gcs = Gs / 2 + 1;
    const int n = (Gs < Gd ? Gs : Gd) / 2;
    const int nyqd = Gd / 2;
    const int64_t row = idx / Gcd;
    const int kk = (int)(idx - row * Gcd);
    const int i = (int)(row / Gd);
    const int j = (int)(row - (int64_t)i * Gd);
    const int ki = i - (i >= nyqd ? Gd : 0);
    const int kj = j - (j >= nyqd ? Gd : 0);
    double2 v;
    v.x = 0.0; v.y = 0.0;
    if (ki > -n && ki < n && kj > -n && kj < n && kk < n) {
        const int is = ki < 0 ? ki + Gs : ki;
        const int js = kj < 0 ? kj + Gs : kj;
        v = src[((int64_t)is * Gs + js) * Gcs + kk];
    }
    return v;
}

// mod(x, boxsize) with Python semantics (ic.py:1396-1398)
PM_HD double mod_box(double x, double L) {
    double r = fmod(x, L);
    if (r < 0) r += L;
    if (r == L) r = 0;
    return r;
}

}  // namespace icops
}  // namespace pm
