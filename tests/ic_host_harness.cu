// CPU build of the per-element code of the initial-condition kernels (concept_b200/csrc/pm_ic_ops.cuh):
// the same __host__ __device__ functions pm_ic.cu's kernels call, looped over sequentially.  Compiled with
// g++ into a shared library and driven through ctypes by tests/test_widen_ic.py (no GPU needed).
#include <cstdint>

#include "pm_copy_ops.cuh"
#include "pm_ic_ops.cuh"

using namespace pm::icops;

extern "C" {

void h_lattice(double* pos, int n, int nxl, double bx, double by, double bz, double cell, int64_t index_bgn) {
    for (int64_t p = 0; p < (int64_t)nxl * n * n; ++p) lattice_point(p, n, bx, by, bz, cell, pos + 3 * (index_bgn + p));
}

void h_potential(const double* noise, double* dst, int G, int njl, int j0, const double* amplitudes, int k2_max,
                 const double* th, int rotate, double lap) {
    const Slab s{G, njl, j0};
    double2* out = reinterpret_cast<double2*>(dst);
    for (int64_t idx = 0; idx < (int64_t)G * njl * (G / 2 + 1); ++idx)
        out[idx] = potential_mode(idx, s, reinterpret_cast<const double2*>(noise), amplitudes, k2_max, th[0], th[1], th[2],
                                  rotate, lap);
}

void h_displace(double* pos, double* mom, const double* grid, int G, int Gp, int nxl, int64_t index_bgn, int dim,
                double pos_factor, double mom_factor) {
    for (int64_t p = 0; p < (int64_t)nxl * G * G; ++p) {
        const double psi = grid[real_index(p, G, Gp)];
        const int64_t q = 3 * (index_bgn + p) + dim;
        if (pos) pos[q] += pos_factor * psi;
        if (mom) mom[q] += mom_factor * psi;
    }
}

void h_export(const double* grid, double* out, int G, int Gp, int nxl) {
    for (int64_t p = 0; p < (int64_t)nxl * G * G; ++p) out[p] = grid[real_index(p, G, Gp)];
}

void h_source(double* grid, const double* d00, const double* d11, const double* d22, const double* d01, const double* d12,
              const double* d02, int G, int Gp, int nxl) {
    for (int64_t p = 0; p < (int64_t)nxl * G * G; ++p)
        grid[real_index(p, G, Gp)] = lpt2_source(d00[p], d11[p], d22[p], d01[p], d12[p], d02[p]);
}

void h_resize(const double* src, double* dst, int Gs, int Gd) {
    double2* out = reinterpret_cast<double2*>(dst);
    for (int64_t idx = 0; idx < (int64_t)Gd * Gd * (Gd / 2 + 1); ++idx)
        out[idx] = resize_mode(idx, reinterpret_cast<const double2*>(src), Gs, Gd);
}

void h_nongaussianity(double* grid, int G, int Gp, int nxl, double f) {
    for (int64_t p = 0; p < (int64_t)nxl * G * G; ++p) {
        const int64_t q = real_index(p, G, Gp);
        grid[q] = nongaussian_point(grid[q], f);
    }
}

void h_lpt_accumulate(double* acc, int64_t n, double factor, const double* a, const double* b, const double* c, int assign) {
    for (int64_t p = 0; p < n; ++p) {
        const double v = factor * lpt_product(a[p], b[p], c ? c[p] : 1.0, c != nullptr);
        acc[p] = assign ? v : acc[p] + v;
    }
}

void h_import(double* grid, const double* in, int G, int Gp, int nxl) {
    for (int64_t p = 0; p < (int64_t)nxl * G * G; ++p) grid[real_index(p, G, Gp)] = in[p];
}

void h_wrap(double* pos, int64_t n3, double L) {
    for (int64_t i = 0; i < n3; ++i) pos[i] = mod_box(pos[i], L);
}

// pm_fourier_copy_modes (csrc/pm_copymodes.cu): '=' or '+=' of slab [Gs][Gs][Gs/2+1] onto [Gd][Gd][Gd/2+1]
void h_copy_modes(const double* src, double* dst, int Gs, int Gd, int deconv_order, const double* th, int rotate,
                  double cell_phase, double scale, const double* tab_x, const double* tab_sin, int accumulate) {
    pm::copyops::CopyParams p;
    p.Gs = Gs; p.Gd = Gd; p.deconv_order = deconv_order; p.rotate = rotate;
    p.th[0] = th[0]; p.th[1] = th[1]; p.th[2] = th[2];
    p.cell_phase = cell_phase; p.scale = scale;
    double2* out = reinterpret_cast<double2*>(dst);
    for (int64_t idx = 0; idx < (int64_t)Gd * Gd * (Gd / 2 + 1); ++idx) {
        double2 v;
        const bool shared = pm::copyops::copy_mode(idx, reinterpret_cast<const double2*>(src), p, tab_x, tab_sin, &v);
        if (!accumulate) out[idx] = v;
        else if (shared) { out[idx].x += v.x; out[idx].y += v.y; }
    }
}

}  // extern "C"
