// pm_ic.cu — device side of the particle initial-condition generator.
//
// Reference (all under src/): ic.py:1199-1399 realize_particles, :670-782 realize_grid, :1447-1589
// carryout_1lpt / carryout_2lpt, :2138-2247 preinitialize_particles, :2249-2283 displace_particles;
// mesh.py:3422-3437 laplacian_inverse, :3470-3510 fourier_diff, mesh.py resize_grid (dealiasing).
//
// The primordial noise itself is drawn on the host (sequential NumPy bit streams, ic.py:928-1163);
// everything downstream of it runs here.  All kernels are streaming, HBM-bound passes over a G³ grid
// or over the particle arrays; grids are fp64 (PM_GRID_F64 contexts only).
#include "pm_internal.cuh"
#include "pm_ic_ops.cuh"

namespace pm {

using icops::Slab;

// pos at lattice points, mom = 0, ids by lattice point (preinitialize_particles, ic.py:2197-2243)
__global__ void __launch_bounds__(256)
ic_lattice_kernel(double* __restrict__ pos, double* __restrict__ mom, int64_t* __restrict__ ids, int n, int nxl,
                  double bx, double by, double bz, double cell, int64_t index_bgn, int64_t id_bgn, int64_t id_plane0) {
    const int64_t total = (int64_t)nxl * n * n;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        icops::lattice_point(p, n, bx, by, bz, cell, pos + 3 * (index_bgn + p));
        double* m = mom + 3 * (index_bgn + p);
        m[0] = 0; m[1] = 0; m[2] = 0;
        if (ids != nullptr) ids[index_bgn + p] = id_bgn + id_plane0 + p;
    }
}

// realize_grid (scalar, Fourier output) + laplacian_inverse in one pass over the local slab
__global__ void __launch_bounds__(256)
ic_potential_kernel(const double2* __restrict__ noise, double2* __restrict__ dst, Slab s,
                    const double* __restrict__ amplitudes, int k2_max, double th0, double th1, double th2, int rotate,
                    double lap) {
    const int64_t total = (int64_t)s.G * s.njl * (s.G / 2 + 1);
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x)
        dst[idx] = icops::potential_mode(idx, s, noise, amplitudes, k2_max, th0, th1, th2, rotate, lap);
}

// displace_particles (ic.py:2249-2283): lattice particle p = (i·G + j)·G + k reads grid point (i, j, k)
__global__ void __launch_bounds__(256)
ic_displace_kernel(double* __restrict__ pos, double* __restrict__ mom, const double* __restrict__ grid, int G, int Gp,
                   int nxl, int64_t index_bgn, int dim, double pos_factor, double mom_factor) {
    const int64_t total = (int64_t)nxl * G * G;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const double psi = __ldcs(grid + icops::real_index(p, G, Gp));
        const int64_t q = 3 * (index_bgn + p) + dim;
        if (pos != nullptr) pos[q] += pos_factor * psi;
        if (mom != nullptr) mom[q] += mom_factor * psi;
    }
}

__global__ void __launch_bounds__(256) ic_wrap_kernel(double* __restrict__ pos, int64_t n3, double L) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n3; i += (int64_t)gridDim.x * blockDim.x)
        pos[i] = icops::mod_box(pos[i], L);
}

// local non-Gaussianity on the real-space grid: x += f·x²
__global__ void __launch_bounds__(256)
ic_nongaussianity_kernel(double* __restrict__ grid, int G, int Gp, int64_t total, double f) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = icops::real_index(p, G, Gp);
        grid[q] = icops::nongaussian_point(grid[q], f);
    }
}

// real grid (padded rows) → compact [nxl][G][G]
__global__ void __launch_bounds__(256)
real_export_kernel(const double* __restrict__ grid, double* __restrict__ out, int G, int Gp, int64_t total) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x)
        out[p] = __ldcs(grid + icops::real_index(p, G, Gp));
}

// source of the 2LPT potential (carryout_2lpt, ic.py:1553-1575)
__global__ void __launch_bounds__(256)
ic_2lpt_source_kernel(double* __restrict__ grid, const double* __restrict__ d00, const double* __restrict__ d11,
                      const double* __restrict__ d22, const double* __restrict__ d01, const double* __restrict__ d12,
                      const double* __restrict__ d02, int G, int Gp, int64_t total) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x)
        grid[icops::real_index(p, G, Gp)] = icops::lpt2_source(d00[p], d11[p], d22[p], d01[p], d12[p], d02[p]);
}

// one term of an LPT source (handle_lpt_term, ic.py:1895-2057) on compact grids: acc (=|+=) factor·a·b[·c]
__global__ void __launch_bounds__(256)
lpt_accumulate_kernel(double* __restrict__ acc, int64_t n, double factor, const double* __restrict__ a,
                      const double* __restrict__ b, const double* __restrict__ c, int assign) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const double v = factor * icops::lpt_product(a[p], b[p], c != nullptr ? c[p] : 1.0, c != nullptr);
        acc[p] = assign ? v : acc[p] + v;
    }
}

// compact [nxl][G][G] → real grid (padded rows)
__global__ void __launch_bounds__(256)
real_import_kernel(double* __restrict__ grid, const double* __restrict__ in, int G, int Gp, int64_t total) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x)
        grid[icops::real_index(p, G, Gp)] = in[p];
}

// resize_grid(…, 'fourier'): dst[k] = src[k] for |k| < min(G_src, G_dst)/2, zero elsewhere
__global__ void __launch_bounds__(256)
fourier_resize_kernel(const double2* __restrict__ src, double2* __restrict__ dst, int Gs, int Gd) {
    const int64_t total = (int64_t)Gd * Gd * (Gd / 2 + 1);
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x)
        dst[idx] = icops::resize_mode(idx, src, Gs, Gd);
}

}  // namespace pm

using namespace pm;

static int ic_check(pm_ctx* c, const char* who) {
    PM_REQUIRE(c != nullptr, "%s: NULL context", who);
    PM_REQUIRE(c->dtype == PM_GRID_F64, "%s: initial conditions need a PM_GRID_F64 context", who);
    return PM_OK;
}

extern "C" {

int pm_ic_lattice(pm_ctx* c, double* pos, double* mom, int64_t* ids, const double* shift, int64_t index_bgn,
                  int64_t id_bgn, int64_t* n_local_out) {
    PM_TRY(ic_check(c, "pm_ic_lattice"));
    PM_REQUIRE(pos != nullptr && mom != nullptr && index_bgn >= 0, "pm_ic_lattice: bad argument");
    const Geom& g = c->g;
    const int64_t total = (int64_t)g.nxl * g.G * g.G;
    const double s0 = shift ? shift[0] : 0.0, s1 = shift ? shift[1] : 0.0, s2 = shift ? shift[2] : 0.0;
    // ℝ[domain_bgn + 0.5*cell_centered + lattice.shift] (ic.py:2206-2216); cell-centred grids only
    PM_LAUNCH(ic_lattice_kernel, kNumSMs * 4, 256, 0, c->stream, pos, mom, ids, g.G, g.nxl, g.x0 + 0.5 + s0, 0 + 0.5 + s1,
              0 + 0.5 + s2, c->boxsize / g.G, index_bgn, id_bgn, (int64_t)g.x0 * g.G * g.G);
    if (n_local_out) *n_local_out = total;
    return PM_OK;
}

int pm_ic_potential(pm_ctx* c, const double* noise, const double* amplitudes, int k2_max, const double* shift,
                    double lap_factor) {
    PM_TRY(ic_check(c, "pm_ic_potential"));
    PM_REQUIRE(noise != nullptr && amplitudes != nullptr && k2_max >= 1, "pm_ic_potential: bad argument");
    const Geom& g = c->g;
    PM_REQUIRE(k2_max >= 3 * (g.G / 2 - 1) * (g.G / 2 - 1), "pm_ic_potential: amplitude table too short (k2_max = %d)", k2_max);
    const double kf = 2 * M_PI / c->boxsize;
    double th[3];
    int rotate = 0;
    for (int d = 0; d < 3; ++d) {
        // realize_grid negates the particle shift (ic.py:692-695); fourier_loop: θ = −2π/G·k·shift' (mesh.py:2873-2888)
        const double s = shift ? -shift[d] : 0.0;
        th[d] = -2 * M_PI / g.G * s;
        if (s != 0.0) rotate = 1;
    }
    PM_LAUNCH(ic_potential_kernel, kNumSMs * 8, 256, 0, c->stream, reinterpret_cast<const double2*>(noise),
              reinterpret_cast<double2*>(c->fourier), Slab{g.G, g.njl, g.j0}, amplitudes, k2_max, th[0], th[1], th[2], rotate,
              -lap_factor / (kf * kf));
    c->space_fourier = true;
    c->grid_in_phi = false;
    c->real_is_zero = false;
    return PM_OK;
}

int pm_ic_displace(pm_ctx* c, double* pos, double* mom, int64_t index_bgn, int dim, double pos_factor,
                   double mom_factor) {
    PM_TRY(ic_check(c, "pm_ic_displace"));
    PM_REQUIRE(dim >= 0 && dim < 3 && index_bgn >= 0, "pm_ic_displace: bad argument");
    PM_REQUIRE(!c->space_fourier, "pm_ic_displace: the slab holds Fourier data (call pm_fft_backward)");
    const Geom& g = c->g;
    PM_LAUNCH(ic_displace_kernel, kNumSMs * 8, 256, 0, c->stream, pos, mom,
              reinterpret_cast<const double*>(c->grid_read()) + (size_t)g.halo * g.G * g.Gp, g.G, g.Gp, g.nxl, index_bgn,
              dim, pos_factor, mom_factor);
    return PM_OK;
}

int pm_ic_wrap(pm_ctx* c, double* pos, int64_t n) {
    PM_REQUIRE(c != nullptr && n >= 0 && (pos != nullptr || n == 0), "pm_ic_wrap: bad argument");
    if (n == 0) return PM_OK;
    PM_LAUNCH(ic_wrap_kernel, kNumSMs * 8, 256, 0, c->stream, pos, 3 * n, c->boxsize);
    return PM_OK;
}

int pm_ic_nongaussianity(pm_ctx* c, double f_nl) {
    PM_TRY(ic_check(c, "pm_ic_nongaussianity"));
    PM_REQUIRE(!c->space_fourier, "pm_ic_nongaussianity: the slab holds Fourier data (call pm_fft_backward)");
    PM_TRY(ensure_in_real(c));
    const Geom& g = c->g;
    PM_LAUNCH(ic_nongaussianity_kernel, kNumSMs * 8, 256, 0, c->stream, c->real_interior<double>(), g.G, g.Gp,
              (int64_t)g.nxl * g.G * g.G, f_nl);
    c->real_is_zero = false;
    return PM_OK;
}

int pm_real_export(pm_ctx* c, double* dev_out) {
    PM_TRY(ic_check(c, "pm_real_export"));
    PM_REQUIRE(dev_out != nullptr, "pm_real_export: NULL output");
    PM_REQUIRE(!c->space_fourier, "pm_real_export: the slab holds Fourier data");
    const Geom& g = c->g;
    const int64_t total = (int64_t)g.nxl * g.G * g.G;
    PM_LAUNCH(real_export_kernel, kNumSMs * 8, 256, 0, c->stream,
              reinterpret_cast<const double*>(c->grid_read()) + (size_t)g.halo * g.G * g.Gp, dev_out, g.G, g.Gp, total);
    return PM_OK;
}

int pm_ic_2lpt_source(pm_ctx* c, const double* d00, const double* d11, const double* d22, const double* d01,
                      const double* d12, const double* d02) {
    PM_TRY(ic_check(c, "pm_ic_2lpt_source"));
    PM_REQUIRE(d00 && d11 && d22 && d01 && d12 && d02, "pm_ic_2lpt_source: NULL input");
    const Geom& g = c->g;
    const int64_t total = (int64_t)g.nxl * g.G * g.G;
    PM_LAUNCH(ic_2lpt_source_kernel, kNumSMs * 8, 256, 0, c->stream, c->real_interior<double>(), d00, d11, d22, d01, d12,
              d02, g.G, g.Gp, total);
    c->space_fourier = false;
    c->grid_in_phi = false;
    c->real_is_zero = false;
    return PM_OK;
}

int pm_lpt_accumulate(pm_ctx* c, double* acc, int64_t n, double factor, const double* a, const double* b,
                      const double* third, int assign) {
    PM_REQUIRE(c != nullptr && acc != nullptr && a != nullptr && b != nullptr && n >= 0, "pm_lpt_accumulate: bad argument");
    if (n == 0) return PM_OK;
    PM_LAUNCH(lpt_accumulate_kernel, kNumSMs * 8, 256, 0, c->stream, acc, n, factor, a, b, third, assign ? 1 : 0);
    return PM_OK;
}

int pm_real_import(pm_ctx* c, const double* dev_in) {
    PM_TRY(ic_check(c, "pm_real_import"));
    PM_REQUIRE(dev_in != nullptr, "pm_real_import: NULL input");
    const Geom& g = c->g;
    PM_LAUNCH(real_import_kernel, kNumSMs * 8, 256, 0, c->stream, c->real_interior<double>(), dev_in, g.G, g.Gp,
              (int64_t)g.nxl * g.G * g.G);
    c->space_fourier = false;
    c->grid_in_phi = false;
    c->real_is_zero = false;
    return PM_OK;
}

int pm_fourier_resize(pm_ctx* src, pm_ctx* dst) {
    PM_TRY(ic_check(src, "pm_fourier_resize"));
    PM_TRY(ic_check(dst, "pm_fourier_resize"));
    PM_REQUIRE(src != dst, "pm_fourier_resize: source and destination are the same context");
    PM_REQUIRE(src->nranks == 1 && dst->nranks == 1, "pm_fourier_resize: single-rank contexts only");
    PM_REQUIRE(src->space_fourier, "pm_fourier_resize: the source slab holds real-space data");
    PM_REQUIRE(src->device == dst->device, "pm_fourier_resize: contexts live on different devices");
    // order the two contexts' streams: dst waits for everything queued on src
    if (src->stream != dst->stream) {
        cudaEvent_t ev;
        PM_CHECK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        PM_CHECK_CUDA(cudaEventRecord(ev, src->stream));
        PM_CHECK_CUDA(cudaStreamWaitEvent(dst->stream, ev, 0));
        PM_CHECK_CUDA(cudaEventDestroy(ev));
    }
    PM_LAUNCH(fourier_resize_kernel, kNumSMs * 8, 256, 0, dst->stream, reinterpret_cast<const double2*>(src->fourier),
              reinterpret_cast<double2*>(dst->fourier), src->g.G, dst->g.G);
    if (src->stream != dst->stream) {
        // and src must not overwrite its slab before dst has read it
        cudaEvent_t ev;
        PM_CHECK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        PM_CHECK_CUDA(cudaEventRecord(ev, dst->stream));
        PM_CHECK_CUDA(cudaStreamWaitEvent(src->stream, ev, 0));
        PM_CHECK_CUDA(cudaEventDestroy(ev));
    }
    dst->space_fourier = true;
    dst->grid_in_phi = false;
    dst->real_is_zero = false;
    return PM_OK;
}

}  // extern "C"
