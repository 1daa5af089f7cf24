// pm_api.cu — the C ABI of libpmgrav.so (include/pmgrav.h): context life-cycle, thin wrappers
// over the kernels, the whole-kick orchestration and the parity taps.
#include "pm_internal.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace pm {

std::atomic<int64_t> g_launches{0};
static thread_local char g_error[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int ensure_force(pm_ctx* c) {
    if (c->force) return PM_OK;
    const size_t bytes = c->real_elems * c->elem_size();
    PM_CHECK_CUDA(cudaMalloc(&c->force, bytes));
    PM_CHECK_CUDA(cudaMemsetAsync(c->force, 0, bytes, c->stream));
    c->bytes_allocated += bytes;
    return PM_OK;
}

int ensure_in_real(pm_ctx* c) {
    if (!c->grid_in_phi) return PM_OK;
    PM_CHECK_CUDA(cudaMemcpyAsync(c->real, c->phi, c->real_elems * c->elem_size(), cudaMemcpyDeviceToDevice, c->stream));
    c->grid_in_phi = false;
    c->real_is_zero = false;
    return PM_OK;
}

static int halo_for_gather(int order, int diff_order, int interlace, int* lo, int* hi) {
    // planes touched along x beyond the slab: interpolation reach + difference reach; the
    // half-cell lattice shift of interlacing moves the footprint by up to one more plane
    const int reach = diff_order == 0 ? 0 : (diff_order <= 2 ? 1 : diff_order / 2);
    const int interp = (order <= 3 ? 1 : 2) + (interlace ? 1 : 0);   // NGP/CIC/TSC: 1; PCS: 2
    *lo = interp + reach;
    *hi = interp + reach;
    if (*lo > kHalo) *lo = kHalo;
    if (*hi > kHalo) *hi = kHalo;
    return PM_OK;
}

}  // namespace pm

using namespace pm;

template <typename T>
static int get_grid_t(pm_ctx* c, int which, double* host_out) {
    const Geom& g = c->g;
    if (which == PM_TAP_FOURIER) {
        const size_t n = c->fourier_elems * 2;
        std::vector<T> tmp(n);
        PM_CHECK_CUDA(cudaMemcpy(tmp.data(), c->fourier, n * sizeof(T), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; ++i) host_out[i] = (double)tmp[i];
        return PM_OK;
    }
    const T* src = reinterpret_cast<const T*>(which == PM_TAP_FORCE ? c->force : c->grid_read()) +
                   (size_t)g.halo * g.G * g.Gp;
    const size_t n = (size_t)g.nxl * g.G * g.Gp;
    std::vector<T> tmp(n);
    PM_CHECK_CUDA(cudaMemcpy(tmp.data(), src, n * sizeof(T), cudaMemcpyDeviceToHost));
    for (int64_t r = 0; r < (int64_t)g.nxl * g.G; ++r)
        for (int k = 0; k < g.G; ++k) host_out[r * g.G + k] = (double)tmp[r * g.Gp + k];
    return PM_OK;
}

template <typename T>
static int set_grid_t(pm_ctx* c, const double* host_in) {
    const Geom& g = c->g;
    const size_t n = (size_t)g.nxl * g.G * g.Gp;
    std::vector<T> tmp(n, (T)0);
    for (int64_t r = 0; r < (int64_t)g.nxl * g.G; ++r)
        for (int k = 0; k < g.G; ++k) tmp[r * g.Gp + k] = (T)host_in[r * g.G + k];
    T* dst = reinterpret_cast<T*>(c->real) + (size_t)g.halo * g.G * g.Gp;
    PM_CHECK_CUDA(cudaMemcpy(dst, tmp.data(), n * sizeof(T), cudaMemcpyHostToDevice));
    c->grid_in_phi = false;
    c->real_is_zero = false;
    return PM_OK;
}

extern "C" {

const char* pm_version(void) { return "pmgrav 0.1 (sm_100a)"; }
const char* pm_last_error(void) { return g_error; }
int64_t pm_launch_count(void) { return g_launches.load(); }

int pm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int pm_create(pm_ctx** out, int gridsize, double boxsize, int grid_dtype, int rank, int nranks,
              int device, void* stream) {
    PM_REQUIRE(out != nullptr, "pm_create: out is NULL");
    *out = nullptr;
    PM_REQUIRE(gridsize >= 2 && gridsize % 2 == 0, "pm_create: gridsize = %d must be even and >= 2", gridsize);
    PM_REQUIRE(boxsize > 0, "pm_create: boxsize must be positive");
    PM_REQUIRE(grid_dtype == PM_GRID_F64 || grid_dtype == PM_GRID_F32, "pm_create: grid_dtype = %d", grid_dtype);
    PM_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "pm_create: rank %d of %d", rank, nranks);
    // mesh.py:3779-3783: the slabs must tile the grid
    PM_REQUIRE(gridsize % nranks == 0, "pm_create: gridsize = %d is not divisible by nranks = %d", gridsize, nranks);
    if (nranks > 1) {
        // mesh.py:1911-1919 analogue: a slab must be at least as thick as the halos it feeds
        PM_REQUIRE(gridsize / nranks >= kHalo, "pm_create: slab of %d planes is thinner than the halo (%d)",
                   gridsize / nranks, kHalo);
    }
    PM_CHECK_CUDA(cudaSetDevice(device));
    pm_ctx* c = new pm_ctx();
    memset(static_cast<void*>(c), 0, sizeof(*c));
    c->boxsize = boxsize;
    c->dtype = grid_dtype;
    c->rank = rank;
    c->nranks = nranks;
    c->device = device;
    Geom& g = c->g;
    g.G = gridsize;
    g.Gc = gridsize / 2 + 1;
    g.Gp = 2 * g.Gc;
    g.nxl = gridsize / nranks;
    g.x0 = rank * g.nxl;
    g.halo = nranks > 1 ? kHalo : 0;
    g.wrap_x = nranks == 1;
    g.njl = gridsize / nranks;
    g.j0 = rank * g.njl;
    // NULL is the legacy default stream: work is ordered with whatever the caller runs there
    c->stream = reinterpret_cast<cudaStream_t>(stream);
    c->own_stream = false;
    const size_t es = c->elem_size();
    c->real_elems = (size_t)(g.nxl + 2 * g.halo) * g.G * g.Gp;
    c->fourier_elems = (size_t)g.G * g.njl * g.Gc;
    auto fail = [&](int code) {
        pm_destroy(c);
        return code;
    };
    // the hand-written transform (pm_fft.cu) keeps two intermediate copies of the Fourier slab behind the grid
    size_t alloc_bytes = c->real_elems * es;
    if (gridsize == 128 || gridsize == 256 || gridsize == 512 || gridsize == 1024) {
        const size_t layout_bytes = (size_t)g.nxl * (gridsize / 2) * gridsize * 2 * es;
        c->f2_off_a = (alloc_bytes + 255) / 256 * 256;
        c->f2_off_b = c->f2_off_a + (layout_bytes + 255) / 256 * 256;
        alloc_bytes = c->f2_off_b + layout_bytes;
    }
    if (nranks > 1) {
        // IPC arena: header | migration mailboxes (2 parities × nranks slots) | P³M ghost positions (2 parities × 2 sides)
        const size_t slab_bytes = c->real_elems * es;
        size_t mailbox_total = std::max<size_t>((size_t)16 << 20, slab_bytes / 4);
        size_t ghost_total = std::max<size_t>((size_t)16 << 20, slab_bytes / 2);
        if (const char* e = getenv("PM_MAILBOX_MB")) mailbox_total = (size_t)std::max(1, atoi(e)) << 20;
        if (const char* e = getenv("PM_GHOST_MB")) ghost_total = (size_t)std::max(1, atoi(e)) << 20;
        c->mailbox_slot_bytes = mailbox_total / (2 * (size_t)nranks) / 256 * 256;
        c->ghost_buf_bytes = ghost_total / 4 / 256 * 256;
        c->off_arena = (alloc_bytes + 255) / 256 * 256;
        c->off_mailbox = c->off_arena + kArenaHeaderBytes;
        c->off_ghost = c->off_mailbox + 2 * (size_t)nranks * c->mailbox_slot_bytes;
        c->arena_bytes = kArenaHeaderBytes + 2 * (size_t)nranks * c->mailbox_slot_bytes + 4 * c->ghost_buf_bytes;
        alloc_bytes = c->off_arena + c->arena_bytes;
    }
    if (cudaMalloc(&c->real, alloc_bytes) != cudaSuccess) {
        set_error("pm_create: cannot allocate %zu bytes for the grid", alloc_bytes);
        return fail(PM_ERR_ALLOC);
    }
    c->bytes_allocated += alloc_bytes;
    cudaMemsetAsync(c->real, 0, c->real_elems * es, c->stream);
    c->real_is_zero = true;
    if (nranks > 1) cudaMemsetAsync(reinterpret_cast<char*>(c->real) + c->off_arena, 0, kArenaHeaderBytes, c->stream);
    if (nranks == 1) {
        c->fourier = c->real;
    } else {
        const size_t fb = c->fourier_elems * 2 * es;
        if (cudaMalloc(&c->fourier, fb) != cudaSuccess || cudaMalloc(&c->sendbuf, fb) != cudaSuccess) {
            set_error("pm_create: cannot allocate the Fourier slab / transpose buffer (%zu bytes each)", fb);
            return fail(PM_ERR_ALLOC);
        }
        c->bytes_allocated += 2 * fb;
    }
    // k-space tables: x_l = k_l·π/G + ε and sin(x_l) over signed wavenumbers (mesh.py:2775-2776)
    std::vector<double> tx(g.G), ts(g.G);
    for (int i = 0; i < g.G; ++i) {
        const int k = i - (i >= g.G / 2 ? g.G : 0);
        tx[i] = k * (M_PI / g.G) + kEps;
        ts[i] = sin(tx[i]);
    }
    if (cudaMalloc(&c->tab_x, sizeof(double) * g.G) != cudaSuccess ||
        cudaMalloc(&c->tab_sin, sizeof(double) * g.G) != cudaSuccess ||
        cudaMalloc(&c->d_scratch, sizeof(double) * 64) != cudaSuccess ||
        cudaMalloc(&c->d_tilectr, sizeof(unsigned long long) * 4) != cudaSuccess ||
        cudaMalloc(&c->d_comm_err, sizeof(int) * 4) != cudaSuccess ||
        cudaMallocHost(&c->h_pinned, 4096) != cudaSuccess ||
        cudaMalloc(&c->d_counts, sizeof(int64_t) * (3 * nranks + (size_t)nranks * nranks + 8 + kNumSMs * 4)) != cudaSuccess) {
        set_error("pm_create: cannot allocate tables");
        return fail(PM_ERR_ALLOC);
    }
    cudaMemcpyAsync(c->tab_x, tx.data(), sizeof(double) * g.G, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(c->tab_sin, ts.data(), sizeof(double) * g.G, cudaMemcpyHostToDevice, c->stream);
    cudaMemsetAsync(c->d_scratch, 0, sizeof(double) * 64, c->stream);
    cudaMemsetAsync(c->d_comm_err, 0, sizeof(int) * 4, c->stream);
    cudaStreamSynchronize(c->stream);   // tx/ts go out of scope
    c->fused_solve = true;
    c->solve_mode = PM_SOLVE_AUTO;
    int s = make_xsolve_tables(c);
    if (s != PM_OK) return fail(s);
    s = make_fft2_tables(c);
    if (s != PM_OK) return fail(s);
    s = make_plans(c);
    if (s != PM_OK) return fail(s);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
        set_error("pm_create: %s", cudaGetErrorString(e));
        return fail(PM_ERR_CUDA);
    }
    *out = c;
    return PM_OK;
}

int pm_destroy(pm_ctx* c) {
    if (!c) return PM_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    destroy_plans(c);
    if (c->peers_ready)
        for (int r = 0; r < c->nranks; ++r)
            if (r != c->rank && c->peer_real[r]) cudaIpcCloseMemHandle(c->peer_real[r]);
    if (c->comm_ready) ncclCommDestroy(c->comm);
    cudaFree(c->xs_tw);
    cudaFree(c->f2_tw);
    cudaFree(c->f2_ctr);
    cudaFree(c->xs_sep);
    cudaFree(c->sr_buf);
    cudaFree(c->sr_tmp);
    if (c->fourier && c->fourier != c->real) cudaFree(c->fourier);
    cudaFree(c->real);
    cudaFree(c->phi);
    cudaFree(c->saved);
    cudaFree(c->force);
    cudaFree(c->sendbuf);
    cudaFree(c->fft_work);
    cudaFree(c->tab_x);
    cudaFree(c->tab_sin);
    cudaFree(c->d_scratch);
    cudaFree(c->d_counts);
    cudaFree(c->d_tilectr);
    cudaFree(c->d_comm_err);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    cudaFree(c->xchg_buf);
    if (c->pipe_ready) {
        cudaStreamDestroy(c->s_h2d);
        cudaStreamDestroy(c->s_d2h);
        for (cudaEvent_t& e : c->ev_pipe) cudaEventDestroy(e);
    }
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return PM_OK;
}

int pm_set_stream(pm_ctx* c, void* stream) {
    PM_REQUIRE(c != nullptr, "pm_set_stream: NULL context");
    PM_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    if (c->own_stream) cudaStreamDestroy(c->stream);
    c->own_stream = false;
    c->stream = reinterpret_cast<cudaStream_t>(stream);
    PM_CHECK_CUFFT(cufftSetStream(c->plan_fwd, c->stream));
    PM_CHECK_CUFFT(cufftSetStream(c->plan_bwd, c->stream));
    if (c->nranks > 1) PM_CHECK_CUFFT(cufftSetStream(c->plan_x, c->stream));
    if (c->plan2_ready) {      // the batched 2-D plans of the cuFFT + x-solve path
        PM_CHECK_CUFFT(cufftSetStream(c->plan2_fwd, c->stream));
        PM_CHECK_CUFFT(cufftSetStream(c->plan2_bwd, c->stream));
    }
    return PM_OK;
}

int pm_sync(pm_ctx* c) {
    PM_REQUIRE(c != nullptr, "pm_sync: NULL context");
    PM_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    return PM_OK;
}

int pm_local_shape(const pm_ctx* c, int64_t* nx_local, int64_t* x_start, int64_t* nj_local, int64_t* j_start) {
    PM_REQUIRE(c != nullptr, "pm_local_shape: NULL context");
    if (nx_local) *nx_local = c->g.nxl;
    if (x_start) *x_start = c->g.x0;
    if (nj_local) *nj_local = c->g.njl;
    if (j_start) *j_start = c->g.j0;
    return PM_OK;
}

int64_t pm_device_bytes(const pm_ctx* c) { return c ? c->bytes_allocated : 0; }

// ---- mesh operators ---------------------------------------------------------
int pm_grid_zero(pm_ctx* c) {
    PM_REQUIRE(c != nullptr, "pm_grid_zero: NULL context");
    // (after a fused solve on one rank the forward transform has already nullified the grid)
    PM_TRY(barrier_before_overwrite(c));     // a neighbour may still be filling its halo from this slab
    if (!c->real_is_zero) PM_CHECK_CUDA(cudaMemsetAsync(c->real, 0, c->real_elems * c->elem_size(), c->stream));
    c->real_is_zero = true;
    c->grid_in_phi = false;
    c->space_fourier = false;
    return PM_OK;
}

int pm_deposit(pm_ctx* c, const double* pos, int64_t n, int order, double contribution, const double* shift) {
    PM_REQUIRE(c != nullptr && (pos != nullptr || n == 0) && n >= 0, "pm_deposit: bad argument");
    PM_REQUIRE(!c->space_fourier, "pm_deposit: the slab holds Fourier data (call pm_grid_zero)");
    if (c->grid_in_phi && !c->real_is_zero) PM_TRY(ensure_in_real(c));
    c->grid_in_phi = false;       // the density grid is `real`
    if (n > 0) c->real_is_zero = false;
    return launch_deposit(c, pos, n, order, contribution, shift);
}

int pm_halo_add(pm_ctx* c) {
    PM_REQUIRE(c != nullptr, "pm_halo_add: NULL context");
    return halo_add(c);
}

int pm_halo_fill(pm_ctx* c) {
    PM_REQUIRE(c != nullptr, "pm_halo_fill: NULL context");
    return halo_fill(c, c->g.halo, c->g.halo, PM_TAP_REAL);
}

int pm_halo_fill_for(pm_ctx* c, int order, int diff_order, int interlace) {
    PM_REQUIRE(c != nullptr, "pm_halo_fill_for: NULL context");
    int lo, hi;
    halo_for_gather(order, diff_order, interlace, &lo, &hi);
    return halo_fill(c, lo, hi, PM_TAP_REAL);
}

int pm_fft_forward(pm_ctx* c) {
    PM_REQUIRE(c != nullptr, "pm_fft_forward: NULL context");
    return fft_forward(c);
}

int pm_fft_backward(pm_ctx* c) {
    PM_REQUIRE(c != nullptr, "pm_fft_backward: NULL context");
    return fft_backward(c);
}

int pm_kspace_potential(pm_ctx* c, double prefactor, int deconv_order, double gauss, double scale) {
    PM_REQUIRE(c != nullptr, "pm_kspace_potential: NULL context");
    return launch_kspace(c, prefactor, deconv_order, gauss, scale, nullptr, -1, false, prefactor != 0);
}

int pm_fourier_operate(pm_ctx* c, int deconv_order, const double* shift, double scale, int diff_dim,
                       int from_saved) {
    PM_REQUIRE(c != nullptr, "pm_fourier_operate: NULL context");
    return launch_kspace(c, 0, deconv_order, 0, scale, shift, diff_dim, from_saved != 0, false);
}

int pm_solve_fused(pm_ctx* c, double prefactor, int deconv_order, double gauss) {
    PM_REQUIRE(c != nullptr, "pm_solve_fused: NULL context");
    return solve_fused(c, prefactor, deconv_order, gauss);
}

int pm_solve_fused_stage(pm_ctx* c, double prefactor, int deconv_order, double gauss, int stage) {
    PM_REQUIRE(c != nullptr, "pm_solve_fused_stage: NULL context");
    return solve_fused(c, prefactor, deconv_order, gauss, stage);
}

int pm_fused_solve_available(const pm_ctx* c) { return c && (fft2_supported(c) || xsolve_supported(c)) ? 1 : 0; }

int pm_set_fused_solve(pm_ctx* c, int mode) {
    PM_REQUIRE(c != nullptr, "pm_set_fused_solve: NULL context");
    PM_REQUIRE(mode >= PM_SOLVE_UNFUSED && mode <= PM_SOLVE_FFT2_L2, "pm_set_fused_solve: mode = %d", mode);
    c->fused_solve = mode != PM_SOLVE_UNFUSED;
    c->solve_mode = mode;
    return PM_OK;
}

int pm_check_async_error(pm_ctx* c) {
    PM_REQUIRE(c != nullptr, "pm_check_async_error: NULL context");
    PM_TRY(fft2_check_error(c));
    int e[2] = {0, 0};
    PM_CHECK_CUDA(cudaMemcpyAsync(e, c->d_comm_err, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PM_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    if (e[0] != 0) {
        set_error("a rank did not arrive at a device barrier in time (results are invalid)");
        return PM_ERR_COMM;
    }
    if (e[1] != 0) {
        set_error("pm_deposit: a particle lies outside this rank's x-slab plus halo — its mass was not deposited "
                  "(particles must be distributed by x-slab: Component.set_particles / pm_exchange)");
        return PM_ERR_ARG;
    }
    return PM_OK;
}

int pm_power_k2(pm_ctx* c, int k2_max, double* power, unsigned long long* count) {
    PM_REQUIRE(c != nullptr, "pm_power_k2: NULL context");
    return launch_power_k2(c, k2_max, power, count);
}

int pm_slab_save(pm_ctx* c) { PM_REQUIRE(c != nullptr, "NULL context"); return slab_copy(c, 0); }
int pm_slab_accumulate(pm_ctx* c) { PM_REQUIRE(c != nullptr, "NULL context"); return slab_copy(c, 1); }
int pm_slab_restore(pm_ctx* c) { PM_REQUIRE(c != nullptr, "NULL context"); return slab_copy(c, 2); }

int pm_diff(pm_ctx* c, int dim, int order) {
    PM_REQUIRE(c != nullptr, "pm_diff: NULL context");
    PM_REQUIRE(!c->space_fourier, "pm_diff: the slab holds Fourier data");
    PM_TRY(launch_diff(c, dim, order));
    // the gather that follows reaches PCS (2 planes) + the interlacing shift (1 plane) beyond the slab
    if (c->nranks > 1) PM_TRY(halo_fill(c, std::min(3, c->g.nxl), std::min(3, c->g.nxl), PM_TAP_FORCE));
    return PM_OK;
}

int pm_gather(pm_ctx* c, int which, const double* pos, double* mom, int64_t n, int order, int dim,
              double factor, const double* shift) {
    PM_REQUIRE(c != nullptr && n >= 0, "pm_gather: bad argument");
    PM_REQUIRE(!c->space_fourier, "pm_gather: the slab holds Fourier data");
    return launch_gather(c, which, pos, mom, n, order, dim, factor, shift);
}

int pm_gather_kick(pm_ctx* c, const double* pos, double* mom, int64_t n, int order, int diff_order,
                   double factor, const double* shift, double* sum_mom2) {
    PM_REQUIRE(c != nullptr && n >= 0, "pm_gather_kick: bad argument");
    PM_REQUIRE(!c->space_fourier, "pm_gather_kick: the slab holds Fourier data (call pm_fft_backward)");
    return launch_gather_kick(c, const_cast<double*>(pos), mom, n, order, diff_order, factor, shift, sum_mom2);
}

int pm_gather_kick_drift(pm_ctx* c, double* pos, double* mom, int64_t n, int order, int diff_order,
                         double factor, const double* shift, double* sum_mom2, double dt_over_mass) {
    PM_REQUIRE(c != nullptr && n >= 0, "pm_gather_kick_drift: bad argument");
    PM_REQUIRE(!c->space_fourier, "pm_gather_kick_drift: the slab holds Fourier data (call pm_fft_backward)");
    return launch_gather_kick(c, pos, mom, n, order, diff_order, factor, shift, sum_mom2, dt_over_mass != 0, dt_over_mass);
}

// ---- particle operators -------------------------------------------------------
int pm_drift(pm_ctx* c, double* pos, const double* mom, int64_t n, double dt_over_mass) {
    PM_REQUIRE(c != nullptr && n >= 0, "pm_drift: bad argument");
    return launch_drift(c, pos, mom, n, dt_over_mass);
}

int pm_sum_mom2(pm_ctx* c, const double* mom, int64_t n, double* out) {
    PM_REQUIRE(c != nullptr && out != nullptr && n >= 0, "pm_sum_mom2: bad argument");
    return launch_sum_mom2(c, mom, n, out);
}

int pm_exchange(pm_ctx* c, double* pos, double* mom, int64_t* ids, int64_t* n_inout, int64_t capacity) {
    PM_REQUIRE(c != nullptr, "pm_exchange: NULL context");
    return exchange_particles(c, pos, mom, ids, nullptr, nullptr, nullptr, n_inout, capacity);
}

int pm_exchange_rungs(pm_ctx* c, double* pos, double* mom, int64_t* ids, double* dmom, signed char* rung,
                      signed char* rung_jumped, int64_t* n_inout, int64_t capacity) {
    PM_REQUIRE(c != nullptr, "pm_exchange_rungs: NULL context");
    return exchange_particles(c, pos, mom, ids, dmom, rung, rung_jumped, n_inout, capacity);
}

// ---- whole-path entry points -----------------------------------------------------
static const double kBccShift[3] = {-0.5, -0.5, -0.5};   // Lattice.shift_amount, mesh.py:85

// kick (+ drift of the same particles with the kicked momenta when dt_over_mass != 0)
static int kick_long_impl(pm_ctx* c, double* pos, double* mom, int64_t n, const pm_kick_params* p,
                          double* sum_mom2, double dt_over_mass) {
    const bool drift = dt_over_mass != 0;
    PM_REQUIRE(c != nullptr && p != nullptr && n >= 0, "pm_kick_long: bad argument");
    PM_REQUIRE(p->interlace == 0 || p->interlace == 1, "pm_kick_long: interlace = %d", p->interlace);
    const int nl = p->interlace ? 2 : 1;
    const double lscale = 1.0 / nl;
    if (c->fused_solve && nl == 1 && p->diff_order != 0 && pm_fused_solve_available(c)) {
        // default path: 2-D transforms + fused x pass (FFT · Green's function · inverse FFT)
        int hlo, hhi;
        halo_for_gather(p->order, p->diff_order, 0, &hlo, &hhi);
        PM_TRY(pm_grid_zero(c));
        PM_TRY(pm_deposit(c, pos, n, p->order, p->contribution, nullptr));
        PM_TRY(halo_add(c));
        PM_TRY(solve_fused(c, p->prefactor, p->deconv_order, p->gauss));
        PM_TRY(halo_fill(c, hlo, hhi, PM_TAP_REAL));
        return launch_gather_kick(c, pos, mom, n, p->order, p->diff_order, p->kick_factor, nullptr, sum_mom2,
                                  drift, dt_over_mass);
    }
    // upstream: interpolate_upstream(..., output_space='Fourier')  (mesh.py:492-616)
    for (int l = 0; l < nl; ++l) {
        const double* shift = l == 0 ? nullptr : kBccShift;
        PM_TRY(pm_grid_zero(c));
        PM_TRY(pm_deposit(c, pos, n, p->order, p->contribution, shift));
        PM_TRY(halo_add(c));
        PM_TRY(fft_forward(c));
        if (nl > 1) {
            // fourier_operate / copy_modes with the lattice: ×1/n_lattices and phase rotation
            PM_TRY(launch_kspace(c, 0, 0, 0, lscale, shift, -1, false, false));
            PM_TRY(slab_copy(c, l == 0 ? 0 : 1));
        }
    }
    if (nl > 1) PM_TRY(slab_copy(c, 2));
    // potential + promoted deconvolutions + Nyquist/origin nullification (interactions.py:2092-2118)
    PM_TRY(launch_kspace(c, p->prefactor, p->deconv_order, p->gauss, 1.0, nullptr, -1, false, true));
    const bool need_copy = nl > 1 || p->diff_order == 0;
    if (need_copy) PM_TRY(slab_copy(c, 0));
    int hlo, hhi;
    halo_for_gather(p->order, p->diff_order, p->interlace, &hlo, &hhi);
    // downstream (interactions.py:2214-2330)
    for (int l = 0; l < nl; ++l) {
        const double* shift = l == 0 ? nullptr : kBccShift;
        if (p->diff_order == 0) {
            for (int dim = 0; dim < 3; ++dim) {
                PM_TRY(launch_kspace(c, 0, 0, 0, lscale, shift, dim, true, false));
                PM_TRY(fft_backward(c));
                PM_TRY(halo_fill(c, hlo, hhi, PM_TAP_REAL));
                PM_TRY(launch_gather(c, PM_TAP_REAL, pos, mom, n, p->order, dim, p->kick_factor, shift));
            }
        } else {
            if (nl > 1) {
                PM_TRY(launch_kspace(c, 0, 0, 0, lscale, shift, -1, true, false));
            }
            PM_TRY(fft_backward(c));
            PM_TRY(halo_fill(c, hlo, hhi, PM_TAP_REAL));
            const bool last = l == nl - 1;
            PM_TRY(launch_gather_kick(c, pos, mom, n, p->order, p->diff_order, p->kick_factor, shift,
                                      last ? sum_mom2 : nullptr));
        }
    }
    if (p->diff_order == 0 && sum_mom2) PM_TRY(launch_sum_mom2(c, mom, n, sum_mom2));
    if (drift) PM_TRY(launch_drift(c, pos, mom, n, dt_over_mass));
    return PM_OK;
}

int pm_kick_long(pm_ctx* c, const double* pos, double* mom, int64_t n, const pm_kick_params* p,
                 double* sum_mom2) {
    return kick_long_impl(c, const_cast<double*>(pos), mom, n, p, sum_mom2, 0.0);
}

int pm_kick_drift(pm_ctx* c, double* pos, double* mom, int64_t n, const pm_kick_params* p,
                  double dt_over_mass, double* sum_mom2) {
    return kick_long_impl(c, pos, mom, n, p, sum_mom2, dt_over_mass);
}

// Host buffers in, host buffers out.  On the default path the transfers are pipelined against the kernels in
// kHostChunks chunks over two copy streams (PCIe is full duplex):
//   H2D pos[k] → deposit[k] …; H2D mom[k] runs under the solve; gather/kick/drift[k] → D2H mom[k], pos[k] …
// so the wall time approaches  H2D(pos) + solve + D2H(pos, mom)  instead of the sum of everything.
constexpr int kHostChunks = 8;

static int ensure_pipe(pm_ctx* c) {
    if (c->pipe_ready) return PM_OK;
    PM_CHECK_CUDA(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
    PM_CHECK_CUDA(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
    for (cudaEvent_t& e : c->ev_pipe) PM_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->pipe_ready = true;
    return PM_OK;
}

int pm_kick_long_host(pm_ctx* c, double* pos_host, double* mom_host, int64_t n, const pm_kick_params* p,
                      double dt_over_mass, double* sum_mom2_host) {
    PM_REQUIRE(c != nullptr && p != nullptr && n >= 0 && pos_host && mom_host, "pm_kick_long_host: bad argument");
    const size_t bytes = sizeof(double) * 3 * (size_t)n;
    // staging in the exchange buffer: pos | mom
    if (2 * bytes > c->xchg_bytes) {
        if (c->xchg_buf) { cudaFree(c->xchg_buf); c->bytes_allocated -= c->xchg_bytes; c->xchg_buf = nullptr; c->xchg_bytes = 0; }
        PM_CHECK_CUDA(cudaMalloc(&c->xchg_buf, 2 * bytes + 256));
        c->xchg_bytes = 2 * bytes + 256;
        c->bytes_allocated += c->xchg_bytes;
    }
    double* dpos = reinterpret_cast<double*>(c->xchg_buf);
    double* dmom = dpos + 3 * n;
    double* dsum = nullptr;
    if (sum_mom2_host) {
        dsum = c->d_scratch;
        PM_CHECK_CUDA(cudaMemsetAsync(dsum, 0, sizeof(double), c->stream));
    }
    const bool drift = dt_over_mass != 0;
    const bool pipelined = c->fused_solve && p->interlace == 0 && p->diff_order != 0 && pm_fused_solve_available(c) &&
                           n >= (int64_t)kHostChunks * 65536;
    if (!pipelined) {
        PM_CHECK_CUDA(cudaMemcpyAsync(dpos, pos_host, bytes, cudaMemcpyHostToDevice, c->stream));
        PM_CHECK_CUDA(cudaMemcpyAsync(dmom, mom_host, bytes, cudaMemcpyHostToDevice, c->stream));
        PM_TRY(kick_long_impl(c, dpos, dmom, n, p, dsum, dt_over_mass));
        if (drift) PM_CHECK_CUDA(cudaMemcpyAsync(pos_host, dpos, bytes, cudaMemcpyDeviceToHost, c->stream));
        PM_CHECK_CUDA(cudaMemcpyAsync(mom_host, dmom, bytes, cudaMemcpyDeviceToHost, c->stream));
    } else {
        PM_TRY(ensure_pipe(c));
        cudaEvent_t* ev_pos = c->ev_pipe;
        cudaEvent_t* ev_mom = c->ev_pipe + 16;
        cudaEvent_t* ev_out = c->ev_pipe + 32;
        const int64_t per = ((n + kHostChunks - 1) / kHostChunks + 2047) / 2048 * 2048;
        auto lo = [&](int k) { return std::min<int64_t>(n, (int64_t)k * per); };
        auto cnt = [&](int k) { return lo(k + 1) - lo(k); };
        // the copy streams start after whatever the context's stream has queued so far
        PM_CHECK_CUDA(cudaEventRecord(ev_out[kHostChunks], c->stream));
        PM_CHECK_CUDA(cudaStreamWaitEvent(c->s_h2d, ev_out[kHostChunks], 0));
        for (int k = 0; k < kHostChunks; ++k) {
            if (cnt(k) > 0)
                PM_CHECK_CUDA(cudaMemcpyAsync(dpos + 3 * lo(k), pos_host + 3 * lo(k), sizeof(double) * 3 * cnt(k),
                                              cudaMemcpyHostToDevice, c->s_h2d));
            PM_CHECK_CUDA(cudaEventRecord(ev_pos[k], c->s_h2d));
        }
        for (int k = 0; k < kHostChunks; ++k) {
            if (cnt(k) > 0)
                PM_CHECK_CUDA(cudaMemcpyAsync(dmom + 3 * lo(k), mom_host + 3 * lo(k), sizeof(double) * 3 * cnt(k),
                                              cudaMemcpyHostToDevice, c->s_h2d));
            PM_CHECK_CUDA(cudaEventRecord(ev_mom[k], c->s_h2d));
        }
        PM_TRY(pm_grid_zero(c));
        for (int k = 0; k < kHostChunks; ++k) {
            PM_CHECK_CUDA(cudaStreamWaitEvent(c->stream, ev_pos[k], 0));
            if (cnt(k) > 0) PM_TRY(pm_deposit(c, dpos + 3 * lo(k), cnt(k), p->order, p->contribution, nullptr));
        }
        int hlo, hhi;
        halo_for_gather(p->order, p->diff_order, 0, &hlo, &hhi);
        PM_TRY(halo_add(c));
        PM_TRY(solve_fused(c, p->prefactor, p->deconv_order, p->gauss));
        PM_TRY(halo_fill(c, hlo, hhi, PM_TAP_REAL));
        for (int k = 0; k < kHostChunks; ++k) {
            PM_CHECK_CUDA(cudaStreamWaitEvent(c->stream, ev_mom[k], 0));
            if (cnt(k) > 0)
                PM_TRY(launch_gather_kick(c, dpos + 3 * lo(k), dmom + 3 * lo(k), cnt(k), p->order, p->diff_order, p->kick_factor,
                                          nullptr, dsum, drift, dt_over_mass));
            PM_CHECK_CUDA(cudaEventRecord(ev_out[k], c->stream));
            PM_CHECK_CUDA(cudaStreamWaitEvent(c->s_d2h, ev_out[k], 0));
            if (cnt(k) > 0) {
                PM_CHECK_CUDA(cudaMemcpyAsync(mom_host + 3 * lo(k), dmom + 3 * lo(k), sizeof(double) * 3 * cnt(k),
                                              cudaMemcpyDeviceToHost, c->s_d2h));
                if (drift)
                    PM_CHECK_CUDA(cudaMemcpyAsync(pos_host + 3 * lo(k), dpos + 3 * lo(k), sizeof(double) * 3 * cnt(k),
                                                  cudaMemcpyDeviceToHost, c->s_d2h));
            }
        }
    }
    if (sum_mom2_host)
        PM_CHECK_CUDA(cudaMemcpyAsync(sum_mom2_host, dsum, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PM_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    if (pipelined) PM_CHECK_CUDA(cudaStreamSynchronize(c->s_d2h));
    return PM_OK;
}

// ---- parity taps ---------------------------------------------------------------
int64_t pm_tap_size(const pm_ctx* c, int which) {
    if (!c) return 0;
    if (which == PM_TAP_FOURIER) return (int64_t)c->fourier_elems * 2;
    return (int64_t)c->g.nxl * c->g.G * c->g.G;
}

int pm_get_grid(pm_ctx* c, int which, double* host_out) {
    PM_REQUIRE(c != nullptr && host_out != nullptr, "pm_get_grid: NULL argument");
    PM_REQUIRE(which >= PM_TAP_REAL && which <= PM_TAP_FORCE, "pm_get_grid: which = %d", which);
    if (which == PM_TAP_FORCE) PM_REQUIRE(c->force != nullptr, "pm_get_grid: no force grid (call pm_diff first)");
    PM_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    return c->dtype == PM_GRID_F64 ? get_grid_t<double>(c, which, host_out) : get_grid_t<float>(c, which, host_out);
}

int pm_set_grid(pm_ctx* c, const double* host_in) {
    PM_REQUIRE(c != nullptr && host_in != nullptr, "pm_set_grid: NULL argument");
    PM_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    c->space_fourier = false;
    return c->dtype == PM_GRID_F64 ? set_grid_t<double>(c, host_in) : set_grid_t<float>(c, host_in);
}

}  // extern "C"
