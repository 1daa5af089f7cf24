// pm_fourier.cu — FFT plans (cuFFT) and the fused k-space kernel.
//
// Reference semantics (file:line under the reference's src/):
//   transforms    fft.c:105-290, mesh.py:4012-4157: unnormalised r2c / c2r, in place, padded z
//   mode set      mesh.py:2615-2890 fourier_loop (Nyquist planes skipped; nullified mesh.py:3591-3622)
//   deconvolution mesh.py:2775-2856   [Π x_l/sin x_l]^D, x_l = k_l·π/G + ε
//   potential     interactions.py:2092-2118   ·(−L²G_N/π)/k² [·exp(−k²(2π r_s/L)²)], origin → 0
//   interlacing / Fourier differentiation   mesh.py:3327-3400 fourier_operate
//
// Layout of the Fourier slab: complex [i][j_local][kk], i ∈ [0,G), j = j0 + j_local,
// kk ∈ [0, G/2].  With one rank this is cuFFT's natural in-place r2c output; with several
// ranks it is what the all-to-all transpose delivers (pm_comm.cu) and the 1-D x transform runs
// on it with stride njl·Gc.  (FFTW-MPI's TRANSPOSED_OUT order [j_local][i][kk], fft.c:55-72,
// is the same data with the first two axes swapped; pm_get_grid documents the tap layout.)
#include "pm_internal.cuh"

#include <cmath>
#include <cstdlib>

namespace pm {

template <typename T> struct Cplx;
template <> struct Cplx<double> { using type = double2; };
template <> struct Cplx<float> { using type = float2; };

struct KParams {
    double prefactor;   // −L²·G_N/π (0: no potential factor)
    double gauss;       // (2π r_s/L)²
    double scale;       // 1/n_lattices
    double th[3];       // −2π/G·shift[d]
    double kfund;       // 2π/L
    int deconv_order;
    int rotate;         // any shift != 0
    int diff_dim;       // -1 none
    int potential;      // apply 1/k² (and nullify origin)
};

__device__ __forceinline__ double ipow(double f, int n) {
    double r = 1.0;
    while (n > 0) {
        if (n & 1) r *= f;
        f *= f;
        n >>= 1;
    }
    return r;
}

template <typename T>
__global__ void __launch_bounds__(256)
kspace_kernel(const typename Cplx<T>::type* __restrict__ src, typename Cplx<T>::type* __restrict__ dst,
              Geom g, KParams p, const double* __restrict__ tab_x, const double* __restrict__ tab_sin) {
    using C = typename Cplx<T>::type;
    const int nyq = g.G / 2;
    const int64_t total = (int64_t)g.G * g.njl * g.Gc;
    // flat index, kk fastest: consecutive threads touch consecutive 16-byte (8-byte) modes
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / g.Gc;
        const int kk = (int)(idx - row * g.Gc);
        const int i = (int)(row / g.njl);
        const int j = g.j0 + (int)(row - (int64_t)i * g.njl);
        C v = src[idx];
        double re = (double)v.x, im = (double)v.y;
        if (i == nyq || j == nyq || kk == nyq) {
            re = 0; im = 0;
        } else {
            const int ki = i - (i >= nyq ? g.G : 0);
            const int kj = j - (j >= nyq ? g.G : 0);
            double factor = 1;
            if (p.deconv_order) {
                // ((xi·xj)·xk)/((si·sj)·sk), then **D  (mesh.py:2795-2856)
                factor = ((tab_x[i] * tab_x[j]) * tab_x[kk]) / ((tab_sin[i] * tab_sin[j]) * tab_sin[kk]);
                factor = ipow(factor, p.deconv_order);
            }
            factor *= p.scale;
            if (p.rotate) {
                const double theta = (ki * p.th[0] + kj * p.th[1]) + kk * p.th[2];
                double sn, cs;
                sincos(theta, &sn, &cs);
                const double r2 = re * cs - im * sn;
                const double i2 = re * sn + im * cs;
                re = r2; im = i2;
            }
            if (p.diff_dim >= 0) {
                const int kl = p.diff_dim == 0 ? ki : (p.diff_dim == 1 ? kj : kk);
                factor *= p.kfund * kl;
                const double t = re;
                re = -im; im = t;
            }
            if (p.potential) {
                const int k2 = (kj * kj + ki * ki) + kk * kk;
                if (k2 == 0) {
                    factor = 0;
                } else if (p.gauss != 0) {
                    factor *= p.prefactor / k2 * exp(k2 * (-p.gauss));
                } else {
                    factor *= p.prefactor / k2;
                }
            }
            re *= factor;
            im *= factor;
        }
        v.x = (T)re; v.y = (T)im;
        dst[idx] = v;
    }
}

int launch_kspace(pm_ctx* c, double prefactor, int deconv_order, double gauss, double scale,
                  const double* shift, int diff_dim, bool from_saved, bool potential) {
    // reading from the saved copy overwrites the working slab, whatever it held
    PM_REQUIRE(from_saved || c->space_fourier, "k-space operation called while the slab holds real-space data");
    PM_REQUIRE(deconv_order >= 0 && deconv_order <= 64, "deconv_order = %d out of range", deconv_order);
    PM_REQUIRE(diff_dim >= -1 && diff_dim < 3, "fourier_operate called with diff_dim = %d not in {-1, 0, 1, 2}", diff_dim);
    if (from_saved) PM_REQUIRE(c->saved != nullptr, "pm_fourier_operate(from_saved) without pm_slab_save");
    KParams p;
    p.prefactor = prefactor;
    p.gauss = gauss;
    p.scale = scale;
    p.kfund = 2 * M_PI / c->boxsize;
    p.deconv_order = deconv_order;
    p.diff_dim = diff_dim;
    p.potential = potential ? 1 : 0;
    p.rotate = 0;
    for (int d = 0; d < 3; ++d) {
        const double s = shift ? shift[d] : 0.0;
        p.th[d] = -2 * M_PI / c->g.G * s;
        if (s != 0.0) p.rotate = 1;
    }
    const int threads = 256;
    const int grid = kNumSMs * 8;
    const void* src = from_saved ? c->saved : c->fourier;
    if (c->dtype == PM_GRID_F64) {
        PM_LAUNCH((kspace_kernel<double>), grid, threads, 0, c->stream,
                  reinterpret_cast<const double2*>(src), reinterpret_cast<double2*>(c->fourier), c->g, p,
                  c->tab_x, c->tab_sin);
    } else {
        PM_LAUNCH((kspace_kernel<float>), grid, threads, 0, c->stream,
                  reinterpret_cast<const float2*>(src), reinterpret_cast<float2*>(c->fourier), c->g, p,
                  c->tab_x, c->tab_sin);
    }
    c->space_fourier = true;
    // the working slab (the interior of `real`) now holds this Fourier data whatever the context held before —
    // in particular a potential that a fused solve had left in `phi` is no longer the current grid
    c->grid_in_phi = false;
    c->real_is_zero = false;
    return PM_OK;
}

// ---------------------------------------------------------------------------
// power per integer k² (compute_powerspec, analysis.py:500-547)
// ---------------------------------------------------------------------------
// Visits the modes of fourier_loop(gridsize, sparse=True, skip_origin=True, k2_max) (mesh.py:2615-2890):
// Nyquist planes and the origin excluded; on the kk = 0 plane only one point of every complex-conjugate
// pair (ki > 0 and (ki = 0, kj > 0) skipped, mesh.py:2813-2826); k² ≤ k2_max.  power[k²] += re² + im²,
// count[k²] += 1 (the multiplicity get_powerspec_bins tallies on the host, analysis.py:303-312).
template <typename T>
__global__ void __launch_bounds__(256)
power_k2_kernel(const typename Cplx<T>::type* __restrict__ slab, Geom g, int k2_max,
                double* __restrict__ power, unsigned long long* __restrict__ count) {
    const int nyq = g.G / 2;
    const int64_t total = (int64_t)g.G * g.njl * g.Gc;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / g.Gc;
        const int kk = (int)(idx - row * g.Gc);
        const int i = (int)(row / g.njl);
        const int j = g.j0 + (int)(row - (int64_t)i * g.njl);
        if (i == nyq || j == nyq || kk == nyq) continue;
        const int ki = i - (i >= nyq ? g.G : 0);
        const int kj = j - (j >= nyq ? g.G : 0);
        if (kk == 0 && (ki > 0 || (ki == 0 && kj >= 0))) continue;   // conjugate pairs and the origin
        const int k2 = (kj * kj + ki * ki) + kk * kk;
        if (k2 > k2_max) continue;
        const typename Cplx<T>::type v = slab[idx];
        const double re = (double)v.x, im = (double)v.y;
        atomicAdd(power + k2, re * re + im * im);
        if (count != nullptr) atomicAdd(count + k2, 1ULL);
    }
}

int launch_power_k2(pm_ctx* c, int k2_max, double* power, unsigned long long* count) {
    PM_REQUIRE(c->space_fourier, "pm_power_k2: the slab holds real-space data (call pm_fft_forward)");
    PM_REQUIRE(k2_max >= 1 && power != nullptr, "pm_power_k2: bad argument");
    const int grid = kNumSMs * 8;
    if (c->dtype == PM_GRID_F64) {
        PM_LAUNCH((power_k2_kernel<double>), grid, 256, 0, c->stream, reinterpret_cast<const double2*>(c->fourier), c->g,
                  k2_max, power, count);
    } else {
        PM_LAUNCH((power_k2_kernel<float>), grid, 256, 0, c->stream, reinterpret_cast<const float2*>(c->fourier), c->g,
                  k2_max, power, count);
    }
    return PM_OK;
}

// ---------------------------------------------------------------------------
// slab save / accumulate / restore
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) slab_add_kernel(T* __restrict__ dst, const T* __restrict__ src, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] += src[i];
}

int ensure_saved(pm_ctx* c) {
    if (c->saved) return PM_OK;
    const size_t bytes = c->fourier_elems * 2 * c->elem_size();
    PM_CHECK_CUDA(cudaMalloc(&c->saved, bytes));
    c->bytes_allocated += bytes;
    return PM_OK;
}

int slab_copy(pm_ctx* c, int mode) {
    PM_TRY(ensure_saved(c));
    const size_t bytes = c->fourier_elems * 2 * c->elem_size();
    if (mode == 0) {
        PM_CHECK_CUDA(cudaMemcpyAsync(c->saved, c->fourier, bytes, cudaMemcpyDeviceToDevice, c->stream));
    } else if (mode == 2) {
        PM_CHECK_CUDA(cudaMemcpyAsync(c->fourier, c->saved, bytes, cudaMemcpyDeviceToDevice, c->stream));
    } else {
        const size_t n = c->fourier_elems * 2;
        if (c->dtype == PM_GRID_F64) {
            PM_LAUNCH((slab_add_kernel<double>), kNumSMs * 8, 256, 0, c->stream,
                      reinterpret_cast<double*>(c->saved), reinterpret_cast<const double*>(c->fourier), n);
        } else {
            PM_LAUNCH((slab_add_kernel<float>), kNumSMs * 8, 256, 0, c->stream,
                      reinterpret_cast<float*>(c->saved), reinterpret_cast<const float*>(c->fourier), n);
        }
    }
    return PM_OK;
}

// ---------------------------------------------------------------------------
// cuFFT plans
// ---------------------------------------------------------------------------
int make_plans(pm_ctx* c) {
    const Geom& g = c->g;
    const bool f64 = c->dtype == PM_GRID_F64;
    size_t ws_f = 0, ws_b = 0, ws_x = 0;
    PM_CHECK_CUFFT(cufftCreate(&c->plan_fwd));
    PM_CHECK_CUFFT(cufftCreate(&c->plan_bwd));
    PM_CHECK_CUFFT(cufftSetAutoAllocation(c->plan_fwd, 0));
    PM_CHECK_CUFFT(cufftSetAutoAllocation(c->plan_bwd, 0));
    if (c->nranks == 1) {
        PM_CHECK_CUFFT(cufftMakePlan3d(c->plan_fwd, g.G, g.G, g.G, f64 ? CUFFT_D2Z : CUFFT_R2C, &ws_f));
        PM_CHECK_CUFFT(cufftMakePlan3d(c->plan_bwd, g.G, g.G, g.G, f64 ? CUFFT_Z2D : CUFFT_C2R, &ws_b));
    } else {
        long long n2[2] = {g.G, g.G};
        long long rembed[2] = {g.G, g.Gp};
        long long cembed[2] = {g.G, g.Gc};
        PM_CHECK_CUFFT(cufftMakePlanMany64(c->plan_fwd, 2, n2, rembed, 1, (long long)g.G * g.Gp, cembed, 1,
                                           (long long)g.G * g.Gc, f64 ? CUFFT_D2Z : CUFFT_R2C, g.nxl, &ws_f));
        PM_CHECK_CUFFT(cufftMakePlanMany64(c->plan_bwd, 2, n2, cembed, 1, (long long)g.G * g.Gc, rembed, 1,
                                           (long long)g.G * g.Gp, f64 ? CUFFT_Z2D : CUFFT_C2R, g.nxl, &ws_b));
        PM_CHECK_CUFFT(cufftCreate(&c->plan_x));
        PM_CHECK_CUFFT(cufftSetAutoAllocation(c->plan_x, 0));
        long long n1[1] = {g.G};
        long long e1[1] = {g.G};
        const long long stride = (long long)g.njl * g.Gc;
        PM_CHECK_CUFFT(cufftMakePlanMany64(c->plan_x, 1, n1, e1, stride, 1, e1, stride, 1,
                                           f64 ? CUFFT_Z2Z : CUFFT_C2C, stride, &ws_x));
    }
    size_t ws2_f = 0, ws2_b = 0;
    c->plan2_ready = false;
    if (c->xs_tw != nullptr) {
        // batched 2-D (y,z) plans; the x direction is handled by the fused x-solve kernel.  PM_FFT_CHUNK
        // (experiment knob) executes them over chunks of x planes.
        int chunk = g.nxl;   // measured on B200 (r01): chunking (8…64 planes) is slower than whole-slab passes
        if (const char* e = getenv("PM_FFT_CHUNK")) chunk = atoi(e);
        if (chunk <= 0 || chunk > g.nxl) chunk = g.nxl;
        while (g.nxl % chunk) --chunk;
        c->fft_chunk = chunk;
        long long n2[2] = {g.G, g.G};
        long long rembed[2] = {g.G, g.Gp};
        long long cembed[2] = {g.G, g.Gc};
        PM_CHECK_CUFFT(cufftCreate(&c->plan2_fwd));
        PM_CHECK_CUFFT(cufftCreate(&c->plan2_bwd));
        PM_CHECK_CUFFT(cufftSetAutoAllocation(c->plan2_fwd, 0));
        PM_CHECK_CUFFT(cufftSetAutoAllocation(c->plan2_bwd, 0));
        PM_CHECK_CUFFT(cufftMakePlanMany64(c->plan2_fwd, 2, n2, rembed, 1, (long long)g.G * g.Gp, cembed, 1,
                                           (long long)g.G * g.Gc, f64 ? CUFFT_D2Z : CUFFT_R2C, chunk, &ws2_f));
        PM_CHECK_CUFFT(cufftMakePlanMany64(c->plan2_bwd, 2, n2, cembed, 1, (long long)g.G * g.Gc, rembed, 1,
                                           (long long)g.G * g.Gp, f64 ? CUFFT_Z2D : CUFFT_C2R, chunk, &ws2_b));
        c->plan2_ready = true;
    }
    size_t ws = ws_f > ws_b ? ws_f : ws_b;
    if (ws_x > ws) ws = ws_x;
    if (ws2_f > ws) ws = ws2_f;
    if (ws2_b > ws) ws = ws2_b;
    c->fft_work_bytes = ws;
    if (ws) {
        PM_CHECK_CUDA(cudaMalloc(&c->fft_work, ws));
        c->bytes_allocated += ws;
    }
    PM_CHECK_CUFFT(cufftSetWorkArea(c->plan_fwd, c->fft_work));
    PM_CHECK_CUFFT(cufftSetWorkArea(c->plan_bwd, c->fft_work));
    PM_CHECK_CUFFT(cufftSetStream(c->plan_fwd, c->stream));
    PM_CHECK_CUFFT(cufftSetStream(c->plan_bwd, c->stream));
    if (c->nranks > 1) {
        PM_CHECK_CUFFT(cufftSetWorkArea(c->plan_x, c->fft_work));
        PM_CHECK_CUFFT(cufftSetStream(c->plan_x, c->stream));
    }
    if (c->plan2_ready) {
        PM_CHECK_CUFFT(cufftSetWorkArea(c->plan2_fwd, c->fft_work));
        PM_CHECK_CUFFT(cufftSetWorkArea(c->plan2_bwd, c->fft_work));
        PM_CHECK_CUFFT(cufftSetStream(c->plan2_fwd, c->stream));
        PM_CHECK_CUFFT(cufftSetStream(c->plan2_bwd, c->stream));
    }
    c->plans_ready = true;
    return PM_OK;
}

void destroy_plans(pm_ctx* c) {
    if (!c->plans_ready) return;
    cufftDestroy(c->plan_fwd);
    cufftDestroy(c->plan_bwd);
    if (c->nranks > 1) cufftDestroy(c->plan_x);
    if (c->plan2_ready) { cufftDestroy(c->plan2_fwd); cufftDestroy(c->plan2_bwd); c->plan2_ready = false; }
    c->plans_ready = false;
}

int fft_forward(pm_ctx* c) {
    PM_REQUIRE(!c->space_fourier, "pm_fft_forward: slab already holds Fourier data");
    PM_TRY(ensure_in_real(c));
    c->real_is_zero = false;
    const bool f64 = c->dtype == PM_GRID_F64;
    if (f64) {
        double* r = c->real_interior<double>();
        PM_CHECK_CUFFT(cufftExecD2Z(c->plan_fwd, r, reinterpret_cast<cufftDoubleComplex*>(r)));
    } else {
        float* r = c->real_interior<float>();
        PM_CHECK_CUFFT(cufftExecR2C(c->plan_fwd, r, reinterpret_cast<cufftComplex*>(r)));
    }
    if (c->nranks > 1) {
        PM_TRY(transpose_forward(c));
        if (f64) {
            auto* f = reinterpret_cast<cufftDoubleComplex*>(c->fourier);
            PM_CHECK_CUFFT(cufftExecZ2Z(c->plan_x, f, f, CUFFT_FORWARD));
        } else {
            auto* f = reinterpret_cast<cufftComplex*>(c->fourier);
            PM_CHECK_CUFFT(cufftExecC2C(c->plan_x, f, f, CUFFT_FORWARD));
        }
    }
    c->space_fourier = true;
    return PM_OK;
}

int fft_backward(pm_ctx* c) {
    PM_REQUIRE(c->space_fourier, "pm_fft_backward: slab holds real-space data");
    const bool f64 = c->dtype == PM_GRID_F64;
    if (c->nranks > 1) {
        if (f64) {
            auto* f = reinterpret_cast<cufftDoubleComplex*>(c->fourier);
            PM_CHECK_CUFFT(cufftExecZ2Z(c->plan_x, f, f, CUFFT_INVERSE));
        } else {
            auto* f = reinterpret_cast<cufftComplex*>(c->fourier);
            PM_CHECK_CUFFT(cufftExecC2C(c->plan_x, f, f, CUFFT_INVERSE));
        }
        PM_TRY(transpose_backward(c));
    }
    if (f64) {
        double* r = c->real_interior<double>();
        PM_CHECK_CUFFT(cufftExecZ2D(c->plan_bwd, reinterpret_cast<cufftDoubleComplex*>(r), r));
    } else {
        float* r = c->real_interior<float>();
        PM_CHECK_CUFFT(cufftExecC2R(c->plan_bwd, reinterpret_cast<cufftComplex*>(r), r));
    }
    c->space_fourier = false;
    return PM_OK;
}

// forward 2-D transforms → fused x pass (FFT · Green · inverse FFT) → inverse 2-D transforms.
// Equivalent to pm_fft_forward + pm_kspace_potential + pm_fft_backward; the slab ends in real space.
int solve_fused(pm_ctx* c, double prefactor, int deconv_order, double gauss, int stage) {
    PM_REQUIRE(!c->space_fourier, "pm_solve_fused: the slab holds Fourier data");
    PM_REQUIRE(stage >= 0 && stage <= 3, "pm_solve_fused_stage: stage = %d", stage);
    int mode = c->solve_mode;
    if (mode == PM_SOLVE_AUTO || mode == PM_SOLVE_UNFUSED)
        mode = fft2_supported(c) ? PM_SOLVE_FFT2_L2 : PM_SOLVE_CUFFT2D_XSOLVE;
    if (mode == PM_SOLVE_FFT2_L2 || mode == PM_SOLVE_FFT2_SPLIT)
        return solve_fft2(c, prefactor, deconv_order, gauss, mode == PM_SOLVE_FFT2_L2, stage);
    PM_REQUIRE(stage == 0, "pm_solve_fused_stage: staged execution needs the hand-written transforms");
    PM_REQUIRE(xsolve_supported(c), "pm_solve_fused: not available for this grid size / rank layout");
    PM_TRY(ensure_in_real(c));
    c->real_is_zero = false;
    const bool f64 = c->dtype == PM_GRID_F64;
    const size_t plane = (size_t)c->g.G * c->g.Gp;
    const int nchunks = c->g.nxl / c->fft_chunk;
    if (f64) {
        double* r = c->real_interior<double>();
        for (int k = 0; k < nchunks; ++k) {
            double* q = r + (size_t)k * c->fft_chunk * plane;
            PM_CHECK_CUFFT(cufftExecD2Z(c->plan2_fwd, q, reinterpret_cast<cufftDoubleComplex*>(q)));
        }
        PM_TRY(xsolve(c, prefactor, deconv_order, gauss));
        for (int k = 0; k < nchunks; ++k) {
            double* q = r + (size_t)k * c->fft_chunk * plane;
            PM_CHECK_CUFFT(cufftExecZ2D(c->plan2_bwd, reinterpret_cast<cufftDoubleComplex*>(q), q));
        }
    } else {
        float* r = c->real_interior<float>();
        for (int k = 0; k < nchunks; ++k) {
            float* q = r + (size_t)k * c->fft_chunk * plane;
            PM_CHECK_CUFFT(cufftExecR2C(c->plan2_fwd, q, reinterpret_cast<cufftComplex*>(q)));
        }
        PM_TRY(xsolve(c, prefactor, deconv_order, gauss));
        for (int k = 0; k < nchunks; ++k) {
            float* q = r + (size_t)k * c->fft_chunk * plane;
            PM_CHECK_CUFFT(cufftExecC2R(c->plan2_bwd, reinterpret_cast<cufftComplex*>(q), q));
        }
    }
    return PM_OK;
}

}  // namespace pm
