/* pm_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C (OpenMP) restatement of CO*N*CEPT's PM long-range kick + drift, loop for loop, used
 *   (1) as a second checker beside oracle/pm_oracle.py, and
 *   (2) as the timed CPU baseline ("port") of bench.py — the compiled reference itself cannot be
 *       built here (needs mpicc, FFTW-MPI, GSL).
 * The product (concept_b200/) never links or calls this file.
 *
 * Parity status: PINNED through tests/test_oracle_c.py, which checks every entry point against
 * the golden vectors produced by the reference itself (the .npz files under tests/golden).
 *
 * Reference loops restated (file:line under the reference's src/):
 *   deposit            mesh.py:1512-1636, weights :5305-5379, loop :5138-5155
 *   Nyquist/origin     mesh.py:3585-3622
 *   k-space factor     mesh.py:2775-2856; interactions.py:2092-2118
 *   finite difference  mesh.py:4874-5030  (three separate force grids, like the reference)
 *   gather + kick      mesh.py:376-459; interactions.py:2384-2387 (one pass per dimension)
 *   drift              species.py:2179-2199; mod commons.py:5102-5131
 *   Σ mom²             analysis.py:3965-3972
 * The FFT is FFTW 3.3.10 in the reference (fft.c; not vendored, absent here): restated as a
 * textbook unnormalised DFT (radix-2 for powers of two, direct O(n²) otherwise), natural
 * [i][j][kk] layout.  Domain decomposition is emulated with OpenMP threads owning x-slabs, the
 * same way the reference's MPI ranks own domains: no atomics, boundary planes coloured even/odd.
 *
 * Build: gcc -O3 -ffast-math -funroll-loops -fopenmp -shared -fPIC (flags of src/Makefile:175-183)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#else
static int omp_get_max_threads(void) { return 1; }
static int omp_get_thread_num(void) { return 0; }
static void omp_set_num_threads(int n) { (void)n; }
#endif

#define EPS 2.220446049250313e-16
#define NGHOSTS_REF 2

typedef struct { double re, im; } cplx;

int pmo_num_threads(void) { return omp_get_max_threads(); }
/* torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU baseline sets its thread count explicitly */
void pmo_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

/* ---------------------------------------------------------------- weights */
static inline int64_t set_weights(int order, double x, double* w) {
    int64_t index;
    if (order == 1) {
        index = (int64_t)(x + 0.5);
        w[0] = 1;
    } else if (order == 2) {
        index = (int64_t)x;
        double dist = x - index;
        w[0] = 1 - dist;
        w[1] = dist;
    } else if (order == 3) {
        index = (int64_t)(x + 0.5);
        double dist = x - index;
        index -= 1;
        double dist2 = dist*dist;
        double w0 = 0.125 + 0.5*(dist2 - dist);
        double w1 = 0.75 - dist2;
        w[0] = w0; w[1] = w1; w[2] = 1 - w0 - w1;
    } else {
        index = (int64_t)x;
        index -= 1;
        double dist = x - index;
        double tmp = 2 - dist, tmp2 = tmp*tmp, tmp3 = tmp*tmp2;
        double w0 = 1./6.*tmp3;
        double w2 = 2./3. - tmp2 + 0.5*tmp3;
        double w3 = 1./6.*((dist - 1)*(dist - 1)*(dist - 1));
        w[0] = w0; w[1] = 1 - w0 - w2 - w3; w[2] = w2; w[3] = w3;
    }
    return index - NGHOSTS_REF;
}

static inline int64_t wrapi(int64_t i, int64_t G) {
    if (i < 0) return i + G;
    if (i >= G) return i - G;
    return i;
}

static void coord_setup(double L, int G, const double* shift, int for_gather, double* off, double* scale) {
    double cellsize = L/G;
    double sgn = for_gather ? 1.0 : -1.0;
    for (int d = 0; d < 3; ++d) {
        double s = shift ? shift[d] : 0.0;
        off[d] = 0.0 - (1 + EPS)*(NGHOSTS_REF - 0.5 + sgn*s)*cellsize;
    }
    *scale = (1/cellsize)*(1 - EPS);
}

/* ---------------------------------------------------------------- deposit */
/* rho[G][G][G] += ; threads own x-slabs of base cells (like MPI domains); slabs are processed in
 * two colours so that the `order`-wide footprints of concurrently running threads never overlap. */
int pmo_deposit(const double* pos, int64_t n, int G, double L, int order, double contribution,
                const double* shift, double* rho) {
    if (order < 1 || order > 4) return -1;
    double off[3], scale;
    coord_setup(L, G, shift, 0, off, &scale);
    int nth = omp_get_max_threads();
    int nslab = 2*nth;
    while (nslab > 1 && G/nslab < 4) nslab--;
    if (nslab < 2 || G/nslab < 4) nslab = 1;
    if (nslab > 1 && (nslab & 1)) nslab--;   /* even count: the periodic wrap pairs the last slab with slab 0 */
    /* bucket particles by slab of their base cell (counting sort of indices) */
    int64_t* count = (int64_t*)calloc((size_t)nslab + 1, sizeof(int64_t));
    int32_t* slab_of = (int32_t*)malloc(sizeof(int32_t)*(size_t)(n > 0 ? n : 1));
    int64_t* perm = (int64_t*)malloc(sizeof(int64_t)*(size_t)(n > 0 ? n : 1));
    for (int64_t p = 0; p < n; ++p) {
        double w[4];
        int64_t ix = wrapi(set_weights(order, (pos[3*p] - off[0])*scale, w), G);
        int s = (int)(ix*nslab/G);
        slab_of[p] = s;
        count[s + 1]++;
    }
    for (int s = 0; s < nslab; ++s) count[s + 1] += count[s];
    int64_t* cursor = (int64_t*)malloc(sizeof(int64_t)*(size_t)nslab);
    memcpy(cursor, count, sizeof(int64_t)*(size_t)nslab);
    for (int64_t p = 0; p < n; ++p) perm[cursor[slab_of[p]]++] = p;
    for (int colour = 0; colour < 2; ++colour) {
        if (nslab == 1 && colour == 1) break;
#pragma omp parallel for schedule(dynamic, 1)
        for (int s = colour; s < nslab; s += 2) {
            for (int64_t q = count[s]; q < count[s + 1]; ++q) {
                int64_t p = perm[q];
                double wx[4], wy[4], wz[4];
                int64_t ix = set_weights(order, (pos[3*p + 0] - off[0])*scale, wx);
                int64_t iy = set_weights(order, (pos[3*p + 1] - off[1])*scale, wy);
                int64_t iz = set_weights(order, (pos[3*p + 2] - off[2])*scale, wz);
                for (int a = 0; a < order; ++a) {
                    double wa = wx[a]*contribution;
                    int64_t ia = wrapi(ix + a, G);
                    for (int b = 0; b < order; ++b) {
                        double wab = wa*wy[b];
                        double* row = rho + (ia*G + wrapi(iy + b, G))*G;
                        for (int c = 0; c < order; ++c) row[wrapi(iz + c, G)] += wab*wz[c];
                    }
                }
            }
        }
    }
    free(count); free(slab_of); free(perm); free(cursor);
    return 0;
}

/* ---------------------------------------------------------------- FFT */
static int is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

typedef struct { int n; cplx* tw; int pow2; } fftplan;

static void plan_make(fftplan* p, int n) {
    p->n = n;
    p->pow2 = is_pow2(n);
    p->tw = (cplx*)malloc(sizeof(cplx)*(size_t)n);
    for (int k = 0; k < n; ++k) {
        double a = -2*M_PI*k/n;
        p->tw[k].re = cos(a);
        p->tw[k].im = sin(a);
    }
}

/* in-place forward (sign = -1) or backward (sign = +1) unnormalised DFT of length n */
static void fft_line(const fftplan* p, cplx* x, cplx* tmp, int sign) {
    int n = p->n;
    if (!p->pow2) {
        for (int k = 0; k < n; ++k) {
            double re = 0, im = 0;
            for (int j = 0; j < n; ++j) {
                int t = (int)(((int64_t)j*k) % n);
                double c = p->tw[t].re, s = sign < 0 ? p->tw[t].im : -p->tw[t].im;
                re += x[j].re*c - x[j].im*s;
                im += x[j].re*s + x[j].im*c;
            }
            tmp[k].re = re; tmp[k].im = im;
        }
        memcpy(x, tmp, sizeof(cplx)*(size_t)n);
        return;
    }
    /* bit reversal */
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { cplx t = x[i]; x[i] = x[j]; x[j] = t; }
    }
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n/len;
        for (int i = 0; i < n; i += len) {
            for (int j = 0; j < half; ++j) {
                double c = p->tw[j*step].re, s = sign < 0 ? p->tw[j*step].im : -p->tw[j*step].im;
                cplx u = x[i + j], v = x[i + j + half];
                double vr = v.re*c - v.im*s, vi = v.re*s + v.im*c;
                x[i + j].re = u.re + vr; x[i + j].im = u.im + vi;
                x[i + j + half].re = u.re - vr; x[i + j + half].im = u.im - vi;
            }
        }
    }
}

/* rho[G][G][G] real -> slab[G][G][Gc] complex, unnormalised (np.fft.rfftn) */
int pmo_fft_forward(const double* rho, int G, cplx* slab) {
    int Gc = G/2 + 1;
    fftplan pl; plan_make(&pl, G);
#pragma omp parallel
    {
        cplx* line = (cplx*)malloc(sizeof(cplx)*(size_t)G*2);
        cplx* tmp = line + G;
#pragma omp for schedule(static)
        for (int64_t r = 0; r < (int64_t)G*G; ++r) {          /* z */
            for (int k = 0; k < G; ++k) { line[k].re = rho[r*G + k]; line[k].im = 0; }
            fft_line(&pl, line, tmp, -1);
            memcpy(slab + r*Gc, line, sizeof(cplx)*(size_t)Gc);
        }
#pragma omp for schedule(static)
        for (int64_t r = 0; r < (int64_t)G*Gc; ++r) {         /* y */
            int64_t i = r/Gc, kk = r % Gc;
            for (int j = 0; j < G; ++j) line[j] = slab[(i*G + j)*Gc + kk];
            fft_line(&pl, line, tmp, -1);
            for (int j = 0; j < G; ++j) slab[(i*G + j)*Gc + kk] = line[j];
        }
#pragma omp for schedule(static)
        for (int64_t r = 0; r < (int64_t)G*Gc; ++r) {         /* x */
            int64_t j = r/Gc, kk = r % Gc;
            for (int i = 0; i < G; ++i) line[i] = slab[((int64_t)i*G + j)*Gc + kk];
            fft_line(&pl, line, tmp, -1);
            for (int i = 0; i < G; ++i) slab[((int64_t)i*G + j)*Gc + kk] = line[i];
        }
        free(line);
    }
    free(pl.tw);
    return 0;
}

/* slab[G][G][Gc] (destroyed) -> phi[G][G][G], unnormalised (irfftn norm='forward') */
int pmo_fft_backward(cplx* slab, int G, double* phi) {
    int Gc = G/2 + 1;
    fftplan pl; plan_make(&pl, G);
#pragma omp parallel
    {
        cplx* line = (cplx*)malloc(sizeof(cplx)*(size_t)G*2);
        cplx* tmp = line + G;
#pragma omp for schedule(static)
        for (int64_t r = 0; r < (int64_t)G*Gc; ++r) {         /* x */
            int64_t j = r/Gc, kk = r % Gc;
            for (int i = 0; i < G; ++i) line[i] = slab[((int64_t)i*G + j)*Gc + kk];
            fft_line(&pl, line, tmp, +1);
            for (int i = 0; i < G; ++i) slab[((int64_t)i*G + j)*Gc + kk] = line[i];
        }
#pragma omp for schedule(static)
        for (int64_t r = 0; r < (int64_t)G*Gc; ++r) {         /* y */
            int64_t i = r/Gc, kk = r % Gc;
            for (int j = 0; j < G; ++j) line[j] = slab[(i*G + j)*Gc + kk];
            fft_line(&pl, line, tmp, +1);
            for (int j = 0; j < G; ++j) slab[(i*G + j)*Gc + kk] = line[j];
        }
#pragma omp for schedule(static)
        for (int64_t r = 0; r < (int64_t)G*G; ++r) {          /* z: Hermitian completion */
            const cplx* h = slab + r*Gc;
            for (int k = 0; k < Gc; ++k) line[k] = h[k];
            line[0].im = 0;
            if ((G & 1) == 0) line[G/2].im = 0;
            for (int k = Gc; k < G; ++k) { line[k].re = h[G - k].re; line[k].im = -h[G - k].im; }
            fft_line(&pl, line, tmp, +1);
            for (int k = 0; k < G; ++k) phi[r*G + k] = line[k].re;
        }
        free(line);
    }
    free(pl.tw);
    return 0;
}

/* ---------------------------------------------------------------- k-space */
static inline double ipow_d(double f, int n) {
    double r = 1;
    while (n > 0) { if (n & 1) r *= f; f *= f; n >>= 1; }
    return r;
}

/* Potential loop of particle_mesh with Nyquist and origin nullification (natural layout) */
int pmo_kspace_potential(cplx* slab, int G, double prefactor, int deconv_order, double gauss, double scale) {
    int Gc = G/2 + 1, nyq = G/2;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < G; ++i) {
        int ki = i - (i >= nyq ? G : 0);
        double xi = ki*(M_PI/G) + EPS, si = sin(xi);
        for (int j = 0; j < G; ++j) {
            int kj = j - (j >= nyq ? G : 0);
            double xj = kj*(M_PI/G) + EPS, sj = sin(xj);
            double num_ij = xi*xj, den_ij = si*sj;
            cplx* row = slab + ((int64_t)i*G + j)*Gc;
            for (int kk = 0; kk < Gc; ++kk) {
                if (i == nyq || j == nyq || kk == nyq) { row[kk].re = 0; row[kk].im = 0; continue; }
                double factor = 1;
                if (deconv_order) {
                    double xk = kk*(M_PI/G) + EPS;
                    factor = (num_ij*xk)/(den_ij*sin(xk));
                    factor = ipow_d(factor, deconv_order);
                }
                factor *= scale;
                int64_t k2 = ((int64_t)kj*kj + (int64_t)ki*ki) + (int64_t)kk*kk;
                if (prefactor != 0) {
                    if (k2 == 0) factor = 0;
                    else if (gauss != 0) factor *= prefactor/k2*exp(k2*(-gauss));
                    else factor *= prefactor/k2;
                }
                row[kk].re *= factor;
                row[kk].im *= factor;
            }
        }
    }
    return 0;
}

/* ---------------------------------------------------------------- finite differences */
int pmo_diff(const double* phi, int G, int dim, int order, double dx, double* out) {
    double c[4] = {0, 0, 0, 0};
    int reach, forward = 0;
    switch (order) {
        case 1: reach = 1; forward = 1; c[0] = 1/dx; break;
        case 2: reach = 1; c[0] = (1./2)/dx; break;
        case 4: reach = 2; c[0] = (2./3)/dx; c[1] = (1./12)/dx; break;
        case 6: reach = 3; c[0] = (3./4)/dx; c[1] = (3./20)/dx; c[2] = (1./60)/dx; break;
        case 8: reach = 4; c[0] = (4./5)/dx; c[1] = (1./5)/dx; c[2] = (4./105)/dx; c[3] = (1./280)/dx; break;
        default: return -1;
    }
    int64_t stride = dim == 0 ? (int64_t)G*G : (dim == 1 ? G : 1);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < G; ++i) {
        for (int j = 0; j < G; ++j) {
            for (int k = 0; k < G; ++k) {
                int64_t idx = ((int64_t)i*G + j)*G + k;
                int pos = dim == 0 ? i : (dim == 1 ? j : k);
                double v = 0;
                for (int m = 1; m <= reach; ++m) {
                    int64_t up = idx + (wrapi(pos + m, G) - pos)*stride;
                    int64_t lo = forward ? idx : idx + (wrapi(pos - m, G) - pos)*stride;
                    double d = c[m - 1]*(phi[up] - phi[lo]);
                    if (m == 1) v = d;
                    else if (m & 1) v = v + d;
                    else v = v - d;
                }
                out[idx] = v;
            }
        }
    }
    return 0;
}

/* ---------------------------------------------------------------- gather */
int pmo_gather(const double* grid, int G, double L, const double* pos, double* mom, int64_t n, int order,
               int dim, double factor, const double* shift) {
    if (order < 1 || order > 4) return -1;
    double off[3], scale;
    coord_setup(L, G, shift, 1, off, &scale);
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < n; ++p) {
        double wx[4], wy[4], wz[4];
        int64_t ix = set_weights(order, (pos[3*p + 0] - off[0])*scale, wx);
        int64_t iy = set_weights(order, (pos[3*p + 1] - off[1])*scale, wy);
        int64_t iz = set_weights(order, (pos[3*p + 2] - off[2])*scale, wz);
        double value = 0;
        for (int a = 0; a < order; ++a) {
            int64_t ia = wrapi(ix + a, G);
            for (int b = 0; b < order; ++b) {
                double wab = wx[a]*wy[b];
                const double* row = grid + (ia*G + wrapi(iy + b, G))*G;
                for (int c = 0; c < order; ++c) value += row[wrapi(iz + c, G)]*(wab*wz[c]);
            }
        }
        if (factor != 1) value *= factor;
        mom[3*p + dim] += value;
    }
    return 0;
}

/* ---------------------------------------------------------------- drift, Σ mom² */
int pmo_drift(double* pos, const double* mom, int64_t n, double dt_over_mass, double L) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < 3*n; ++i) {
        double x = fmod(pos[i] + mom[i]*dt_over_mass, L);
        if (x < 0) x += L;
        if (x == L) x = 0;
        pos[i] = x;
    }
    return 0;
}

double pmo_sum_mom2(const double* mom, int64_t n) {
    double s = 0;
#pragma omp parallel for reduction(+:s) schedule(static)
    for (int64_t i = 0; i < 3*n; ++i) s += mom[i]*mom[i];
    return s;
}

/* ---------------------------------------------------------------- one long-range kick (no interlacing,
 * real-space differentiation: the default path, interactions.py:2281-2330) */
int pmo_kick_long(const double* pos, double* mom, int64_t n, int G, double L, int order, int diff_order,
                  int deconv_order, double contribution, double prefactor, double gauss, double kick_factor,
                  double* work_real /* 2·G³ doubles */, cplx* work_slab /* G²·(G/2+1) */) {
    int64_t G3 = (int64_t)G*G*G;
    double* rho = work_real;
    double* force = work_real + G3;
    memset(rho, 0, sizeof(double)*(size_t)G3);
    int s = pmo_deposit(pos, n, G, L, order, contribution, NULL, rho);
    if (s) return s;
    pmo_fft_forward(rho, G, work_slab);
    pmo_kspace_potential(work_slab, G, prefactor, deconv_order, gauss, 1.0);
    pmo_fft_backward(work_slab, G, rho);   /* rho now holds φ */
    for (int dim = 0; dim < 3; ++dim) {
        s = pmo_diff(rho, G, dim, diff_order, L/G, force);
        if (s) return s;
        pmo_gather(force, G, L, pos, mom, n, order, dim, kick_factor, NULL);
    }
    return 0;
}
