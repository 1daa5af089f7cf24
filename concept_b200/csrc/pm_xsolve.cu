// pm_xsolve.cu — the fused x pass of the Poisson solve:
//
//     forward FFT along x  →  Green's function / deconvolution / Nyquist+origin nullification
//                          →  inverse FFT along x,           in ONE kernel, in place.
//
// The (y,z) transforms are batched 2-D cuFFT plans per x plane; this kernel replaces the third
// forward pass, the k-space kernel and the first inverse pass (3 sweeps over the 1 GB slab → 1).
// Layout: complex [i_local][j][kk] inside every rank's real buffer (what the in-place 2-D r2c leaves).
// An x line (fixed j, kk) is strided by G·Gc elements and — with several ranks — spread over the ranks'
// buffers: rank r holds i ∈ [r·nxl, (r+1)·nxl).  The kernel reads and writes those rows directly through
// peer pointers (CUDA IPC mappings over NVLink/NVSwitch), so the slab "transpose" of FFTW-MPI
// (fft.c:34-73, 240-257) never materialises: transfer and math are fused tile by tile.
//
// FFT: N = 8^S points (S = 2, 3: N = 64, 512), radix-8 Cooley–Tukey in registers, N/8 threads per line,
// S−1 shared-memory exchanges each way, natural order in and out.  A CTA of 256 threads works on a tile of
// COLS = 4 adjacent kk columns (64-byte row segments in fp64) × 256/(4·N/8) j rows.
//
// Reference semantics of the factor: mesh.py:2775-2856, interactions.py:2092-2118, mesh.py:3585-3622.
#include "pm_internal.cuh"

#include <algorithm>
#include <cmath>

namespace pm {

template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

struct XSolveParams {
    void* base[kMaxPeers];   // rank r's first interior plane, complex [il][j][kk]
    const double2* tw;       // exp(−2πi·m/N), m < N
    const double* sep;       // separable factor per axis index l < G: (x_l/sin x_l)^D · exp(−gauss·k_l²)
    double prefactor;
    int G, Gc, nxl, j0, njl;
    int nxl_shift;           // nxl = 1 << nxl_shift (power-of-two slabs: rank of plane i is i >> nxl_shift)
};

// 8-point DFT in registers; DIR = −1 forward (W = e^{−2πi/8}), +1 inverse.
template <int DIR, typename T>
__device__ __forceinline__ void dft8(T (&r)[8], T (&i)[8]) {
    const T h = (T)0.70710678118654752440;
    T a0r = r[0] + r[4], a0i = i[0] + i[4], a4r = r[0] - r[4], a4i = i[0] - i[4];
    T a1r = r[1] + r[5], a1i = i[1] + i[5], a5r = r[1] - r[5], a5i = i[1] - i[5];
    T a2r = r[2] + r[6], a2i = i[2] + i[6], a6r = r[2] - r[6], a6i = i[2] - i[6];
    T a3r = r[3] + r[7], a3i = i[3] + i[7], a7r = r[3] - r[7], a7i = i[3] - i[7];
    T t;
    if (DIR < 0) {
        t = (a5r + a5i) * h; a5i = (a5i - a5r) * h; a5r = t;        // ·(1−i)/√2
        t = a6i; a6i = -a6r; a6r = t;                               // ·(−i)
        t = (a7i - a7r) * h; a7i = (-a7r - a7i) * h; a7r = t;       // ·(−1−i)/√2
    } else {
        t = (a5r - a5i) * h; a5i = (a5r + a5i) * h; a5r = t;        // ·(1+i)/√2
        t = -a6i; a6i = a6r; a6r = t;                               // ·(+i)
        t = (-a7r - a7i) * h; a7i = (a7r - a7i) * h; a7r = t;       // ·(−1+i)/√2
    }
    T b0r = a0r + a2r, b0i = a0i + a2i, b2r = a0r - a2r, b2i = a0i - a2i;
    T b1r = a1r + a3r, b1i = a1i + a3i, b3r = a1r - a3r, b3i = a1i - a3i;
    T b4r = a4r + a6r, b4i = a4i + a6i, b6r = a4r - a6r, b6i = a4i - a6i;
    T b5r = a5r + a7r, b5i = a5i + a7i, b7r = a5r - a7r, b7i = a5i - a7i;
    if (DIR < 0) {
        t = b3i; b3i = -b3r; b3r = t;
        t = b7i; b7i = -b7r; b7r = t;
    } else {
        t = -b3i; b3i = b3r; b3r = t;
        t = -b7i; b7i = b7r; b7r = t;
    }
    r[0] = b0r + b1r; i[0] = b0i + b1i; r[4] = b0r - b1r; i[4] = b0i - b1i;
    r[2] = b2r + b3r; i[2] = b2i + b3i; r[6] = b2r - b3r; i[6] = b2i - b3i;
    r[1] = b4r + b5r; i[1] = b4i + b5i; r[5] = b4r - b5r; i[5] = b4i - b5i;
    r[3] = b6r + b7r; i[3] = b6i + b7i; r[7] = b6r - b7r; i[7] = b6i - b7i;
}

template <int DIR, typename T, typename V>
__device__ __forceinline__ void twiddle(T& r, T& i, const V* stw, int m) {
    const V w = stw[m];
    const T wi = DIR < 0 ? w.y : -w.y;
    const T t = fma(r, w.x, -(i * wi));
    i = fma(r, wi, i * w.x);
    r = t;
}

// Per-mode factor of particle_mesh's potential loop (interactions.py:2092-2118), in separable form:
//   [Π_l x_l/sin x_l]^D · exp(−gauss·k²) = Π_l sep[l],  times prefactor/k².
// (kspace_kernel keeps the reference's exact operation order; the two agree to a few ulp.)
__device__ __forceinline__ double green_factor(const XSolveParams& p, double sep_jk, int i, int kj2_kk2, bool line_nyq) {
    const int nyq = p.G >> 1;
    if (line_nyq || i == nyq) return 0.0;
    const int ki = i - (i >= nyq ? p.G : 0);
    const int k2 = kj2_kk2 + ki * ki;
    if (k2 == 0) return 0.0;
    return (p.sep[i] * sep_jk) * (p.prefactor * __drcp_rn((double)k2));
}

template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    if constexpr (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem));
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

constexpr int kXCols = 4;          // kk columns per tile
constexpr int kXThreads = 256;

// S = 3: N = 512, 64 threads per line, 1 j row per tile.   S = 2: N = 64, 8 threads per line, 8 j rows.
template <typename T, int S>
__global__ void __launch_bounds__(kXThreads, 2)
xsolve_kernel(XSolveParams p) {
    using V = typename Vec2<T>::type;
    constexpr int N = S == 3 ? 512 : 64;
    constexpr int TPL = N / 8;                         // threads per line
    constexpr int JROWS = kXThreads / (kXCols * TPL);  // j rows per tile
    constexpr int LP = S == 3 ? 578 : 74;              // smem pitch per line (complex), ≡ 2 mod 8
    // twiddles, laid out so that consecutive threads read consecutive entries (no bank conflicts):
    //   stw1[k1·TPL + tl] = W^(tl·k1)         first stage, forward
    //   stwq[k1·TPL + tl] = W^(q(tl)·k1)      first stage, inverse (q = natural-order owner of the thread's outputs)
    //   stw2[c·8 + b]     = W^(8·b·c)         middle stage (S = 3)
    __shared__ V stw1[N];
    __shared__ V stwq[S == 3 ? N : 1];
    __shared__ V stw2[S == 3 ? 64 : 1];
    extern __shared__ __align__(16) unsigned char xs_raw[];
    V* sbuf = reinterpret_cast<V*>(xs_raw);
    // prefetch slots: the next tile's 8 values of this thread, written by cp.async (LDGSTS) while the
    // current tile is being transformed; each thread only ever touches its own slots
    V* pf = sbuf + (kXThreads / TPL) * LP + threadIdx.x;

    for (int m = threadIdx.x; m < N; m += kXThreads) {
        const int k1 = m / TPL, t = m - k1 * TPL;
        double2 w = p.tw[(t * k1) % N];
        V v; v.x = (T)w.x; v.y = (T)w.y;
        stw1[m] = v;
        if constexpr (S == 3) {
            const int qq = (t >> 3) + 8 * (t & 7);
            w = p.tw[(qq * k1) % N];
            v.x = (T)w.x; v.y = (T)w.y;
            stwq[m] = v;
            if (m < 64) {
                w = p.tw[(8 * (m & 7) * (m >> 3)) % N];
                v.x = (T)w.x; v.y = (T)w.y;
                stw2[m] = v;
            }
        }
    }
    const int col = threadIdx.x & (kXCols - 1);
    const int tl = (threadIdx.x >> 2) % TPL;           // thread within line
    const int jr = (threadIdx.x >> 2) / TPL;           // j row within tile
    V* line = sbuf + (jr * kXCols + col) * LP;
    const int ktiles = (p.Gc - 1) / kXCols;            // kk = 0 … G/2−1 ; the kk = G/2 column is zero-filled below
    const int jtiles = (p.njl + JROWS - 1) / JROWS;
    const int64_t ntiles = (int64_t)ktiles * jtiles;
    const size_t rowstride = (size_t)p.G * p.Gc;       // elements between consecutive i
    __syncthreads();

    // issue the asynchronous loads of one tile into this thread's prefetch slots
    auto prefetch = [&](int64_t tile) {
        const int jt = (int)(tile / ktiles);
        const int kk = (int)(tile - (int64_t)jt * ktiles) * kXCols + col;
        const int jl = jt * JROWS + jr;
        if (jl < p.njl) {
            const size_t inplane = (size_t)(p.j0 + jl) * p.Gc + kk;
#pragma unroll
            for (int n1 = 0; n1 < 8; ++n1) {
                const int i = tl + TPL * n1;
                const int r = i >> p.nxl_shift;
                const int il = i & (p.nxl - 1);
                cp_async<sizeof(V)>(pf + n1 * kXThreads,
                                    reinterpret_cast<const V*>(p.base[r]) + (size_t)il * rowstride + inplane);
            }
        }
        cp_async_commit();
    };
    if (blockIdx.x < ntiles) prefetch(blockIdx.x);

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int jt = (int)(tile / ktiles);
        const int kk = (int)(tile - (int64_t)jt * ktiles) * kXCols + col;
        const int jl = jt * JROWS + jr;
        const bool active = jl < p.njl;
        const int j = p.j0 + (active ? jl : 0);
        const size_t inplane = (size_t)j * p.Gc + kk;
        T vr[8], vi[8];
        // ---- x[n2 + TPL·n1], n2 = tl, from the prefetch slots -------------------------------------
        cp_async_wait_all();
#pragma unroll
        for (int n1 = 0; n1 < 8; ++n1) {
            V v; v.x = 0; v.y = 0;
            if (active) v = pf[n1 * kXThreads];
            vr[n1] = v.x; vi[n1] = v.y;
        }
        int q;   // this thread ends up holding X[q + TPL·d], d = 0…7
        // ---- forward ---------------------------------------------------------------------
        dft8<-1>(vr, vi);
        if (tile + gridDim.x < ntiles) prefetch(tile + gridDim.x);   // overlaps with the rest of this tile
        if constexpr (S == 3) {
            const int a = tl >> 3, b = tl & 7;
#pragma unroll
            for (int k1 = 0; k1 < 8; ++k1) {
                if (k1) twiddle<-1>(vr[k1], vi[k1], stw1, k1 * TPL + tl);
                V v; v.x = vr[k1]; v.y = vi[k1];
                line[k1 * 72 + a * 9 + b] = v;
            }
            __syncthreads();
            const int k1 = tl >> 3;   // (b = tl & 7 reused)
#pragma unroll
            for (int aa = 0; aa < 8; ++aa) {
                const V v = line[k1 * 72 + aa * 9 + b];
                vr[aa] = v.x; vi[aa] = v.y;
            }
            __syncthreads();
            dft8<-1>(vr, vi);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (c) twiddle<-1>(vr[c], vi[c], stw2, c * 8 + b);
                V v; v.x = vr[c]; v.y = vi[c];
                line[k1 * 72 + b * 9 + c] = v;
            }
            __syncthreads();
            const int c = tl & 7;
#pragma unroll
            for (int bb = 0; bb < 8; ++bb) {
                const V v = line[k1 * 72 + bb * 9 + c];
                vr[bb] = v.x; vi[bb] = v.y;
            }
            __syncthreads();
            dft8<-1>(vr, vi);
            q = k1 + 8 * c;
        } else {
#pragma unroll
            for (int k1 = 0; k1 < 8; ++k1) {
                if (k1) twiddle<-1>(vr[k1], vi[k1], stw1, k1 * TPL + tl);
                V v; v.x = vr[k1]; v.y = vi[k1];
                line[k1 * 9 + tl] = v;
            }
            __syncthreads();
#pragma unroll
            for (int n2 = 0; n2 < 8; ++n2) {
                const V v = line[tl * 9 + n2];
                vr[n2] = v.x; vi[n2] = v.y;
            }
            __syncthreads();
            dft8<-1>(vr, vi);
            q = tl;
        }
        // ---- Green's function ----------------------------------------------------------------
        {
            const int nyq = p.G >> 1;
            const int kj = j - (j >= nyq ? p.G : 0);
            const bool line_nyq = (j == nyq) || (kk == nyq);
            const double sep_jk = p.sep[j] * p.sep[kk];
            const int kj2_kk2 = kj * kj + kk * kk;
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                const T f = (T)green_factor(p, sep_jk, q + TPL * d, kj2_kk2, line_nyq);
                vr[d] *= f; vi[d] *= f;
            }
        }
        // ---- inverse: n2 = q, n1 = d ------------------------------------------------------------
        dft8<+1>(vr, vi);
        if constexpr (S == 3) {
            const int a = q >> 3, b = q & 7;
#pragma unroll
            for (int k1 = 0; k1 < 8; ++k1) {
                if (k1) twiddle<+1>(vr[k1], vi[k1], stwq, k1 * TPL + tl);
                V v; v.x = vr[k1]; v.y = vi[k1];
                line[k1 * 72 + a * 9 + b] = v;
            }
            __syncthreads();
            const int k1 = tl >> 3, b2 = tl & 7;
#pragma unroll
            for (int aa = 0; aa < 8; ++aa) {
                const V v = line[k1 * 72 + aa * 9 + b2];
                vr[aa] = v.x; vi[aa] = v.y;
            }
            __syncthreads();
            dft8<+1>(vr, vi);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (c) twiddle<+1>(vr[c], vi[c], stw2, c * 8 + b2);
                V v; v.x = vr[c]; v.y = vi[c];
                line[k1 * 72 + b2 * 9 + c] = v;
            }
            __syncthreads();
            const int c = tl & 7;
#pragma unroll
            for (int bb = 0; bb < 8; ++bb) {
                const V v = line[k1 * 72 + bb * 9 + c];
                vr[bb] = v.x; vi[bb] = v.y;
            }
            __syncthreads();
            dft8<+1>(vr, vi);
            q = k1 + 8 * c;
        } else {
#pragma unroll
            for (int k1 = 0; k1 < 8; ++k1) {
                if (k1) twiddle<+1>(vr[k1], vi[k1], stw1, k1 * TPL + tl);
                V v; v.x = vr[k1]; v.y = vi[k1];
                line[k1 * 9 + q] = v;
            }
            __syncthreads();
#pragma unroll
            for (int n2 = 0; n2 < 8; ++n2) {
                const V v = line[tl * 9 + n2];
                vr[n2] = v.x; vi[n2] = v.y;
            }
            __syncthreads();
            dft8<+1>(vr, vi);
            q = tl;
        }
        // ---- store y[q + TPL·d] back in place ----------------------------------------------------
        if (active) {
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                const int i = q + TPL * d;
                const int r = i >> p.nxl_shift;
                const int il = i & (p.nxl - 1);
                V v; v.x = vr[d]; v.y = vi[d];
                reinterpret_cast<V*>(p.base[r])[(size_t)il * rowstride + inplane] = v;
            }
        }
    }
}

// kk = G/2 column of the local planes → 0 (Nyquist plane, mesh.py:3615-3622)
template <typename V>
__global__ void __launch_bounds__(256) zero_nyquist_column_kernel(V* __restrict__ planes, int64_t rows, int Gc) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        V z; z.x = 0; z.y = 0;
        planes[r * Gc + (Gc - 1)] = z;
    }
}

bool xsolve_supported(const pm_ctx* c) {
    if (!(c->g.G == 512 || c->g.G == 64)) return false;
    if (c->nranks > kMaxPeers) return false;
    if (c->g.nxl & (c->g.nxl - 1)) return false;   // slabs must be a power of two thick
    if (c->nranks > 1 && !c->peers_ready) return false;
    return true;
}

// separable per-axis factor (x_l/sin x_l)^D · exp(−gauss·k_l²), x_l = k_l·π/G + ε (mesh.py:2775-2776),
// cached for one (deconv_order, gauss)
int update_sep_table(pm_ctx* c, int deconv_order, double gauss) {
    const Geom& g = c->g;
    if (c->xs_sep != nullptr && c->xs_sep_deconv == deconv_order && c->xs_sep_gauss == gauss) return PM_OK;
    std::vector<double> sep(g.G);
    for (int l = 0; l < g.G; ++l) {
        const int k = l - (l >= g.G / 2 ? g.G : 0);
        const double x = k * (M_PI / g.G) + kEps;
        double v = 1.0;
        for (int e = 0; e < deconv_order; ++e) v *= x / sin(x);
        sep[l] = v * exp(-gauss * (double)k * (double)k);
    }
    if (c->xs_sep == nullptr) PM_CHECK_CUDA(cudaMalloc(&c->xs_sep, sizeof(double) * g.G));
    PM_CHECK_CUDA(cudaStreamSynchronize(c->stream));   // a previous launch may still read the table
    PM_CHECK_CUDA(cudaMemcpy(c->xs_sep, sep.data(), sizeof(double) * g.G, cudaMemcpyHostToDevice));
    c->xs_sep_deconv = deconv_order;
    c->xs_sep_gauss = gauss;
    return PM_OK;
}

template <typename T>
static int xsolve_t(pm_ctx* c, double prefactor, int deconv_order, double gauss) {
    using V = typename Vec2<T>::type;
    const Geom& g = c->g;
    XSolveParams p;
    for (int r = 0; r < kMaxPeers; ++r) p.base[r] = nullptr;
    if (c->nranks == 1) {
        p.base[0] = c->real_interior<T>();
    } else {
        for (int r = 0; r < c->nranks; ++r)
            p.base[r] = reinterpret_cast<T*>(c->peer_real[r]) + (size_t)g.halo * g.G * g.Gp;
    }
    PM_TRY(update_sep_table(c, deconv_order, gauss));
    p.tw = c->xs_tw;
    p.sep = c->xs_sep;
    p.prefactor = prefactor;
    p.G = g.G; p.Gc = g.Gc; p.nxl = g.nxl; p.j0 = g.j0; p.njl = g.njl;
    p.nxl_shift = 0;
    while ((1 << p.nxl_shift) < g.nxl) ++p.nxl_shift;
    // own planes: Nyquist column
    PM_LAUNCH((zero_nyquist_column_kernel<V>), kNumSMs, 256, 0, c->stream,
              reinterpret_cast<V*>(c->real_interior<T>()), (int64_t)g.nxl * g.G, g.Gc);
    if (c->nranks > 1) PM_TRY(device_barrier(c));   // every rank's 2-D spectra are complete
    const int S = g.G == 512 ? 3 : 2;
    const int lines = S == 3 ? kXCols : kXCols * 8;
    const int LP = S == 3 ? 578 : 74;
    const size_t smem = ((size_t)lines * LP + 8 * kXThreads) * sizeof(V);
    const int jrows = S == 3 ? 1 : 8;
    const int64_t ntiles = (int64_t)((g.Gc - 1) / kXCols) * ((g.njl + jrows - 1) / jrows);
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)kNumSMs * 2);
    if (S == 3) {
        PM_CHECK_CUDA(cudaFuncSetAttribute(xsolve_kernel<T, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PM_LAUNCH((xsolve_kernel<T, 3>), grid, kXThreads, smem, c->stream, p);
    } else {
        PM_CHECK_CUDA(cudaFuncSetAttribute(xsolve_kernel<T, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PM_LAUNCH((xsolve_kernel<T, 2>), grid, kXThreads, smem, c->stream, p);
    }
    if (c->nranks > 1) PM_TRY(device_barrier(c));   // all peers have written our planes
    return PM_OK;
}

int xsolve(pm_ctx* c, double prefactor, int deconv_order, double gauss) {
    return c->dtype == PM_GRID_F64 ? xsolve_t<double>(c, prefactor, deconv_order, gauss)
                                   : xsolve_t<float>(c, prefactor, deconv_order, gauss);
}

int make_xsolve_tables(pm_ctx* c) {
    const int N = c->g.G;
    if (!(N == 512 || N == 64)) return PM_OK;
    std::vector<double2> tw(N);
    for (int m = 0; m < N; ++m) {
        const long double a = -2.0L * 3.14159265358979323846264338327950288L * m / N;
        tw[m].x = (double)cosl(a);
        tw[m].y = (double)sinl(a);
    }
    PM_CHECK_CUDA(cudaMalloc(&c->xs_tw, sizeof(double2) * N));
    PM_CHECK_CUDA(cudaMemcpy(c->xs_tw, tw.data(), sizeof(double2) * N, cudaMemcpyHostToDevice));
    return PM_OK;
}

}  // namespace pm
