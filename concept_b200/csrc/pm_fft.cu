// pm_fft.cu — hand-written slab transform for the fused Poisson solve (sm_100a, HBM/L2/shared-memory
// bound; no tensor cores: fp64/fp32 butterflies on the FMA pipes).
//
//   fft2d_kernel<DIR=−1>   per x plane: r2c along z (ZFwd tiles), then c2c along y (YPass tiles)
//   xsolve2_kernel         c2c along x · Green's function · inverse c2c along x, in place, addressing
//                          every rank's slab through peer pointers (no materialised transpose, fft.c:34-73)
//   fft2d_kernel<DIR=+1>   per x plane: inverse c2c along y, then c2r along z
//
// Replaces fft.c's FFTW-MPI plans (fft.c:105-290) + the potential loop (interactions.py:2092-2118) on the
// default gravity path.  Tile operations and their index math live in pm_fftops.cuh (CPU-checked by
// tests/test_fftcore_host.py); this file adds the persistent scheduling around them.
//
// Data path of one tile: 128-byte row segments  --ld.global.cg-->  registers  --first radix-8 stage-->
// ONE shared-memory tile updated in place by the middle stages  --last stage-->  registers  --st.global.
// (First version staged tiles with cp.async: measured 16 shared-memory wavefronts per 16-byte LDGSTS
// warp instruction, 4x an STS.128 — 38 % of the shared-memory pipe, which is the scarce resource here.)
// Several small CTAs per SM instead of one big one: their LDS / FP64 / STS / barrier phases interleave.
//
// Scheduling: persistent CTAs take tiles from a global ticket counter IN ORDER.  In the 2-D kernels the
// ticket order interleaves the first pass of plane p + lag with the second pass of plane p, and a
// second-pass tile waits (per-plane completion counter, release/acquire) until all first-pass tiles of
// its plane are stored.  The intermediate plane therefore never leaves L2 (2 MB per plane, a few planes
// in flight, 126 MB L2): DRAM sees one read and one write of the slab per 2-D transform instead of two.
#include "pm_internal.cuh"
#include "pm_fftops.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace pm {

using namespace fftc;

constexpr int kFftThreads = 256;
constexpr int kFft2dOcc = 3;      // CTAs per SM: independent barrier domains hide each other's LDS/DP/STS phases
constexpr int kXSolveOcc = 2;

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// the 128-byte lines touched by [p, p + bytes)
__device__ __forceinline__ void prefetch_l2_span(const void* p, int bytes) {
    const char* a = reinterpret_cast<const char*>(p);
    prefetch_l2(a);
    if ((reinterpret_cast<uintptr_t>(a) & 127) + bytes > 128) prefetch_l2(a + bytes - 1);
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---------------------------------------------------------------------------------------------
// persistent, ticket-ordered tile loop
// ---------------------------------------------------------------------------------------------
// Job interface:
//   int  decode(unsigned slot)        item (>= 0), −2: empty slot, −1: past the end
//   const unsigned* dep(int item)     completion counter the item waits for (nullptr: none)
//   unsigned dep_need()
//   void process(int item, V* work)   phases with __syncthreads() in between; global loads in the first,
//                                     global stores in the last
//   unsigned* signal(int item)        counter to bump once the item's stores are complete (nullptr: none)
//   void prefetch(int item)           prefetch.global.L2 over the item's input rows
// Thread 0 always holds the ticket after the current one (its atomic latency hides behind the tile), and
// the whole CTA prefetches that next tile into L2 while it works on the current one: the direct
// global->register loads of a tile's first stage then hit L2 (~4x shorter latency), so the few
// registers a thread can spare for loads in flight are enough to keep HBM busy.
// A CTA blocks on a dependency only before it starts a tile, dependencies point to strictly lower
// tickets and first-pass tiles have none, so the CTA with the lowest blocked ticket can always proceed.
template <class Job, typename V>
__device__ __forceinline__ void run_tiles(Job& job, unsigned* ticket, V* work, int* err) {
    __shared__ int s_item, s_ahead;
    const int tid = threadIdx.x;
    unsigned slot_next = 0;
    if (tid == 0) slot_next = atomicAdd(ticket, 1u);
    for (;;) {
        __syncthreads();                  // the previous tile is done with `work` and s_item
        if (tid == 0) {
            int item = job.decode(slot_next);
            while (item == -2) item = job.decode(atomicAdd(ticket, 1u));
            int ahead = -1;
            if (item >= 0) {
                slot_next = atomicAdd(ticket, 1u);
                ahead = job.decode(slot_next);
                const unsigned* d = job.dep(item);
                if (d != nullptr) {
                    unsigned spins = 0;
                    while (ld_acquire(d) < job.dep_need()) {
                        __nanosleep(100);
                        if (++spins > (1u << 24)) { atomicExch(err, 1); break; }   // seconds: give up loudly, never hang
                    }
                }
            }
            s_item = item;
            s_ahead = ahead;
        }
        __syncthreads();
        const int item = s_item;
        if (item < 0) break;
        const int ahead = s_ahead;
        if (ahead >= 0) job.prefetch(ahead);
        job.process(item, work);
        unsigned* sig = job.signal(item);
        if (sig != nullptr) {             // uniform over the CTA
            __syncthreads();              // every thread's stores are issued …
            if (tid == 0) {
                __threadfence();          // … and made visible device-wide before the counter moves
                atomicAdd(sig, 1u);
            }
        }
    }
}

// copy the B | C (| R) tables to shared memory; visible after the first barrier of run_tiles
template <typename V>
__device__ __forceinline__ Twiddles<V> load_twiddles(V* smem, const void* gmem, int G, bool with_r) {
    const int n = 64 + G + (with_r ? G / 8 + 1 : 0);
    for (int m = threadIdx.x; m < n; m += kFftThreads) smem[m] = reinterpret_cast<const V*>(gmem)[m];
    Twiddles<V> tw;
    tw.B = smem;
    tw.C = smem + 64;
    tw.R = smem + 64 + G;
    return tw;
}

// ---------------------------------------------------------------------------------------------
// 2-D (y,z) transforms of the local planes
// ---------------------------------------------------------------------------------------------
struct Fft2dParams {
    void* interior;       // first interior plane
    const void* tw;       // B | C | R twiddle tables (pm_fftcore.cuh)
    int nplanes;
    int mode;             // 0: both passes, dependency-ordered (L2-resident); 1: first pass only; 2: second pass only
    int lag;              // planes between the first and the second pass in ticket order (mode 0)
    unsigned* ticket;
    unsigned* done;       // [nplanes] finished first-pass tiles
    int* err;
};

template <typename T, int G, int DIR>
struct Fft2dJob {
    using S = SlabFFT<T, G>;
    using V = typename S::V;
    // forward: first pass z (kZTilesPerPlane tiles), second y;  inverse: first y, second z
    static constexpr int nA = DIR < 0 ? S::kZTilesPerPlane : S::kYTilesPerPlane;
    static constexpr int nB = DIR < 0 ? S::kYTilesPerPlane : S::kZTilesPerPlane;
    const Fft2dParams& p;
    Twiddles<V> tw;
    __device__ Fft2dJob(const Fft2dParams& p_, const Twiddles<V>& tw_) : p(p_), tw(tw_) {}

    // item = kind << 30 | plane << 10 | tile
    __device__ __forceinline__ int decode(unsigned slot) const {
        if (p.mode == 0) {
            const unsigned per = nA + nB;
            const unsigned blk = slot / per, r = slot - blk * per;
            if (blk >= (unsigned)(p.nplanes + p.lag)) return -1;
            if (r < (unsigned)nA) return blk < (unsigned)p.nplanes ? (int)((blk << 10) | r) : -2;
            const int pl = (int)blk - p.lag;
            return (pl >= 0 && pl < p.nplanes) ? (int)((1u << 30) | ((unsigned)pl << 10) | (r - nA)) : -2;
        }
        const unsigned n = p.mode == 1 ? nA : nB;
        const unsigned pl = slot / n, t = slot - pl * n;
        if (pl >= (unsigned)p.nplanes) return -1;
        return (int)(((p.mode == 1 ? 0u : 1u) << 30) | (pl << 10) | t);
    }
    __device__ __forceinline__ const unsigned* dep(int item) const {
        return (p.mode == 0 && (item >> 30)) ? p.done + ((item >> 10) & 0xfffff) : nullptr;
    }
    __device__ __forceinline__ unsigned dep_need() const { return nA; }
    __device__ __forceinline__ unsigned* signal(int item) const {
        return (p.mode == 0 && !(item >> 30)) ? p.done + ((item >> 10) & 0xfffff) : nullptr;
    }
    __device__ __forceinline__ T* plane(int item) const {
        return reinterpret_cast<T*>(p.interior) + (size_t)((item >> 10) & 0xfffff) * G * S::Gp;
    }
    __device__ __forceinline__ bool is_z(int item) const { return (DIR < 0) == ((item >> 30) == 0); }

    __device__ __forceinline__ void prefetch(int item) const {
        if (p.mode == 0 && (item >> 30)) return;     // second pass in one launch: its input was just written, it is in L2
        const int t = item & 1023;
        if (is_z(item)) {                            // CZ contiguous rows
            const char* base = reinterpret_cast<const char*>(plane(item) + (size_t)t * S::CZ * S::Gp);
            const int bytes = S::CZ * S::Gp * (int)sizeof(T);
            for (int o = threadIdx.x * 128; o < bytes + 127; o += kFftThreads * 128) prefetch_l2(base + min(o, bytes - 1));
        } else {                                     // CY·sizeof(V) = 128 bytes of every row
            const V* col = reinterpret_cast<const V*>(plane(item)) + t * S::CY;
            for (int j = threadIdx.x; j < G; j += kFftThreads) prefetch_l2_span(col + (size_t)j * S::Gc, S::CY * (int)sizeof(V));
        }
    }
    template <class Op>
    __device__ __forceinline__ void run(const Op& op, V* work) const {
#pragma unroll
        for (int ph = 0; ph < Op::kPhases; ++ph) {
            op.phase(ph, work, tw, threadIdx.x, kFftThreads);
            if (ph + 1 < Op::kPhases) __syncthreads();
        }
    }
    __device__ __forceinline__ void process(int item, V* work) const {
        const int t = item & 1023;
        if (is_z(item)) {
            if constexpr (DIR < 0) run(typename S::ZFwd{plane(item), t * S::CZ}, work);
            else run(typename S::ZInv{plane(item), t * S::CZ}, work);
        } else {
            run(typename S::template YPass<DIR>{reinterpret_cast<V*>(plane(item)), t * S::CY}, work);
        }
    }
};

template <typename T, int G, int DIR>
__global__ void __launch_bounds__(kFftThreads, kFft2dOcc) fft2d_kernel(const __grid_constant__ Fft2dParams p) {
    using S = SlabFFT<T, G>;
    using V = typename S::V;
    extern __shared__ __align__(16) unsigned char fft_smem[];
    V* work = reinterpret_cast<V*>(fft_smem);
    const Twiddles<V> tw = load_twiddles(work + S::kWorkElems, p.tw, G, true);
    Fft2dJob<T, G, DIR> job(p, tw);
    run_tiles(job, p.ticket, work, p.err);
}

// ---------------------------------------------------------------------------------------------
// x solve
// ---------------------------------------------------------------------------------------------
template <typename T, int G>
struct XSolveKParams {
    typename SlabFFT<T, G>::XGeom xg;
    const void* tw;
    int j0, njl, self_rank;
    unsigned* ticket;
    int* err;
};

template <typename T, int G>
struct XSolveJob {
    using S = SlabFFT<T, G>;
    using V = typename S::V;
    const typename S::XGeom* xg;   // in shared memory
    Twiddles<V> tw;
    int j0, njl, self_rank;
    __device__ __forceinline__ int decode(unsigned slot) const {
        return slot < (unsigned)(njl * S::kYTilesPerPlane) ? (int)slot : -1;
    }
    __device__ __forceinline__ const unsigned* dep(int) const { return nullptr; }
    __device__ __forceinline__ unsigned dep_need() const { return 0; }
    __device__ __forceinline__ unsigned* signal(int) const { return nullptr; }
    __device__ __forceinline__ void prefetch(int item) const {
        const int jl = item / S::kYTilesPerPlane, t = item - jl * S::kYTilesPerPlane;
        const int nxl = 1 << xg->nxl_shift;
        for (int il = threadIdx.x; il < nxl; il += kFftThreads)   // peer planes bypass the local L2: own planes only
            prefetch_l2_span(xg->at(self_rank * nxl + il, j0 + jl, t * S::CY), S::CY * (int)sizeof(V));
    }
    __device__ __forceinline__ void process(int item, V* work) const {
        const int jl = item / S::kYTilesPerPlane, t = item - jl * S::kYTilesPerPlane;
        const typename S::XSolve o{xg, j0 + jl, t * S::CY};
#pragma unroll
        for (int ph = 0; ph < S::XSolve::kPhases; ++ph) {
            o.phase(ph, work, tw, threadIdx.x, kFftThreads);
            if (ph + 1 < S::XSolve::kPhases) __syncthreads();
        }
    }
};

template <typename T, int G>
__global__ void __launch_bounds__(kFftThreads, kXSolveOcc) xsolve2_kernel(const __grid_constant__ XSolveKParams<T, G> p) {
    using S = SlabFFT<T, G>;
    using V = typename S::V;
    extern __shared__ __align__(16) unsigned char fft_smem[];
    V* work = reinterpret_cast<V*>(fft_smem);
    const Twiddles<V> tw = load_twiddles(work + S::kYTileElems, p.tw, G, false);
    double* sep = reinterpret_cast<double*>(work + S::kYTileElems + 64 + G);
    __shared__ typename S::XGeom s_xg;
    for (int m = threadIdx.x; m < G; m += kFftThreads) sep[m] = p.xg.sep[m];
    if (threadIdx.x == 0) {
        s_xg = p.xg;
        s_xg.sep = sep;
    }
    XSolveJob<T, G> job{&s_xg, tw, p.j0, p.njl, p.self_rank};
    run_tiles(job, p.ticket, work, p.err);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool fft2_supported(const pm_ctx* c) {
    const int G = c->g.G;
    if (!(G == 128 || G == 256 || G == 512)) return false;
    if (c->nranks > kMaxFftPeers) return false;
    if (c->g.nxl & (c->g.nxl - 1)) return false;   // slabs must be a power of two thick
    if (c->g.nxl >= (1 << 20)) return false;
    if (c->nranks > 1 && !c->peers_ready) return false;
    return true;
}

template <typename T>
static int make_fft2_tables_t(pm_ctx* c) {
    using V = typename Vec2<T>::type;
    const int G = c->g.G;
    const int n = 64 + G + G / 8 + 1;     // B | C | R (pm_fftcore.cuh)
    std::vector<V> tw(n);
    auto w = [](long double num, long double den) {
        const long double a = -2.0L * 3.14159265358979323846264338327950288L * num / den;
        V v; v.x = (T)cosl(a); v.y = (T)sinl(a);
        return v;
    };
    for (int a = 0; a < 8; ++a) for (int b = 0; b < 8; ++b) tw[a * 8 + b] = w(a * b, 64);
    for (int a = 0; a < G / 64; ++a) for (int q = 0; q < 64; ++q) tw[64 + a * 64 + q] = w(a * q, G);
    for (int k = 0; k <= G / 8; ++k) tw[64 + G + k] = w(k, G);
    PM_CHECK_CUDA(cudaMalloc(&c->f2_tw, sizeof(V) * n));
    PM_CHECK_CUDA(cudaMemcpy(c->f2_tw, tw.data(), sizeof(V) * n, cudaMemcpyHostToDevice));
    return PM_OK;
}

int make_fft2_tables(pm_ctx* c) {
    const int G = c->g.G;
    if (!(G == 128 || G == 256 || G == 512)) return PM_OK;
    PM_TRY(c->dtype == PM_GRID_F64 ? make_fft2_tables_t<double>(c) : make_fft2_tables_t<float>(c));
    // counters: [0] ticket fwd, [1] ticket x, [2] ticket inv, [3] unused, [4 …) done fwd, then done inv;
    // one more entry after those f2_nctr: the sticky error flag
    c->f2_nctr = 4 + 2 * (size_t)c->g.nxl;
    PM_CHECK_CUDA(cudaMalloc(&c->f2_ctr, sizeof(unsigned) * (c->f2_nctr + 1)));
    PM_CHECK_CUDA(cudaMemset(c->f2_ctr, 0, sizeof(unsigned) * (c->f2_nctr + 1)));
    c->f2_lag = 10;   // planes; > CTAs in flight / tiles per plane (444 / 96)
    if (const char* e = getenv("PM_FFT_LAG")) c->f2_lag = std::max(1, atoi(e));
    return PM_OK;
}

template <typename T, int G>
static size_t fft2d_smem() {
    using S = SlabFFT<T, G>;
    return sizeof(typename S::V) * ((size_t)S::kWorkElems + twiddle_entries<G>());
}
template <typename T, int G>
static size_t xsolve2_smem() {
    using S = SlabFFT<T, G>;
    return sizeof(typename S::V) * ((size_t)S::kYTileElems + 64 + G) + sizeof(double) * G;
}

template <typename T, int G, int DIR>
static int launch_fft2d(pm_ctx* c, int mode) {
    using S = SlabFFT<T, G>;
    Fft2dParams p;
    p.interior = c->real_interior<T>();
    p.tw = c->f2_tw;
    p.nplanes = c->g.nxl;
    p.mode = mode;
    p.lag = c->f2_lag;
    p.ticket = c->f2_ctr + (DIR < 0 ? 0 : 2);
    p.done = c->f2_ctr + 4 + (DIR < 0 ? 0 : c->g.nxl);
    p.err = reinterpret_cast<int*>(c->f2_ctr + c->f2_nctr);
    const size_t smem = fft2d_smem<T, G>();
    PM_CHECK_CUDA(cudaFuncSetAttribute(fft2d_kernel<T, G, DIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t tiles = (int64_t)c->g.nxl * (mode == 0 ? S::kZTilesPerPlane + S::kYTilesPerPlane
                                                         : ((mode == 1) == (DIR < 0) ? S::kZTilesPerPlane : S::kYTilesPerPlane));
    const int grid = (int)std::min<int64_t>(tiles, (int64_t)kNumSMs * kFft2dOcc);
    if (mode != 0) PM_CHECK_CUDA(cudaMemsetAsync(p.ticket, 0, sizeof(unsigned), c->stream));
    PM_LAUNCH((fft2d_kernel<T, G, DIR>), grid, kFftThreads, smem, c->stream, p);
    return PM_OK;
}

template <typename T, int G>
static int launch_xsolve2(pm_ctx* c, double prefactor) {
    using S = SlabFFT<T, G>;
    using V = typename S::V;
    const Geom& g = c->g;
    XSolveKParams<T, G> p;
    for (int r = 0; r < kMaxFftPeers; ++r) p.xg.base[r] = nullptr;
    if (c->nranks == 1) {
        p.xg.base[0] = reinterpret_cast<V*>(c->real_interior<T>());
    } else {
        for (int r = 0; r < c->nranks; ++r)
            p.xg.base[r] = reinterpret_cast<V*>(reinterpret_cast<T*>(c->peer_real[r]) + (size_t)g.halo * g.G * g.Gp);
    }
    p.xg.nxl_shift = 0;
    while ((1 << p.xg.nxl_shift) < g.nxl) ++p.xg.nxl_shift;
    p.xg.sep = c->xs_sep;
    p.xg.prefactor = prefactor;
    p.tw = c->f2_tw;
    p.j0 = g.j0;
    p.njl = g.njl;
    p.self_rank = c->rank;
    p.ticket = c->f2_ctr + 1;
    p.err = reinterpret_cast<int*>(c->f2_ctr + c->f2_nctr);
    const size_t smem = xsolve2_smem<T, G>();
    PM_CHECK_CUDA(cudaFuncSetAttribute(xsolve2_kernel<T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t tiles = (int64_t)g.njl * S::kYTilesPerPlane;
    const int grid = (int)std::min<int64_t>(tiles, (int64_t)kNumSMs * kXSolveOcc);
    PM_LAUNCH((xsolve2_kernel<T, G>), grid, kFftThreads, smem, c->stream, p);
    return PM_OK;
}

// stage: 0 = whole solve; 1 = forward 2-D transforms, 2 = x solve, 3 = inverse 2-D transforms (for timing
// the three kernels separately; the caller runs 1, 2, 3 in order)
template <typename T, int G>
static int solve_fft2_tg(pm_ctx* c, double prefactor, bool l2_fused, int stage) {
    if (stage == 0 || stage == 1) {
        PM_CHECK_CUDA(cudaMemsetAsync(c->f2_ctr, 0, sizeof(unsigned) * c->f2_nctr, c->stream));
        if (l2_fused) {
            PM_TRY((launch_fft2d<T, G, -1>(c, 0)));
        } else {
            PM_TRY((launch_fft2d<T, G, -1>(c, 1)));
            PM_TRY((launch_fft2d<T, G, -1>(c, 2)));
        }
    }
    if (stage == 0 || stage == 2) {
        if (c->nranks > 1) PM_TRY(device_barrier(c));   // every rank's 2-D spectra are complete
        PM_TRY((launch_xsolve2<T, G>(c, prefactor)));
        if (c->nranks > 1) PM_TRY(device_barrier(c));   // all peers have written our planes
    }
    if (stage == 0 || stage == 3) {
        if (l2_fused) {
            PM_TRY((launch_fft2d<T, G, +1>(c, 0)));
        } else {
            PM_TRY((launch_fft2d<T, G, +1>(c, 1)));
            PM_TRY((launch_fft2d<T, G, +1>(c, 2)));
        }
    }
    return PM_OK;
}

// forward 2-D transforms → fused x pass → inverse 2-D transforms with the hand-written kernels.
// l2_fused: dependency-ordered single launch per 2-D transform.
int solve_fft2(pm_ctx* c, double prefactor, int deconv_order, double gauss, bool l2_fused, int stage) {
    PM_REQUIRE(fft2_supported(c) && c->f2_tw != nullptr, "hand-written FFT path not available for this grid size / rank layout");
    PM_TRY(update_sep_table(c, deconv_order, gauss));
    const bool f64 = c->dtype == PM_GRID_F64;
    int s;
    switch (c->g.G) {
        case 128: s = f64 ? solve_fft2_tg<double, 128>(c, prefactor, l2_fused, stage) : solve_fft2_tg<float, 128>(c, prefactor, l2_fused, stage); break;
        case 256: s = f64 ? solve_fft2_tg<double, 256>(c, prefactor, l2_fused, stage) : solve_fft2_tg<float, 256>(c, prefactor, l2_fused, stage); break;
        default:  s = f64 ? solve_fft2_tg<double, 512>(c, prefactor, l2_fused, stage) : solve_fft2_tg<float, 512>(c, prefactor, l2_fused, stage); break;
    }
    return s;
}

// the give-up flag of run_tiles (a dependency wait that ran into its spin limit); host-synchronising
int fft2_check_error(pm_ctx* c) {
    if (c->f2_ctr == nullptr) return PM_OK;
    int e = 0;
    PM_CHECK_CUDA(cudaMemcpyAsync(&e, c->f2_ctr + c->f2_nctr, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PM_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    PM_REQUIRE(e == 0, "hand-written FFT: a tile dependency was not satisfied in time (results are invalid)");
    return PM_OK;
}

}  // namespace pm
