// pm_fft.cu — hand-written slab transform for the fused Poisson solve (sm_100a, HBM/L2-bound; no
// tensor cores: fp64/fp32 butterflies on the FMA pipes, 128-byte row segments staged through shared
// memory with cp.async).
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
// Scheduling: one persistent CTA per SM takes tiles from a global ticket counter IN ORDER.  In the 2-D
// kernels the ticket order interleaves the first pass of plane p + lag with the second pass of plane p,
// and a second-pass tile waits (per-plane completion counter, release/acquire) until all first-pass tiles
// of its plane are stored.  The intermediate plane therefore never leaves L2 (2 MB per plane, a few
// planes in flight, 126 MB L2): DRAM sees one read and one write of the slab per 2-D transform instead of
// two.  A CTA only ever blocks on a dependency while it holds no unfinished tile, and dependencies point
// to strictly lower tickets, so the wait graph cannot close.
#include "pm_internal.cuh"
#include "pm_fftops.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace pm {

using namespace fftc;

constexpr int kFftThreads = 512;

template <int BYTES>
__device__ __forceinline__ void cp_async_g2s(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    if constexpr (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
struct CpAsync {
    template <typename V> __device__ __forceinline__ void operator()(V* dst, const V* src) const {
        cp_async_g2s<sizeof(V)>(dst, src);
    }
};
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---------------------------------------------------------------------------------------------
// persistent, ticket-ordered, double-buffered tile pipeline
// ---------------------------------------------------------------------------------------------
// Job interface:
//   int  decode(unsigned slot)        item (>= 0), −2: empty slot, −1: past the end
//   const unsigned* dep(int item)     completion counter the item waits for (nullptr: none)
//   unsigned dep_need()
//   void issue(int item, V* raw)      cp.async loads of the raw tile
//   void process(int item, const V* raw, V* work)   phases with __syncthreads() in between, ends in global stores
//   unsigned* signal(int item)        counter to bump once the item's stores are complete (nullptr: none)
template <class Job, typename V>
__device__ __forceinline__ void run_pipeline(Job& job, unsigned* ticket, V* ring0, V* ring1, V* work, int* err) {
    __shared__ int s_next, s_ready;
    const int tid = threadIdx.x;
    auto fetch = [&]() -> int {
        for (;;) {
            const int item = job.decode(atomicAdd(ticket, 1u));
            if (item != -2) return item;
        }
    };
    auto ready = [&](int item) -> bool {
        const unsigned* d = job.dep(item);
        return d == nullptr || ld_acquire(d) >= job.dep_need();
    };
    auto publish = [&](unsigned* sig) {
        if (tid == 0 && sig != nullptr) {
            __threadfence();
            atomicAdd(sig, 1u);
        }
    };
    if (tid == 0) s_next = fetch();
    __syncthreads();
    int next = s_next;
    bool next_loaded = false;
    unsigned* pending = nullptr;
    int buf = 0;
    while (next >= 0) {
        V* rawc = buf ? ring1 : ring0;
        V* rawn = buf ? ring0 : ring1;
        if (!next_loaded) {
            // blocking wait: this CTA holds no unfinished tile here, and the tiles waited for have lower tickets
            if (tid == 0) {
                unsigned spins = 0;
                while (!ready(next)) {
                    __nanosleep(200);
                    if (++spins > (1u << 23)) { atomicExch(err, 1); break; }   // ~seconds: give up loudly, never hang
                }
            }
            __syncthreads();
            job.issue(next, rawc);
            cp_commit();
        }
        const int cur = next;
        if (tid == 0) {
            const int n2 = fetch();
            s_next = n2;
            s_ready = (n2 >= 0 && ready(n2)) ? 1 : 0;
        }
        cp_wait_all();
        __syncthreads();          // cur's tile landed; s_next visible; the previous tile's stores are ordered before this
        publish(pending);
        pending = nullptr;
        next = s_next;
        next_loaded = false;
        if (next >= 0 && s_ready) {
            job.issue(next, rawn);
            cp_commit();
            next_loaded = true;
        }
        job.process(cur, rawc, work);
        pending = job.signal(cur);
        if (!next_loaded) {       // about to block or to leave: publish the finished tile first
            __syncthreads();
            publish(pending);
            pending = nullptr;
        }
        buf ^= 1;
    }
}

// ---------------------------------------------------------------------------------------------
// 2-D (y,z) transforms of the local planes
// ---------------------------------------------------------------------------------------------
struct Fft2dParams {
    void* interior;       // first interior plane
    const void* tw;       // e^{−2πi m/G} as V[G]
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
    const V* tw;
    __device__ Fft2dJob(const Fft2dParams& p_, const V* tw_) : p(p_), tw(tw_) {}

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

    __device__ __forceinline__ void issue(int item, V* raw) const {
        CpAsync cp;
        const int t = item & 1023;
        if (is_z(item)) {
            if constexpr (DIR < 0) { typename S::ZFwd op{plane(item), t * S::CZ}; op.load(raw, threadIdx.x, kFftThreads, cp); }
            else { typename S::ZInv op{plane(item), t * S::CZ}; op.load(raw, threadIdx.x, kFftThreads, cp); }
        } else {
            typename S::template YPass<DIR> op{reinterpret_cast<V*>(plane(item)), t * S::CY};
            op.load(raw, threadIdx.x, kFftThreads, cp);
        }
    }
    template <class Op>
    __device__ __forceinline__ void run(const Op& op, const V* raw, V* work) const {
#pragma unroll
        for (int ph = 0; ph < Op::kPhases; ++ph) {
            op.phase(ph, raw, work, tw, threadIdx.x, kFftThreads);
            if (ph + 1 < Op::kPhases) __syncthreads();
        }
    }
    __device__ __forceinline__ void process(int item, const V* raw, V* work) const {
        const int t = item & 1023;
        if (is_z(item)) {
            if constexpr (DIR < 0) { typename S::ZFwd op{plane(item), t * S::CZ}; run(op, raw, work); }
            else { typename S::ZInv op{plane(item), t * S::CZ}; run(op, raw, work); }
        } else {
            typename S::template YPass<DIR> op{reinterpret_cast<V*>(plane(item)), t * S::CY};
            run(op, raw, work);
        }
    }
};

template <typename T, int G, int DIR>
__global__ void __launch_bounds__(kFftThreads, 1) fft2d_kernel(const __grid_constant__ Fft2dParams p) {
    using S = SlabFFT<T, G>;
    using V = typename S::V;
    extern __shared__ __align__(16) unsigned char fft_smem[];
    V* tw = reinterpret_cast<V*>(fft_smem);
    V* ring0 = tw + G;
    V* ring1 = ring0 + S::kRawElems;
    V* work = ring1 + S::kRawElems;
    for (int m = threadIdx.x; m < G; m += kFftThreads) tw[m] = reinterpret_cast<const V*>(p.tw)[m];
    Fft2dJob<T, G, DIR> job(p, tw);
    run_pipeline(job, p.ticket, ring0, ring1, work, p.err);
}

// ---------------------------------------------------------------------------------------------
// x solve
// ---------------------------------------------------------------------------------------------
template <typename T, int G>
struct XSolveKParams {
    typename SlabFFT<T, G>::XGeom xg;
    const void* tw;
    int j0, njl;
    unsigned* ticket;
    int* err;
};

template <typename T, int G>
struct XSolveJob {
    using S = SlabFFT<T, G>;
    using V = typename S::V;
    const typename S::XGeom* xg;   // in shared memory
    const V* tw;
    int j0, njl;
    __device__ __forceinline__ int decode(unsigned slot) const {
        return slot < (unsigned)(njl * S::kYTilesPerPlane) ? (int)slot : -1;
    }
    __device__ __forceinline__ const unsigned* dep(int) const { return nullptr; }
    __device__ __forceinline__ unsigned dep_need() const { return 0; }
    __device__ __forceinline__ unsigned* signal(int) const { return nullptr; }
    __device__ __forceinline__ typename S::XSolve op(int item) const {
        const int jl = item / S::kYTilesPerPlane, t = item - jl * S::kYTilesPerPlane;
        return typename S::XSolve{xg, j0 + jl, t * S::CY};
    }
    __device__ __forceinline__ void issue(int item, V* raw) const {
        CpAsync cp;
        op(item).load(raw, threadIdx.x, kFftThreads, cp);
    }
    __device__ __forceinline__ void process(int item, const V* raw, V* work) const {
        const typename S::XSolve o = op(item);
#pragma unroll
        for (int ph = 0; ph < S::XSolve::kPhases; ++ph) {
            o.phase(ph, raw, work, tw, threadIdx.x, kFftThreads);
            if (ph + 1 < S::XSolve::kPhases) __syncthreads();
        }
    }
};

template <typename T, int G>
__global__ void __launch_bounds__(kFftThreads, 1) xsolve2_kernel(const __grid_constant__ XSolveKParams<T, G> p) {
    using S = SlabFFT<T, G>;
    using V = typename S::V;
    extern __shared__ __align__(16) unsigned char fft_smem[];
    V* tw = reinterpret_cast<V*>(fft_smem);
    V* ring0 = tw + G;
    V* ring1 = ring0 + S::kYTileElems;
    V* work = ring1 + S::kYTileElems;
    double* sep = reinterpret_cast<double*>(work + S::kYTileElems);
    __shared__ typename S::XGeom s_xg;
    for (int m = threadIdx.x; m < G; m += kFftThreads) {
        tw[m] = reinterpret_cast<const V*>(p.tw)[m];
        sep[m] = p.xg.sep[m];
    }
    if (threadIdx.x == 0) {
        s_xg = p.xg;
        s_xg.sep = sep;
    }
    __syncthreads();
    XSolveJob<T, G> job{&s_xg, tw, p.j0, p.njl};
    run_pipeline(job, p.ticket, ring0, ring1, work, p.err);
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
    std::vector<V> tw(G);
    for (int m = 0; m < G; ++m) {
        const long double a = -2.0L * 3.14159265358979323846264338327950288L * m / G;
        tw[m].x = (T)cosl(a);
        tw[m].y = (T)sinl(a);
    }
    PM_CHECK_CUDA(cudaMalloc(&c->f2_tw, sizeof(V) * G));
    PM_CHECK_CUDA(cudaMemcpy(c->f2_tw, tw.data(), sizeof(V) * G, cudaMemcpyHostToDevice));
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
    c->f2_lag = 6;
    if (const char* e = getenv("PM_FFT_LAG")) c->f2_lag = std::max(1, atoi(e));
    return PM_OK;
}

template <typename T, int G>
static size_t fft2d_smem() {
    using S = SlabFFT<T, G>;
    return sizeof(typename S::V) * ((size_t)G + 2 * S::kRawElems + S::kWorkElems);
}
template <typename T, int G>
static size_t xsolve2_smem() {
    using S = SlabFFT<T, G>;
    return sizeof(typename S::V) * ((size_t)G + 3 * S::kYTileElems) + sizeof(double) * G;
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
    const int grid = (int)std::min<int64_t>(tiles, kNumSMs);
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
    p.ticket = c->f2_ctr + 1;
    p.err = reinterpret_cast<int*>(c->f2_ctr + c->f2_nctr);
    const size_t smem = xsolve2_smem<T, G>();
    PM_CHECK_CUDA(cudaFuncSetAttribute(xsolve2_kernel<T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t tiles = (int64_t)g.njl * S::kYTilesPerPlane;
    const int grid = (int)std::min<int64_t>(tiles, kNumSMs);
    PM_LAUNCH((xsolve2_kernel<T, G>), grid, kFftThreads, smem, c->stream, p);
    return PM_OK;
}

template <typename T, int G>
static int solve_fft2_tg(pm_ctx* c, double prefactor, bool l2_fused) {
    PM_CHECK_CUDA(cudaMemsetAsync(c->f2_ctr, 0, sizeof(unsigned) * c->f2_nctr, c->stream));
    if (l2_fused) {
        PM_TRY((launch_fft2d<T, G, -1>(c, 0)));
    } else {
        PM_TRY((launch_fft2d<T, G, -1>(c, 1)));
        PM_TRY((launch_fft2d<T, G, -1>(c, 2)));
    }
    if (c->nranks > 1) PM_TRY(device_barrier(c));   // every rank's 2-D spectra are complete
    PM_TRY((launch_xsolve2<T, G>(c, prefactor)));
    if (c->nranks > 1) PM_TRY(device_barrier(c));   // all peers have written our planes
    if (l2_fused) {
        PM_TRY((launch_fft2d<T, G, +1>(c, 0)));
    } else {
        PM_TRY((launch_fft2d<T, G, +1>(c, 1)));
        PM_TRY((launch_fft2d<T, G, +1>(c, 2)));
    }
    return PM_OK;
}

// forward 2-D transforms → fused x pass → inverse 2-D transforms with the hand-written kernels.
// l2_fused: dependency-ordered single launch per 2-D transform (fp64 only: the 8-byte cp.async of the
// fp32 tiles goes through L1, which is not coherent across the in-kernel hand-over).
int solve_fft2(pm_ctx* c, double prefactor, int deconv_order, double gauss, bool l2_fused) {
    PM_REQUIRE(fft2_supported(c) && c->f2_tw != nullptr, "hand-written FFT path not available for this grid size / rank layout");
    PM_TRY(update_sep_table(c, deconv_order, gauss));
    const bool f64 = c->dtype == PM_GRID_F64;
    if (!f64) l2_fused = false;
    int s;
    switch (c->g.G) {
        case 128: s = f64 ? solve_fft2_tg<double, 128>(c, prefactor, l2_fused) : solve_fft2_tg<float, 128>(c, prefactor, l2_fused); break;
        case 256: s = f64 ? solve_fft2_tg<double, 256>(c, prefactor, l2_fused) : solve_fft2_tg<float, 256>(c, prefactor, l2_fused); break;
        default:  s = f64 ? solve_fft2_tg<double, 512>(c, prefactor, l2_fused) : solve_fft2_tg<float, 512>(c, prefactor, l2_fused); break;
    }
    return s;
}

// the pipeline's give-up flag (a dependency wait that ran into its spin limit); host-synchronising
int fft2_check_error(pm_ctx* c) {
    if (c->f2_ctr == nullptr) return PM_OK;
    int e = 0;
    PM_CHECK_CUDA(cudaMemcpyAsync(&e, c->f2_ctr + c->f2_nctr, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PM_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    PM_REQUIRE(e == 0, "hand-written FFT: a tile dependency was not satisfied in time (results are invalid)");
    return PM_OK;
}

}  // namespace pm
