// pm_fft.cu — hand-written slab transform for the fused Poisson solve (sm_100a, HBM/L2/shared-memory
// bound; no tensor cores: fp64/fp32 butterflies on the FMA pipes).
//
//   fft2d_kernel<DIR=−1>   per x plane: r2c along z (ZFwd tiles: real rows -> A), then c2c along y (YFwd: A -> B)
//   xsolve2_kernel         c2c along x · Green's function · inverse c2c along x (B of all ranks -> A of all
//                          ranks through peer pointers; no materialised transpose, fft.c:34-73)
//   fft2d_kernel<DIR=+1>   per x plane: inverse c2c along y (YInv: A -> real rows), then c2r along z (ZInv)
//
// Replaces fft.c's FFTW-MPI plans (fft.c:105-290) + the potential loop (interactions.py:2092-2118) on the
// default gravity path.  Tile operations, layouts and their index math live in pm_fftops.cuh (CPU-checked
// by tests/test_fftcore_host.py); this file adds the asynchronous tile pipeline around them.
//
// Data path of a y or x tile: the intermediate layouts A and B make the tile ONE contiguous chunk of global
// memory (per rank), so a single thread moves it with cp.async.bulk (the TMA engine; completion on an
// mbarrier) into one of the CTA's two 32 KB buffers while the other buffer is being transformed in place; the
// last stage stores aligned 64-byte segments straight from registers.  No load instruction, register or
// scoreboard is spent on the input, and three such CTAs per SM interleave their LDS / FP64 / STS / barrier
// phases.  z tiles read and write whole contiguous rows directly; their "load" is a cp.async.bulk.prefetch.L2
// of the next tile's rows.  A blocks that a forward y tile has consumed are dropped from L2
// (discard.global.L2) instead of being written back, and on one rank the forward z tiles nullify the density
// rows they have consumed (self-cleaning grid, pm_internal.cuh).
// (History, measured on B200: cp.async/LDGSTS staging costs 16 shared-memory wavefronts per 16-byte warp
// instruction — 38 % of the shared-memory pipe; direct ld.global into registers leaves only ~32 KB per SM
// in flight and the passes latency-bound at 0.50-0.55 ms each.)
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

#ifndef PM_FFT2D_THREADS
#define PM_FFT2D_THREADS 128
#endif
#ifndef PM_FFT2D_DOUBLE_BUFFER
#define PM_FFT2D_DOUBLE_BUFFER 0
#endif
#ifndef PM_FFT_Z_RPW_FWD
#define PM_FFT_Z_RPW_FWD 1
#endif
#ifndef PM_FFT_Z_RPW_INV
#define PM_FFT_Z_RPW_INV 2
#endif
#ifndef PM_XSOLVE_THREADS
#define PM_XSOLVE_THREADS 128
#endif
#ifndef PM_XSOLVE_DOUBLE_BUFFER
#define PM_XSOLVE_DOUBLE_BUFFER 1
#endif
constexpr bool kFft2dDouble = PM_FFT2D_DOUBLE_BUFFER != 0;   // second tile buffer (bulk copy of the next y tile under the current one)
#ifndef PM_FFT2D_OCC
#define PM_FFT2D_OCC 4
#endif
#ifndef PM_FFT2D_OCC_F32
#define PM_FFT2D_OCC_F32 PM_FFT2D_OCC     /* fp32 grids: half the registers per value */
#endif
#ifndef PM_XSOLVE_OCC
#define PM_XSOLVE_OCC 3
#endif
#ifndef PM_XSOLVE_OCC_F32
#define PM_XSOLVE_OCC_F32 2
#endif
// Kernel shapes per grid size.  Small CTAs, several per SM (independent barrier domains hide each other's LDS / FP64 / STS
// phases); a y or x tile is G × 64 bytes of columns, so G = 1024 doubles the tile (64 KB) and takes twice the threads.
// Measured on B200 at 512³ fp64 (profiles/r02_fft_config_sweep.md): 2-D transforms 128 threads × 4 CTAs, one tile
// buffer; x solve 128 threads × 3 CTAs, two tile buffers.
template <typename T, int G>
struct FftCfg {
    static constexpr bool kBig = G >= 1024;
    static constexpr bool kF32 = sizeof(T) == 4;      // fp32 x solve: 256 threads × 2 CTAs measured faster (0.38 vs 0.43 ms)
    static constexpr int kThreads2d = kBig ? 256 : PM_FFT2D_THREADS;     // every warp owns a z row
    static constexpr int kOcc2d = kBig ? 2 : (kF32 ? PM_FFT2D_OCC_F32 : PM_FFT2D_OCC);
    static constexpr int kThreadsX = (kBig || kF32) ? 256 : PM_XSOLVE_THREADS;
    static constexpr int kOccX = kBig ? 1 : (kF32 ? PM_XSOLVE_OCC_F32 : PM_XSOLVE_OCC);      // radix-16 stage: 16 complex values + their twiddles per thread
    // z rows per warp (their loads are in flight together): measured 1 for the forward pass (0.82 vs 0.87 ms), 2 for the
    // inverse (0.73 vs 0.81 ms)
    static constexpr int kRowsFwd = PM_FFT_Z_RPW_FWD, kRowsInv = PM_FFT_Z_RPW_INV;
    static constexpr bool kDoubleX = kBig ? true : (PM_XSOLVE_DOUBLE_BUFFER != 0);
};

// ---- PTX helpers: mbarrier + bulk asynchronous copy ----------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// orders this CTA's earlier generic-proxy accesses to shared memory before later async-proxy (bulk copy) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global bulk copy (bulk-group completion); the source must stay untouched until wait_group.read
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(__cvta_generic_to_global(gdst)), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ bool mbar_test(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded: a copy that never lands raises the error flag instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity, int* err) {
    unsigned spins = 0;
    while (!mbar_test(bar, parity)) {
        if (++spins > 64) __nanosleep(40);
        if (spins > (1u << 22)) { atomicExch(err, 2); break; }
    }
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---------------------------------------------------------------------------------------------
// persistent, ticket-ordered, double-buffered tile loop
// ---------------------------------------------------------------------------------------------
// Job interface:
//   int  decode(unsigned slot)        item (>= 0), −2: empty slot, −1: past the end
//   const unsigned* dep(int item)     completion counter the item waits for (nullptr: none)
//   unsigned dep_need()
//   void issue(int item, V* buf, uint64_t* bar)   thread 0: expect_tx + bulk copies of the item's tile
//   void process(int item, V* buf)    phases with __syncthreads() in between, in place in buf; the last
//                                     phase stores to global memory from registers
//   unsigned* signal(int item)        counter to bump once the item's stores are complete (nullptr: none)
// While tile t is processed in one buffer, thread 0 has already taken the next ticket and — if that
// tile's dependency is met — started its copy into the other buffer.  A CTA blocks on a dependency only
// between tiles (holding no unfinished work); dependencies point to strictly lower tickets and first-pass
// tiles have none, so the CTA with the lowest blocked ticket can always proceed.
template <bool DOUBLE, class Job, typename V>
__device__ __forceinline__ void run_tiles(Job& job, unsigned* ticket, V* buf0, V* buf1, int* err) {
    __shared__ __align__(8) uint64_t mbar[2];
    __shared__ int s_next[2], s_loaded[2];
    const int tid = threadIdx.x;
    auto fetch = [&]() -> int {
        int item;
        do { item = job.decode(atomicAdd(ticket, 1u)); } while (item == -2);
        if (item >= 0 && *reinterpret_cast<volatile int*>(err) != 0) item = -1;   // something already failed: drain
        return item;
    };
    auto ready = [&](int item) -> bool {
        const unsigned* d = job.dep(item);
        return d == nullptr || ld_acquire(d) >= job.dep_need();
    };
    auto wait_dep = [&](int item) {
        unsigned spins = 0;
        while (!ready(item)) {
            __nanosleep(100);
            if (++spins > (1u << 24)) { atomicExch(err, 1); break; }   // seconds: give up loudly, never hang
        }
    };
    auto issue = [&](int item, V* buf, uint64_t* bar) {
        fence_proxy_async();
        job.issue(item, buf, bar);
    };
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_mbar_init();
        const int item = fetch();
        if (item >= 0) {
            wait_dep(item);
            issue(item, buf0, &mbar[0]);
        }
        s_next[1] = item;
    }
    __syncthreads();
    int cur = s_next[1];
    for (int it = 0; cur >= 0; ++it) {
        const int b = it & 1;
        V* bufc = (DOUBLE && b) ? buf1 : buf0;
        V* bufn = (DOUBLE && !b) ? buf1 : buf0;
        if (tid == 0) {
            // take the next ticket now; start its copy at once if it does not need the buffer in use (second buffer, or a
            // tile that is read directly and only prefetched into L2) and its dependency is met
            const int nxt = fetch();
            int loaded = 0;
            if (nxt >= 0 && (DOUBLE || !job.uses_buffer(nxt)) && ready(nxt)) {
                issue(nxt, bufn, &mbar[b ^ 1]);     // bufn: tile it−1 finished before the barrier that ended it
                loaded = 1;
            }
            s_next[b] = nxt;
            s_loaded[b] = loaded;
        }
        mbar_wait(&mbar[b], (unsigned)(it >> 1) & 1u, err);   // buffer b's (it/2)-th fill
        job.process(cur, bufc);
        __syncthreads();                 // end of tile: every thread's stores are issued, bufc is free, s_next[b] visible
        unsigned* sig = job.signal(cur);
        if (tid == 0 && sig != nullptr) {
            __threadfence();             // … and visible device-wide before the counter moves
            atomicAdd(sig, 1u);
        }
        const int nxt = s_next[b];
        if (nxt >= 0 && !s_loaded[b] && tid == 0) {
            wait_dep(nxt);
            issue(nxt, bufn, &mbar[b ^ 1]);
        }
        cur = nxt;
    }
}

// copy the twiddle tables to shared memory; visible after the first barrier of run_tiles.  The butterflies form their
// twiddles as powers of ONE root (tw_powers), so only rows 0 … 2 of the C table are ever read (TWS ∈ {1, 2}):
// B[64] | C rows 0-2 [192] | R[G/8 + 1]
constexpr int kTwCRows = 3;
template <typename V>
__device__ __forceinline__ Twiddles<V> load_twiddles(V* smem, const void* gmem, int G, bool with_r, int nthr) {
    const V* g = reinterpret_cast<const V*>(gmem);
    const int crows = G / 64 < kTwCRows ? G / 64 : kTwCRows;
    for (int m = threadIdx.x; m < 64 + 64 * crows; m += nthr) smem[m] = g[m];
    if (with_r)
        for (int m = threadIdx.x; m < G / 8 + 1; m += nthr) smem[64 + 64 * kTwCRows + m] = g[64 + G + m];
    Twiddles<V> tw;
    tw.B = smem;
    tw.C = smem + 64;
    tw.R = smem + 64 + 64 * kTwCRows;
    return tw;
}
template <int G> constexpr int smem_twiddle_entries() { return 64 + 64 * kTwCRows + G / 8 + 1; }

template <class Op, typename V, typename T, int NREGS, int NTHR>
__device__ __forceinline__ void run_phases(const Op& op, V* buf, const Twiddles<V>& tw) {
    T rg[NREGS];
#pragma unroll
    for (int ph = 0; ph < Op::kPhases; ++ph) {
        op.phase(ph, buf, tw, threadIdx.x, NTHR, rg);
        if (ph + 1 < Op::kPhases) {
            if constexpr (Op::kWarpPrivate) __syncwarp();     // the warp works on its own row: no block barrier
            else __syncthreads();
        }
    }
}

// tiles that are read directly: ask the copy engine to pull the chunk into L2 and complete the phase at once
template <class Op>
__device__ __forceinline__ void issue_prefetch(const Op& op, uint64_t* bar) {
    const TileLoad l = op.load(0);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(__cvta_generic_to_global(l.src)), "r"((unsigned)l.bytes) : "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <class Op, typename V>
__device__ __forceinline__ void issue_loads(const Op& op, int nloads, V* buf, uint64_t* bar) {
    unsigned total = 0;
    for (int k = 0; k < nloads; ++k) total += (unsigned)op.load(k).bytes;
    mbar_expect_tx(bar, total);
    for (int k = 0; k < nloads; ++k) {
        const TileLoad l = op.load(k);
        bulk_g2s(reinterpret_cast<char*>(buf) + l.dst_bytes, l.src, (unsigned)l.bytes, bar);
    }
}

// ---------------------------------------------------------------------------------------------
// 2-D (y,z) transforms of the local planes
// ---------------------------------------------------------------------------------------------
struct Fft2dParams {
    void* interior;       // first interior plane of the padded real slab (forward: input; inverse: output)
    int clear_input;      // forward: nullify the real rows once read
    void* a;              // A: V[nxl][NKT][G][CY]
    void* b;              // B: V[NKT][G][nxl][CY]
    const void* tw;       // B | C | R twiddle tables (pm_fftcore.cuh)
    int nplanes;
    int mode;             // 0: both passes, dependency-ordered (L2-resident); 1: first pass only; 2: second pass only
    int lag;              // planes between the first and the second pass in ticket order (mode 0)
    int discard;          // forward y tiles drop the consumed A block from L2
    unsigned* ticket;
    unsigned* done;       // [nplanes] finished first-pass tiles
    int* err;
};

template <typename T, int G, int DIR>
struct Fft2dJob {
    using S = SlabFFT<T, G, FftCfg<T, G>::kThreads2d, (DIR < 0 ? FftCfg<T, G>::kRowsFwd : FftCfg<T, G>::kRowsInv)>;
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
    __device__ __forceinline__ bool is_z(int item) const { return (DIR < 0) == ((item >> 30) == 0); }
    __device__ __forceinline__ bool uses_buffer(int item) const { return !is_z(item); }   // z tiles are read directly (L2 prefetch only)
    __device__ __forceinline__ int plane_of(int item) const { return (item >> 10) & 0xfffff; }
    __device__ __forceinline__ T* plane(int item) const {
        return reinterpret_cast<T*>(p.interior) + (size_t)plane_of(item) * G * S::Gp;
    }
    __device__ __forceinline__ V* a_plane(int item) const {
        return reinterpret_cast<V*>(p.a) + (size_t)plane_of(item) * S::NKT * G * S::CY;
    }
    __device__ __forceinline__ typename S::ZFwd zfwd(int item) const { return typename S::ZFwd{plane(item), a_plane(item), (item & 1023) * S::CZ, p.clear_input != 0}; }
    __device__ __forceinline__ typename S::ZInv zinv(int item) const { return typename S::ZInv{plane(item), (item & 1023) * S::CZ}; }
    __device__ __forceinline__ typename S::YFwd yfwd(int item) const {
        const int kt = item & 1023;
        return typename S::YFwd{a_plane(item) + (size_t)kt * G * S::CY, reinterpret_cast<V*>(p.b), plane_of(item), kt, p.nplanes};
    }
    __device__ __forceinline__ typename S::YInv yinv(int item) const {
        const int kt = item & 1023;
        return typename S::YInv{a_plane(item) + (size_t)kt * G * S::CY, plane(item), kt};
    }
    __device__ __forceinline__ void issue(int item, V* buf, uint64_t* bar) const {
        if (is_z(item)) {
            if constexpr (DIR < 0) issue_prefetch(zfwd(item), bar);
            else issue_prefetch(zinv(item), bar);
        } else {
            if constexpr (DIR < 0) issue_loads(yfwd(item), 1, buf, bar);
            else issue_loads(yinv(item), 1, buf, bar);
        }
    }
    __device__ __forceinline__ void process(int item, V* buf) const {
        if (is_z(item)) {
            // (self-cleaning density grid: the forward z tiles nullify the rows they have consumed — plain stores;
            // bulk stores from a block of zeros in shared memory were measured slower, 1.01 vs 0.97 ms)
            if constexpr (DIR < 0) run_phases<typename S::ZFwd, V, T, S::kRegs, FftCfg<T, G>::kThreads2d>(zfwd(item), buf, tw);
            else run_phases<typename S::ZInv, V, T, S::kRegs, FftCfg<T, G>::kThreads2d>(zinv(item), buf, tw);
        } else {
            if constexpr (DIR < 0) {
                // the A block is in shared memory now and nobody reads it again before the x solve rewrites
                // it: drop its (dirty) L2 lines instead of letting them be written back to HBM
                const typename S::YFwd op = yfwd(item);
                const char* blk = reinterpret_cast<const char*>(op.a_tile);
                if (p.discard)
                for (int o = threadIdx.x * 128; o < G * S::CY * (int)sizeof(V); o += FftCfg<T, G>::kThreads2d * 128)
                    asm volatile("discard.global.L2 [%0], 128;" ::"l"(__cvta_generic_to_global(blk + o)) : "memory");
                run_phases<typename S::YFwd, V, T, S::kRegs, FftCfg<T, G>::kThreads2d>(op, buf, tw);
            } else {
                run_phases<typename S::YInv, V, T, S::kRegs, FftCfg<T, G>::kThreads2d>(yinv(item), buf, tw);
            }
        }
    }
};

template <typename T, int G, int DIR>
__global__ void __launch_bounds__(FftCfg<T, G>::kThreads2d, FftCfg<T, G>::kOcc2d) fft2d_kernel(const __grid_constant__ Fft2dParams p) {
    using S = typename Fft2dJob<T, G, DIR>::S;
    using V = typename S::V;
    extern __shared__ __align__(128) unsigned char fft_smem[];
    V* buf0 = reinterpret_cast<V*>(fft_smem);
    V* buf1 = buf0 + (kFft2dDouble ? S::kBufElems : 0);
    const Twiddles<V> tw = load_twiddles(buf1 + S::kBufElems, p.tw, G, true, FftCfg<T, G>::kThreads2d);
    Fft2dJob<T, G, DIR> job(p, tw);
    run_tiles<kFft2dDouble>(job, p.ticket, buf0, buf1, p.err);
}

// ---------------------------------------------------------------------------------------------
// x solve
// ---------------------------------------------------------------------------------------------
template <typename T, int G>
struct XSolveKParams {
    typename SlabFFT<T, G, FftCfg<T, G>::kThreadsX>::XGeom xg;
    const void* tw;
    int j0, njl;
    unsigned* ticket;
    int* err;
    int bulk_out;      // in place: stage a tile's results in shared memory and send each rank its piece as one bulk copy
};

template <typename T, int G>
struct XSolveJob {
    using S = SlabFFT<T, G, FftCfg<T, G>::kThreadsX>;
    using V = typename S::V;
    const typename S::XGeom* xg;   // in shared memory
    Twiddles<V> tw;
    int j0, njl;
    V* stage;      // nullptr: results go out as 64-byte stores from registers
    // column tile fastest: consecutive tickets share a j row
    __device__ __forceinline__ int decode(unsigned slot) const {
        return slot < (unsigned)(njl * S::NKT) ? (int)slot : -1;
    }
    __device__ __forceinline__ const unsigned* dep(int) const { return nullptr; }
    __device__ __forceinline__ unsigned dep_need() const { return 0; }
    __device__ __forceinline__ unsigned* signal(int) const { return nullptr; }
    __device__ __forceinline__ bool uses_buffer(int) const { return true; }
    __device__ __forceinline__ typename S::XSolve op(int item) const {
        const int jl = item / S::NKT, kt = item - jl * S::NKT;
        return typename S::XSolve{xg, j0 + jl, kt, stage};
    }
    __device__ __forceinline__ void issue(int item, V* buf, uint64_t* bar) const {
        const typename S::XSolve o = op(item);
        issue_loads(o, o.nloads(), buf, bar);
    }
    __device__ __forceinline__ void process(int item, V* buf) const {
        const typename S::XSolve o = op(item);
        const bool bulk = stage != nullptr && xg->in_place;
        // the copy engine has read the previous tile's pieces out of the staging tile (the barriers between the phases
        // order this wait before the last phase writes there again)
        if (bulk && (int)threadIdx.x < xg->nranks) bulk_wait_read();
        run_phases<typename S::XSolve, V, T, S::kRegs, FftCfg<T, G>::kThreadsX>(o, buf, tw);
        if (bulk) {
            fence_proxy_async();
            __syncthreads();
            if ((int)threadIdx.x < xg->nranks) {
                const int r = threadIdx.x, nxl = 1 << xg->nxl_shift;
                bulk_s2g(const_cast<V*>(xg->b[r]) + S::b_index(o.kt, o.j, 0, 0, nxl), stage + (size_t)r * nxl * S::CY,
                         (unsigned)(nxl * S::CY * sizeof(V)));
            }
        }
    }
};

template <typename T, int G>
__global__ void __launch_bounds__(FftCfg<T, G>::kThreadsX, FftCfg<T, G>::kOccX) xsolve2_kernel(const __grid_constant__ XSolveKParams<T, G> p) {
    using S = SlabFFT<T, G, FftCfg<T, G>::kThreadsX>;
    constexpr bool kXSolveDouble = FftCfg<T, G>::kDoubleX;
    constexpr int kFftThreads = FftCfg<T, G>::kThreadsX;
    using V = typename S::V;
    extern __shared__ __align__(128) unsigned char fft_smem[];
    V* buf0 = reinterpret_cast<V*>(fft_smem);
    V* buf1 = buf0 + (kXSolveDouble ? S::kYTileElems : 0);
    const Twiddles<V> tw = load_twiddles(buf1 + S::kYTileElems, p.tw, G, false, kFftThreads);
    double* sep = reinterpret_cast<double*>(buf1 + S::kYTileElems + 64 + 64 * kTwCRows);
    __shared__ typename S::XGeom s_xg;
    for (int m = threadIdx.x; m < G; m += kFftThreads) sep[m] = p.xg.sep[m];
    if (threadIdx.x == 0) {
        s_xg = p.xg;
        s_xg.sep = sep;
    }
    V* stage = nullptr;
    if (p.bulk_out) stage = reinterpret_cast<V*>((reinterpret_cast<uintptr_t>(sep + G) + 127) & ~(uintptr_t)127);
    XSolveJob<T, G> job{&s_xg, tw, p.j0, p.njl, stage};
    run_tiles<kXSolveDouble>(job, p.ticket, buf0, buf1, p.err);
    if (stage != nullptr && (int)threadIdx.x < p.xg.nranks) bulk_wait_all();
}

// B [kt][j][il][c] -> A [il][kt][j][c] on this rank (64-byte elements; 16 × 16 of them through shared memory so that both
// sides move 1 KB runs): the hand-over to the inverse y pass when the x solve wrote its results in place into B.
__global__ void __launch_bounds__(256)
b_to_a_kernel(const uint4* __restrict__ b, uint4* __restrict__ a, int nxl, int rows /* NKT·G */) {
    __shared__ uint4 tile[16][16 * 4 + 1];
    const int tiles_il = nxl / 16, tiles_r = rows / 16;
    for (int t = blockIdx.x; t < tiles_il * tiles_r; t += gridDim.x) {
        const int il0 = (t % tiles_il) * 16, r0 = (t / tiles_il) * 16;
        __syncthreads();
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {         // row r0 + rr: 16 planes × 4 uint4 contiguous
            const int rr = e / 64, w = e % 64;
            tile[rr][w] = b[((size_t)(r0 + rr) * nxl + il0) * 4 + w];
        }
        __syncthreads();
        for (int e = threadIdx.x; e < 16 * 64; e += 256) {         // plane il0 + pl: 16 rows × 4 uint4 contiguous
            const int pl = e / 64, w = e % 64;
            a[((size_t)(il0 + pl) * rows + r0) * 4 + w] = tile[w / 4][pl * 4 + (w % 4)];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool fft2_supported(const pm_ctx* c) {
    const int G = c->g.G;
    if (!(G == 128 || G == 256 || G == 512 || G == 1024)) return false;
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
    if (!(G == 128 || G == 256 || G == 512 || G == 1024)) return PM_OK;
    PM_TRY(c->dtype == PM_GRID_F64 ? make_fft2_tables_t<double>(c) : make_fft2_tables_t<float>(c));
    // counters: [0] ticket fwd, [1] ticket x, [2] ticket inv, [3] unused, [4 …) done fwd, then done inv;
    // one more entry after those f2_nctr: the sticky error flag
    c->f2_nctr = 4 + 2 * (size_t)c->g.nxl;
    PM_CHECK_CUDA(cudaMalloc(&c->f2_ctr, sizeof(unsigned) * (c->f2_nctr + 1)));
    PM_CHECK_CUDA(cudaMemset(c->f2_ctr, 0, sizeof(unsigned) * (c->f2_nctr + 1)));
    // the intermediate layouts A and B live behind the real slab in the same allocation (pm_create),
    // so that one CUDA-IPC handle exposes all three to the peers
    c->f2_a = reinterpret_cast<char*>(c->real) + c->f2_off_a;
    c->f2_b = reinterpret_cast<char*>(c->real) + c->f2_off_b;
    // planes between the two passes in ticket order; > CTAs in flight / tiles per plane.  Measured on B200 at 512³ fp64
    // (profiles/r02_fft_config_sweep.md): forward 2/3/4/5/6/8/10 -> 0.92/0.87/0.84/0.82/0.82/0.86/0.89 ms,
    // inverse (two rows per warp) 6/8/10/12/16/20 -> 0.78/0.74/0.73/0.72/0.77/0.80 ms
    c->f2_lag = 5;
    c->f2_lag_inv = 12;
    if (const char* e = getenv("PM_FFT_LAG")) c->f2_lag = c->f2_lag_inv = std::max(1, atoi(e));
    if (const char* e = getenv("PM_FFT_LAG_FWD")) c->f2_lag = std::max(1, atoi(e));
    if (const char* e = getenv("PM_FFT_LAG_INV")) c->f2_lag_inv = std::max(1, atoi(e));
    return PM_OK;
}

template <typename T, int G, int DIR>
static size_t fft2d_smem() {
    using S = typename Fft2dJob<T, G, DIR>::S;
    return sizeof(typename S::V) * ((size_t)(kFft2dDouble ? 2 : 1) * S::kBufElems + smem_twiddle_entries<G>());
}
template <typename T, int G>
static size_t xsolve2_smem() {
    using S = SlabFFT<T, G, FftCfg<T, G>::kThreadsX>;
    return sizeof(typename S::V) * ((size_t)(FftCfg<T, G>::kDoubleX ? 2 : 1) * S::kYTileElems + 64 + 64 * kTwCRows) + sizeof(double) * G;
}

template <typename T, int G, int DIR>
static int launch_fft2d(pm_ctx* c, int mode) {
    using S = typename Fft2dJob<T, G, DIR>::S;
    Fft2dParams p;
    // one rank: the density grid cleans itself and the potential goes to `phi` (pm_internal.cuh)
    const bool self_clean = c->nranks == 1 && c->phi != nullptr;
    p.interior = (DIR > 0 && self_clean) ? c->phi : static_cast<void*>(c->real_interior<T>());
    p.clear_input = (DIR < 0 && self_clean) ? 1 : 0;
    p.a = c->f2_a;
    p.b = c->f2_b;
    p.tw = c->f2_tw;
    p.nplanes = c->g.nxl;
    p.mode = mode;
    p.lag = DIR < 0 ? c->f2_lag : c->f2_lag_inv;
    static const int no_discard = getenv("PM_FFT_NO_DISCARD") ? atoi(getenv("PM_FFT_NO_DISCARD")) : 0;
    p.discard = !no_discard;
    p.ticket = c->f2_ctr + (DIR < 0 ? 0 : 2);
    p.done = c->f2_ctr + 4 + (DIR < 0 ? 0 : c->g.nxl);
    p.err = reinterpret_cast<int*>(c->f2_ctr + c->f2_nctr);
    const size_t smem = fft2d_smem<T, G, DIR>();
    PM_CHECK_CUDA(cudaFuncSetAttribute(fft2d_kernel<T, G, DIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t tiles = (int64_t)c->g.nxl * (mode == 0 ? S::kZTilesPerPlane + S::kYTilesPerPlane
                                                         : ((mode == 1) == (DIR < 0) ? S::kZTilesPerPlane : S::kYTilesPerPlane));
    const int grid = (int)std::min<int64_t>(tiles, (int64_t)kNumSMs * FftCfg<T, G>::kOcc2d);
    if (mode != 0) PM_CHECK_CUDA(cudaMemsetAsync(p.ticket, 0, sizeof(unsigned), c->stream));
    constexpr int kThreads = FftCfg<T, G>::kThreads2d;
    PM_LAUNCH((fft2d_kernel<T, G, DIR>), grid, kThreads, smem, c->stream, p);
    return PM_OK;
}

template <typename T, int G>
static int launch_xsolve2(pm_ctx* c, double prefactor) {
    using S = SlabFFT<T, G, FftCfg<T, G>::kThreadsX>;
    using V = typename S::V;
    const Geom& g = c->g;
    XSolveKParams<T, G> p;
    for (int r = 0; r < kMaxFftPeers; ++r) { p.xg.a[r] = nullptr; p.xg.b[r] = nullptr; }
    // (timing experiments only — the results are wrong: PM_X_LOCAL bit 0 keeps all stores, bit 1 all loads on this rank)
    static const int x_local = getenv("PM_X_LOCAL") ? atoi(getenv("PM_X_LOCAL")) : 0;
    for (int r = 0; r < c->nranks; ++r) {
        char* base = reinterpret_cast<char*>(c->nranks == 1 ? c->real : c->peer_real[r]);
        char* own = reinterpret_cast<char*>(c->real);
        p.xg.a[r] = reinterpret_cast<V*>(((x_local & 1) ? own : base) + c->f2_off_a);
        p.xg.b[r] = reinterpret_cast<const V*>(((x_local & 2) ? own : base) + c->f2_off_b);
    }
    p.xg.nranks = c->nranks;
    // several ranks: results return to the B pieces they came from — contiguous per peer, so a tile's results are staged in
    // shared memory and every peer gets its nxl·64 bytes as ONE bulk copy — then one local re-layout pass B -> A.
    // Measured on 8 B200 (x solve incl. barriers and re-layout): G = 512  0.505 ms (64-byte stores from registers straight into
    // the peers' A) -> 0.433 ms; G = 1024  3.44 ms (in place, 64-byte stores) -> 3.16 ms; on 2 B200: 1.017 -> 0.965, 8.14 -> 7.42 ms.
    static const int x_inplace = getenv("PM_X_INPLACE") ? atoi(getenv("PM_X_INPLACE")) : -1;
    p.xg.in_place = x_inplace >= 0 ? (x_inplace != 0 && g.nxl % 16 == 0) : (c->nranks > 1 && g.nxl % 16 == 0);
    p.xg.nxl_shift = 0;
    while ((1 << p.xg.nxl_shift) < g.nxl) ++p.xg.nxl_shift;
    p.xg.sep = c->xs_sep;
    p.xg.prefactor = prefactor;
    p.tw = c->f2_tw;
    p.j0 = g.j0;
    p.njl = g.njl;
    p.ticket = c->f2_ctr + 1;
    p.err = reinterpret_cast<int*>(c->f2_ctr + c->f2_nctr);
    static const int x_bulk = getenv("PM_X_BULK") ? atoi(getenv("PM_X_BULK")) : 1;
    p.bulk_out = (p.xg.in_place && x_bulk) ? 1 : 0;
    const size_t smem = xsolve2_smem<T, G>() + (p.bulk_out ? 128 + sizeof(V) * S::kYTileElems : 0);
    PM_CHECK_CUDA(cudaFuncSetAttribute(xsolve2_kernel<T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t tiles = (int64_t)g.njl * S::NKT;
    const int grid = (int)std::min<int64_t>(tiles, (int64_t)kNumSMs * FftCfg<T, G>::kOccX);
    constexpr int kThreads = FftCfg<T, G>::kThreadsX;
    PM_LAUNCH((xsolve2_kernel<T, G>), grid, kThreads, smem, c->stream, p);
    c->f2_x_in_place = p.xg.in_place != 0;
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
        if (c->f2_x_in_place) {
            using S = SlabFFT<T, G, FftCfg<T, G>::kThreadsX>;
            const int rows = S::NKT * G;
            PM_LAUNCH(b_to_a_kernel, kNumSMs * 8, 256, 0, c->stream, reinterpret_cast<const uint4*>(c->f2_b),
                      reinterpret_cast<uint4*>(c->f2_a), c->g.nxl, rows);
        }
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
    if (c->nranks == 1) {
        if (stage == 0 || stage == 1) PM_TRY(ensure_in_real(c));     // the input is the density in `real`
        if (c->phi == nullptr && getenv("PM_NO_SELF_CLEAN") == nullptr) {
            if (cudaMalloc(&c->phi, c->real_elems * c->elem_size()) == cudaSuccess) c->bytes_allocated += c->real_elems * c->elem_size();
            else { c->phi = nullptr; cudaGetLastError(); }           // no room: keep the in-place scheme
        }
    }
    int s;
    switch (c->g.G) {
        case 128: s = f64 ? solve_fft2_tg<double, 128>(c, prefactor, l2_fused, stage) : solve_fft2_tg<float, 128>(c, prefactor, l2_fused, stage); break;
        case 256: s = f64 ? solve_fft2_tg<double, 256>(c, prefactor, l2_fused, stage) : solve_fft2_tg<float, 256>(c, prefactor, l2_fused, stage); break;
        case 512: s = f64 ? solve_fft2_tg<double, 512>(c, prefactor, l2_fused, stage) : solve_fft2_tg<float, 512>(c, prefactor, l2_fused, stage); break;
        default:  s = f64 ? solve_fft2_tg<double, 1024>(c, prefactor, l2_fused, stage) : solve_fft2_tg<float, 1024>(c, prefactor, l2_fused, stage); break;
    }
    if (s == PM_OK && c->nranks == 1 && c->phi != nullptr) {
        if (stage == 0 || stage == 1) c->real_is_zero = true;     // the forward z pass has nullified what it read
        if (stage == 0 || stage == 3) c->grid_in_phi = true;      // the potential is in `phi`
    } else if (s == PM_OK && (stage == 0 || stage == 1)) {
        c->real_is_zero = false;
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
