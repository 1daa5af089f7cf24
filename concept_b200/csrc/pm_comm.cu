// pm_comm.cu — the exchange steps of the x-slab decomposition over NCCL (NVLink 5 / NVSwitch).
//
// Reference counterparts (file:line under the reference's src/):
//   halo add / fill   communicate_ghosts(grid, '+=' | '=')   communication.py:563-660
//   slab transpose    FFTW-MPI's internal all-to-all          fft.c:34-73, 240-257
//   allreduce         analysis.py:3971, main.py:1202
// One rank owns G/P consecutive x planes and, in Fourier space, G/P consecutive j rows; the
// domain<->slab redistribution of mesh.py:2138-2411 does not exist in this layout.
#include "pm_internal.cuh"

#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace pm {

template <typename T> struct NcclType;
template <> struct NcclType<double> { static constexpr ncclDataType_t v = ncclDouble; };
template <> struct NcclType<float> { static constexpr ncclDataType_t v = ncclFloat; };

// ---------------------------------------------------------------------------
// halo planes
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) add_planes_kernel(T* __restrict__ dst, const T* __restrict__ src, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] += src[i];
}

constexpr int kDepositHalo = 2;  // PCS touches 2 planes beyond the slab on either side

template <typename T>
static int halo_add_t(pm_ctx* c) {
    const Geom& g = c->g;
    const size_t plane = (size_t)g.G * g.Gp;
    const size_t cnt = plane * kDepositHalo;
    T* base = reinterpret_cast<T*>(c->real);
    T* lo_halo = base + (size_t)(g.halo - kDepositHalo) * plane;          // my planes x0-2 .. x0-1
    T* hi_halo = base + (size_t)(g.halo + g.nxl) * plane;                 // my planes x0+nxl .. +1
    T* first = base + (size_t)g.halo * plane;                             // interior, low end
    T* last = base + (size_t)(g.halo + g.nxl - kDepositHalo) * plane;     // interior, high end
    // staging: reuse the far halo planes that the deposit never touches? keep it simple: sendbuf
    T* stage = reinterpret_cast<T*>(c->sendbuf);
    const int next = (c->rank + 1) % c->nranks, prev = (c->rank + c->nranks - 1) % c->nranks;
    PM_CHECK_NCCL(ncclGroupStart());
    PM_CHECK_NCCL(ncclSend(hi_halo, cnt, NcclType<T>::v, next, c->comm, c->stream));
    PM_CHECK_NCCL(ncclSend(lo_halo, cnt, NcclType<T>::v, prev, c->comm, c->stream));
    PM_CHECK_NCCL(ncclRecv(stage, cnt, NcclType<T>::v, prev, c->comm, c->stream));        // prev's hi halo
    PM_CHECK_NCCL(ncclRecv(stage + cnt, cnt, NcclType<T>::v, next, c->comm, c->stream));  // next's lo halo
    PM_CHECK_NCCL(ncclGroupEnd());
    PM_LAUNCH((add_planes_kernel<T>), kNumSMs * 4, 256, 0, c->stream, first, stage, cnt);
    PM_LAUNCH((add_planes_kernel<T>), kNumSMs * 4, 256, 0, c->stream, last, stage + cnt, cnt);
    return PM_OK;
}

// The same over the CUDA-IPC mappings of the neighbours' slabs (NVLink/NVSwitch peer loads, no NCCL launch, no
// staging copy): one kernel adds prev's upper halo planes onto my first interior planes and next's lower halo planes
// onto my last ones.  A flag barrier before it says that every rank's deposit is complete; the halo planes are only
// overwritten again by the next pm_grid_zero, which is at least one barrier later on every rank.
template <typename T>
__global__ void __launch_bounds__(256)
halo_add_peer_kernel(T* __restrict__ first, const T* __restrict__ prev_hi, T* __restrict__ last, const T* __restrict__ next_lo,
                     size_t n) {
    // n elements per side, a multiple of 2 (rows are Gp = G + 2 long); 16-byte vectors for f64
    using V = typename std::conditional<sizeof(T) == 8, double2, float2>::type;
    const size_t n2 = n / 2;
    V* f2 = reinterpret_cast<V*>(first);
    V* l2 = reinterpret_cast<V*>(last);
    const V* p2 = reinterpret_cast<const V*>(prev_hi);
    const V* q2 = reinterpret_cast<const V*>(next_lo);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < 2 * n2; i += (size_t)gridDim.x * blockDim.x) {
        if (i < n2) { V a = f2[i]; const V b = p2[i]; a.x += b.x; a.y += b.y; f2[i] = a; }
        else { const size_t k = i - n2; V a = l2[k]; const V b = q2[k]; a.x += b.x; a.y += b.y; l2[k] = a; }
    }
}

template <typename T>
static int halo_add_peer_t(pm_ctx* c) {
    const Geom& g = c->g;
    const size_t plane = (size_t)g.G * g.Gp;
    const size_t cnt = plane * kDepositHalo;
    const int next = (c->rank + 1) % c->nranks, prev = (c->rank + c->nranks - 1) % c->nranks;
    T* base = reinterpret_cast<T*>(c->real);
    const T* pbase = reinterpret_cast<const T*>(c->peer_real[prev]);
    const T* nbase = reinterpret_cast<const T*>(c->peer_real[next]);
    T* first = base + (size_t)g.halo * plane;
    T* last = base + (size_t)(g.halo + g.nxl - kDepositHalo) * plane;
    const T* prev_hi = pbase + (size_t)(g.halo + g.nxl) * plane;              // prev's planes x0 .. x0+1 of mine
    const T* next_lo = nbase + (size_t)(g.halo - kDepositHalo) * plane;       // next's planes below its slab = my last ones
    PM_TRY(device_barrier(c));
    PM_LAUNCH((halo_add_peer_kernel<T>), kNumSMs * 4, 256, 0, c->stream, first, prev_hi, last, next_lo, cnt);
    return PM_OK;
}

int halo_add(pm_ctx* c) {
    if (c->nranks == 1) return PM_OK;
    PM_REQUIRE(c->g.nxl >= kDepositHalo, "slab thinner than the deposit halo");
    if (c->peers_ready && getenv("PM_HALO_NCCL") == nullptr)
        return c->dtype == PM_GRID_F64 ? halo_add_peer_t<double>(c) : halo_add_peer_t<float>(c);
    PM_REQUIRE(c->comm_ready, "pm_halo_add: call pm_comm_init first");
    return c->dtype == PM_GRID_F64 ? halo_add_t<double>(c) : halo_add_t<float>(c);
}

template <typename T>
static int halo_fill_t(pm_ctx* c, int planes_lo, int planes_hi, int which) {
    const Geom& g = c->g;
    const size_t plane = (size_t)g.G * g.Gp;
    T* base = reinterpret_cast<T*>(which == PM_TAP_FORCE ? c->force : c->real);
    const int next = (c->rank + 1) % c->nranks, prev = (c->rank + c->nranks - 1) % c->nranks;
    PM_CHECK_NCCL(ncclGroupStart());
    // my first `planes_hi` interior planes become prev's high halo; my last `planes_lo` its next's low halo
    PM_CHECK_NCCL(ncclSend(base + (size_t)g.halo * plane, plane * planes_hi, NcclType<T>::v, prev, c->comm, c->stream));
    PM_CHECK_NCCL(ncclSend(base + (size_t)(g.halo + g.nxl - planes_lo) * plane, plane * planes_lo, NcclType<T>::v, next, c->comm, c->stream));
    PM_CHECK_NCCL(ncclRecv(base + (size_t)(g.halo + g.nxl) * plane, plane * planes_hi, NcclType<T>::v, next, c->comm, c->stream));
    PM_CHECK_NCCL(ncclRecv(base + (size_t)(g.halo - planes_lo) * plane, plane * planes_lo, NcclType<T>::v, prev, c->comm, c->stream));
    PM_CHECK_NCCL(ncclGroupEnd());
    return PM_OK;
}

// communicate_ghosts(grid, '=') over the peer mappings: my halo planes are read straight from the neighbours' interiors.
__global__ void __launch_bounds__(256)
halo_fill_peer_kernel(double2* __restrict__ lo_dst, const double2* __restrict__ lo_src, size_t n_lo,
                      double2* __restrict__ hi_dst, const double2* __restrict__ hi_src, size_t n_hi) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_lo + n_hi; i += (size_t)gridDim.x * blockDim.x) {
        if (i < n_lo) lo_dst[i] = lo_src[i];
        else hi_dst[i - n_lo] = hi_src[i - n_lo];
    }
}

template <typename T>
static int halo_fill_peer_t(pm_ctx* c, int planes_lo, int planes_hi) {
    const Geom& g = c->g;
    const size_t plane = (size_t)g.G * g.Gp;       // G·(G+2) elements: a multiple of 4 for even G, so 16-byte vectors fit
    const int next = (c->rank + 1) % c->nranks, prev = (c->rank + c->nranks - 1) % c->nranks;
    T* base = reinterpret_cast<T*>(c->real);
    const T* pbase = reinterpret_cast<const T*>(c->peer_real[prev]);
    const T* nbase = reinterpret_cast<const T*>(c->peer_real[next]);
    const size_t per16 = 16 / sizeof(T);
    PM_REQUIRE(plane % per16 == 0, "halo_fill: plane size not a multiple of 16 bytes");
    PM_TRY(device_barrier(c));                      // every rank's potential is complete
    PM_LAUNCH(halo_fill_peer_kernel, kNumSMs * 4, 256, 0, c->stream,
              reinterpret_cast<double2*>(base + (size_t)(g.halo - planes_lo) * plane),
              reinterpret_cast<const double2*>(pbase + (size_t)(g.halo + g.nxl - planes_lo) * plane), plane * planes_lo / per16,
              reinterpret_cast<double2*>(base + (size_t)(g.halo + g.nxl) * plane),
              reinterpret_cast<const double2*>(nbase + (size_t)g.halo * plane), plane * planes_hi / per16);
    c->peers_may_read = true;   // the neighbours read my interior the same way: no overwriting before the next barrier
    return PM_OK;
}

int halo_fill(pm_ctx* c, int planes_lo, int planes_hi, int which) {
    if (c->nranks == 1) return PM_OK;
    PM_REQUIRE(planes_lo <= c->g.halo && planes_hi <= c->g.halo && planes_lo <= c->g.nxl && planes_hi <= c->g.nxl,
               "halo_fill: %d/%d planes exceed halo %d or slab %d", planes_lo, planes_hi, c->g.halo, c->g.nxl);
    if (c->peers_ready && which == PM_TAP_REAL && !c->grid_in_phi && getenv("PM_HALO_NCCL") == nullptr)
        return c->dtype == PM_GRID_F64 ? halo_fill_peer_t<double>(c, planes_lo, planes_hi)
                                       : halo_fill_peer_t<float>(c, planes_lo, planes_hi);
    PM_REQUIRE(c->comm_ready, "pm_halo_fill: call pm_comm_init first");
    return c->dtype == PM_GRID_F64 ? halo_fill_t<double>(c, planes_lo, planes_hi, which)
                                   : halo_fill_t<float>(c, planes_lo, planes_hi, which);
}

// ---------------------------------------------------------------------------
// slab transpose
// ---------------------------------------------------------------------------
// 2-D spectra in the real buffer: complex [il][j][kk]  ->  send blocks [r][il][jl][kk], j = r·njl + jl
template <typename C>
__global__ void __launch_bounds__(256)
pack_kernel(const C* __restrict__ src, C* __restrict__ dst, int nxl, int G, int Gc, int njl, bool unpack) {
    const int64_t total = (int64_t)nxl * G * Gc;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int kk = (int)(idx % Gc);
        const int j = (int)((idx / Gc) % G);
        const int il = (int)(idx / ((int64_t)Gc * G));
        const int r = j / njl, jl = j - r * njl;
        const int64_t blk = (((int64_t)r * nxl + il) * njl + jl) * Gc + kk;
        if (unpack) const_cast<C*>(src)[idx] = dst[blk];
        else dst[blk] = src[idx];
    }
}

template <typename T, typename C>
static int transpose_t(pm_ctx* c, bool forward) {
    const Geom& g = c->g;
    const int P = c->nranks;
    const size_t blk = (size_t)g.nxl * g.njl * g.Gc;   // complex elements per (src, dst) pair
    C* spectra = reinterpret_cast<C*>(c->real_interior<T>());
    C* stage = reinterpret_cast<C*>(c->sendbuf);
    C* slab = reinterpret_cast<C*>(c->fourier);
    if (forward) {
        PM_LAUNCH((pack_kernel<C>), kNumSMs * 8, 256, 0, c->stream, spectra, stage, g.nxl, g.G, g.Gc, g.njl, false);
    }
    C* from = forward ? stage : slab;
    C* to = forward ? slab : stage;
    PM_CHECK_NCCL(ncclGroupStart());
    for (int r = 0; r < P; ++r) {
        if (r == c->rank) continue;
        PM_CHECK_NCCL(ncclSend(from + r * blk, 2 * blk, NcclType<T>::v, r, c->comm, c->stream));
        PM_CHECK_NCCL(ncclRecv(to + r * blk, 2 * blk, NcclType<T>::v, r, c->comm, c->stream));
    }
    PM_CHECK_NCCL(ncclGroupEnd());
    PM_CHECK_CUDA(cudaMemcpyAsync(to + c->rank * blk, from + c->rank * blk, blk * sizeof(C),
                                  cudaMemcpyDeviceToDevice, c->stream));
    if (!forward) {
        PM_LAUNCH((pack_kernel<C>), kNumSMs * 8, 256, 0, c->stream, spectra, stage, g.nxl, g.G, g.Gc, g.njl, true);
    }
    return PM_OK;
}

int transpose_forward(pm_ctx* c) {
    PM_REQUIRE(c->comm_ready, "FFT transpose: call pm_comm_init first");
    return c->dtype == PM_GRID_F64 ? transpose_t<double, double2>(c, true) : transpose_t<float, float2>(c, true);
}

int transpose_backward(pm_ctx* c) {
    PM_REQUIRE(c->comm_ready, "FFT transpose: call pm_comm_init first");
    return c->dtype == PM_GRID_F64 ? transpose_t<double, double2>(c, false) : transpose_t<float, float2>(c, false);
}

// Stream-ordered barrier over all ranks through the IPC arenas: lane r announces this rank's epoch in rank r's
// header (a release store over NVLink) and waits until rank r's epoch has arrived here.  Everything this rank's
// stream did before the barrier kernel — including its stores into peer memory — is complete when the kernel starts,
// and the acquire loads order the peers' data before whatever this stream launches next.  One 32-thread kernel,
// a few microseconds, instead of an NCCL all-reduce used as a barrier.  The wait is bounded: a peer that never arrives
// raises the sticky flag (pm_check_async_error) instead of hanging the GPU.
struct BarrierPeers { unsigned long long* flag[kMaxPeers]; };   // &header(r)->bar[my rank].v

__global__ void __launch_bounds__(32)
flag_barrier_kernel(BarrierPeers peers, const ArenaHeader* __restrict__ mine, int nranks, unsigned long long epoch, int* err) {
    const int r = threadIdx.x;
    if (r >= nranks) return;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peers.flag[r]), "l"(epoch) : "memory");
    const unsigned long long* f = &mine->bar[r].v;
    unsigned long long v;
    unsigned spins = 0;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
        if (v >= epoch) break;
        if (++spins > 64) __nanosleep(200);
        if (spins > (1u << 24)) { atomicExch(err, 3); break; }   // seconds
    }
}

int device_barrier(pm_ctx* c) {
    if (c->nranks == 1) return PM_OK;
    c->peers_may_read = false;
    if (c->peers_ready && getenv("PM_BARRIER_NCCL") == nullptr) {
        BarrierPeers bp;
        for (int r = 0; r < kMaxPeers; ++r) bp.flag[r] = nullptr;
        for (int r = 0; r < c->nranks; ++r) bp.flag[r] = &arena_header(c, r)->bar[c->rank].v;
        const unsigned long long epoch = ++c->bar_epoch;
        PM_LAUNCH(flag_barrier_kernel, 1, 32, 0, c->stream, bp, arena_header(c, c->rank), c->nranks, epoch, c->d_comm_err);
        return PM_OK;
    }
    PM_REQUIRE(c->comm_ready, "device barrier: call pm_comm_init first");
    double* token = c->d_scratch + 32;
    PM_CHECK_NCCL(ncclAllReduce(token, token, 1, ncclDouble, ncclSum, c->comm, c->stream));
    return PM_OK;
}

int barrier_before_overwrite(pm_ctx* c) {
    if (c->nranks == 1 || !c->peers_may_read) return PM_OK;
    return device_barrier(c);
}

}  // namespace pm

// ---------------------------------------------------------------------------
// C ABI: peer mappings of the slabs (CUDA IPC) for the fused x-solve
// ---------------------------------------------------------------------------
extern "C" int pm_ipc_get_handle(pm_ctx* c, void* handle_out_64) {
    PM_REQUIRE(c && handle_out_64, "pm_ipc_get_handle: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    cudaIpcMemHandle_t h;
    PM_CHECK_CUDA(cudaIpcGetMemHandle(&h, c->real));
    memcpy(handle_out_64, &h, sizeof(h));
    return PM_OK;
}

extern "C" int pm_ipc_open_peers(pm_ctx* c, const void* handles_nranks_x_64) {
    PM_REQUIRE(c && handles_nranks_x_64, "pm_ipc_open_peers: NULL argument");
    if (c->nranks == 1) return PM_OK;
    PM_REQUIRE(c->nranks <= pm::kMaxPeers, "pm_ipc_open_peers: more than %d ranks", pm::kMaxPeers);
    PM_REQUIRE(!c->peers_ready, "pm_ipc_open_peers: already opened");
    PM_CHECK_CUDA(cudaSetDevice(c->device));
    const unsigned char* hs = reinterpret_cast<const unsigned char*>(handles_nranks_x_64);
    for (int r = 0; r < c->nranks; ++r) {
        if (r == c->rank) {
            c->peer_real[r] = c->real;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, hs + (size_t)r * 64, sizeof(h));
        PM_CHECK_CUDA(cudaIpcOpenMemHandle(&c->peer_real[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    c->peers_ready = true;
    return PM_OK;
}

// ---------------------------------------------------------------------------
// C ABI: communicator
// ---------------------------------------------------------------------------
extern "C" int pm_comm_unique_id(void* id_out_128) {
    PM_REQUIRE(id_out_128 != nullptr, "pm_comm_unique_id: NULL");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    PM_CHECK_NCCL(ncclGetUniqueId(&id));
    memcpy(id_out_128, &id, sizeof(id));
    return PM_OK;
}

extern "C" int pm_comm_init(pm_ctx* c, const void* id_128) {
    PM_REQUIRE(c && id_128, "pm_comm_init: NULL argument");
    if (c->nranks == 1) return PM_OK;
    PM_REQUIRE(!c->comm_ready, "pm_comm_init: communicator already initialised");
    ncclUniqueId id;
    memcpy(&id, id_128, sizeof(id));
    PM_CHECK_CUDA(cudaSetDevice(c->device));
    PM_CHECK_NCCL(ncclCommInitRank(&c->comm, c->nranks, id, c->rank));
    c->comm_ready = true;
    return PM_OK;
}

extern "C" int pm_allreduce_sum(pm_ctx* c, double* dev_values, int n) {
    PM_REQUIRE(c && dev_values && n >= 0, "pm_allreduce_sum: bad argument");
    if (c->nranks == 1 || n == 0) return PM_OK;
    PM_REQUIRE(c->comm_ready, "pm_allreduce_sum: call pm_comm_init first");
    PM_CHECK_NCCL(ncclAllReduce(dev_values, dev_values, n, ncclDouble, ncclSum, c->comm, c->stream));
    return PM_OK;
}
