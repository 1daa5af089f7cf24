// pm_internal.cuh — shared definitions for libpmgrav.so (sm_100a only).
// Not part of the public ABI (that is include/pmgrav.h).
#pragma once

#include <cuda_runtime.h>
#include <cufft.h>
#include <nccl.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "pmgrav.h"

namespace pm {

// machine_ϵ and the reference's ghost count that enters the (1±ε) coordinate fudge
// (reference mesh.py:1577-1606, commons.py:1814, :4411-4432)
constexpr double kEps = 2.220446049250313e-16;
constexpr int kNghostsRef = 2;
constexpr int kNumSMs = 148;  // B200
// x-halo planes kept on each side of a slab when nranks > 1:
// PCS reach (2) + interlacing shift (1) + 8th-order difference reach (4)
constexpr int kHalo = 7;
constexpr int kMaxPeers = 16;   // ranks whose slabs one kernel can address through peer pointers

// IPC-mapped arena behind the slab (nranks > 1): what the ranks hand each other without NCCL or the host —
// barrier flags, the migration mailboxes (one slot per source rank, double-buffered by exchange parity) and the
// read-only ghost particles of the P³M short-range force.  Every word a peer writes sits on its own 64-byte line.
struct alignas(64) ArenaWord { unsigned long long v; unsigned long long pad[7]; };
struct ArenaHeader {
    ArenaWord bar[kMaxPeers];             // barrier epoch announced by rank s
    ArenaWord mb_count[2][kMaxPeers];     // particles rank s has put into its slot of my mailbox (per parity)
    ArenaWord mb_unsent[2][kMaxPeers];    // movers rank s could not send this round (any destination)
    ArenaWord ghost_count[2][2];          // [parity][side]: ghost positions in my lower / upper ghost buffer
};
constexpr size_t kArenaHeaderBytes = 8192;
static_assert(sizeof(ArenaHeader) <= kArenaHeaderBytes, "arena header");

extern std::atomic<int64_t> g_launches;
void set_error(const char* fmt, ...);

#define PM_CHECK_CUDA(expr)                                                         \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            pm::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__,            \
                          cudaGetErrorName(_e), cudaGetErrorString(_e));            \
            return PM_ERR_CUDA;                                                     \
        }                                                                           \
    } while (0)

#define PM_CHECK_CUFFT(expr)                                                        \
    do {                                                                            \
        cufftResult _e = (expr);                                                    \
        if (_e != CUFFT_SUCCESS) {                                                  \
            pm::set_error("%s:%d cuFFT error %d", __FILE__, __LINE__, (int)_e);     \
            return PM_ERR_CUDA;                                                     \
        }                                                                           \
    } while (0)

#define PM_CHECK_NCCL(expr)                                                         \
    do {                                                                            \
        ncclResult_t _e = (expr);                                                   \
        if (_e != ncclSuccess) {                                                    \
            pm::set_error("%s:%d NCCL error %s", __FILE__, __LINE__,                \
                          ncclGetErrorString(_e));                                  \
            return PM_ERR_COMM;                                                     \
        }                                                                           \
    } while (0)

#define PM_REQUIRE(cond, ...)                                                       \
    do {                                                                            \
        if (!(cond)) {                                                              \
            pm::set_error(__VA_ARGS__);                                             \
            return PM_ERR_ARG;                                                      \
        }                                                                           \
    } while (0)

#define PM_TRY(expr)                                                                \
    do {                                                                            \
        int _s = (expr);                                                            \
        if (_s != PM_OK) return _s;                                                 \
    } while (0)

// Count + launch + check.  Every kernel of this library goes through here.
#define PM_LAUNCH(kernel, grid, block, smem, stream, ...)                           \
    do {                                                                            \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                 \
        pm::g_launches.fetch_add(1, std::memory_order_relaxed);                     \
        PM_CHECK_CUDA(cudaGetLastError());                                          \
    } while (0)

// Geometry of the local slab, passed by value to kernels
struct Geom {
    int G;        // global cells per side
    int Gp;       // padded z length of the real grid: 2*(G/2+1)
    int Gc;       // complex z length: G/2+1
    int nxl;      // local x planes
    int x0;       // global index of first local plane
    int halo;     // halo planes on each side in x (0 when nranks == 1)
    int wrap_x;   // 1: x wraps periodically in-kernel (nranks == 1)
    int njl;      // local j rows of the Fourier slab
    int j0;       // global index of first local j row
};

}  // namespace pm

struct pm_ctx {
    pm::Geom g;
    double boxsize;
    int dtype;      // PM_GRID_F64 / PM_GRID_F32
    int rank, nranks, device;
    cudaStream_t stream;
    bool own_stream;
    // buffers
    void* real;       // [(nxl+2*halo)][G][Gp] of T ; the interior doubles as the in-place FFT slab
    void* fourier;    // == interior of `real` when nranks == 1, else separate [G][njl][Gc] complex
    void* saved;      // lazily allocated copy of the Fourier slab
    void* force;      // lazily allocated scratch force grid, same shape as `real`
    void* sendbuf;    // nranks > 1: packed transpose blocks
    void* fft_work;   // shared cuFFT work area
    size_t fft_work_bytes;
    size_t real_elems;     // elements of T in `real`
    size_t fourier_elems;  // complex elements in the Fourier slab
    // k-space tables (device): x_l and sin(x_l) for l over signed wavenumbers / kk
    double* tab_x;     // [G]   x(k_signed(i)) = k·π/G + ε
    double* tab_sin;   // [G]
    // cuFFT
    cufftHandle plan_fwd, plan_bwd;        // nranks == 1: 3-D ; else batched 2-D
    cufftHandle plan_x;                    // nranks > 1: strided 1-D c2c along i
    cufftHandle plan2_fwd, plan2_bwd;      // nranks == 1: batched 2-D (y,z) plans for the fused x-solve path
    bool plan2_ready;
    int fft_chunk;                         // x planes per 2-D plan execution
    bool plans_ready;
    // NCCL
    ncclComm_t comm;
    bool comm_ready;
    // scratch for reductions / exchange
    double* d_scratch;        // small device scratch (>= 64 doubles)
    int64_t* d_counts;        // exchange counters
    unsigned long long* d_tilectr;   // dynamic tile counters of the particle kernels (4 entries)
    void* xchg_buf;           // staging for migrating particles
    size_t xchg_bytes;
    int64_t bytes_allocated;
    // fused x-solve (pm_xsolve.cu)
    double2* xs_tw;           // exp(−2πi·m/G)
    double* xs_sep;           // separable k-space factor per axis index (cached for one (deconv, gauss))
    int xs_sep_deconv;
    double xs_sep_gauss;
    bool fused_solve;         // use the fused path when supported
    int solve_mode;           // PM_SOLVE_* (pmgrav.h): which implementation pm_solve_fused / pm_kick_long use
    // hand-written slab transform (pm_fft.cu)
    void* f2_tw;              // twiddle tables B | C | R in the grid's precision
    void* f2_a;               // intermediate layouts A and B (pm_fftops.cuh), inside the `real` allocation
    void* f2_b;
    size_t f2_off_a, f2_off_b; // their byte offsets from `real` (the same on every rank)
    unsigned* f2_ctr;         // tickets, per-plane completion counters, error flag (last entry)
    size_t f2_nctr;
    int f2_lag, f2_lag_inv;   // planes between the two passes of the forward / inverse 2-D transform in ticket order
    bool f2_x_in_place;       // the last x solve left its results in B (several ranks): re-layout before the inverse y pass
    void* peer_real[pm::kMaxPeers];   // IPC mappings of every rank's `real` buffer (own pointer for self)
    bool peers_ready;
    // IPC arena (ArenaHeader | mailboxes | ghost buffers) inside the `real` allocation, same offsets on every rank
    size_t off_arena, arena_bytes;
    size_t off_mailbox, mailbox_slot_bytes;   // 2 parities × nranks slots
    size_t off_ghost, ghost_buf_bytes;        // 2 parities × 2 sides
    unsigned long long bar_epoch;     // barriers issued so far (every rank issues the same sequence)
    unsigned long long xchg_epoch;    // exchanges issued so far (mailbox parity)
    unsigned long long ghost_epoch;   // ghost exchanges issued so far (ghost-buffer parity)
    bool xchg_pending;                // an exchange ran out of particle-buffer room: arrivals wait in the mailbox
    int64_t xchg_pending_n;
    int xchg_pending_parity;
    bool peers_may_read;              // a peer may still be reading `real` through its mapping (halo fill): barrier before overwriting
    int* d_comm_err;                  // sticky flag of the flag barrier (a peer that never arrived)
    void* h_pinned;                   // small pinned host block for asynchronous read-backs
    // P3M short range scratch (pm_shortrange.cu)
    void* sr_buf;
    size_t sr_bytes;
    void* sr_tmp;
    size_t sr_tmp_bytes;
    bool sr_want_stats;       // the pair kernel counts pairs and candidates (pm_shortrange_stats)
    // pm_kick_long_host: copy streams and events of the chunked H2D / compute / D2H pipeline
    cudaStream_t s_h2d, s_d2h;
    cudaEvent_t ev_pipe[3 * 16];
    bool pipe_ready;
    // Self-cleaning density grid (one rank, hand-written transform): the forward z pass zeroes the rows of
    // `real` it has consumed and the inverse transforms deliver the potential into `phi`, so the next deposit
    // finds `real` already nullified and the 1 GB memset of get_buffer(nullify=True) disappears from the cycle.
    void* phi;                // lazily allocated, same shape as `real`
    bool real_is_zero;        // `real` is known to hold zeros
    bool grid_in_phi;         // the current real-space grid (what gathers and taps read) is `phi`
    // state flags
    bool space_fourier;       // working slab currently holds Fourier data

    size_t elem_size() const { return dtype == PM_GRID_F64 ? 8 : 4; }
    // the buffer that currently holds the real-space grid for readers (gather, diff, taps)
    void* grid_read() const { return grid_in_phi ? phi : real; }
    template <typename T> T* real_interior() const {
        return reinterpret_cast<T*>(real) + (size_t)g.halo * g.G * g.Gp;
    }
};

namespace pm {

// implemented in pm_mesh.cu
int launch_deposit(pm_ctx* c, const double* pos, int64_t n, int order, double contribution,
                   const double* shift);
int launch_gather(pm_ctx* c, int which, const double* pos, double* mom, int64_t n, int order,
                  int dim, double factor, const double* shift);
int launch_gather_kick(pm_ctx* c, double* pos, double* mom, int64_t n, int order,
                       int diff_order, double factor, const double* shift, double* sum_mom2,
                       bool drift = false, double drift_dt = 0.0);
int launch_diff(pm_ctx* c, int dim, int order);
int launch_halo_wrap_add(pm_ctx* c);   // single-rank no-op; multi-rank local part of halo add
// implemented in pm_fourier.cu
int make_plans(pm_ctx* c);
void destroy_plans(pm_ctx* c);
int fft_forward(pm_ctx* c);
int fft_backward(pm_ctx* c);
int launch_kspace(pm_ctx* c, double prefactor, int deconv_order, double gauss, double scale,
                  const double* shift, int diff_dim, bool from_saved, bool potential);
int slab_copy(pm_ctx* c, int mode);  // 0 save, 1 accumulate, 2 restore
int launch_power_k2(pm_ctx* c, int k2_max, double* power, unsigned long long* count);
// implemented in pm_particles.cu
int launch_drift(pm_ctx* c, double* pos, const double* mom, int64_t n, double dt_over_mass);
int launch_sum_mom2(pm_ctx* c, const double* mom, int64_t n, double* out);
int exchange_particles(pm_ctx* c, double* pos, double* mom, int64_t* ids, double* dmom, signed char* rung,
                       signed char* rung_jumped, int64_t* n_inout, int64_t capacity);
// implemented in pm_comm.cu
int halo_add(pm_ctx* c);
int halo_fill(pm_ctx* c, int planes_lo, int planes_hi, int which);
int transpose_forward(pm_ctx* c);   // real-buffer 2-D spectra -> Fourier slab (all-to-all)
int transpose_backward(pm_ctx* c);

int device_barrier(pm_ctx* c);      // stream-ordered barrier over all ranks (flags in the peers' arenas; NCCL before the mappings exist)
int barrier_before_overwrite(pm_ctx* c);   // barrier only if a peer may still be reading this rank's slab
inline pm::ArenaHeader* arena_header(const pm_ctx* c, int r) {
    return reinterpret_cast<pm::ArenaHeader*>(reinterpret_cast<char*>(c->peer_real[r]) + c->off_arena);
}
// implemented in pm_xsolve.cu
bool xsolve_supported(const pm_ctx* c);
int xsolve(pm_ctx* c, double prefactor, int deconv_order, double gauss);
int make_xsolve_tables(pm_ctx* c);
int update_sep_table(pm_ctx* c, int deconv_order, double gauss);
// implemented in pm_fft.cu
bool fft2_supported(const pm_ctx* c);
int make_fft2_tables(pm_ctx* c);
int solve_fft2(pm_ctx* c, double prefactor, int deconv_order, double gauss, bool l2_fused, int stage = 0);
int fft2_check_error(pm_ctx* c);
int solve_fused(pm_ctx* c, double prefactor, int deconv_order, double gauss, int stage = 0);   // pm_fourier.cu

int ensure_saved(pm_ctx* c);
int ensure_force(pm_ctx* c);
int ensure_in_real(pm_ctx* c);   // un-fused operators work on `real`: bring the grid back from `phi` if needed

}  // namespace pm
