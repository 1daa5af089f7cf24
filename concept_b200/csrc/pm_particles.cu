// pm_particles.cu — streaming particle kernels: drift, Σ mom², slab migration.
//
// Reference semantics (file:line under the reference's src/):
//   drift     species.py:2179-2199  pos = mod(pos + mom·Δ, boxsize);  mod: commons.py:5102-5131
//   v_rms     analysis.py:3965-3972 Σ mom²
//   exchange  communication.py:135-517 (owner = which_domain, :756-772) — here for x-slabs
#include "pm_internal.cuh"

#include <algorithm>

namespace pm {

// np.mod(x, L) with the reference's x == L → 0 guard
__device__ __forceinline__ double mod_box(double x, double L) {
    double r = fmod(x, L);
    if (r < 0) r += L;
    if (r == L) r = 0;
    return r;
}

__global__ void __launch_bounds__(256)
drift_kernel(double* __restrict__ pos, const double* __restrict__ mom, int64_t n3, double dt_over_mass,
             double L) {
    // 3N contiguous doubles; 128-bit vector body + scalar tail
    const int64_t n2 = n3 >> 1;
    double2* p2 = reinterpret_cast<double2*>(pos);
    const double2* m2 = reinterpret_cast<const double2*>(mom);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2;
         i += (int64_t)gridDim.x * blockDim.x) {
        double2 p = p2[i];
        const double2 m = __ldg(m2 + i);
        p.x = mod_box(p.x + m.x * dt_over_mass, L);
        p.y = mod_box(p.y + m.y * dt_over_mass, L);
        p2[i] = p;
    }
    if ((n3 & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        pos[n3 - 1] = mod_box(pos[n3 - 1] + mom[n3 - 1] * dt_over_mass, L);
    }
}

__global__ void __launch_bounds__(256)
drift_kernel_scalar(double* __restrict__ pos, const double* __restrict__ mom, int64_t n3,
                    double dt_over_mass, double L) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n3;
         i += (int64_t)gridDim.x * blockDim.x)
        pos[i] = mod_box(pos[i] + mom[i] * dt_over_mass, L);
}

int launch_drift(pm_ctx* c, double* pos, const double* mom, int64_t n, double dt_over_mass) {
    if (n == 0) return PM_OK;
    const int64_t n3 = 3 * n;
    const bool aligned = ((reinterpret_cast<uintptr_t>(pos) | reinterpret_cast<uintptr_t>(mom)) & 15) == 0;
    const int grid = kNumSMs * 8;
    if (aligned) {
        PM_LAUNCH(drift_kernel, grid, 256, 0, c->stream, pos, mom, n3, dt_over_mass, c->boxsize);
    } else {
        PM_LAUNCH(drift_kernel_scalar, grid, 256, 0, c->stream, pos, mom, n3, dt_over_mass, c->boxsize);
    }
    return PM_OK;
}

__global__ void __launch_bounds__(256)
sum_mom2_kernel(const double* __restrict__ mom, int64_t n3, double* __restrict__ out) {
    __shared__ double sred[8];
    double acc = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n3;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double m = __ldg(mom + i);
        acc += m * m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < 8; ++w) t += sred[w];
        atomicAdd(out, t);
    }
}

int launch_sum_mom2(pm_ctx* c, const double* mom, int64_t n, double* out) {
    if (n == 0) return PM_OK;
    PM_LAUNCH(sum_mom2_kernel, kNumSMs * 4, 256, 0, c->stream, mom, 3 * n, out);
    return PM_OK;
}

// ---------------------------------------------------------------------------
// slab migration
// ---------------------------------------------------------------------------
// exchange(component) (communication.py:135-517) for x-slabs, entirely on the device and without NCCL:
//   pack      every particle whose owner is another rank is written straight into that rank's mailbox slot for this
//             source (peer stores over NVLink through the CUDA-IPC mapping) and its index is recorded as a hole;
//   publish   the per-destination counts go into the peers' arena headers; flag barrier;
//   plan / partition / fill   arrivals drop into the holes the movers left (the reference fills holes the same way,
//             communication.py:430-517); surplus arrivals are appended, surplus holes are filled from the tail.
// The host learns the new particle count from one 32-byte read-back at the end (the only synchronisation; the
// GPU never waits for the host).  A mailbox slot that is too small makes the exchange take another round; a particle
// buffer that is too small leaves the arrivals in the mailbox and returns PM_ERR_OVERFLOW on that rank only — the
// caller grows its buffers and calls again, which then only unpacks (no collective step is repeated, so nothing hangs).
// Particles far from the slab faces never move: in the (x-major) lattice or cell order the movers and therefore the
// holes sit at both ends of the array, where the arrivals belong, so the memory order the mesh kernels rely on survives.
// Owner of a particle: the rank whose x-slab holds its cell, cell = int(x·G/L).
__device__ __forceinline__ int owner_of(double x, double cells_per_len, int G, int nxl) {
    int cell = (int)(x * cells_per_len);
    if (cell < 0) cell = 0;
    if (cell >= G) cell = G - 1;
    return cell / nxl;
}

struct XchgArrays {        // device arrays that travel with a particle (any but pos/mom may be NULL)
    double* pos; double* mom; int64_t* ids; double* dmom; signed char* rung; signed char* rungj;
};

// a mailbox slot of capacity K records, structure of arrays
struct SlotView { double* pos; double* mom; int64_t* ids; double* dmom; signed char* rung; signed char* rungj; };
constexpr size_t kRecordBytes = 24 + 24 + 8 + 24 + 2;
__host__ __device__ __forceinline__ SlotView slot_view(char* base, int64_t K) {
    SlotView v;
    v.pos = reinterpret_cast<double*>(base);
    v.mom = v.pos + 3 * K;
    v.ids = reinterpret_cast<int64_t*>(v.mom + 3 * K);
    v.dmom = reinterpret_cast<double*>(v.ids + K);
    v.rung = reinterpret_cast<signed char*>(v.dmom + 3 * K);
    v.rungj = v.rung + K;
    return v;
}

__device__ __forceinline__ void copy_record(const XchgArrays& a, int64_t i, const SlotView& v, int64_t k) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { v.pos[3 * k + d] = a.pos[3 * i + d]; v.mom[3 * k + d] = a.mom[3 * i + d]; }
    if (a.ids) v.ids[k] = a.ids[i];
    if (a.dmom) {
#pragma unroll
        for (int d = 0; d < 3; ++d) v.dmom[3 * k + d] = a.dmom[3 * i + d];
    }
    if (a.rung) { v.rung[k] = a.rung[i]; v.rungj[k] = a.rungj[i]; }
}
__device__ __forceinline__ void take_record(const XchgArrays& a, int64_t i, const SlotView& v, int64_t k) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { a.pos[3 * i + d] = v.pos[3 * k + d]; a.mom[3 * i + d] = v.mom[3 * k + d]; }
    if (a.ids) a.ids[i] = v.ids[k];
    if (a.dmom) {
#pragma unroll
        for (int d = 0; d < 3; ++d) a.dmom[3 * i + d] = v.dmom[3 * k + d];
    }
    if (a.rung) { a.rung[i] = v.rung[k]; a.rungj[i] = v.rungj[k]; }
}
__device__ __forceinline__ void move_record(const XchgArrays& a, int64_t from, int64_t to) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { a.pos[3 * to + d] = a.pos[3 * from + d]; a.mom[3 * to + d] = a.mom[3 * from + d]; }
    if (a.ids) a.ids[to] = a.ids[from];
    if (a.dmom) {
#pragma unroll
        for (int d = 0; d < 3; ++d) a.dmom[3 * to + d] = a.dmom[3 * from + d];
    }
    if (a.rung) { a.rung[to] = a.rung[from]; a.rungj[to] = a.rungj[from]; }
}

// device-side bookkeeping of one exchange (zeroed by a memset before the pack kernel)
struct XchgState {
    unsigned long long out_count[kMaxPeers];   // movers found per destination (may exceed the slot capacity)
    unsigned long long unsent;                 // movers that did not fit into a slot this round
    unsigned long long nholes;
    unsigned long long nlow;                   // holes below the new end of the array
    unsigned long long cursor;                 // tail particles moved so far
    // plan
    long long n_new, arrivals, err, unsent_total;
};

struct PeerSlots { char* slot[kMaxPeers]; };   // rank r's mailbox slot for THIS source (current parity)
struct PeerHeaders { ArenaHeader* hdr[kMaxPeers]; };

__global__ void __launch_bounds__(256)
xchg_pack_kernel(XchgArrays a, int64_t n, double cells_per_len, int G, int nxl, int rank, PeerSlots ps, int64_t K,
                 XchgState* __restrict__ st, unsigned int* __restrict__ holes) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = owner_of(a.pos[3 * i], cells_per_len, G, nxl);
        if (r == rank) continue;
        const unsigned long long k = atomicAdd(&st->out_count[r], 1ULL);
        if (k >= (unsigned long long)K) {            // slot full: this mover waits for the next round
            atomicAdd(&st->unsent, 1ULL);
            continue;
        }
        copy_record(a, i, slot_view(ps.slot[r], K), (int64_t)k);
        holes[atomicAdd(&st->nholes, 1ULL)] = (unsigned int)i;
    }
}

__global__ void __launch_bounds__(32)
xchg_publish_kernel(PeerHeaders ph, int rank, int nranks, int parity, int64_t K, const XchgState* __restrict__ st) {
    const int r = threadIdx.x;
    if (r >= nranks) return;
    const unsigned long long cnt = st->out_count[r] < (unsigned long long)K ? st->out_count[r] : (unsigned long long)K;
    ph.hdr[r]->mb_count[parity][rank].v = (r == rank) ? 0ULL : cnt;
    ph.hdr[r]->mb_unsent[parity][rank].v = st->unsent;
}

__global__ void xchg_plan_kernel(const ArenaHeader* __restrict__ mine, int nranks, int parity, int64_t n, int64_t capacity,
                                 XchgState* __restrict__ st) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    long long arrivals = 0, unsent = 0;
    for (int s = 0; s < nranks; ++s) {
        arrivals += (long long)mine->mb_count[parity][s].v;
        unsent += (long long)mine->mb_unsent[parity][s].v;
    }
    st->arrivals = arrivals;
    st->unsent_total = unsent;
    st->n_new = n - (long long)st->nholes + arrivals;
    st->err = st->n_new > capacity ? 1 : 0;
}

// holes below min(n, n_new) will be filled; holes in the tail [n_new, n) just disappear with the tail
__global__ void __launch_bounds__(256)
xchg_partition_kernel(int64_t n, XchgState* __restrict__ st, const unsigned int* __restrict__ holes,
                      unsigned int* __restrict__ lowholes, unsigned char* __restrict__ tail_is_hole) {
    if (st->err) return;
    const long long n_new = st->n_new;
    const long long lim = n_new < n ? n_new : n;
    const long long H = (long long)st->nholes;
    for (long long h = blockIdx.x * (long long)blockDim.x + threadIdx.x; h < H; h += (long long)gridDim.x * blockDim.x) {
        const unsigned int i = holes[h];
        if ((long long)i < lim) lowholes[atomicAdd(&st->nlow, 1ULL)] = i;
        else tail_is_hole[(long long)i - n_new] = 1;
    }
}

struct MySlots { char* slot[kMaxPeers]; };     // my mailbox slots by source rank (current parity)

__global__ void __launch_bounds__(256)
xchg_fill_kernel(XchgArrays a, int64_t n, int nranks, int parity, const ArenaHeader* __restrict__ mine, MySlots ms, int64_t K,
                 XchgState* __restrict__ st, const unsigned int* __restrict__ lowholes,
                 const unsigned char* __restrict__ tail_is_hole) {
    if (st->err) return;
    const long long A = st->arrivals, n_new = st->n_new, nlow = (long long)st->nlow;
    const long long T = n_new < n ? n - n_new : 0;     // tail to be vacated
    for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < A + T; w += (long long)gridDim.x * blockDim.x) {
        if (w < A) {
            // arrival w: which source, which record
            long long k = w;
            int s = 0;
            for (; s < nranks; ++s) {
                const long long cs = (long long)mine->mb_count[parity][s].v;
                if (k < cs) break;
                k -= cs;
            }
            const long long dest = w < nlow ? (long long)lowholes[w] : (long long)n + (w - nlow);
            take_record(a, dest, slot_view(ms.slot[s], K), k);
        } else {
            const long long t = w - A;                  // tail particle n_new + t
            if (tail_is_hole[t]) continue;
            const unsigned long long j = atomicAdd(&st->cursor, 1ULL);
            move_record(a, n_new + t, (long long)lowholes[A + (long long)j]);
        }
    }
}

// grows the scratch buffer; the first `keep` bytes (the state and hole list of a pending exchange) survive
static int ensure_xchg_scratch(pm_ctx* c, size_t need, size_t keep) {
    if (need <= c->xchg_bytes) return PM_OK;
    const size_t bytes = need + need / 8 + 4096;
    void* fresh = nullptr;
    PM_CHECK_CUDA(cudaMalloc(&fresh, bytes));
    if (c->xchg_buf) {
        if (keep) PM_CHECK_CUDA(cudaMemcpyAsync(fresh, c->xchg_buf, std::min(keep, c->xchg_bytes), cudaMemcpyDeviceToDevice, c->stream));
        PM_CHECK_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(c->xchg_buf);
        c->bytes_allocated -= c->xchg_bytes;
    }
    c->xchg_buf = fresh;
    c->xchg_bytes = bytes;
    c->bytes_allocated += bytes;
    return PM_OK;
}

// unpack half of an exchange round: plan, partition, fill, read-back.  *n_inout is updated on success.
static int xchg_unpack(pm_ctx* c, const XchgArrays& a, int64_t n, int64_t capacity, int parity, XchgState* st,
                       unsigned int* holes, unsigned int* lowholes, unsigned char* tail_mask, int64_t* n_out,
                       long long* unsent_total) {
    const int P = c->nranks;
    const int64_t K = (int64_t)(c->mailbox_slot_bytes / kRecordBytes) / 8 * 8;
    MySlots ms;
    for (int s = 0; s < kMaxPeers; ++s) ms.slot[s] = nullptr;
    for (int s = 0; s < P; ++s)
        ms.slot[s] = reinterpret_cast<char*>(c->real) + c->off_mailbox + ((size_t)parity * P + s) * c->mailbox_slot_bytes;
    const ArenaHeader* mine = arena_header(c, c->rank);
    // nlow and cursor start from zero (a repeated unpack after PM_ERR_OVERFLOW runs the same plan again)
    PM_CHECK_CUDA(cudaMemsetAsync(&st->nlow, 0, 2 * sizeof(unsigned long long), c->stream));
    PM_CHECK_CUDA(cudaMemsetAsync(tail_mask, 0, (size_t)n + 16, c->stream));
    PM_LAUNCH(xchg_plan_kernel, 1, 32, 0, c->stream, mine, P, parity, n, capacity, st);
    PM_LAUNCH(xchg_partition_kernel, kNumSMs * 2, 256, 0, c->stream, n, st, holes, lowholes, tail_mask);
    PM_LAUNCH(xchg_fill_kernel, kNumSMs * 2, 256, 0, c->stream, a, n, P, parity, mine, ms, K, st, lowholes, tail_mask);
    long long* h = reinterpret_cast<long long*>(c->h_pinned);
    PM_CHECK_CUDA(cudaMemcpyAsync(h, &st->n_new, 4 * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    PM_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    *unsent_total = h[3];
    if (h[2] != 0) {
        set_error("pm_exchange: %lld particles after migration exceed capacity %lld (grow the buffers and call again)",
                  h[0], (long long)capacity);
        return PM_ERR_OVERFLOW;
    }
    *n_out = (int64_t)h[0];
    return PM_OK;
}

int exchange_particles(pm_ctx* c, double* pos, double* mom, int64_t* ids, double* dmom, signed char* rung,
                       signed char* rung_jumped, int64_t* n_inout, int64_t capacity) {
    PM_REQUIRE(n_inout != nullptr, "pm_exchange: n_inout is NULL");
    if (c->nranks == 1) return PM_OK;
    PM_REQUIRE(c->peers_ready, "pm_exchange: call pm_ipc_open_peers first (the migration runs over the peer mappings)");
    PM_REQUIRE((rung == nullptr) == (rung_jumped == nullptr), "pm_exchange: rung and rung_jumped go together");
    const int P = c->nranks;
    int64_t n = *n_inout;
    PM_REQUIRE(n <= capacity && n < ((int64_t)1 << 32), "pm_exchange: n = %lld exceeds capacity %lld", (long long)n, (long long)capacity);
    const double cells_per_len = c->g.G / c->boxsize;
    const int64_t K = (int64_t)(c->mailbox_slot_bytes / kRecordBytes) / 8 * 8;
    PM_REQUIRE(K > 0, "pm_exchange: mailbox slots too small");
    const XchgArrays a{pos, mom, ids, dmom, rung, rung_jumped};
    // scratch: state | holes[cap] | lowholes[cap] | tail mask[cap]
    const size_t cap = (size_t)capacity + 64;
    PM_TRY(ensure_xchg_scratch(c, 1024 + cap * 9 + 64, c->xchg_pending ? 1024 + 4 * ((size_t)c->xchg_pending_n + 64) : 0));
    char* sb = reinterpret_cast<char*>(c->xchg_buf);
    XchgState* st = reinterpret_cast<XchgState*>(sb);
    unsigned int* holes = reinterpret_cast<unsigned int*>(sb + 1024);      // at most n entries, n <= capacity
    unsigned int* lowholes = holes + cap;
    unsigned char* tail_mask = reinterpret_cast<unsigned char*>(lowholes + cap);
    static_assert(sizeof(XchgState) <= 1024, "XchgState");
    if (c->xchg_pending) {
        // the previous call ran out of room: its arrivals are still in the mailbox and the plan is unchanged
        long long unsent = 0;
        PM_TRY(xchg_unpack(c, a, c->xchg_pending_n, capacity, c->xchg_pending_parity, st, holes, lowholes, tail_mask, &n, &unsent));
        c->xchg_pending = false;
        *n_inout = n;
        if (unsent == 0) return PM_OK;
    }
    for (int round = 0; round < 256; ++round) {
        const int parity = (int)(c->xchg_epoch++ & 1);
        PeerSlots ps;
        PeerHeaders ph;
        for (int r = 0; r < kMaxPeers; ++r) { ps.slot[r] = nullptr; ph.hdr[r] = nullptr; }
        for (int r = 0; r < P; ++r) {
            ps.slot[r] = reinterpret_cast<char*>(c->peer_real[r]) + c->off_mailbox + ((size_t)parity * P + c->rank) * c->mailbox_slot_bytes;
            ph.hdr[r] = arena_header(c, r);
        }
        PM_CHECK_CUDA(cudaMemsetAsync(st, 0, sizeof(XchgState), c->stream));
        if (n > 0)
            PM_LAUNCH(xchg_pack_kernel, kNumSMs * 4, 256, 0, c->stream, a, n, cells_per_len, c->g.G, c->g.nxl, c->rank, ps, K, st, holes);
        PM_LAUNCH(xchg_publish_kernel, 1, 32, 0, c->stream, ph, c->rank, P, parity, K, st);
        PM_TRY(device_barrier(c));
        long long unsent = 0;
        const int s = xchg_unpack(c, a, n, capacity, parity, st, holes, lowholes, tail_mask, &n, &unsent);
        if (s == PM_ERR_OVERFLOW) {
            c->xchg_pending = true;
            c->xchg_pending_n = n;
            c->xchg_pending_parity = parity;
            return s;
        }
        PM_TRY(s);
        *n_inout = n;
        if (unsent == 0) return PM_OK;     // the same total on every rank: all ranks leave the loop together
    }
    set_error("pm_exchange: movers left after 256 rounds (mailbox slots of %lld particles; raise PM_MAILBOX_MB)", (long long)K);
    return PM_ERR_OVERFLOW;
}

}  // namespace pm
