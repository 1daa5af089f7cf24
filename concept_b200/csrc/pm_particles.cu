// pm_particles.cu — streaming particle kernels: drift, Σ mom², slab migration.
//
// Reference semantics (file:line under the reference's src/):
//   drift     species.py:2179-2199  pos = mod(pos + mom·Δ, boxsize);  mod: commons.py:5102-5131
//   v_rms     analysis.py:3965-3972 Σ mom²
//   exchange  communication.py:135-517 (owner = which_domain, :756-772) — here for x-slabs
#include "pm_internal.cuh"

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
// Owner of a particle: the rank whose x-slab holds its cell, cell = int(x·G/L).
__device__ __forceinline__ int owner_of(double x, double cells_per_len, int G, int nxl) {
    int cell = (int)(x * cells_per_len);
    if (cell < 0) cell = 0;
    if (cell >= G) cell = G - 1;
    return cell / nxl;
}

constexpr int kXchgBlocks = kNumSMs * 4;   // contiguous chunks of the particle array, one per CTA
constexpr int kXchgThreads = 256;

// Pass 1: counts[r] = particles bound for rank r (global); stay[b] = stayers in chunk b.
__global__ void __launch_bounds__(kXchgThreads)
owner_count_kernel(const double* __restrict__ pos, int64_t n, double cells_per_len, int G, int nxl,
                   int rank, int nranks, unsigned long long* __restrict__ counts,
                   unsigned int* __restrict__ stay) {
    extern __shared__ unsigned int scount[];
    for (int r = threadIdx.x; r < nranks; r += blockDim.x) scount[r] = 0;
    __syncthreads();
    const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo = blockIdx.x * chunk, hi = min(n, lo + chunk);
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x)
        atomicAdd(&scount[owner_of(pos[3 * i], cells_per_len, G, nxl)], 1u);
    __syncthreads();
    for (int r = threadIdx.x; r < nranks; r += blockDim.x)
        if (scount[r]) atomicAdd(&counts[r], (unsigned long long)scount[r]);
    if (threadIdx.x == 0) stay[blockIdx.x] = scount[rank];
}

// exclusive scan of the per-chunk stayer counts (one CTA; kXchgBlocks is small)
__global__ void __launch_bounds__(1024) stay_scan_kernel(unsigned int* __restrict__ stay, int nblocks) {
    __shared__ unsigned int s[1024];
    unsigned int carry = 0;
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned int v = i < nblocks ? stay[i] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < 1024; off <<= 1) {
            const unsigned int t = threadIdx.x >= off ? s[threadIdx.x - off] : 0;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nblocks) stay[i] = carry + s[threadIdx.x] - v;
        carry += s[1023];
        __syncthreads();
    }
}

// Pass 2: ORDER-PRESERVING partition.  Stayers keep their relative order (the memory order of the
// particles — lattice / cell-sorted — is what makes deposit and gather cache-friendly, and it must not
// be scrambled by every migration); movers are grouped by destination rank, order irrelevant.
__global__ void __launch_bounds__(kXchgThreads)
owner_scatter_kernel(const double* __restrict__ pos, const double* __restrict__ mom,
                     const int64_t* __restrict__ ids, int64_t n, double cells_per_len, int G, int nxl, int rank,
                     const unsigned int* __restrict__ stay_offset,
                     const unsigned long long* __restrict__ offsets, unsigned long long* __restrict__ cursor,
                     double* __restrict__ pos_out, double* __restrict__ mom_out, int64_t* __restrict__ ids_out) {
    __shared__ unsigned int warp_tot[kXchgThreads / 32];
    __shared__ unsigned int running;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = stay_offset[blockIdx.x];
    __syncthreads();
    const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo = blockIdx.x * chunk, hi = min(n, lo + chunk);
    for (int64_t base = lo; base < hi; base += blockDim.x) {
        const int64_t i = base + threadIdx.x;
        const bool valid = i < hi;
        double x = 0, y = 0, z = 0;
        int r = -1;
        if (valid) {
            x = pos[3 * i]; y = pos[3 * i + 1]; z = pos[3 * i + 2];
            r = owner_of(x, cells_per_len, G, nxl);
        }
        const bool stays = valid && r == rank;
        const unsigned ball = __ballot_sync(0xffffffffu, stays);
        if (lane == 0) warp_tot[warp] = __popc(ball);
        __syncthreads();
        unsigned int before = 0, total = 0;
        for (int w = 0; w < kXchgThreads / 32; ++w) {
            const unsigned int t = warp_tot[w];
            if (w < warp) before += t;
            total += t;
        }
        int64_t slot = -1;
        if (stays) {
            slot = (int64_t)running + before + __popc(ball & ((1u << lane) - 1));
        } else if (valid) {
            const unsigned peers = __match_any_sync(__activemask(), r);
            const int leader = __ffs(peers) - 1;
            unsigned long long b0 = 0;
            if (lane == leader) b0 = atomicAdd(&cursor[r], (unsigned long long)__popc(peers));
            b0 = __shfl_sync(peers, b0, leader);
            slot = (int64_t)(offsets[r] + b0 + __popc(peers & ((1u << lane) - 1)));
        }
        if (slot >= 0) {
            pos_out[3 * slot] = x; pos_out[3 * slot + 1] = y; pos_out[3 * slot + 2] = z;
            mom_out[3 * slot] = mom[3 * i]; mom_out[3 * slot + 1] = mom[3 * i + 1]; mom_out[3 * slot + 2] = mom[3 * i + 2];
            if (ids) ids_out[slot] = ids[i];
        }
        __syncthreads();
        if (threadIdx.x == 0) running += total;
        __syncthreads();
    }
}

int exchange_particles(pm_ctx* c, double* pos, double* mom, int64_t* ids, int64_t* n_inout,
                       int64_t capacity) {
    PM_REQUIRE(n_inout != nullptr, "pm_exchange: n_inout is NULL");
    if (c->nranks == 1) return PM_OK;
    PM_REQUIRE(c->comm_ready, "pm_exchange: call pm_comm_init first");
    const int P = c->nranks;
    const int64_t n = *n_inout;
    PM_REQUIRE(n <= capacity, "pm_exchange: n = %lld exceeds capacity %lld", (long long)n, (long long)capacity);
    const double cells_per_len = c->g.G / c->boxsize;
    // device counters: counts[P], offsets[P], cursor[P], recv matrix [P*P]
    auto* d_counts = reinterpret_cast<unsigned long long*>(c->d_counts);
    auto* d_offsets = d_counts + P;
    auto* d_cursor = d_offsets + P;
    auto* d_matrix = d_cursor + P;
    auto* d_stay = reinterpret_cast<unsigned int*>(d_matrix + (size_t)P * P);   // [kXchgBlocks]
    PM_CHECK_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(unsigned long long) * 3 * P, c->stream));
    PM_CHECK_CUDA(cudaMemsetAsync(d_stay, 0, sizeof(unsigned int) * kXchgBlocks, c->stream));
    if (n > 0) {
        PM_LAUNCH(owner_count_kernel, kXchgBlocks, kXchgThreads, P * sizeof(unsigned int), c->stream, pos, n,
                  cells_per_len, c->g.G, c->g.nxl, c->rank, P, d_counts, d_stay);
        PM_LAUNCH(stay_scan_kernel, 1, 1024, 0, c->stream, d_stay, kXchgBlocks);
    }
    // everyone learns everyone's send counts
    PM_CHECK_NCCL(ncclAllGather(d_counts, d_matrix, P, ncclUint64, c->comm, c->stream));
    std::vector<unsigned long long> matrix((size_t)P * P);
    PM_CHECK_CUDA(cudaMemcpyAsync(matrix.data(), d_matrix, sizeof(unsigned long long) * P * P,
                                  cudaMemcpyDeviceToHost, c->stream));
    PM_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    const unsigned long long* mine = &matrix[(size_t)c->rank * P];
    // staging layout: [own | r = 0..P-1 except own]
    std::vector<unsigned long long> offsets(P);
    unsigned long long cur = mine[c->rank];
    offsets[c->rank] = 0;
    for (int r = 0; r < P; ++r) {
        if (r == c->rank) continue;
        offsets[r] = cur;
        cur += mine[r];
    }
    int64_t n_new = (int64_t)mine[c->rank];
    for (int r = 0; r < P; ++r)
        if (r != c->rank) n_new += (int64_t)matrix[(size_t)r * P + c->rank];
    if (n_new > capacity) {
        set_error("pm_exchange: %lld particles after migration exceed capacity %lld", (long long)n_new,
                  (long long)capacity);
        return PM_ERR_OVERFLOW;
    }
    // staging buffers (pos, mom, ids) for n particles
    const size_t need = (size_t)n * (3 + 3 + 1) * 8;
    if (need > c->xchg_bytes) {
        if (c->xchg_buf) { cudaFree(c->xchg_buf); c->bytes_allocated -= c->xchg_bytes; }
        const size_t bytes = need + need / 8 + 4096;
        PM_CHECK_CUDA(cudaMalloc(&c->xchg_buf, bytes));
        c->xchg_bytes = bytes;
        c->bytes_allocated += bytes;
    }
    double* spos = reinterpret_cast<double*>(c->xchg_buf);
    double* smom = spos + 3 * n;
    int64_t* sids = reinterpret_cast<int64_t*>(smom + 3 * n);
    PM_CHECK_CUDA(cudaMemcpyAsync(d_offsets, offsets.data(), sizeof(unsigned long long) * P,
                                  cudaMemcpyHostToDevice, c->stream));
    if (n > 0) {
        PM_LAUNCH(owner_scatter_kernel, kXchgBlocks, kXchgThreads, 0, c->stream, pos, mom, ids, n, cells_per_len,
                  c->g.G, c->g.nxl, c->rank, d_stay, d_offsets, d_cursor, spos, smom, ids ? sids : nullptr);
    }
    // own particles back to the front of the live arrays
    const int64_t n_own = (int64_t)mine[c->rank];
    PM_CHECK_CUDA(cudaMemcpyAsync(pos, spos, sizeof(double) * 3 * n_own, cudaMemcpyDeviceToDevice, c->stream));
    PM_CHECK_CUDA(cudaMemcpyAsync(mom, smom, sizeof(double) * 3 * n_own, cudaMemcpyDeviceToDevice, c->stream));
    if (ids) PM_CHECK_CUDA(cudaMemcpyAsync(ids, sids, sizeof(int64_t) * n_own, cudaMemcpyDeviceToDevice, c->stream));
    // movers: grouped send/recv, arrivals appended behind the own particles
    int64_t tail = n_own;
    PM_CHECK_NCCL(ncclGroupStart());
    for (int r = 0; r < P; ++r) {
        if (r == c->rank) continue;
        const int64_t ns = (int64_t)mine[r];
        const int64_t nr = (int64_t)matrix[(size_t)r * P + c->rank];
        if (ns) {
            PM_CHECK_NCCL(ncclSend(spos + 3 * offsets[r], 3 * ns, ncclDouble, r, c->comm, c->stream));
            PM_CHECK_NCCL(ncclSend(smom + 3 * offsets[r], 3 * ns, ncclDouble, r, c->comm, c->stream));
            if (ids) PM_CHECK_NCCL(ncclSend(sids + offsets[r], ns, ncclInt64, r, c->comm, c->stream));
        }
        if (nr) {
            PM_CHECK_NCCL(ncclRecv(pos + 3 * tail, 3 * nr, ncclDouble, r, c->comm, c->stream));
            PM_CHECK_NCCL(ncclRecv(mom + 3 * tail, 3 * nr, ncclDouble, r, c->comm, c->stream));
            if (ids) PM_CHECK_NCCL(ncclRecv(ids + tail, nr, ncclInt64, r, c->comm, c->stream));
            tail += nr;
        }
    }
    PM_CHECK_NCCL(ncclGroupEnd());
    *n_inout = n_new;
    return PM_OK;
}

}  // namespace pm
