// pm_shortrange.cu — P³M short-range pair force and the rung (adaptive time-step) bookkeeping.
//
// Reference semantics (file:line under the reference's src/):
//   pair kick      gravity.py:263-354  Δmom_i += (x_i − x_j)·table[int(r²·(T−1)/r²_max)]·G·m_i·m_j·ᔑdt_rungs[…][rung_i]
//                  for every pair within the range, receivers on active rungs only; inactive particles
//                  still supply (interactions.py:1693-1731); periodic minimum image
//   table          gravity.py:373-421 (built on the host exactly as there; passed in as a device array)
//   tiles          species.py:439-850, interactions.py:848-1351: the reference pairs tiles/subtiles of
//                  size ≥ range; here a uniform cell list with cells ≥ range and a gather-form kernel
//                  (one thread per receiver, no atomics on Δmom)
//   apply_Δmom / convert_Δmom_to_acc   species.py:2253-2325
//   get_rung / assign_rungs / flag_rung_jumps / apply_rung_jumps   species.py:2340-2545
#include "pm_internal.cuh"

#include <algorithm>
#include <cmath>

#include <cub/device/device_scan.cuh>

namespace pm {

// Cell list: cells of side ≥ range/S (S = 2 when the box allows it), z fastest, so that the (2S+1) cells of a z run are one
// contiguous stretch of the cell-sorted positions.  One rank: cells cover the periodic box in all three dimensions.  Several
// ranks: in x the cells cover the slab plus one `range` of read-only ghost particles on either side (the analogue of
// sendrecv_component with tile selection, interactions.py:518-590, communication.py:847-1175); ghosts that come across the
// periodic boundary arrive already shifted by ±L, so x needs no minimum image there.
struct CellGeom {
    int ncx, nc;            // cells along x, cells along y and z
    int S;                  // neighbour reach in cells
    int periodic_x;         // one rank: x wraps like y and z
    double inv_cell_x, inv_cell;
    double x_origin;        // x of the lower edge of cell 0
    double L;
};

__device__ __forceinline__ int cell_1d(double u, double inv_cell, int n) {
    int c = (int)(u * inv_cell);
    if (c < 0) c = 0;
    if (c >= n) c = n - 1;
    return c;
}
__device__ __forceinline__ int cell_index(double x, double y, double z, const CellGeom& g) {
    return (cell_1d(x - g.x_origin, g.inv_cell_x, g.ncx) * g.nc + cell_1d(y, g.inv_cell, g.nc)) * g.nc + cell_1d(z, g.inv_cell, g.nc);
}

// the particles of the pair kernel: local ones followed by the ghosts of the lower and the upper neighbour
struct SrParticles {
    const double* pos;                      // local, AoS
    int64_t n;
    const double* ghost[2];                 // ghost positions (AoS), NULL on one rank
    const unsigned long long* nghost[2];    // their counts (device: written by the neighbours)
    int64_t ghost_cap;
};
__device__ __forceinline__ int64_t sr_total(const SrParticles& p, int64_t* g0, int64_t* g1) {
    *g0 = p.ghost[0] ? min((int64_t)*p.nghost[0], p.ghost_cap) : 0;
    *g1 = p.ghost[1] ? min((int64_t)*p.nghost[1], p.ghost_cap) : 0;
    return p.n + *g0 + *g1;
}
__device__ __forceinline__ const double* sr_pos(const SrParticles& p, int64_t i, int64_t g0) {
    if (i < p.n) return p.pos + 3 * i;
    i -= p.n;
    if (i < g0) return p.ghost[0] + 3 * i;
    return p.ghost[1] + 3 * (i - g0);
}

__global__ void __launch_bounds__(256)
cell_count_kernel(SrParticles p, CellGeom g, int* __restrict__ cell_of, int* __restrict__ count) {
    int64_t g0, g1;
    const int64_t total = sr_total(p, &g0, &g1);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const double* q = sr_pos(p, i, g0);
        const int c = cell_index(q[0], q[1], q[2], g);
        cell_of[i] = c;
        atomicAdd(&count[c], 1);
    }
}

__global__ void __launch_bounds__(256)
cell_scatter_kernel(SrParticles p, const int* __restrict__ cell_of, const int* __restrict__ offset, int* __restrict__ cursor,
                    double* __restrict__ pos_s, size_t stride, int* __restrict__ idx_s) {
    int64_t g0, g1;
    const int64_t total = sr_total(p, &g0, &g1);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const double* q = sr_pos(p, i, g0);
        const int c = cell_of[i];
        const int slot = offset[c] + atomicAdd(&cursor[c], 1);
        pos_s[slot] = q[0];                    // structure of arrays: the lanes of a warp walk neighbouring stretches of the
        pos_s[stride + slot] = q[1];           // sorted list, so one load instruction touches a few 128-byte lines per
        pos_s[2 * stride + slot] = q[2];       // coordinate instead of one 32-byte record per lane
        idx_s[slot] = i < p.n ? (int)i : -1;      // ghosts only supply
    }
}

// receivers of this call in cell order: slots of local particles on active rungs
__global__ void __launch_bounds__(256)
active_list_kernel(const int* __restrict__ idx_s, const int* __restrict__ offset, int ncell_total,
                   const signed char* __restrict__ rung, int lowest_active_rung, int* __restrict__ list,
                   unsigned int* __restrict__ nlist) {
    const int total = offset[ncell_total];
    const int lane = threadIdx.x & 31;
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < total; base += gridDim.x * blockDim.x) {
        const int s = base + lane;
        bool act = false;
        if (s < total) {
            const int i = idx_s[s];
            act = i >= 0 && rung[i] >= lowest_active_rung;
        }
        const unsigned m = __ballot_sync(0xffffffffu, act);
        if (m == 0) continue;
        unsigned b0 = 0;
        if (lane == 0) b0 = atomicAdd(nlist, (unsigned)__popc(m));
        b0 = __shfl_sync(0xffffffffu, b0, 0);
        if (act) list[b0 + __popc(m & ((1u << lane) - 1))] = s;
    }
}

struct PairParams {
    double range2;      // range²
    double scaling;     // (T − 1)/r²_max
    double L;
    int lowest_active_rung;
    int zero_bin;       // index of the 0.0 appended to the device copy of the table
};

// One thread per receiver.  Gather form: no atomics on Δmom; a pair is evaluated from both sides (its two
// receivers may sit on different rungs, or on different ranks).  The (2S+1)² z runs around the receiver's cell are
// contiguous stretches of pos_s; runs that wrap around the box carry their image shift, so the pair loop itself has no
// minimum-image logic: x⃗ = (x⃗_i − x⃗_j) − shift, r² = x·x + y·y + z·z in the reference's order (gravity.py:306-327).
template <bool STATS>
__global__ void __launch_bounds__(128)
shortrange_kernel(const double* __restrict__ xs, size_t stride, const int* __restrict__ idx_s, const int* __restrict__ offset,
                  const int* __restrict__ list, const unsigned int* __restrict__ nlist, CellGeom g, PairParams pp,
                  const double* __restrict__ table, const signed char* __restrict__ rung_jumped,
                  const double* __restrict__ factors, double* __restrict__ dmom, unsigned long long* __restrict__ stats) {
    const double* __restrict__ ys = xs + stride;
    const double* __restrict__ zs = xs + 2 * stride;
    const unsigned int nrecv = *nlist;
    unsigned long long hits = 0, cands = 0;
    for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < nrecv; t += gridDim.x * blockDim.x) {
        const int s = list ? list[t] : (int)t;
        const int i = idx_s[s];
        if (i < 0) continue;
        const double xi = xs[s], yi = ys[s], zi = zs[s];
        const int cx = cell_1d(xi - g.x_origin, g.inv_cell_x, g.ncx), cy = cell_1d(yi, g.inv_cell, g.nc), cz = cell_1d(zi, g.inv_cell, g.nc);
        double ax = 0, ay = 0, az = 0;
        const bool zwrap = cz - g.S < 0 || cz + g.S >= g.nc;
        for (int dx = -g.S; dx <= g.S; ++dx) {
            int nx = cx + dx;
            double sx = 0;
            if (nx < 0) { if (!g.periodic_x) continue; nx += g.ncx; sx = -g.L; }
            else if (nx >= g.ncx) { if (!g.periodic_x) continue; nx -= g.ncx; sx = g.L; }
            for (int dy = -g.S; dy <= g.S; ++dy) {
                int ny = cy + dy;
                double sy = 0;
                if (ny < 0) { ny += g.nc; sy = -g.L; }
                else if (ny >= g.nc) { ny -= g.nc; sy = g.L; }
                const int row = (nx * g.nc + ny) * g.nc;
                const int nseg = zwrap ? 2 * g.S + 1 : 1;
                for (int seg = 0; seg < nseg; ++seg) {
                    int jb, je;
                    double sz = 0;
                    if (!zwrap) {
                        jb = offset[row + cz - g.S];
                        je = offset[row + cz + g.S + 1];
                    } else {
                        int nz = cz - g.S + seg;
                        if (nz < 0) { nz += g.nc; sz = -g.L; }
                        else if (nz >= g.nc) { nz -= g.nc; sz = g.L; }
                        jb = offset[row + nz];
                        je = offset[row + nz + 1];
                    }
                    const double bx = xi - sx, by = yi - sy, bz = zi - sz;     // (x_i − shift) − x_j
                    if (STATS) cands += (unsigned long long)(je - jb);
                    // Branch-free pair loop, four candidates in flight: a candidate beyond the range (or past the end of
                    // the run) reads the zero appended to the table, so it adds exactly nothing; the receiver itself has
                    // x⃗ = 0 and adds nothing either.  r² keeps the reference's operation order — it decides the table bin.
                    for (int j0 = jb; j0 < je; j0 += 4) {
                        double x[4], y[4], z[4], f[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {      // (the array has slack behind its last particle: reading past a run is harmless)
                            x[u] = bx - xs[j0 + u]; y[u] = by - ys[j0 + u]; z[u] = bz - zs[j0 + u];
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const double r2 = x[u] * x[u] + y[u] * y[u] + z[u] * z[u];
                            const bool hit = (r2 <= pp.range2) & (j0 + u < je);
                            const int bin = hit ? (int)(r2 * pp.scaling) : pp.zero_bin;
                            f[u] = __ldg(table + bin);
                            if (STATS) hits += hit ? 1u : 0u;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) { ax = fma(x[u], f[u], ax); ay = fma(y[u], f[u], ay); az = fma(z[u], f[u], az); }
                    }
                }
            }
        }
        if (STATS) --hits;     // the receiver itself
        const double factor = factors[rung_jumped[i]];
        dmom[3 * (size_t)i] = ax * factor;
        dmom[3 * (size_t)i + 1] = ay * factor;
        dmom[3 * (size_t)i + 2] = az * factor;
    }
    if (STATS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            hits += __shfl_xor_sync(0xffffffffu, hits, o);
            cands += __shfl_xor_sync(0xffffffffu, cands, o);
        }
        if ((threadIdx.x & 31) == 0) { atomicAdd(stats, hits); atomicAdd(stats + 1, cands); }
    }
}

// ---- ghosts over the peer mappings -------------------------------------------------------------------------------
struct GhostTargets {
    double* buf[2];                 // [0]: the lower neighbour's UPPER ghost buffer, [1]: the upper neighbour's LOWER ghost buffer
    double shift[2];                // added to x on the way (±L across the periodic boundary, else 0)
};

__global__ void __launch_bounds__(256)
ghost_pack_kernel(const double* __restrict__ pos, int64_t n, double x_lo, double x_hi, double range, GhostTargets gt,
                  int64_t cap, unsigned long long* __restrict__ counts /* [2] local */) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = pos[3 * i];
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const bool near_face = side == 0 ? (x - x_lo < range) : (x_hi - x <= range);
            if (!near_face) continue;
            const unsigned long long k = atomicAdd(&counts[side], 1ULL);
            if (k >= (unsigned long long)cap) continue;      // counted all the same: the receiver sees the overflow
            double* q = gt.buf[side] + 3 * k;
            q[0] = x + gt.shift[side];
            q[1] = pos[3 * i + 1];
            q[2] = pos[3 * i + 2];
        }
    }
}

__global__ void ghost_publish_kernel(ArenaHeader* lower, ArenaHeader* upper, int parity, const unsigned long long* __restrict__ counts) {
    if (threadIdx.x == 0) lower->ghost_count[parity][1].v = counts[0];    // my low-face particles are its upper ghosts
    if (threadIdx.x == 1) upper->ghost_count[parity][0].v = counts[1];
}

static int ensure_bytes(pm_ctx* c, void** buf, size_t* have, size_t need) {
    if (need <= *have) return PM_OK;
    if (*buf) { cudaFree(*buf); c->bytes_allocated -= *have; }
    PM_CHECK_CUDA(cudaMalloc(buf, need + need / 8));
    *have = need + need / 8;
    c->bytes_allocated += *have;
    // never-written slots must hold finite numbers: the pair loop reads up to three slots past a run (and masks them)
    PM_CHECK_CUDA(cudaMemsetAsync(*buf, 0, *have, c->stream));
    return PM_OK;
}

int shortrange(pm_ctx* c, const double* pos, int64_t n, const signed char* rung, const signed char* rung_jumped,
               int lowest_active_rung, const double* factors_host, int nfactors, double range,
               const double* table_dev, int tablesize, double maxr2, double* dmom) {
    PM_REQUIRE(nfactors > 0 && nfactors <= 64, "pm_shortrange: nfactors = %d", nfactors);
    const int P = c->nranks;
    const double L = c->boxsize, W = L / P;
    // ---- ghosts (several ranks): every rank takes part, also one without particles
    SrParticles sp;
    sp.pos = pos; sp.n = n;
    sp.ghost[0] = sp.ghost[1] = nullptr;
    sp.nghost[0] = sp.nghost[1] = nullptr;
    sp.ghost_cap = 0;
    int64_t max_total = n;
    unsigned long long* d_stats = reinterpret_cast<unsigned long long*>(c->d_counts) + 3 * P + (size_t)P * P;   // [4] scratch words
    if (P > 1) {
        PM_REQUIRE(c->peers_ready, "pm_shortrange: call pm_ipc_open_peers first (ghost particles travel over the peer mappings)");
        PM_REQUIRE(W >= 2 * range, "pm_shortrange: slabs of width %g are thinner than twice the short-range range %g", W, range);
        const int parity = (int)(c->ghost_epoch++ & 1);
        const int lower = (c->rank + P - 1) % P, upper = (c->rank + 1) % P;
        const int64_t cap = (int64_t)(c->ghost_buf_bytes / 24);
        auto ghost_buf = [&](int r, int side) {
            return reinterpret_cast<double*>(reinterpret_cast<char*>(c->peer_real[r]) + c->off_ghost +
                                             ((size_t)parity * 2 + side) * c->ghost_buf_bytes);
        };
        GhostTargets gt;
        gt.buf[0] = ghost_buf(lower, 1);
        gt.buf[1] = ghost_buf(upper, 0);
        gt.shift[0] = c->rank == 0 ? L : 0.0;            // across x = 0: they appear beyond the upper face of rank P − 1
        gt.shift[1] = c->rank == P - 1 ? -L : 0.0;
        unsigned long long* d_gc = d_stats + 2;
        PM_CHECK_CUDA(cudaMemsetAsync(d_gc, 0, 2 * sizeof(unsigned long long), c->stream));
        if (n > 0)
            PM_LAUNCH(ghost_pack_kernel, kNumSMs * 4, 256, 0, c->stream, pos, n, c->rank * W, (c->rank + 1) * W, range, gt, cap, d_gc);
        PM_LAUNCH(ghost_publish_kernel, 1, 32, 0, c->stream, arena_header(c, lower), arena_header(c, upper), parity, d_gc);
        PM_TRY(device_barrier(c));
        sp.ghost[0] = ghost_buf(c->rank, 0);
        sp.ghost[1] = ghost_buf(c->rank, 1);
        sp.nghost[0] = &arena_header(c, c->rank)->ghost_count[parity][0].v;
        sp.nghost[1] = &arena_header(c, c->rank)->ghost_count[parity][1].v;
        sp.ghost_cap = cap;
        max_total = n + 2 * cap;
    }
    PM_REQUIRE(max_total < ((int64_t)1 << 31), "pm_shortrange: too many particles for 32-bit cell lists");
    if (max_total == 0) return PM_OK;
    // ---- cell geometry
    CellGeom g;
    g.L = L;
    g.S = 2;
    g.nc = (int)floor(L / (range / g.S));
    if (g.nc < 2 * g.S + 1) { g.S = 1; g.nc = (int)floor(L / range); }
    // the reference requires at least 4 tiles across the box (species.py:3971-3975); 3 is the minimum for
    // a 27-cell neighbourhood without double counting
    PM_REQUIRE(g.nc >= 3, "pm_shortrange: range %g is too large for the box %g (need boxsize/range >= 3)", range, L);
    if (g.nc > 400) g.nc = 400;
    g.inv_cell = g.nc / L;
    if (P == 1) {
        g.periodic_x = 1;
        g.ncx = g.nc;
        g.inv_cell_x = g.inv_cell;
        g.x_origin = 0;
    } else {
        g.periodic_x = 0;
        const double cell = L / g.nc;                          // same cell size as in y and z (≥ range/S)
        g.x_origin = c->rank * W - range;
        g.ncx = (int)ceil((W + 2 * range) / cell);
        g.inv_cell_x = 1.0 / cell;
    }
    const size_t ncell = (size_t)g.ncx * g.nc * g.nc;
    // scratch: pos_s[3·total] factors[64] | cell_of[total] idx_s[total] list[total] count[ncell+1] offset[ncell+1] nlist
    const size_t tot = (size_t)max_total;
    const size_t stride = (tot + 8 + 15) / 16 * 16;        // doubles per coordinate array (slack behind the last particle)
    const size_t need = sizeof(double) * (3 * stride + 64 + (size_t)tablesize + 8) +
                        sizeof(int) * (3 * tot + 2 * (ncell + 1) + 8) + 256;
    PM_TRY(ensure_bytes(c, &c->sr_buf, &c->sr_bytes, need));
    double* pos_s = reinterpret_cast<double*>(c->sr_buf);       // x | y | z of the cell-sorted particles
    double* d_factors = pos_s + 3 * stride;
    double* d_table = d_factors + 64;          // the caller's table followed by a zero (what a non-pair reads)
    int* cell_of = reinterpret_cast<int*>(d_table + tablesize + 8);
    PM_CHECK_CUDA(cudaMemcpyAsync(d_table, table_dev, sizeof(double) * tablesize, cudaMemcpyDeviceToDevice, c->stream));
    PM_CHECK_CUDA(cudaMemsetAsync(d_table + tablesize, 0, sizeof(double) * 8, c->stream));
    int* idx_s = cell_of + tot;
    int* list = idx_s + tot;
    int* count = list + tot;
    int* offset = count + (ncell + 1);
    unsigned int* nlist = reinterpret_cast<unsigned int*>(offset + (ncell + 1));
    PM_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (ncell + 1), c->stream));
    PM_CHECK_CUDA(cudaMemcpyAsync(d_factors, factors_host, sizeof(double) * nfactors, cudaMemcpyHostToDevice, c->stream));
    const int blocks = (int)std::min<int64_t>((max_total + 255) / 256, (int64_t)kNumSMs * 8);
    PM_LAUNCH(cell_count_kernel, blocks, 256, 0, c->stream, sp, g, cell_of, count);
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, count, offset, (int)(ncell + 1), c->stream);
    PM_TRY(ensure_bytes(c, &c->sr_tmp, &c->sr_tmp_bytes, tmp_bytes));
    PM_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(c->sr_tmp, tmp_bytes, count, offset, (int)(ncell + 1), c->stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PM_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (ncell + 1), c->stream));   // reuse as cursor
    PM_LAUNCH(cell_scatter_kernel, blocks, 256, 0, c->stream, sp, cell_of, offset, count, pos_s, stride, idx_s);
    // (receivers in blocks of 4³ cells instead of the z-fastest cell order were measured too: no gain, 16.8 vs 16.5 ms)
    PM_CHECK_CUDA(cudaMemsetAsync(nlist, 0, sizeof(unsigned int), c->stream));
    PM_LAUNCH(active_list_kernel, blocks, 256, 0, c->stream, idx_s, offset, (int)ncell, rung, lowest_active_rung, list, nlist);
    PairParams pp;
    pp.range2 = range * range;
    pp.scaling = (tablesize - 1) / maxr2;
    pp.L = L;
    pp.lowest_active_rung = lowest_active_rung;
    pp.zero_bin = tablesize;
    PM_CHECK_CUDA(cudaMemsetAsync(d_stats, 0, 2 * sizeof(unsigned long long), c->stream));
    const int pblocks = (int)std::min<int64_t>((std::max<int64_t>(n, 1) + 127) / 128, (int64_t)kNumSMs * 16);
    if (c->sr_want_stats)
        PM_LAUNCH(shortrange_kernel<true>, pblocks, 128, 0, c->stream, pos_s, stride, idx_s, offset, list, nlist, g, pp, d_table, rung_jumped,
                  d_factors, dmom, d_stats);
    else
        PM_LAUNCH(shortrange_kernel<false>, pblocks, 128, 0, c->stream, pos_s, stride, idx_s, offset, list, nlist, g, pp, d_table, rung_jumped,
                  d_factors, dmom, d_stats);
    return PM_OK;
}

// ---------------------------------------------------------------------------
// Δmom application and conversion to acceleration
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
apply_dmom_kernel(double* __restrict__ mom, double* __restrict__ dmom, int64_t n, const signed char* __restrict__ rung,
                  const signed char* __restrict__ rung_jumped, int lowest_active_rung, const double* __restrict__ conv,
                  int apply) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (rung[i] < lowest_active_rung) continue;
        const double f = conv[rung_jumped[i]];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double dm = dmom[3 * i + d];
            if (apply) mom[3 * i + d] += dm;          // apply_Δmom, species.py:2253-2266
            dmom[3 * i + d] = dm * f;                 // convert_Δmom_to_acc, species.py:2290-2325
        }
    }
}

int apply_dmom(pm_ctx* c, double* mom, double* dmom, int64_t n, const signed char* rung, const signed char* rung_jumped,
               int lowest_active_rung, const double* conv_host, int nconv, int apply) {
    PM_REQUIRE(nconv > 0 && nconv <= 64, "pm_apply_dmom: nconv = %d", nconv);
    if (n == 0) return PM_OK;
    // d_scratch layout (64 doubles): [0,31) per-call tables, [31] flag, [32] barrier token, [33,64) rung counts
    double* d_conv = c->d_scratch;
    PM_REQUIRE(nconv <= 31, "pm_apply_dmom: at most 31 conversion factors");
    PM_CHECK_CUDA(cudaMemcpyAsync(d_conv, conv_host, sizeof(double) * nconv, cudaMemcpyHostToDevice, c->stream));
    PM_LAUNCH(apply_dmom_kernel, kNumSMs * 4, 256, 0, c->stream, mom, dmom, n, rung, rung_jumped, lowest_active_rung,
              d_conv, apply);
    return PM_OK;
}

// ---------------------------------------------------------------------------
// rungs
// ---------------------------------------------------------------------------
__device__ __forceinline__ signed char get_rung_dev(const double* acc, int64_t i, double rung_factor, int n_rungs,
                                                    signed char current) {
    // Component.get_rung, species.py:2340-2360
    const double ax = acc[3 * i], ay = acc[3 * i + 1], az = acc[3 * i + 2];
    const double acc2 = ax * ax + ay * ay + az * az;
    if (acc2 == 0) return current;
    const double rf = rung_factor + 0.25 * log2(acc2);
    if (rf < 0) return 0;
    if (rf > n_rungs - 1) return (signed char)(n_rungs - 1);
    return (signed char)(1 + (signed char)rf);
}

__global__ void __launch_bounds__(256)
assign_rungs_kernel(const double* __restrict__ acc, int64_t n, double rung_factor, int n_rungs,
                    signed char* __restrict__ rung, signed char* __restrict__ rung_jumped,
                    unsigned long long* __restrict__ rungs_N) {
    __shared__ unsigned int cnt[32];
    if (threadIdx.x < 32) cnt[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const signed char r = get_rung_dev(acc, i, rung_factor, n_rungs, rung[i]);
        rung[i] = r;
        rung_jumped[i] = r;
        atomicAdd(&cnt[r], 1u);
    }
    __syncthreads();
    if (threadIdx.x < n_rungs && cnt[threadIdx.x]) atomicAdd(&rungs_N[threadIdx.x], (unsigned long long)cnt[threadIdx.x]);
}

__global__ void __launch_bounds__(256)
flag_rung_jumps_kernel(const double* __restrict__ acc, int64_t n, const signed char* __restrict__ rung,
                       signed char* __restrict__ rung_jumped, int lowest_active_rung, double rf_up, double rf_down,
                       const double* __restrict__ dt1, int n_rungs, int* __restrict__ any) {
    // Component.flag_rung_jumps, species.py:2461-2510
    int local_any = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const signed char r = rung[i];
        if (r < lowest_active_rung) continue;
        if (dt1[r] == 0) continue;
        signed char ought = get_rung_dev(acc, i, rf_up, n_rungs, r);
        if (ought > r) {
            local_any = 1;
            rung_jumped[i] = (signed char)(r + 2 * n_rungs);
        } else {
            const int down = r + n_rungs;
            if (dt1[down] == -1) continue;
            ought = get_rung_dev(acc, i, rf_down, n_rungs, r);
            if (ought < r) {
                local_any = 1;
                rung_jumped[i] = (signed char)down;
            }
        }
    }
    if (__any_sync(0xffffffffu, local_any) && (threadIdx.x & 31) == 0) atomicOr(any, 1);
}

__global__ void __launch_bounds__(256)
apply_rung_jumps_kernel(int64_t n, signed char* __restrict__ rung, signed char* __restrict__ rung_jumped, int n_rungs,
                        unsigned long long* __restrict__ rungs_N) {
    // Component.apply_rung_jumps, species.py:2523-2545 (rungs_N recounted from scratch)
    __shared__ unsigned int cnt[32];
    if (threadIdx.x < 32) cnt[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        signed char r = rung[i];
        const signed char rj = rung_jumped[i];
        if (rj >= n_rungs) {
            r += (rj >= 2 * n_rungs) ? 1 : -1;
            rung[i] = r;
            rung_jumped[i] = r;
        }
        atomicAdd(&cnt[r], 1u);
    }
    __syncthreads();
    if (threadIdx.x < n_rungs && cnt[threadIdx.x]) atomicAdd(&rungs_N[threadIdx.x], (unsigned long long)cnt[threadIdx.x]);
}

}  // namespace pm

using namespace pm;

extern "C" {

int pm_shortrange(pm_ctx* c, const double* pos, int64_t n, const signed char* rung, const signed char* rung_jumped,
                  int lowest_active_rung, const double* factors_host, int nfactors, double range,
                  const double* table_dev, int tablesize, double maxr2, double* dmom) {
    PM_REQUIRE(c && (n == 0 || (pos && rung && rung_jumped && factors_host && table_dev && dmom)), "pm_shortrange: NULL argument");
    return shortrange(c, pos, n, rung, rung_jumped, lowest_active_rung, factors_host, nfactors, range, table_dev,
                      tablesize, maxr2, dmom);
}

int pm_shortrange_stats(pm_ctx* c, int enable, int64_t* pairs_out, int64_t* candidates_out) {
    PM_REQUIRE(c != nullptr, "pm_shortrange_stats: NULL context");
    if (pairs_out || candidates_out) {
        unsigned long long h[2] = {0, 0};
        const unsigned long long* d = reinterpret_cast<unsigned long long*>(c->d_counts) + 3 * c->nranks + (size_t)c->nranks * c->nranks;
        PM_CHECK_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        PM_CHECK_CUDA(cudaStreamSynchronize(c->stream));
        if (pairs_out) *pairs_out = (int64_t)h[0];
        if (candidates_out) *candidates_out = (int64_t)h[1];
    }
    c->sr_want_stats = enable != 0;
    return PM_OK;
}

int pm_apply_dmom(pm_ctx* c, double* mom, double* dmom, int64_t n, const signed char* rung,
                  const signed char* rung_jumped, int lowest_active_rung, const double* conv_host, int nconv, int apply) {
    PM_REQUIRE(c && (n == 0 || (mom && dmom && rung && rung_jumped && conv_host)), "pm_apply_dmom: NULL argument");
    return apply_dmom(c, mom, dmom, n, rung, rung_jumped, lowest_active_rung, conv_host, nconv, apply);
}

int pm_assign_rungs(pm_ctx* c, const double* acc, int64_t n, double rung_factor, int n_rungs, signed char* rung,
                    signed char* rung_jumped, int64_t* rungs_N_host) {
    PM_REQUIRE(c && rungs_N_host && n_rungs > 0 && n_rungs <= 32, "pm_assign_rungs: bad argument");
    auto* d_N = reinterpret_cast<unsigned long long*>(c->d_scratch + 32) + 1;
    PM_CHECK_CUDA(cudaMemsetAsync(d_N, 0, sizeof(unsigned long long) * n_rungs, c->stream));
    if (n) PM_LAUNCH(assign_rungs_kernel, kNumSMs * 4, 256, 0, c->stream, acc, n, rung_factor, n_rungs, rung, rung_jumped, d_N);
    PM_CHECK_CUDA(cudaMemcpyAsync(rungs_N_host, d_N, sizeof(int64_t) * n_rungs, cudaMemcpyDeviceToHost, c->stream));
    PM_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    return PM_OK;
}

int pm_flag_rung_jumps(pm_ctx* c, const double* acc, int64_t n, const signed char* rung, signed char* rung_jumped,
                       int lowest_active_rung, double rung_factor_up, double rung_factor_down,
                       const double* dt1_host, int n_rungs, int* any_host) {
    PM_REQUIRE(c && dt1_host && any_host && n_rungs > 0 && 3 * n_rungs - 1 <= 31, "pm_flag_rung_jumps: bad argument");
    double* d_dt1 = c->d_scratch;           // [0, 3·n_rungs − 1)
    int* d_any = reinterpret_cast<int*>(c->d_scratch + 31);
    PM_CHECK_CUDA(cudaMemcpyAsync(d_dt1, dt1_host, sizeof(double) * (3 * n_rungs - 1), cudaMemcpyHostToDevice, c->stream));
    PM_CHECK_CUDA(cudaMemsetAsync(d_any, 0, sizeof(int), c->stream));
    if (n) PM_LAUNCH(flag_rung_jumps_kernel, kNumSMs * 4, 256, 0, c->stream, acc, n, rung, rung_jumped,
                     lowest_active_rung, rung_factor_up, rung_factor_down, d_dt1, n_rungs, d_any);
    PM_CHECK_CUDA(cudaMemcpyAsync(any_host, d_any, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PM_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    return PM_OK;
}

int pm_apply_rung_jumps(pm_ctx* c, int64_t n, signed char* rung, signed char* rung_jumped, int n_rungs,
                        int64_t* rungs_N_host) {
    PM_REQUIRE(c && rungs_N_host && n_rungs > 0 && n_rungs <= 32, "pm_apply_rung_jumps: bad argument");
    auto* d_N = reinterpret_cast<unsigned long long*>(c->d_scratch + 32) + 1;
    PM_CHECK_CUDA(cudaMemsetAsync(d_N, 0, sizeof(unsigned long long) * n_rungs, c->stream));
    if (n) PM_LAUNCH(apply_rung_jumps_kernel, kNumSMs * 4, 256, 0, c->stream, n, rung, rung_jumped, n_rungs, d_N);
    PM_CHECK_CUDA(cudaMemcpyAsync(rungs_N_host, d_N, sizeof(int64_t) * n_rungs, cudaMemcpyDeviceToHost, c->stream));
    PM_CHECK_CUDA(cudaStreamSynchronize(c->stream));
    return PM_OK;
}

}  // extern "C"
