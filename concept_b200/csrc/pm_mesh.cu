// pm_mesh.cu — particle <-> mesh kernels: mass deposit, finite-difference gradient,
// force gather + kick.  sm_100a, HBM/L2-bound scatter/gather work: no tensor cores.
//
// Reference semantics (file:line under the reference's src/):
//   coordinates   mesh.py:1577-1606 (deposit), :405-432 (gather; lattice shift sign flipped)
//   weights       mesh.py:5305-5379 (set_weights_NGP/CIC/TSC/PCS)
//   3-D weight    mesh.py:5138-5155: (wx·multiplier)·wy·wz
//   deposit       mesh.py:1596-1632   grid[index] += contribution_weighted
//   FD gradient   mesh.py:4961-5015   (coefficients (c/Δx) formed on the host exactly as there)
//   gather/kick   mesh.py:427-459; interactions.py:2384-2387
#include "pm_internal.cuh"

#include <algorithm>
#include <cstdlib>

namespace pm {

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
struct Coord {
    double off[3];   // offset_x/y/z of the reference
    double scale;    // (1/cellsize)*(1 - machine_ϵ)
};

static Coord make_coord(const pm_ctx* c, const double* shift, bool for_gather) {
    // mesh.py:1577-1589 / :405-420 with domain_bgn = 0, cell_centered = True
    Coord k;
    const double cellsize = c->boxsize / c->g.G;
    const double sgn = for_gather ? +1.0 : -1.0;
    for (int d = 0; d < 3; ++d) {
        const double s = shift ? shift[d] : 0.0;
        k.off[d] = 0.0 - (1 + kEps) * (kNghostsRef - 0.5 + sgn * s) * cellsize;
    }
    k.scale = (1 / cellsize) * (1 - kEps);
    return k;
}

// 1-D weights; returns the (ghost-free, unwrapped) index of the first cell.
template <int ORDER>
__device__ __forceinline__ int weights_1d(double x, double (&w)[ORDER]) {
    int index;
    if constexpr (ORDER == 1) {
        index = (int)(x + 0.5);
        w[0] = 1.0;
    } else if constexpr (ORDER == 2) {
        index = (int)x;
        const double dist = x - index;
        w[0] = 1 - dist;
        w[1] = dist;
    } else if constexpr (ORDER == 3) {
        index = (int)(x + 0.5);
        const double dist = x - index;
        index -= 1;
        const double dist2 = dist * dist;
        const double w0 = 0.125 + 0.5 * (dist2 - dist);
        const double w1 = 0.75 - dist2;
        w[0] = w0;
        w[1] = w1;
        w[2] = 1 - w0 - w1;
    } else {
        index = (int)x;
        index -= 1;
        const double dist = x - index;
        const double tmp = 2 - dist;
        const double tmp2 = tmp * tmp;
        const double tmp3 = tmp * tmp2;
        const double w0 = 1. / 6. * tmp3;
        const double w2 = 2. / 3. - tmp2 + 0.5 * tmp3;
        const double dm1 = dist - 1;
        const double w3 = 1. / 6. * (dm1 * dm1 * dm1);
        w[0] = w0;
        w[1] = 1 - w0 - w2 - w3;
        w[2] = w2;
        w[3] = w3;
    }
    return index - kNghostsRef;
}

// np.mod(x, L) with the reference's x == L → 0 guard (commons.py:5102-5131); same as pm_particles.cu
__device__ __forceinline__ double mod_box(double x, double L) {
    double r = fmod(x, L);
    if (r < 0) r += L;
    if (r == L) r = 0;
    return r;
}

__device__ __forceinline__ int wrap(int i, int G) {
    // |i| excursions are at most a handful of cells beyond [0, G)
    if (i < 0) i += G;
    else if (i >= G) i -= G;
    return i;
}

// x plane of the local buffer for global (unwrapped) plane index i, or -1 if outside
__device__ __forceinline__ int local_plane(int i, const Geom& g) {
    if (g.wrap_x) return wrap(i, g.G);
    const int l = i - g.x0 + g.halo;
    return (l >= 0 && l < g.nxl + 2 * g.halo) ? l : -1;
}

// Coalesced 128-bit staging of BLOCK particles (AoS double[3]) through shared memory.
template <int BLOCK>
__device__ __forceinline__ void stage_particles(const double* __restrict__ src, int64_t first,
                                                int count, double* s) {
    const double* p = src + first * 3;
    const int nd = count * 3;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        const double2* p2 = reinterpret_cast<const double2*>(p);
        double2* s2 = reinterpret_cast<double2*>(s);
        const int n2 = nd >> 1;
        for (int t = threadIdx.x; t < n2; t += BLOCK) s2[t] = __ldcs(p2 + t);   // streaming: do not displace the grid in L2
        if ((nd & 1) && threadIdx.x == 0) s[nd - 1] = __ldcs(p + nd - 1);
    } else {
        for (int t = threadIdx.x; t < nd; t += BLOCK) s[t] = __ldcs(p + t);
    }
}

// Dynamic, in-order chunk scheduler: CTAs take the next chunk of kChunkTiles·BLOCK consecutive particles
// from a global counter, so the set of chunks in flight stays a compact window of the (cell-ordered)
// particle array and the grid planes it touches stay L2-resident.  (A static blockIdx-strided loop lets
// fast and slow CTAs drift tens of grid planes apart — measured: the potential grid was read 1.85x from
// DRAM.)  One block-wide atomic per chunk, not per tile: with a ticket per 256-particle tile 64 % of the
// deposit's warp stalls were the barrier around the ticket (ncu, r01).  Inside a chunk the warps run
// free; consecutive tiles of a chunk share grid rows, which the SM's L1 keeps for the gather.
#ifndef PM_DEP_CHUNK_TILES
#define PM_DEP_CHUNK_TILES 2   /* 512 particles per ticket (measured 256 … 8192: 0.78, 0.62, 0.64, 0.68, 0.71, 0.74 ms) */
#endif
constexpr int kChunkTiles = PM_DEP_CHUNK_TILES;

__device__ __forceinline__ int64_t next_chunk(unsigned long long* counter, int64_t* s_slot) {
    __syncthreads();
    if (threadIdx.x == 0) *s_slot = (int64_t)atomicAdd(counter, 1ULL);
    __syncthreads();
    return *s_slot;
}

// ---------------------------------------------------------------------------
// deposit
// ---------------------------------------------------------------------------
constexpr int kDepBlock = 256;

__device__ __forceinline__ void red_add_v2(float* addr8, float a, float b) {     // addr8: 8-byte aligned
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(__cvta_generic_to_global(addr8)), "f"(a), "f"(b) : "memory");
}

template <int ORDER, typename T>
__global__ void __launch_bounds__(kDepBlock)
deposit_kernel(const double* __restrict__ pos, int64_t n, T* __restrict__ grid, Geom g, Coord co,
               double contribution, unsigned long long* __restrict__ tile_counter, int* __restrict__ lost) {
    __shared__ int64_t s_slot;
    constexpr int64_t kChunk = (int64_t)kDepBlock * kChunkTiles;
    const int64_t nchunks = (n + kChunk - 1) / kChunk;
    for (int64_t chunk = next_chunk(tile_counter, &s_slot); chunk < nchunks; chunk = next_chunk(tile_counter, &s_slot)) {
        const int64_t end = min(n, (chunk + 1) * kChunk);
        for (int64_t ip = chunk * kChunk + threadIdx.x; ip < end; ip += kDepBlock) {
            const double* pp = pos + ip * 3;   // streaming: do not displace the grid in L2
            const double x = (__ldcs(pp + 0) - co.off[0]) * co.scale;
            const double y = (__ldcs(pp + 1) - co.off[1]) * co.scale;
            const double z = (__ldcs(pp + 2) - co.off[2]) * co.scale;
            double wx[ORDER], wy[ORDER], wz[ORDER];
            const int ix = weights_1d<ORDER>(x, wx);
            const int iy = weights_1d<ORDER>(y, wy);
            const int iz = weights_1d<ORDER>(z, wz);
            int jy[ORDER], kz[ORDER];
#pragma unroll
            for (int b = 0; b < ORDER; ++b) {
                jy[b] = wrap(iy + b, g.G) * g.Gp;
                kz[b] = wrap(iz + b, g.G);
            }
#pragma unroll
            for (int a = 0; a < ORDER; ++a) {
                const int lx = local_plane(ix + a, g);
                if (lx < 0) {       // several ranks: a plane beyond the halo — the particle is not in this rank's slab and its
                    *lost = 1;      // mass would vanish silently; sticky flag, reported by pm_check_async_error
                    continue;
                }
                const double wa = wx[a] * contribution;
                T* plane = grid + (size_t)lx * g.G * g.Gp;
#pragma unroll
                for (int b = 0; b < ORDER; ++b) {
                    const double wab = wa * wy[b];
                    T* row = plane + jy[b];
                    if constexpr (sizeof(T) == 4 && ORDER >= 2) {
                        // fp32 grid: the ORDER cells of a z run go out as 8-byte-aligned pairs (red.global.add.v2.f32,
                        // REDG.E.ADD.F32x2) plus the odd ends — 18 instead of 27 reductions per TSC particle.  PTX has no
                        // vector form for f64.
                        if (iz >= 0 && iz + ORDER <= g.G) {
                            float v[ORDER];
#pragma unroll
                            for (int cc = 0; cc < ORDER; ++cc) v[cc] = (float)(wab * wz[cc]);
                            float* r0 = reinterpret_cast<float*>(row) + iz;
                            if (iz & 1) {
                                atomicAdd(r0, v[0]);
#pragma unroll
                                for (int cc = 1; cc + 1 < ORDER; cc += 2) red_add_v2(r0 + cc, v[cc], v[cc + 1]);
                                if ((ORDER & 1) == 0) atomicAdd(r0 + ORDER - 1, v[ORDER - 1]);
                            } else {
#pragma unroll
                                for (int cc = 0; cc + 1 < ORDER; cc += 2) red_add_v2(r0 + cc, v[cc], v[cc + 1]);
                                if (ORDER & 1) atomicAdd(r0 + ORDER - 1, v[ORDER - 1]);
                            }
                            continue;
                        }
                    }
#pragma unroll
                    for (int cc = 0; cc < ORDER; ++cc) {
                        atomicAdd(row + kz[cc], (T)(wab * wz[cc]));
                    }
                }
            }
        }
    }
}

template <typename T>
static int deposit_dispatch(pm_ctx* c, const double* pos, int64_t n, int order, double contribution,
                            const Coord& co) {
    const int64_t nchunks = (n + (int64_t)kDepBlock * kChunkTiles - 1) / ((int64_t)kDepBlock * kChunkTiles);
    const int grid = (int)std::min<int64_t>(nchunks, (int64_t)kNumSMs * 8);
    T* gptr = reinterpret_cast<T*>(c->real);
    unsigned long long* ctr = c->d_tilectr;
    int* lost = c->d_comm_err + 1;
    PM_CHECK_CUDA(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), c->stream));
    switch (order) {
        case 1: PM_LAUNCH((deposit_kernel<1, T>), grid, kDepBlock, 0, c->stream, pos, n, gptr, c->g, co, contribution, ctr, lost); break;
        case 2: PM_LAUNCH((deposit_kernel<2, T>), grid, kDepBlock, 0, c->stream, pos, n, gptr, c->g, co, contribution, ctr, lost); break;
        case 3: PM_LAUNCH((deposit_kernel<3, T>), grid, kDepBlock, 0, c->stream, pos, n, gptr, c->g, co, contribution, ctr, lost); break;
        case 4: PM_LAUNCH((deposit_kernel<4, T>), grid, kDepBlock, 0, c->stream, pos, n, gptr, c->g, co, contribution, ctr, lost); break;
    }
    return PM_OK;
}

int launch_deposit(pm_ctx* c, const double* pos, int64_t n, int order, double contribution,
                   const double* shift) {
    PM_REQUIRE(order >= 1 && order <= 4,
               "pm_deposit called with order = %d not in {1 (NGP), 2 (CIC), 3 (TSC), 4 (PCS)}", order);
    if (n == 0) return PM_OK;
    const Coord co = make_coord(c, shift, false);
    return c->dtype == PM_GRID_F64 ? deposit_dispatch<double>(c, pos, n, order, contribution, co)
                                   : deposit_dispatch<float>(c, pos, n, order, contribution, co);
}

// ---------------------------------------------------------------------------
// finite differences
// ---------------------------------------------------------------------------
struct FD {
    int reach;        // neighbours used on each side
    int forward;      // order 1: (φ[+1] - φ[0])/Δx
    double c[4];      // |coefficient|/Δx ; signs alternate + - + -
};

static int make_fd(int order, double dx, FD* fd) {
    // mesh.py:4961-5015; coefficients formed as (c)/Δx like the ℝ[...] constants there
    fd->forward = 0;
    for (double& v : fd->c) v = 0;
    switch (order) {
        case 1: fd->reach = 1; fd->forward = 1; fd->c[0] = 1 / dx; break;
        case 2: fd->reach = 1; fd->c[0] = (1. / 2) / dx; break;
        case 4: fd->reach = 2; fd->c[0] = (2. / 3) / dx; fd->c[1] = (1. / 12) / dx; break;
        case 6: fd->reach = 3; fd->c[0] = (3. / 4) / dx; fd->c[1] = (3. / 20) / dx; fd->c[2] = (1. / 60) / dx; break;
        case 8: fd->reach = 4; fd->c[0] = (4. / 5) / dx; fd->c[1] = (1. / 5) / dx; fd->c[2] = (4. / 105) / dx; fd->c[3] = (1. / 280) / dx; break;
        default:
            set_error("diff order = %d not in {1, 2, 4, 6, 8}", order);
            return PM_ERR_ARG;
    }
    return PM_OK;
}

// Combine the ±n differences exactly in the reference's association:
//   order 2:  c0*(u1-l1)
//   order 4:  + c0*(u1-l1) - c1*(u2-l2)
//   order 6:  c0*(u1-l1) - c1*(u2-l2) + c2*(u3-l3)
//   order 8:  c0*(u1-l1) - c1*(u2-l2) + c2*(u3-l3) - c3*(u4-l4)
template <int REACH>
__device__ __forceinline__ double fd_combine(const double (&d)[REACH], const FD& fd) {
    double v = fd.c[0] * d[0];
    if constexpr (REACH >= 2) v = v - fd.c[1] * d[1];
    if constexpr (REACH >= 3) v = v + fd.c[2] * d[2];
    if constexpr (REACH >= 4) v = v - fd.c[3] * d[3];
    return v;
}

// Explicit force grid (parity tap and the un-fused path): interior planes only.
template <int REACH, typename T>
__global__ void __launch_bounds__(256)
diff_kernel(const T* __restrict__ phi, T* __restrict__ out, Geom g, FD fd, int dim) {
    const int64_t total = (int64_t)g.nxl * g.G * g.G;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(idx % g.G);
        const int j = (int)((idx / g.G) % g.G);
        const int il = (int)(idx / ((int64_t)g.G * g.G));
        const int i = g.x0 + il;
        double d[REACH];
#pragma unroll
        for (int m = 1; m <= REACH; ++m) {
            int iu = i, ju = j, ku = k, idn = i, jd = j, kd = k;
            if (dim == 0) { iu += m; idn -= fd.forward ? 0 : m; }
            if (dim == 1) { ju = wrap(j + m, g.G); jd = fd.forward ? j : wrap(j - m, g.G); }
            if (dim == 2) { ku = wrap(k + m, g.G); kd = fd.forward ? k : wrap(k - m, g.G); }
            const int lu = local_plane(iu, g), ld = local_plane(idn, g);
            const double up = (double)phi[((size_t)lu * g.G + ju) * g.Gp + ku];
            const double lo = (double)phi[((size_t)ld * g.G + jd) * g.Gp + kd];
            d[m - 1] = up - lo;
        }
        const int lc = local_plane(i, g);
        out[((size_t)lc * g.G + j) * g.Gp + k] = (T)fd_combine<REACH>(d, fd);
    }
}

int launch_diff(pm_ctx* c, int dim, int order) {
    PM_REQUIRE(dim >= 0 && dim < 3, "pm_diff called with dim = %d not in {0, 1, 2}", dim);
    FD fd;
    PM_TRY(make_fd(order, c->boxsize / c->g.G, &fd));
    PM_TRY(ensure_force(c));
    const int grid = kNumSMs * 8;
#define PM_DIFF_CASE(R, T)                                                                       \
    PM_LAUNCH((diff_kernel<R, T>), grid, 256, 0, c->stream, reinterpret_cast<const T*>(c->grid_read()), \
              reinterpret_cast<T*>(c->force), c->g, fd, dim)
    if (c->dtype == PM_GRID_F64) {
        switch (fd.reach) {
            case 1: PM_DIFF_CASE(1, double); break;
            case 2: PM_DIFF_CASE(2, double); break;
            case 3: PM_DIFF_CASE(3, double); break;
            case 4: PM_DIFF_CASE(4, double); break;
        }
    } else {
        switch (fd.reach) {
            case 1: PM_DIFF_CASE(1, float); break;
            case 2: PM_DIFF_CASE(2, float); break;
            case 3: PM_DIFF_CASE(3, float); break;
            case 4: PM_DIFF_CASE(4, float); break;
        }
    }
#undef PM_DIFF_CASE
    return PM_OK;
}

// ---------------------------------------------------------------------------
// gather from an explicit grid (one dimension)
// ---------------------------------------------------------------------------
constexpr int kGatBlock = 256;    // gather_kernel (one dimension, explicit grid)
constexpr int kGkBlock = 128;     // gather_kick_kernel: ~100 registers per thread — small CTAs keep 20 warps per SM

template <int ORDER, typename T>
__global__ void __launch_bounds__(kGatBlock)
gather_kernel(const T* __restrict__ grid, const double* __restrict__ pos, double* __restrict__ mom,
              int64_t n, Geom g, Coord co, int dim, double factor) {
    __shared__ __align__(16) double spos[kGatBlock * 3];
    const int64_t ntiles = (n + kGatBlock - 1) / kGatBlock;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t first = tile * kGatBlock;
        const int count = (int)min((int64_t)kGatBlock, n - first);
        __syncthreads();
        stage_particles<kGatBlock>(pos, first, count, spos);
        __syncthreads();
        if ((int)threadIdx.x >= count) continue;
        const double x = (spos[threadIdx.x * 3 + 0] - co.off[0]) * co.scale;
        const double y = (spos[threadIdx.x * 3 + 1] - co.off[1]) * co.scale;
        const double z = (spos[threadIdx.x * 3 + 2] - co.off[2]) * co.scale;
        double wx[ORDER], wy[ORDER], wz[ORDER];
        const int ix = weights_1d<ORDER>(x, wx);
        const int iy = weights_1d<ORDER>(y, wy);
        const int iz = weights_1d<ORDER>(z, wz);
        double value = 0;
#pragma unroll
        for (int a = 0; a < ORDER; ++a) {
            const int lx = local_plane(ix + a, g);
            if (lx < 0) continue;
#pragma unroll
            for (int b = 0; b < ORDER; ++b) {
                const double wab = wx[a] * wy[b];
                const T* row = grid + ((size_t)lx * g.G + wrap(iy + b, g.G)) * g.Gp;
#pragma unroll
                for (int cc = 0; cc < ORDER; ++cc) {
                    value += (double)row[wrap(iz + cc, g.G)] * (wab * wz[cc]);
                }
            }
        }
        if (factor != 1) value *= factor;
        mom[(first + threadIdx.x) * 3 + dim] += value;
    }
}

int launch_gather(pm_ctx* c, int which, const double* pos, double* mom, int64_t n, int order,
                  int dim, double factor, const double* shift) {
    PM_REQUIRE(order >= 1 && order <= 4, "pm_gather called with order = %d not in {1, 2, 3, 4}", order);
    PM_REQUIRE(dim >= 0 && dim < 3, "pm_gather called with dim = %d not in {0, 1, 2}", dim);
    PM_REQUIRE(which == PM_TAP_REAL || which == PM_TAP_FORCE, "pm_gather: which = %d", which);
    if (which == PM_TAP_FORCE) PM_REQUIRE(c->force != nullptr, "pm_gather: no force grid (call pm_diff first)");
    if (n == 0) return PM_OK;
    const Coord co = make_coord(c, shift, true);
    const int64_t ntiles = (n + kGatBlock - 1) / kGatBlock;
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)kNumSMs * 16);
    const void* src = which == PM_TAP_REAL ? c->grid_read() : c->force;
#define PM_GATHER_CASE(O, T)                                                                           \
    PM_LAUNCH((gather_kernel<O, T>), grid, kGatBlock, 0, c->stream, reinterpret_cast<const T*>(src),  \
              pos, mom, n, c->g, co, dim, factor)
    if (c->dtype == PM_GRID_F64) {
        switch (order) {
            case 1: PM_GATHER_CASE(1, double); break;
            case 2: PM_GATHER_CASE(2, double); break;
            case 3: PM_GATHER_CASE(3, double); break;
            case 4: PM_GATHER_CASE(4, double); break;
        }
    } else {
        switch (order) {
            case 1: PM_GATHER_CASE(1, float); break;
            case 2: PM_GATHER_CASE(2, float); break;
            case 3: PM_GATHER_CASE(3, float); break;
            case 4: PM_GATHER_CASE(4, float); break;
        }
    }
#undef PM_GATHER_CASE
    return PM_OK;
}

// ---------------------------------------------------------------------------
// fused gradient + gather + kick (+ Σ mom²) (+ drift)
// ---------------------------------------------------------------------------
// Per particle the finite differences need the ORDER³ cells under the interpolation stencil plus, per
// dimension, 2·REACH "arm" cells beyond it on every line: ORDER³ + 3·ORDER²·2·REACH distinct loads (32
// for CIC/order 2) instead of the 6·REACH·ORDER³ (48) of the straightforward double loop.  Warps run
// independently (no shared-memory staging, no barrier inside a chunk).  x and y sums run over their
// line index innermost, so their summation order differs from the reference's (a, b, c) nest by a
// reassociation (last-bit differences; the stated kick tolerance is 1e-9).
#ifndef PM_GK_CHUNK_TILES
#define PM_GK_CHUNK_TILES 4
#endif
constexpr int kGkChunkTiles = PM_GK_CHUNK_TILES;  // 512 consecutive particles per ticket (measured 512 … 8192: 0.73, 0.75, 0.77, 0.77, 0.87 ms)

// CIC with a 2-point difference: cap the kernel at 80 registers (6 CTAs = 24 warps per SM).  The kernel waits on
// L1/L2 gathers 78 % of the time (ncu), so warps in flight matter more than the 50 bytes of spills: measured
// 0.88 -> 0.77 ms; a cap of 64 registers (8 CTAs) spills too much (0.92 ms).  The wider stencils keep their registers.
template <int ORDER, int REACH, typename T, bool DRIFT>
__global__ void __launch_bounds__(kGkBlock, (ORDER <= 2 && REACH == 1) ? 6 : 1)
gather_kick_kernel(const T* __restrict__ phi, double* __restrict__ pos,
                   double* __restrict__ mom, int64_t n, Geom g, Coord co, FD fd, double factor,
                   double* __restrict__ sum_mom2, unsigned long long* __restrict__ tile_counter,
                   double drift_dt, double boxsize) {
    __shared__ double sred[kGkBlock / 32];
    __shared__ int64_t s_slot;
    constexpr int W = ORDER + 2 * REACH;   // cells touched per axis
    constexpr int64_t kChunk = (int64_t)kGkBlock * kGkChunkTiles;
    double mom2_acc = 0;
    const int64_t nchunks = (n + kChunk - 1) / kChunk;
    for (int64_t chunk = next_chunk(tile_counter, &s_slot); chunk < nchunks; chunk = next_chunk(tile_counter, &s_slot)) {
        const int64_t end = min(n, (chunk + 1) * kChunk);
        for (int64_t ip = chunk * kChunk + threadIdx.x; ip < end; ip += kGkBlock) {
            double* pp = pos + ip * 3;
            const double px = __ldcs(pp), py = __ldcs(pp + 1), pz = __ldcs(pp + 2);
            const double x = (px - co.off[0]) * co.scale;
            const double y = (py - co.off[1]) * co.scale;
            const double z = (pz - co.off[2]) * co.scale;
            double wx[ORDER], wy[ORDER], wz[ORDER];
            const int ix = weights_1d<ORDER>(x, wx);
            const int iy = weights_1d<ORDER>(y, wy);
            const int iz = weights_1d<ORDER>(z, wz);
            // element offsets of the W cells along each axis (pre-multiplied by strides)
            size_t ox[W];
            int oy[W], oz[W];
#pragma unroll
            for (int t = 0; t < W; ++t) {
                const int lx = local_plane(ix + t - REACH, g);
                ox[t] = (size_t)(lx < 0 ? 0 : lx) * g.G * g.Gp;
                oy[t] = wrap(iy + t - REACH, g.G) * g.Gp;
                oz[t] = wrap(iz + t - REACH, g.G);
            }
            // the cells under the interpolation stencil
            T core[ORDER][ORDER][ORDER];
#pragma unroll
            for (int a = 0; a < ORDER; ++a)
#pragma unroll
                for (int b = 0; b < ORDER; ++b)
#pragma unroll
                    for (int cc = 0; cc < ORDER; ++cc)
                        core[a][b][cc] = phi[ox[a + REACH] + oy[b + REACH] + oz[cc + REACH]];
            double vx = 0, vy = 0, vz = 0;
            // ---- x: lines along a for every (b, c)
#pragma unroll
            for (int b = 0; b < ORDER; ++b) {
#pragma unroll
                for (int cc = 0; cc < ORDER; ++cc) {
                    T line[W];
#pragma unroll
                    for (int t = 0; t < W; ++t)
                        line[t] = (t >= REACH && t < REACH + ORDER) ? core[t - REACH][b][cc]
                                                                  : phi[ox[t] + oy[b + REACH] + oz[cc + REACH]];
#pragma unroll
                    for (int a = 0; a < ORDER; ++a) {
                        const double w = (wx[a] * wy[b]) * wz[cc];
                        double d[REACH];
#pragma unroll
                        for (int m = 1; m <= REACH; ++m)
                            d[m - 1] = (double)line[a + REACH + m] - (double)(fd.forward ? line[a + REACH] : line[a + REACH - m]);
                        vx += fd_combine<REACH>(d, fd) * w;
                    }
                }
            }
            // ---- y: lines along b for every (a, c)
#pragma unroll
            for (int a = 0; a < ORDER; ++a) {
#pragma unroll
                for (int cc = 0; cc < ORDER; ++cc) {
                    T line[W];
#pragma unroll
                    for (int t = 0; t < W; ++t)
                        line[t] = (t >= REACH && t < REACH + ORDER) ? core[a][t - REACH][cc]
                                                                  : phi[ox[a + REACH] + oy[t] + oz[cc + REACH]];
#pragma unroll
                    for (int b = 0; b < ORDER; ++b) {
                        const double w = (wx[a] * wy[b]) * wz[cc];
                        double d[REACH];
#pragma unroll
                        for (int m = 1; m <= REACH; ++m)
                            d[m - 1] = (double)line[b + REACH + m] - (double)(fd.forward ? line[b + REACH] : line[b + REACH - m]);
                        vy += fd_combine<REACH>(d, fd) * w;
                    }
                }
            }
            // ---- z: lines along c for every (a, b)
#pragma unroll
            for (int a = 0; a < ORDER; ++a) {
#pragma unroll
                for (int b = 0; b < ORDER; ++b) {
                    T line[W];
#pragma unroll
                    for (int t = 0; t < W; ++t)
                        line[t] = (t >= REACH && t < REACH + ORDER) ? core[a][b][t - REACH]
                                                                  : phi[ox[a + REACH] + oy[b + REACH] + oz[t]];
#pragma unroll
                    for (int cc = 0; cc < ORDER; ++cc) {
                        const double w = (wx[a] * wy[b]) * wz[cc];
                        double d[REACH];
#pragma unroll
                        for (int m = 1; m <= REACH; ++m)
                            d[m - 1] = (double)line[cc + REACH + m] - (double)(fd.forward ? line[cc + REACH] : line[cc + REACH - m]);
                        vz += fd_combine<REACH>(d, fd) * w;
                    }
                }
            }
            if (factor != 1) { vx *= factor; vy *= factor; vz *= factor; }
            double* m = mom + ip * 3;
            const double mx = __ldcs(m) + vx, my = __ldcs(m + 1) + vy, mz = __ldcs(m + 2) + vz;
            __stcs(m, mx); __stcs(m + 1, my); __stcs(m + 2, mz);
            mom2_acc += mx * mx + my * my + mz * mz;
            if constexpr (DRIFT) {
                // Component.drift (species.py:2191-2196) with the freshly kicked momenta
                __stcs(pp, mod_box(px + mx * drift_dt, boxsize));
                __stcs(pp + 1, mod_box(py + my * drift_dt, boxsize));
                __stcs(pp + 2, mod_box(pz + mz * drift_dt, boxsize));
            }
        }
    }
    if (sum_mom2 != nullptr) {
        __syncthreads();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mom2_acc += __shfl_xor_sync(0xffffffffu, mom2_acc, o);
        if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = mom2_acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0;
            for (int w = 0; w < kGkBlock / 32; ++w) t += sred[w];
            atomicAdd(sum_mom2, t);
        }
    }
}

// ---------------------------------------------------------------------------
// fused gradient + gather + kick, streaming form
// ---------------------------------------------------------------------------
// The force on a particle is linear in the potential: with 1-D interpolation weights w[a] on ORDER cells and the centred
// difference D[a] = Σ_m ±c_m·(φ[a+m] − φ[a−m]) (mesh.py:4961-5015),
//   Σ_a w[a]·D[a] = Σ_u φ[u]·w'[u],   w'[u] = Σ_m ±c_m·(w[u−m] − w[u+m])      over W = ORDER + 2·REACH cells,
// so the three force components are three separable sums over the W³ neighbourhood (minus the parts where two axes are
// both outside the interpolation stencil).  Row by row along z: S0 = Σ_c wz[c]·φ, Sz = Σ_c wz'[c]·φ, then
//   F_x += (wx'[t]·wy[b])·S0,   F_y += (wx[t]·wy'[b])·S0,   F_z += (wx[t]·wy[b])·Sz.
// The same ORDER³ + 3·ORDER²·2·REACH distinct loads as the line form above, but nothing lives longer than one row:
// ~50 registers instead of 100+ (TSC/PCS ran at one 128-thread CTA per SM).  The sums are a reassociation of the
// reference's nest (differences are taken after the weighting); the stated kick tolerance is 1e-9 of the largest kick.
template <int ORDER, int REACH>
__device__ __forceinline__ void diff_weights(const double (&w)[ORDER], const FD& fd, double (&wd)[ORDER + 2 * REACH]) {
    constexpr int W = ORDER + 2 * REACH;
#pragma unroll
    for (int u = 0; u < W; ++u) {
        // cell u − REACH relative to the first stencil cell
        double v = 0;
#pragma unroll
        for (int m = 1; m <= REACH; ++m) {
            const int lo = u - REACH - m, hi = u - REACH + m;           // stencil cells whose difference touches this cell
            const double wl = (lo >= 0 && lo < ORDER) ? w[lo] : 0.0;     // … as their upper neighbour (+)
            const double wh = (hi >= 0 && hi < ORDER) ? w[hi] : 0.0;     // … as their lower neighbour (−)
            const double wc = (u - REACH >= 0 && u - REACH < ORDER) ? w[u - REACH] : 0.0;
            const double term = fd.forward ? (wl - wc) : (wl - wh);     // order 1: (φ[a+1] − φ[a])/Δx
            v += ((m & 1) ? fd.c[m - 1] : -fd.c[m - 1]) * term;
        }
        wd[u] = v;
    }
}

#ifndef PM_GK2_OCC
#define PM_GK2_OCC 4
#endif
// CTAs per SM (register cap).  Measured on B200 at 4 / 5 / 6 (profiles/r02_fft_config_sweep.md): TSC fp32, order-2 differences
// 1.08 / 1.13 / 1.15 ms; TSC fp64, order-4 differences 2.56 / 2.31 / 3.30 ms; PCS fp64, order-4 differences 3.74 / 5.52 / 9.10 ms.
template <int ORDER, int REACH, typename T>
constexpr int gk2_occupancy() { return (ORDER == 3 && REACH == 2 && sizeof(T) == 8) ? 5 : PM_GK2_OCC; }
template <int ORDER, int REACH, typename T, bool DRIFT>
__global__ void __launch_bounds__(kGkBlock, gk2_occupancy<ORDER, REACH, T>())
gather_kick2_kernel(const T* __restrict__ phi, double* __restrict__ pos,
                    double* __restrict__ mom, int64_t n, Geom g, Coord co, FD fd, double factor,
                    double* __restrict__ sum_mom2, unsigned long long* __restrict__ tile_counter,
                    double drift_dt, double boxsize) {
    __shared__ double sred[kGkBlock / 32];
    __shared__ int64_t s_slot;
    constexpr int W = ORDER + 2 * REACH;
    constexpr int64_t kChunk = (int64_t)kGkBlock * kGkChunkTiles;
    double mom2_acc = 0;
    const int64_t nchunks = (n + kChunk - 1) / kChunk;
    for (int64_t chunk = next_chunk(tile_counter, &s_slot); chunk < nchunks; chunk = next_chunk(tile_counter, &s_slot)) {
        const int64_t end = min(n, (chunk + 1) * kChunk);
        for (int64_t ip = chunk * kChunk + threadIdx.x; ip < end; ip += kGkBlock) {
            double* pp = pos + ip * 3;
            const double px = __ldcs(pp), py = __ldcs(pp + 1), pz = __ldcs(pp + 2);
            double wx[ORDER], wy[ORDER], wz[ORDER], dx[W], dy[W], dz[W];
            const int ix = weights_1d<ORDER>((px - co.off[0]) * co.scale, wx);
            const int iy = weights_1d<ORDER>((py - co.off[1]) * co.scale, wy);
            const int iz = weights_1d<ORDER>((pz - co.off[2]) * co.scale, wz);
            diff_weights<ORDER, REACH>(wx, fd, dx);
            diff_weights<ORDER, REACH>(wy, fd, dy);
            diff_weights<ORDER, REACH>(wz, fd, dz);
            int oz[W];
#pragma unroll
            for (int t = 0; t < W; ++t) oz[t] = wrap(iz + t - REACH, g.G);
            double vx = 0, vy = 0, vz = 0;
#pragma unroll
            for (int t = 0; t < W; ++t) {
                const bool tcore = t >= REACH && t < REACH + ORDER;
                const int lx = local_plane(ix + t - REACH, g);
                const T* plane = phi + (size_t)(lx < 0 ? 0 : lx) * g.G * g.Gp;
#pragma unroll
                for (int b = 0; b < W; ++b) {
                    const bool bcore = b >= REACH && b < REACH + ORDER;
                    if (!tcore && !bcore) continue;
                    const T* row = plane + (size_t)wrap(iy + b - REACH, g.G) * g.Gp;
                    double s0 = 0, sz = 0;
                    if (tcore && bcore) {
#pragma unroll
                        for (int cc = 0; cc < W; ++cc) {
                            const double v = (double)row[oz[cc]];
                            sz += dz[cc] * v;
                            if (cc >= REACH && cc < REACH + ORDER) s0 += wz[cc - REACH] * v;
                        }
                        vz += (wx[t - REACH] * wy[b - REACH]) * sz;
                        vx += (dx[t] * wy[b - REACH]) * s0;
                        vy += (wx[t - REACH] * dy[b]) * s0;
                    } else {
#pragma unroll
                        for (int cc = 0; cc < ORDER; ++cc) s0 += wz[cc] * (double)row[oz[cc + REACH]];
                        if (bcore) vx += (dx[t] * wy[b - REACH]) * s0;       // x arm
                        else vy += (wx[t - REACH] * dy[b]) * s0;             // y arm
                    }
                }
            }
            if (factor != 1) { vx *= factor; vy *= factor; vz *= factor; }
            double* m = mom + ip * 3;
            const double mx = __ldcs(m) + vx, my = __ldcs(m + 1) + vy, mz = __ldcs(m + 2) + vz;
            __stcs(m, mx); __stcs(m + 1, my); __stcs(m + 2, mz);
            mom2_acc += mx * mx + my * my + mz * mz;
            if constexpr (DRIFT) {
                // Component.drift (species.py:2191-2196) with the freshly kicked momenta
                __stcs(pp, mod_box(px + mx * drift_dt, boxsize));
                __stcs(pp + 1, mod_box(py + my * drift_dt, boxsize));
                __stcs(pp + 2, mod_box(pz + mz * drift_dt, boxsize));
            }
        }
    }
    if (sum_mom2 != nullptr) {
        __syncthreads();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mom2_acc += __shfl_xor_sync(0xffffffffu, mom2_acc, o);
        if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = mom2_acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0;
            for (int w = 0; w < kGkBlock / 32; ++w) t += sred[w];
            atomicAdd(sum_mom2, t);
        }
    }
}

template <typename T>
static int gather_kick_dispatch(pm_ctx* c, double* pos, double* mom, int64_t n, int order,
                                const FD& fd, double factor, const Coord& co, double* sum_mom2,
                                bool drift, double drift_dt) {
    const int64_t nchunks = (n + (int64_t)kGkBlock * kGkChunkTiles - 1) / ((int64_t)kGkBlock * kGkChunkTiles);
    const int grid = (int)std::min<int64_t>(nchunks, (int64_t)kNumSMs * 8);
    const T* phi = reinterpret_cast<const T*>(c->grid_read());
    unsigned long long* ctr = c->d_tilectr + 1;
    PM_CHECK_CUDA(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), c->stream));
    // the streaming form for the wide stencils (TSC, PCS, differences of order ≥ 4); CIC with two-point differences keeps
    // the line form (measured on B200, profiles/r02_gather_forms.md); PM_GATHER_FORM=1|2 forces one of them
    static const int form_env = getenv("PM_GATHER_FORM") ? atoi(getenv("PM_GATHER_FORM")) : 0;
    const bool streaming = form_env ? form_env == 2 : !(order <= 2 && fd.reach == 1);
#define PM_GK(O, R)                                                                                     \
    do {                                                                                                \
        if (streaming) {                                                                                \
            if (drift)                                                                                  \
                PM_LAUNCH((gather_kick2_kernel<O, R, T, true>), grid, kGkBlock, 0, c->stream, phi, pos, mom, n, \
                          c->g, co, fd, factor, sum_mom2, ctr, drift_dt, c->boxsize);                   \
            else                                                                                        \
                PM_LAUNCH((gather_kick2_kernel<O, R, T, false>), grid, kGkBlock, 0, c->stream, phi, pos, mom, n, \
                          c->g, co, fd, factor, sum_mom2, ctr, drift_dt, c->boxsize);                   \
        } else if (drift)                                                                               \
            PM_LAUNCH((gather_kick_kernel<O, R, T, true>), grid, kGkBlock, 0, c->stream, phi, pos, mom, n, \
                      c->g, co, fd, factor, sum_mom2, ctr, drift_dt, c->boxsize);                       \
        else                                                                                            \
            PM_LAUNCH((gather_kick_kernel<O, R, T, false>), grid, kGkBlock, 0, c->stream, phi, pos, mom, n, \
                      c->g, co, fd, factor, sum_mom2, ctr, drift_dt, c->boxsize);                       \
    } while (0)
#define PM_GK_ORDER(R)                                  \
    switch (order) {                                    \
        case 1: PM_GK(1, R); break;                     \
        case 2: PM_GK(2, R); break;                     \
        case 3: PM_GK(3, R); break;                     \
        case 4: PM_GK(4, R); break;                     \
    }
    switch (fd.reach) {
        case 1: PM_GK_ORDER(1); break;
        case 2: PM_GK_ORDER(2); break;
        case 3: PM_GK_ORDER(3); break;
        case 4: PM_GK_ORDER(4); break;
    }
#undef PM_GK_ORDER
#undef PM_GK
    return PM_OK;
}

int launch_gather_kick(pm_ctx* c, double* pos, double* mom, int64_t n, int order,
                       int diff_order, double factor, const double* shift, double* sum_mom2,
                       bool drift, double drift_dt) {
    PM_REQUIRE(order >= 1 && order <= 4, "pm_gather_kick called with order = %d not in {1, 2, 3, 4}", order);
    FD fd;
    PM_TRY(make_fd(diff_order, c->boxsize / c->g.G, &fd));
    if (n == 0) return PM_OK;
    const Coord co = make_coord(c, shift, true);
    return c->dtype == PM_GRID_F64
               ? gather_kick_dispatch<double>(c, pos, mom, n, order, fd, factor, co, sum_mom2, drift, drift_dt)
               : gather_kick_dispatch<float>(c, pos, mom, n, order, fd, factor, co, sum_mom2, drift, drift_dt);
}

}  // namespace pm
