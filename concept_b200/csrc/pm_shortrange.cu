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

#include <cub/device/device_scan.cuh>

namespace pm {

struct CellGeom {
    int nc;             // cells per side
    double inv_cell;    // nc / L
    double L;
};

__device__ __forceinline__ int cell_coord(double x, const CellGeom& g) {
    int c = (int)(x * g.inv_cell);
    if (c < 0) c = 0;
    if (c >= g.nc) c = g.nc - 1;
    return c;
}

__global__ void __launch_bounds__(256)
cell_count_kernel(const double* __restrict__ pos, int64_t n, CellGeom g, int* __restrict__ cell_of,
                  int* __restrict__ count) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (cell_coord(pos[3 * i], g) * g.nc + cell_coord(pos[3 * i + 1], g)) * g.nc + cell_coord(pos[3 * i + 2], g);
        cell_of[i] = c;
        atomicAdd(&count[c], 1);
    }
}

__global__ void __launch_bounds__(256)
cell_scatter_kernel(const double* __restrict__ pos, int64_t n, const int* __restrict__ cell_of,
                    const int* __restrict__ offset, int* __restrict__ cursor, double* __restrict__ pos_s,
                    int* __restrict__ idx_s) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = cell_of[i];
        const int slot = offset[c] + atomicAdd(&cursor[c], 1);
        pos_s[3 * slot] = pos[3 * i];
        pos_s[3 * slot + 1] = pos[3 * i + 1];
        pos_s[3 * slot + 2] = pos[3 * i + 2];
        idx_s[slot] = (int)i;
    }
}

struct PairParams {
    double range2;      // range²
    double scaling;     // (T − 1)/r²_max
    double L, half_L;
    int lowest_active_rung;
};

// One thread per receiver (in cell order).  Gather form: no atomics on Δmom.
__global__ void __launch_bounds__(128)
shortrange_kernel(const double* __restrict__ pos_s, const int* __restrict__ idx_s, const int* __restrict__ offset,
                  int64_t n, CellGeom g, PairParams pp, const double* __restrict__ table,
                  const signed char* __restrict__ rung, const signed char* __restrict__ rung_jumped,
                  const double* __restrict__ factors, double* __restrict__ dmom) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int i = idx_s[s];
    if (rung[i] < pp.lowest_active_rung) return;      // inactive receivers get nothing (interactions.py:1700-1706)
    const double xi = pos_s[3 * s], yi = pos_s[3 * s + 1], zi = pos_s[3 * s + 2];
    const int cx = cell_coord(xi, g), cy = cell_coord(yi, g), cz = cell_coord(zi, g);
    double sx = 0, sy = 0, sz = 0;
    for (int dx = -1; dx <= 1; ++dx) {
        const int nx = (cx + dx + g.nc) % g.nc;
        for (int dy = -1; dy <= 1; ++dy) {
            const int ny = (cy + dy + g.nc) % g.nc;
            for (int dz = -1; dz <= 1; ++dz) {
                const int nz = (cz + dz + g.nc) % g.nc;
                const int c = (nx * g.nc + ny) * g.nc + nz;
                const int jb = offset[c], je = offset[c + 1];
                for (int j = jb; j < je; ++j) {
                    if (j == s) continue;
                    double x = xi - pos_s[3 * j], y = yi - pos_s[3 * j + 1], z = zi - pos_s[3 * j + 2];
                    // periodic minimum image (periodic_offset_x/y/z of the reference's tile pairing)
                    if (x > pp.half_L) x -= pp.L; else if (x < -pp.half_L) x += pp.L;
                    if (y > pp.half_L) y -= pp.L; else if (y < -pp.half_L) y += pp.L;
                    if (z > pp.half_L) z -= pp.L; else if (z < -pp.half_L) z += pp.L;
                    const double r2 = x * x + y * y + z * z;
                    if (r2 > pp.range2) continue;
                    const double f = table[(int)(r2 * pp.scaling)];
                    sx += x * f; sy += y * f; sz += z * f;
                }
            }
        }
    }
    const double factor = factors[rung_jumped[i]];
    dmom[3 * (size_t)i] = sx * factor;
    dmom[3 * (size_t)i + 1] = sy * factor;
    dmom[3 * (size_t)i + 2] = sz * factor;
}

static int ensure_bytes(pm_ctx* c, void** buf, size_t* have, size_t need) {
    if (need <= *have) return PM_OK;
    if (*buf) { cudaFree(*buf); c->bytes_allocated -= *have; }
    PM_CHECK_CUDA(cudaMalloc(buf, need));
    *have = need;
    c->bytes_allocated += need;
    return PM_OK;
}

int shortrange(pm_ctx* c, const double* pos, int64_t n, const signed char* rung, const signed char* rung_jumped,
               int lowest_active_rung, const double* factors_host, int nfactors, double range,
               const double* table_dev, int tablesize, double maxr2, double* dmom) {
    PM_REQUIRE(c->nranks == 1, "pm_shortrange: the P3M short-range force is single-GPU in this round");
    PM_REQUIRE(n < (int64_t)1 << 31, "pm_shortrange: too many particles for 32-bit cell lists");
    PM_REQUIRE(nfactors > 0 && nfactors <= 64, "pm_shortrange: nfactors = %d", nfactors);
    if (n == 0) return PM_OK;
    CellGeom g;
    g.L = c->boxsize;
    g.nc = (int)floor(c->boxsize / range);
    // the reference requires at least 4 tiles across the box (species.py:3971-3975); 3 is the minimum for
    // a 27-cell neighbourhood without double counting
    PM_REQUIRE(g.nc >= 3, "pm_shortrange: range %g is too large for the box %g (need boxsize/range >= 3)", range, c->boxsize);
    if (g.nc > 256) g.nc = 256;
    g.inv_cell = g.nc / c->boxsize;
    const size_t ncell = (size_t)g.nc * g.nc * g.nc;
    // scratch: cell_of[n] idx_s[n] count[ncell+1] offset[ncell+1] pos_s[3n] factors[64]
    const size_t need = sizeof(int) * (2 * (size_t)n + 2 * (ncell + 1)) + sizeof(double) * (3 * (size_t)n + 64) + 256;
    PM_TRY(ensure_bytes(c, &c->sr_buf, &c->sr_bytes, need));
    double* pos_s = reinterpret_cast<double*>(c->sr_buf);
    double* d_factors = pos_s + 3 * n;
    int* cell_of = reinterpret_cast<int*>(d_factors + 64);
    int* idx_s = cell_of + n;
    int* count = idx_s + n;
    int* offset = count + (ncell + 1);
    PM_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (ncell + 1), c->stream));
    PM_CHECK_CUDA(cudaMemcpyAsync(d_factors, factors_host, sizeof(double) * nfactors, cudaMemcpyHostToDevice, c->stream));
    PM_LAUNCH(cell_count_kernel, kNumSMs * 4, 256, 0, c->stream, pos, n, g, cell_of, count);
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, count, offset, (int)(ncell + 1), c->stream);
    PM_TRY(ensure_bytes(c, &c->sr_tmp, &c->sr_tmp_bytes, tmp_bytes));
    PM_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(c->sr_tmp, tmp_bytes, count, offset, (int)(ncell + 1), c->stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PM_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * (ncell + 1), c->stream));   // reuse as cursor
    PM_LAUNCH(cell_scatter_kernel, kNumSMs * 4, 256, 0, c->stream, pos, n, cell_of, offset, count, pos_s, idx_s);
    PairParams pp;
    pp.range2 = range * range;
    pp.scaling = (tablesize - 1) / maxr2;
    pp.L = c->boxsize;
    pp.half_L = 0.5 * c->boxsize;
    pp.lowest_active_rung = lowest_active_rung;
    PM_LAUNCH(shortrange_kernel, (unsigned)((n + 127) / 128), 128, 0, c->stream, pos_s, idx_s, offset, n, g, pp,
              table_dev, rung, rung_jumped, d_factors, dmom);
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
