// pm_sort.cu — reorder the particle arrays by grid cell (x plane, y row, z) so that consecutive particles
// touch neighbouring grid rows: the deposit and the gather rely on that locality (L1/L2 reuse of grid rows;
// measured 3x slower deposit and 9x slower gather on randomly ordered particles).  Simulations start from
// lattice order and lose it slowly; the reference re-sorts its particles by tile for the same reason
// (Component.tile_sort, species.py:2657-2780).  Stable LSD radix sort of (cell key, index) pairs (CUB), then
// one gather pass per array.
#include "pm_internal.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace pm {

__global__ void __launch_bounds__(256)
cell_key_kernel(const double* __restrict__ pos, int64_t n, double cells_per_len, int G,
                unsigned long long* __restrict__ keys, unsigned int* __restrict__ idx) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int c[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            int v = (int)(pos[i * 3 + d] * cells_per_len);
            c[d] = v < 0 ? 0 : (v >= G ? G - 1 : v);
        }
        keys[i] = ((unsigned long long)c[0] * G + c[1]) * G + c[2];
        idx[i] = (unsigned int)i;
    }
}

template <typename T, int W>
__global__ void __launch_bounds__(256)
permute_kernel(const T* __restrict__ src, T* __restrict__ dst, const unsigned int* __restrict__ idx, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const size_t s = (size_t)idx[i] * W;
#pragma unroll
        for (int w = 0; w < W; ++w) dst[i * W + w] = src[s + w];
    }
}

int sort_particles(pm_ctx* c, double* pos, double* mom, int64_t* ids, int64_t n) {
    PM_REQUIRE(n >= 0 && n < (int64_t)1 << 32, "pm_sort_particles: n = %lld out of range", (long long)n);
    if (n < 2) return PM_OK;
    const int G = c->g.G;
    int bits = 1;
    while (((unsigned long long)1 << bits) < (unsigned long long)G * G * G) ++bits;
    // scratch: keys ×2, indices ×2, CUB workspace, one (N,3) double array for the permutation
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (unsigned int*)nullptr, (unsigned int*)nullptr, (int)n, 0, bits, c->stream);
    const size_t kb = ((size_t)n * 8 + 255) / 256 * 256, ib = ((size_t)n * 4 + 255) / 256 * 256;
    const size_t ab = ((size_t)n * 24 + 255) / 256 * 256;
    const size_t need = 2 * kb + 2 * ib + ab + cub_bytes + 256;
    if (need > c->sr_tmp_bytes) {
        if (c->sr_tmp) { cudaFree(c->sr_tmp); c->bytes_allocated -= c->sr_tmp_bytes; c->sr_tmp = nullptr; c->sr_tmp_bytes = 0; }
        PM_CHECK_CUDA(cudaMalloc(&c->sr_tmp, need));
        c->sr_tmp_bytes = need;
        c->bytes_allocated += need;
    }
    char* base = reinterpret_cast<char*>(c->sr_tmp);
    auto* keys_in = reinterpret_cast<unsigned long long*>(base);
    auto* keys_out = reinterpret_cast<unsigned long long*>(base + kb);
    auto* idx_in = reinterpret_cast<unsigned int*>(base + 2 * kb);
    auto* idx_out = reinterpret_cast<unsigned int*>(base + 2 * kb + ib);
    double* tmp = reinterpret_cast<double*>(base + 2 * kb + 2 * ib);
    void* cub_ws = base + 2 * kb + 2 * ib + ab;
    const int grid = kNumSMs * 8;
    PM_LAUNCH(cell_key_kernel, grid, 256, 0, c->stream, pos, n, G / c->boxsize, G, keys_in, idx_in);
    PM_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, keys_in, keys_out, idx_in, idx_out, (int)n, 0, bits, c->stream));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    PM_LAUNCH((permute_kernel<double, 3>), grid, 256, 0, c->stream, pos, tmp, idx_out, n);
    PM_CHECK_CUDA(cudaMemcpyAsync(pos, tmp, sizeof(double) * 3 * n, cudaMemcpyDeviceToDevice, c->stream));
    PM_LAUNCH((permute_kernel<double, 3>), grid, 256, 0, c->stream, mom, tmp, idx_out, n);
    PM_CHECK_CUDA(cudaMemcpyAsync(mom, tmp, sizeof(double) * 3 * n, cudaMemcpyDeviceToDevice, c->stream));
    if (ids != nullptr) {
        int64_t* t64 = reinterpret_cast<int64_t*>(tmp);
        PM_LAUNCH((permute_kernel<int64_t, 1>), grid, 256, 0, c->stream, ids, t64, idx_out, n);
        PM_CHECK_CUDA(cudaMemcpyAsync(ids, t64, sizeof(int64_t) * n, cudaMemcpyDeviceToDevice, c->stream));
    }
    return PM_OK;
}

}  // namespace pm

extern "C" int pm_sort_particles(pm_ctx* c, double* pos, double* mom, int64_t* ids, int64_t n) {
    PM_REQUIRE(c != nullptr && (n == 0 || (pos != nullptr && mom != nullptr)), "pm_sort_particles: bad argument");
    return pm::sort_particles(c, pos, mom, ids, n);
}
