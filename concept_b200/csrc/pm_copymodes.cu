// pm_copymodes.cu — Fourier slabs between grids of different size: the reference's copy_modes
// (mesh.py:980-1322), which particle_mesh uses when components have their own upstream/downstream grid sizes
// (add_upstream_to_global_slabs, mesh.py:618-710; interactions.py:2120-2140).  One rank per context: on several
// ranks the modes of one x-slab land in other ranks' slabs (the reference's subslab exchange, mesh.py:1105-1230),
// which is not built.  One streaming pass over the destination slab: 16 B read + 16 B written (+16 B read for '+=')
// per shared mode.
#include "pm_internal.cuh"
#include "pm_copy_ops.cuh"

namespace pm {

__global__ void __launch_bounds__(256)
copy_modes_kernel(const double2* __restrict__ src, double2* __restrict__ dst, copyops::CopyParams p,
                  const double* __restrict__ tab_x, const double* __restrict__ tab_sin, int accumulate) {
    const int64_t total = (int64_t)p.Gd * p.Gd * (p.Gd / 2 + 1);
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        double2 v;
        const bool shared = copyops::copy_mode(idx, src, p, tab_x, tab_sin, &v);
        if (!accumulate) {
            dst[idx] = v;
        } else if (shared) {
            double2 d = dst[idx];
            d.x += v.x; d.y += v.y;
            dst[idx] = d;
        }
    }
}

}  // namespace pm

using namespace pm;

extern "C" int pm_fourier_copy_modes(pm_ctx* src, pm_ctx* dst, int deconv_order, const double* shift, double scale,
                                     int src_saved, int dst_saved, int accumulate) {
    PM_REQUIRE(src != nullptr && dst != nullptr && src != dst, "pm_fourier_copy_modes: need two different contexts");
    PM_REQUIRE(src->dtype == PM_GRID_F64 && dst->dtype == PM_GRID_F64, "pm_fourier_copy_modes: PM_GRID_F64 contexts only");
    PM_REQUIRE(src->nranks == 1 && dst->nranks == 1, "pm_fourier_copy_modes: component-specific grid sizes need one rank per context");
    PM_REQUIRE(src->device == dst->device, "pm_fourier_copy_modes: contexts live on different devices");
    PM_REQUIRE(src->g.G != dst->g.G, "pm_fourier_copy_modes: equal grid sizes (use pm_fourier_operate)");
    PM_REQUIRE(deconv_order >= 0 && deconv_order <= 64, "pm_fourier_copy_modes: deconv_order = %d out of range", deconv_order);
    if (src_saved) PM_REQUIRE(src->saved != nullptr, "pm_fourier_copy_modes: the source has no saved slab (pm_slab_save)");
    else PM_REQUIRE(src->space_fourier, "pm_fourier_copy_modes: the source slab holds real-space data");
    if (dst_saved) {
        PM_REQUIRE(!accumulate || dst->saved != nullptr, "pm_fourier_copy_modes: '+=' onto a saved slab that does not exist");
        PM_TRY(ensure_saved(dst));
    } else {
        PM_REQUIRE(!accumulate || dst->space_fourier, "pm_fourier_copy_modes: '+=' onto a slab that holds real-space data");
    }
    copyops::CopyParams p;
    p.Gs = src->g.G;
    p.Gd = dst->g.G;
    p.deconv_order = deconv_order;
    p.rotate = 0;
    for (int d = 0; d < 3; ++d) {
        const double s = shift ? shift[d] : 0.0;
        p.th[d] = -2 * M_PI / p.Gs * s;
        if (s != 0.0) p.rotate = 1;
    }
    p.cell_phase = M_PI / p.Gd - M_PI / p.Gs;
    p.scale = scale;
    if (src->stream != dst->stream) {
        cudaEvent_t ev;
        PM_CHECK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        PM_CHECK_CUDA(cudaEventRecord(ev, src->stream));
        PM_CHECK_CUDA(cudaStreamWaitEvent(dst->stream, ev, 0));
        PM_CHECK_CUDA(cudaEventDestroy(ev));
    }
    PM_LAUNCH(copy_modes_kernel, kNumSMs * 8, 256, 0, dst->stream,
              reinterpret_cast<const double2*>(src_saved ? src->saved : src->fourier),
              reinterpret_cast<double2*>(dst_saved ? dst->saved : dst->fourier), p, src->tab_x, src->tab_sin,
              accumulate ? 1 : 0);
    if (src->stream != dst->stream) {
        cudaEvent_t ev;
        PM_CHECK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        PM_CHECK_CUDA(cudaEventRecord(ev, dst->stream));
        PM_CHECK_CUDA(cudaStreamWaitEvent(src->stream, ev, 0));
        PM_CHECK_CUDA(cudaEventDestroy(ev));
    }
    if (!dst_saved) {
        dst->space_fourier = true;
        dst->grid_in_phi = false;
        dst->real_is_zero = false;
    }
    return PM_OK;
}
