// pm_copymodes.cu — Fourier slabs between grids of different size: the reference's copy_modes
// (mesh.py:980-1322), which particle_mesh uses when components have their own upstream/downstream grid sizes
// (add_upstream_to_global_slabs, mesh.py:618-710; interactions.py:2120-2140).
// One rank: one streaming pass over the destination slab, 16 B read + 16 B written (+16 B read for '+=') per shared mode.
// Several ranks: both slabs are distributed over their j rows ([i][j_local][kk]), so the row kj of the shared cube
// |k| < min(Gs, Gd)/2 lives on rank js/njl_s of the source and belongs on rank jd/njl_d of the destination — the reference's
// subslab exchange (mesh.py:1105-1230).  Every rank packs the rows it holds for each destination rank ((2n−1) × n modes
// per row), one grouped ncclSend/ncclRecv moves them, and the receiver applies the source grid's deconvolution and the
// phases while writing them into its slab.
#include "pm_internal.cuh"
#include "pm_copy_ops.cuh"

#include <algorithm>
#include <vector>

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

// ---- several ranks ------------------------------------------------------------------------------------------------
// row r of the send buffer: modes (ki, kj_r, kk), ki = a − (n − 1) for a < 2n − 1, kk < n, from the source slab row jl[r]
__global__ void __launch_bounds__(256)
copy_modes_pack_kernel(const double2* __restrict__ src, double2* __restrict__ out, const int* __restrict__ jl, int nrows,
                       int n, int Gs, int njl_s) {
    const int Gcs = Gs / 2 + 1, W = 2 * n - 1;
    const int64_t total = (int64_t)nrows * W * n;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int kk = (int)(idx % n);
        const int a = (int)((idx / n) % W);
        const int r = (int)(idx / ((int64_t)n * W));
        const int ki = a - (n - 1);
        const int is = ki < 0 ? ki + Gs : ki;
        out[idx] = src[((int64_t)is * njl_s + jl[r]) * Gcs + kk];
    }
}

// row r of the receive buffer belongs to mode row kj[r] and destination slab row jl[r]
__global__ void __launch_bounds__(256)
copy_modes_unpack_kernel(const double2* __restrict__ in, double2* __restrict__ dst, const int* __restrict__ kj_of,
                         const int* __restrict__ jl, int nrows, int n, copyops::CopyParams p, int njl_d,
                         const double* __restrict__ tab_x, const double* __restrict__ tab_sin, int accumulate) {
    const int Gcd = p.Gd / 2 + 1, W = 2 * n - 1;
    const int64_t total = (int64_t)nrows * W * n;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int kk = (int)(idx % n);
        const int a = (int)((idx / n) % W);
        const int r = (int)(idx / ((int64_t)n * W));
        const int ki = a - (n - 1), kj = kj_of[r];
        const int is = ki < 0 ? ki + p.Gs : ki, js = kj < 0 ? kj + p.Gs : kj;
        const int id = ki < 0 ? ki + p.Gd : ki;
        const double2 v = copyops::mode_value(in[idx], ki, kj, kk, is, js, p, tab_x, tab_sin);
        double2* q = dst + ((int64_t)id * njl_d + jl[r]) * Gcd + kk;
        if (accumulate) { double2 d = *q; d.x += v.x; d.y += v.y; *q = d; }
        else *q = v;
    }
}

// stream-ordered scratch that is released on every way out of the function
struct AsyncScratch {
    void* ptr = nullptr;
    cudaStream_t stream;
    explicit AsyncScratch(cudaStream_t s) : stream(s) {}
    ~AsyncScratch() { if (ptr) cudaFreeAsync(ptr, stream); }
    AsyncScratch(const AsyncScratch&) = delete;
    AsyncScratch& operator=(const AsyncScratch&) = delete;
    cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&ptr, bytes ? bytes : 16, stream); }
};

static int copy_modes_ranks(pm_ctx* src, pm_ctx* dst, const copyops::CopyParams& p, const double2* from, double2* onto,
                            bool accumulate) {
    const int P = dst->nranks, me = dst->rank;
    const int n = std::min(p.Gs, p.Gd) / 2, W = 2 * n - 1;
    const int njl_s = src->g.njl, njl_d = dst->g.njl;
    // rows in the order both sides enumerate them: kj = −n+1 … n−1, grouped by the peer
    std::vector<std::vector<int>> send_jl(P), recv_kj(P), recv_jl(P);
    for (int kj = -n + 1; kj < n; ++kj) {
        const int js = kj < 0 ? kj + p.Gs : kj, jd = kj < 0 ? kj + p.Gd : kj;
        const int rs = js / njl_s, rd = jd / njl_d;
        if (rs == me) send_jl[rd].push_back(js - rs * njl_s);
        if (rd == me) { recv_kj[rs].push_back(kj); recv_jl[rs].push_back(jd - rd * njl_d); }
    }
    std::vector<int> h_send, h_kj, h_jl;
    std::vector<size_t> send_off(P + 1, 0), recv_off(P + 1, 0);
    for (int r = 0; r < P; ++r) {
        h_send.insert(h_send.end(), send_jl[r].begin(), send_jl[r].end());
        h_kj.insert(h_kj.end(), recv_kj[r].begin(), recv_kj[r].end());
        h_jl.insert(h_jl.end(), recv_jl[r].begin(), recv_jl[r].end());
        send_off[r + 1] = send_off[r] + send_jl[r].size();
        recv_off[r + 1] = recv_off[r] + recv_jl[r].size();
    }
    const size_t row_elems = (size_t)W * n;            // complex values per row
    const size_t ns = send_off[P], nr = recv_off[P];
    cudaStream_t st = dst->stream;
    AsyncScratch b_send(st), b_recv(st), b_idx(st);
    PM_CHECK_CUDA(b_send.alloc(sizeof(double2) * ns * row_elems));
    PM_CHECK_CUDA(b_recv.alloc(sizeof(double2) * nr * row_elems));
    PM_CHECK_CUDA(b_idx.alloc(sizeof(int) * (ns + 2 * nr)));
    double2 *d_send = static_cast<double2*>(b_send.ptr), *d_recv = static_cast<double2*>(b_recv.ptr);
    int* d_idx = static_cast<int*>(b_idx.ptr);
    if (ns) PM_CHECK_CUDA(cudaMemcpyAsync(d_idx, h_send.data(), sizeof(int) * ns, cudaMemcpyHostToDevice, st));
    if (nr) {
        PM_CHECK_CUDA(cudaMemcpyAsync(d_idx + ns, h_kj.data(), sizeof(int) * nr, cudaMemcpyHostToDevice, st));
        PM_CHECK_CUDA(cudaMemcpyAsync(d_idx + ns + nr, h_jl.data(), sizeof(int) * nr, cudaMemcpyHostToDevice, st));
    }
    PM_CHECK_CUDA(cudaStreamSynchronize(st));          // the index vectors above are pageable host memory about to go out of scope
    if (!accumulate) PM_CHECK_CUDA(cudaMemsetAsync(onto, 0, sizeof(double2) * dst->fourier_elems, st));   // modes outside the shared cube
    if (ns) PM_LAUNCH(copy_modes_pack_kernel, kNumSMs * 4, 256, 0, st, from, d_send, d_idx, (int)ns, n, p.Gs, njl_s);
    PM_CHECK_NCCL(ncclGroupStart());
    ncclResult_t posted = ncclSuccess;
    for (int r = 0; r < P && posted == ncclSuccess; ++r) {
        const size_t cs = (send_off[r + 1] - send_off[r]) * row_elems * 2, cr = (recv_off[r + 1] - recv_off[r]) * row_elems * 2;
        if (cs) posted = ncclSend(d_send + send_off[r] * row_elems, cs, ncclDouble, r, dst->comm, st);
        if (cr && posted == ncclSuccess) posted = ncclRecv(d_recv + recv_off[r] * row_elems, cr, ncclDouble, r, dst->comm, st);
    }
    PM_CHECK_NCCL(ncclGroupEnd());      // the group is closed whatever happened inside
    PM_CHECK_NCCL(posted);
    if (nr)
        PM_LAUNCH(copy_modes_unpack_kernel, kNumSMs * 4, 256, 0, st, d_recv, onto, d_idx + ns, d_idx + ns + nr, (int)nr, n, p, njl_d,
                  src->tab_x, src->tab_sin, accumulate ? 1 : 0);
    return PM_OK;      // the scratch is released in stream order behind the unpack kernel
}

}  // namespace pm

using namespace pm;

extern "C" int pm_fourier_copy_modes(pm_ctx* src, pm_ctx* dst, int deconv_order, const double* shift, double scale,
                                     int src_saved, int dst_saved, int accumulate) {
    PM_REQUIRE(src != nullptr && dst != nullptr && src != dst, "pm_fourier_copy_modes: need two different contexts");
    PM_REQUIRE(src->dtype == PM_GRID_F64 && dst->dtype == PM_GRID_F64, "pm_fourier_copy_modes: PM_GRID_F64 contexts only");
    PM_REQUIRE(src->nranks == dst->nranks && src->rank == dst->rank, "pm_fourier_copy_modes: the two contexts are distributed differently");
    PM_REQUIRE(dst->nranks == 1 || dst->comm_ready, "pm_fourier_copy_modes: the destination context has no communicator (pm_comm_init)");
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
    const double2* from = reinterpret_cast<const double2*>(src_saved ? src->saved : src->fourier);
    double2* onto = reinterpret_cast<double2*>(dst_saved ? dst->saved : dst->fourier);
    if (dst->nranks == 1)
        PM_LAUNCH(copy_modes_kernel, kNumSMs * 8, 256, 0, dst->stream, from, onto, p, src->tab_x, src->tab_sin, accumulate ? 1 : 0);
    else
        PM_TRY(copy_modes_ranks(src, dst, p, from, onto, accumulate != 0));
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
