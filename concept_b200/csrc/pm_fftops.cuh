// pm_fftops.cuh — the tile operations of the hand-written slab transform (pm_fft.cu), built from
// the stages in pm_fftcore.cuh.  Each operation is a sequence of phases separated by a block barrier;
// phases take (tid, nthr) so the CPU harness (tests/fft_host_harness.cu) can walk through a whole
// 3-D solve with exactly this code.  Phase 0 reads its inputs straight from global memory
// (ld.global.cg, 128-byte row segments) into registers; the last phase stores from registers; in
// between the tile lives in ONE shared-memory buffer that every stage updates in place.
//
// Slab layout (per rank): real T[nxl][G][Gp], Gp = G + 2; in place complex V[nxl][G][Gc], Gc = G/2 + 1
// (what fft.c:124 calls the padded slab).  The kk = G/2 column is never read or transformed: the
// potential nullifies that Nyquist plane (mesh.py:3615-3622); the forward z pass stores a zero there.
//
//   ZFwd   CZ rows of one plane: r2c along z               (contiguous 4 KB rows)
//   YPass  CY adjacent kk columns of one plane: c2c along y (forward or inverse; 128 B row segments)
//   XSolve CY adjacent kk columns of one j row, all planes (of all ranks): forward c2c along x,
//          Green's function (interactions.py:2092-2118, mesh.py:2775-2856, :3585-3622), inverse c2c
//   ZInv   CZ rows of one plane: c2r along z
#pragma once

#include "pm_fftcore.cuh"

namespace pm {
namespace fftc {

constexpr int kMaxFftPeers = 16;

// streaming global load: L2 only (the 2-D kernels hand tiles from one SM to another inside a launch,
// which L1 would not notice; nothing here is re-read anyway)
template <typename V>
PM_HD V ld_stream(const V* p) {
#ifdef __CUDA_ARCH__
    return __ldcg(p);
#else
    return *p;
#endif
}

template <typename T, int G_>
struct SlabFFT {
    using V = typename Vec2<T>::type;
    using TW = Twiddles<V>;
    static constexpr int G = G_;
    static constexpr int M = G / 2;          // complex points of the packed real transform
    static constexpr int Gc = M + 1;
    static constexpr int Gp = 2 * Gc;
    static constexpr int CY = 128 / (int)sizeof(V);   // columns per y/x tile (8 in fp64, 16 in fp32)
    static constexpr int CZ = 8;                      // rows per z tile
    using LZ = RowLayout<M, CZ>;
    using LY = ColLayout<CY>;
    static constexpr int kZTileElems = LZ::PITCH * CZ;
    static constexpr int kYTileElems = G * CY;
    static constexpr int kWorkElems = (kZTileElems > kYTileElems) ? kZTileElems : kYTileElems;
    static constexpr int kZTilesPerPlane = G / CZ;
    static constexpr int kYTilesPerPlane = M / CY;    // kk = 0 … G/2−1
    static_assert(G % 64 == 0 && M % 64 == 0 && G / 64 <= 8, "G must be 128, 256 or 512");

    struct RowSource {     // complex element k of row c
        const T* plane; int row0;
        PM_HD V operator()(int c, int k) const {
            return ld_stream(reinterpret_cast<const V*>(plane + (size_t)(row0 + c) * Gp) + k);
        }
    };
    struct RowSink {
        T* plane; int row0;
        PM_HD void operator()(int c, int k, T r, T i) const {
            V v; v.x = r; v.y = i;
            reinterpret_cast<V*>(plane + (size_t)(row0 + c) * Gp)[k] = v;
        }
    };

    // ------------------------------------------------------------------ z forward (r2c)
    struct ZFwd {
        T* plane;     // first real of the x plane
        int row0;     // first of the CZ rows
        struct ToTile {
            V* tile;
            PM_HD void operator()(int c, int idx, T r, T i) const { V v; v.x = r; v.y = i; tile[LZ::idx(idx, c)] = v; }
        };
        static constexpr int kPhases = 4;
        PM_HD void phase(int ph, V* work, const TW& tw, int tid, int nthr) const {
            if (ph == 0) dit_stageA<LZ, T, M, -1>(RowSource{plane, row0}, work, tid, nthr);
            else if (ph == 1) dit_stageB<LZ, T, M, -1>(work, tw.B, tid, nthr);
            else if (ph == 2) dit_stageC<LZ, T, M, 2, -1>(work, tw.C, tid, nthr, ToTile{work});
            else r2c_post<LZ, T, M>(work, tw.R, tid, nthr, RowSink{plane, row0});
        }
    };

    // ------------------------------------------------------------------ z inverse (c2r)
    struct ZInv {
        T* plane;
        int row0;
        static constexpr int kPhases = 4;
        PM_HD void phase(int ph, V* work, const TW& tw, int tid, int nthr) const {
            if (ph == 0) c2r_pre<LZ, T, M>(RowSource{plane, row0}, work, tw.R, tid, nthr);
            else if (ph == 1) dit_stageA_inplace<LZ, T, M, +1>(work, tid, nthr);
            else if (ph == 2) dit_stageB<LZ, T, M, +1>(work, tw.B, tid, nthr);
            else dit_stageC<LZ, T, M, 2, +1>(work, tw.C, tid, nthr, RowSink{plane, row0});   // z_m = x_2m + i·x_2m+1
        }
    };

    // ------------------------------------------------------------------ y pass (c2c, either direction)
    template <int DIR>
    struct YPass {
        V* plane;     // complex view of the x plane: [G][Gc]
        int kk0;      // first of the CY columns
        struct Source {
            const V* plane; int kk0;
            PM_HD V operator()(int c, int j) const { return ld_stream(plane + (size_t)j * Gc + kk0 + c); }
        };
        struct ToPlane {
            V* plane; int kk0;
            PM_HD void operator()(int c, int j, T r, T i) const {
                V v; v.x = r; v.y = i;
                plane[(size_t)j * Gc + kk0 + c] = v;
            }
        };
        static constexpr int kPhases = 3;
        PM_HD void phase(int ph, V* work, const TW& tw, int tid, int nthr) const {
            if (ph == 0) dit_stageA<LY, T, G, DIR>(Source{plane, kk0}, work, tid, nthr);
            else if (ph == 1) dit_stageB<LY, T, G, DIR>(work, tw.B, tid, nthr);
            else dit_stageC<LY, T, G, 1, DIR>(work, tw.C, tid, nthr, ToPlane{plane, kk0});
        }
    };

    // ------------------------------------------------------------------ x solve
    struct XGeom {
        V* base[kMaxFftPeers];   // rank r's first interior plane as complex [nxl][G][Gc]
        int nxl_shift;           // planes per rank = 1 << nxl_shift
        const double* sep;       // separable factor per axis index l < G: (x_l/sin x_l)^D·exp(−gauss·k_l²)
        double prefactor;        // −L²·G_N/π
        PM_HD V* at(int i, int j, int kk) const {
            const int r = i >> nxl_shift, il = i & ((1 << nxl_shift) - 1);
            return base[r] + ((size_t)il * G + j) * Gc + kk;
        }
    };
    struct XSolve {
        const XGeom* g;
        int j;        // global j row
        int kk0;      // first of the CY columns
        struct Source {
            const XGeom* g; int j, kk0;
            PM_HD V operator()(int c, int i) const { return ld_stream(g->at(i, j, kk0 + c)); }
        };
        struct ToSlab {
            const XGeom* g; int j, kk0;
            PM_HD void operator()(int c, int i, T r, T im) const {
                V v; v.x = r; v.y = im;
                *g->at(i, j, kk0 + c) = v;
            }
        };
        // per-mode factor in separable form: Π_l sep[l] · prefactor/k²; zero on the Nyquist planes and at
        // the origin.  (kspace_kernel in pm_fourier.cu keeps the reference's operation order; the two
        // agree to a few ulp.)
        PM_HD double factor(int i, double sep_jk, int kj2_kk2, bool line_nyq) const {
            if (line_nyq || i == M) return 0.0;
            const int ki = i - (i >= M ? G : 0);
            const int k2 = kj2_kk2 + ki * ki;
            if (k2 == 0) return 0.0;
#ifdef __CUDA_ARCH__
            return (g->sep[i] * sep_jk) * (g->prefactor * __drcp_rn((double)k2));
#else
            return (g->sep[i] * sep_jk) * (g->prefactor * (1.0 / (double)k2));
#endif
        }
        static constexpr int kPhases = 5;
        PM_HD void phase(int ph, V* work, const TW& tw, int tid, int nthr) const {
            constexpr int R1 = G / 64;
            if (ph == 0) dit_stageA<LY, T, G, -1>(Source{g, j, kk0}, work, tid, nthr);
            else if (ph == 1) dit_stageB<LY, T, G, -1>(work, tw.B, tid, nthr);
            else if (ph == 2) {
                // forward stage C, Green's function, inverse stage 1 — all on the same R1 registers
                const int kj = j - (j >= M ? G : 0);
                for (int b = tid; b < 64 * CY; b += nthr) {
                    int c, q;
                    LY::template decode<64>(b, c, q);
                    T r[R1], im[R1];
#pragma unroll
                    for (int a1 = 0; a1 < R1; ++a1) {
                        const V v = work[LY::idx(64 * a1 + q, c)];
                        r[a1] = v.x; im[a1] = v.y;
                        if (a1) cmul<-1>(r[a1], im[a1], tw.C[a1 * 64 + q]);
                    }
                    dftR<R1, -1>(r, im);
                    const int kk = kk0 + c;
                    const double sep_jk = g->sep[j] * g->sep[kk];
                    const int kj2_kk2 = kj * kj + kk * kk;
                    const bool line_nyq = (j == M) || (kk == M);
#pragma unroll
                    for (int b1 = 0; b1 < R1; ++b1) {
                        const T f = (T)factor(64 * b1 + q, sep_jk, kj2_kk2, line_nyq);
                        r[b1] *= f; im[b1] *= f;
                    }
                    dif_stage1_regs<LY, T, G, 1, +1>(r, im, work, tw.C, c, q);
                }
            } else if (ph == 3) dif_stage2<LY, T, G, +1>(work, tw.B, tid, nthr);
            else dif_stage3<LY, T, G, +1>(work, tid, nthr, ToSlab{g, j, kk0});
        }
    };
};

}  // namespace fftc
}  // namespace pm
