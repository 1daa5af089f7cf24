// pm_fftops.cuh — the tile operations of the hand-written slab transform (pm_fft.cu), built from
// the stages in pm_fftcore.cuh.  Each operation names the contiguous global chunks that make up its
// tile (`loads`; the kernel moves them with cp.async.bulk, the CPU harness with memcpy) and a sequence
// of phases separated by a block barrier.  Phases take (tid, nthr) and a small per-thread register
// file that survives the barriers, so tests/fft_host_harness.cu can walk through a whole 3-D solve
// with exactly this code.  The tile is processed IN PLACE in the buffer the copy filled.
//
// Buffers per rank (complex type V, real type T, CY = 64 bytes / sizeof(V) columns per tile,
// NKT = (G/2)/CY column tiles; the kk = G/2 Nyquist column, which the potential nullifies
// (mesh.py:3615-3622), is never stored):
//   real  T[nxl][G][Gp]            padded real slab, Gp = G + 2 (what fft.c:124 allocates)
//   A     V[nxl][NKT][G][CY]       a (plane, column tile) block = one contiguous y-pass tile
//   B     V[NKT][G][nxl][CY]       a (column tile, j) block     = one contiguous x-pass tile per rank
// Every y/x tile is therefore one contiguous chunk (32 KB in fp64 at G = 512; nxl·64 bytes per peer in
// the x pass) that the copy engine moves, and every store to A or B is an aligned 64-byte segment.  The
// z tiles read and write whole contiguous rows directly (their `load` only names the chunk to prefetch):
//
//   ZFwd   8 rows of a plane:    real rows --r2c along z-->  A                (A ← 16-byte pieces, 64 B runs)
//   YFwd   A block              --c2c along y-->             B
//   XSolve B blocks of all ranks --c2c along x · Green's function · inverse c2c along x-->  A (of all ranks)
//          (interactions.py:2092-2118, mesh.py:2775-2856, :3585-3622; the FFTW-MPI transpose of
//          fft.c:34-73 never materialises)
//   YInv   A block              --inverse c2c along y-->     real rows (complex view)
//   ZInv   8 rows of a plane    --c2r along z-->             real rows
#pragma once

#include "pm_fftcore.cuh"

namespace pm {
namespace fftc {

constexpr int kMaxFftPeers = 16;



// streaming global load: L2 only (the 2-D kernels hand rows from one SM to another inside a launch, which
// L1 would not notice; nothing here is re-read anyway)
template <typename V>
PM_HD V ld_stream(const V* p) {
#ifdef __CUDA_ARCH__
    return __ldcg(p);
#else
    return *p;
#endif
}

struct TileLoad {
    const void* src;   // global
    int dst_bytes;     // byte offset inside the tile buffer
    int bytes;         // multiple of 16; src and dst 16-byte aligned
};

// Bytes of adjacent kk columns per y/x tile row (and per store segment into A, B and the complex rows): 64 or 128.
// G = 1024 keeps 64 (a 128-byte tile would be 128 KB).
#ifndef PM_FFT_CY_BYTES
#define PM_FFT_CY_BYTES 64
#endif
constexpr int cy_bytes(int G) { return G >= 1024 ? 64 : PM_FFT_CY_BYTES; }
// several rows per warp only while the z tile still fits into the y-tile buffer: RPW·(NTHR/32)·(G/2) ≤ G·CY, CY ≥ 4
constexpr bool rpw_fits(int G, int NTHR, int RPW) { return RPW * (NTHR / 32) * (G / 2) <= G * (cy_bytes(G) / 16); }

template <typename T, int G_, int NTHR_, int RPW_ = 1>
struct SlabFFT {
    using V = typename Vec2<T>::type;
    using TW = Twiddles<V>;
    static constexpr int G = G_;
    static constexpr int NTHR = NTHR_;
    static constexpr int M = G / 2;          // complex points of the packed real transform
    static constexpr int Gc = M + 1;
    static constexpr int Gp = 2 * Gc;
    static constexpr int CY = cy_bytes(G_) / (int)sizeof(V);   // columns per y/x tile: 4 in fp64, 8 in fp32 (64-byte groups)
    // rows per warp, their loads in flight together; RPW·(NTHR/32) rows are as large as a y tile (G × 64 bytes) when
    // NTHR = G/4·RPW/2 — at G = 1024 (256 threads) a second row would double the tile buffer
    static constexpr int RPW = rpw_fits(G_, NTHR_, RPW_) ? RPW_ : 1;
    static constexpr int CZ = RPW * (NTHR / 32);     // rows per z tile: every warp owns whole rows (no block barrier inside)
    static constexpr int NKT = M / CY;               // column tiles
    using LY = ColSwz<CY>;
    using LZ = RowSwz<M, RPW, Gc>;                   // the warp's private rows
    static constexpr int kYTileElems = G * CY;
    static constexpr int kZTileElems = CZ * M;
    static constexpr int kBufElems = (kZTileElems > kYTileElems) ? kZTileElems : kYTileElems;
    static constexpr int kZTilesPerPlane = G / CZ;
    static constexpr int kYTilesPerPlane = NKT;
    static constexpr int kNbtY = ((G / 8) * CY + NTHR - 1) / NTHR;            // first-stage butterflies per thread, y/x tiles
    static constexpr int kRegs = 16 * kNbtY;
    static_assert(G % 64 == 0 && M % 64 == 0 && G / 64 <= 16, "G must be 128, 256, 512 or 1024");

    static PM_HD size_t a_index(int il, int kt, int j, int c) { return (((size_t)il * NKT + kt) * G + j) * CY + c; }
    static PM_HD size_t b_index(int kt, int j, int il, int c, int nxl) { return (((size_t)kt * G + j) * nxl + il) * CY + c; }

    struct RowSink {     // complex slot k of row `row + c` of a padded real plane
        T* plane; int row;
        PM_HD void operator()(int c, int k, T r, T i) const {
            V v; v.x = r; v.y = i;
            reinterpret_cast<V*>(plane + (size_t)(row + c) * Gp)[k] = v;
        }
    };

    struct RowSource {     // complex element k of row `row + c`, straight from global memory (L2: the rows were prefetched)
        const T* plane; int row;
        PM_HD V operator()(int c, int k) const {
            return ld_stream(reinterpret_cast<const V*>(plane + (size_t)(row + c) * Gp) + k);
        }
    };

    // ------------------------------------------------------------------ z forward (r2c): real rows -> A
    // The rows are read directly (coalesced 512-byte warp loads, contiguous 4 KB rows) — staging them
    // through the copy engine would cost the z pass an extra shared-memory round trip and barrier.
    struct ZFwd {
        T* plane;         // first real of the x plane
        V* a_plane;       // A block row of this plane: V[NKT][G][CY]
        int row0;         // first of the CZ rows
        bool clear;       // nullify the rows once they are read (self-cleaning density grid)
        static constexpr bool kBulk = false;
        static constexpr bool kWarpPrivate = true;      // phases are separated by __syncwarp(), not by a block barrier
        PM_HD TileLoad load(int) const { return TileLoad{plane + (size_t)row0 * Gp, 0, CZ * Gp * (int)sizeof(T)}; }
        struct ToA {
            V* a_plane; int row;
            PM_HD void operator()(int c, int k, T r, T i) const {
                if (k >= M) return;     // the Nyquist column is not stored
                V v; v.x = r; v.y = i;
                a_plane[((size_t)(k / CY) * G + row + c) * CY + (k % CY)] = v;
            }
        };
        static constexpr int kPhases = 3;
        PM_HD void phase(int ph, V* tile, const TW& tw, int tid, int, T (&)[kRegs]) const {
            const int warp = tid >> 5, lane = tid & 31, row = row0 + warp * RPW;
            V* mine = tile + warp * (RPW * M);
            if (ph == 0) dit_stageA<LZ, T, M, -1>(RowSource{plane, row}, mine, lane, 32);
            else if (ph == 1) {
                if (clear) {     // every lane of the warp has consumed the rows by now (they are contiguous)
                    V zero; zero.x = 0; zero.y = 0;
                    V* cells = reinterpret_cast<V*>(plane + (size_t)row * Gp);
                    for (int e = lane; e < RPW * Gc; e += 32) cells[e] = zero;
                }
                dit_stageB<LZ, T, M, -1>(mine, tw.B, lane, 32);
            }
            else r2c_stageC_post<LZ, T, M>(mine, tw.C, tw.R, lane, 32, ToA{a_plane, row});   // last stage + real post-processing
        }
    };

    // ------------------------------------------------------------------ z inverse (c2r): real rows in place
    struct ZInv {
        T* plane;
        int row0;
        static constexpr bool kBulk = false;
        static constexpr bool kWarpPrivate = true;
        PM_HD TileLoad load(int) const { return TileLoad{plane + (size_t)row0 * Gp, 0, CZ * Gp * (int)sizeof(T)}; }
        static constexpr int kPhases = 3;
        PM_HD void phase(int ph, V* tile, const TW& tw, int tid, int, T (&)[kRegs]) const {
            const int warp = tid >> 5, lane = tid & 31, row = row0 + warp * RPW;
            V* mine = tile + warp * (RPW * M);
            if (ph == 0) c2r_pre_stageA<LZ, T, M>(RowSource{plane, row}, mine, tw.R, lane, 32);   // real pre-processing + first stage
            else if (ph == 1) dit_stageB<LZ, T, M, +1>(mine, tw.B, lane, 32);
            else dit_stageC<LZ, T, M, 2, +1>(mine, tw.C, lane, 32, RowSink{plane, row});   // z_m = x_2m + i·x_2m+1
        }
    };

    // ------------------------------------------------------------------ y forward: A block -> B
    struct YFwd {
        const V* a_tile;  // A block (plane, kt): V[G][CY]
        V* b;             // this rank's B
        int il, kt, nxl;
        static constexpr bool kBulk = true;
        static constexpr bool kWarpPrivate = false;
        PM_HD TileLoad load(int) const { return TileLoad{a_tile, 0, G * CY * (int)sizeof(V)}; }
        struct ToB {
            V* b; int il, kt, nxl;
            PM_HD void operator()(int c, int j, T r, T i) const {
                V v; v.x = r; v.y = i;
                b[b_index(kt, j, il, c, nxl)] = v;
            }
        };
        static constexpr int kPhases = 4;
        PM_HD void phase(int ph, V* tile, const TW& tw, int tid, int nthr, T (&rg)[kRegs]) const {
            T(&ra)[16 * kNbtY] = reinterpret_cast<T(&)[16 * kNbtY]>(rg[0]);
            if (ph == 0) dit_stageA_load<LY, T, G, NTHR, kNbtY>(tile, tid, ra);
            else if (ph == 1) dit_stageA_store<LY, T, G, -1, NTHR, kNbtY>(tile, tid, ra);
            else if (ph == 2) dit_stageB<LY, T, G, -1>(tile, tw.B, tid, nthr);
            else dit_stageC<LY, T, G, 1, -1>(tile, tw.C, tid, nthr, ToB{b, il, kt, nxl});
        }
    };

    // ------------------------------------------------------------------ y inverse: A block -> real rows
    struct YInv {
        const V* a_tile;
        T* plane;         // padded real plane, written through its complex view [G][Gc]
        int kt;
        static constexpr bool kBulk = true;
        static constexpr bool kWarpPrivate = false;
        PM_HD TileLoad load(int) const { return TileLoad{a_tile, 0, G * CY * (int)sizeof(V)}; }
        struct ToRows {
            T* plane; int kt;
            PM_HD void operator()(int c, int j, T r, T i) const {
                V v; v.x = r; v.y = i;
                reinterpret_cast<V*>(plane + (size_t)j * Gp)[kt * CY + c] = v;
            }
        };
        static constexpr int kPhases = 4;
        PM_HD void phase(int ph, V* tile, const TW& tw, int tid, int nthr, T (&rg)[kRegs]) const {
            T(&ra)[16 * kNbtY] = reinterpret_cast<T(&)[16 * kNbtY]>(rg[0]);
            if (ph == 0) dit_stageA_load<LY, T, G, NTHR, kNbtY>(tile, tid, ra);
            else if (ph == 1) dit_stageA_store<LY, T, G, +1, NTHR, kNbtY>(tile, tid, ra);
            else if (ph == 2) dit_stageB<LY, T, G, +1>(tile, tw.B, tid, nthr);
            else dit_stageC<LY, T, G, 1, +1>(tile, tw.C, tid, nthr, ToRows{plane, kt});
        }
    };

    // ------------------------------------------------------------------ x solve: B (all ranks) -> A (all ranks)
    struct XGeom {
        V* a[kMaxFftPeers];          // rank r's A
        const V* b[kMaxFftPeers];    // rank r's B
        int nranks;
        int in_place;                // results go back into the B piece they came from (contiguous per rank) instead of A
        int nxl_shift;               // planes per rank = 1 << nxl_shift
        const double* sep;           // separable factor per axis index l < G: (x_l/sin x_l)^D·exp(−gauss·k_l²)
        double prefactor;            // −L²·G_N/π
    };
    struct XSolve {
        const XGeom* g;
        int j;        // global j row
        int kt;       // column tile
        V* stage = nullptr;   // in place with a staging tile: the results are laid out like the loaded tile (rank-major, each
                              // rank's nxl·CY elements contiguous) and the caller sends every piece to its owner in one bulk copy
        static constexpr bool kBulk = true;
        static constexpr bool kWarpPrivate = false;
        PM_HD int nloads() const { return g->nranks; }
        PM_HD TileLoad load(int r) const {
            const int nxl = 1 << g->nxl_shift;
            return TileLoad{g->b[r] + b_index(kt, j, 0, 0, nxl), r * nxl * CY * (int)sizeof(V), nxl * CY * (int)sizeof(V)};
        }
        struct ToA {
            const XGeom* g; int j, kt;
            PM_HD void operator()(int c, int i, T r, T im) const {
                V v; v.x = r; v.y = im;
                const int rk = i >> g->nxl_shift, il = i & ((1 << g->nxl_shift) - 1);
                // in place: the nxl·64-byte piece of rank rk's B that this tile loaded — consecutive planes are consecutive
                // 64-byte segments, a few pages per peer.  (Straight into A, consecutive planes lie one plane of A apart:
                // at G = 1024 that is 8 MB, and 64-byte stores scattered like that over a peer's 8 GB ran at 5 GB/s.)
                if (g->in_place) const_cast<V*>(g->b[rk])[b_index(kt, j, il, c, 1 << g->nxl_shift)] = v;
                else g->a[rk][a_index(il, kt, j, c)] = v;
            }
        };
        struct ToStage {
            V* stage;
            PM_HD void operator()(int c, int i, T r, T im) const {
                V v; v.x = r; v.y = im;
                stage[(size_t)i * CY + c] = v;
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
        static constexpr int kPhases = 6;
        PM_HD void phase(int ph, V* tile, const TW& tw, int tid, int nthr, T (&rg)[kRegs]) const {
            constexpr int R1 = G / 64;
            T(&ra)[16 * kNbtY] = reinterpret_cast<T(&)[16 * kNbtY]>(rg[0]);
            if (ph == 0) dit_stageA_load<LY, T, G, NTHR, kNbtY>(tile, tid, ra);
            else if (ph == 1) dit_stageA_store<LY, T, G, -1, NTHR, kNbtY>(tile, tid, ra);
            else if (ph == 2) dit_stageB<LY, T, G, -1>(tile, tw.B, tid, nthr);
            else if (ph == 3) {
                // forward stage C, Green's function, inverse stage 1 — all on the same R1 registers
                const int kj = j - (j >= M ? G : 0);
                for (int b = tid; b < 64 * CY; b += nthr) {
                    int c, q;
                    LY::template decode<64>(b, c, q);
                    T r[R1], im[R1], wr[R1], wi[R1];
                    tw_powers<R1>(&tw.C[64 + q], wr, wi);     // ωG^(a1·q): forward stage C and, conjugated, inverse stage 1
#pragma unroll
                    for (int a1 = 0; a1 < R1; ++a1) {
                        const V v = tile[LY::idx_s64(q, a1, c)];
                        r[a1] = v.x; im[a1] = v.y;
                        if (a1) cmul2<-1>(r[a1], im[a1], wr[a1], wi[a1]);
                    }
                    dftR<R1, -1>(r, im);
                    const int kk = kt * CY + c;
                    const double sep_jk = g->sep[j] * g->sep[kk];
                    const int kj2_kk2 = kj * kj + kk * kk;
                    const bool line_nyq = (j == M) || (kk == M);
#pragma unroll
                    for (int b1 = 0; b1 < R1; ++b1) {
                        const T f = (T)factor(64 * b1 + q, sep_jk, kj2_kk2, line_nyq);
                        r[b1] *= f; im[b1] *= f;
                    }
                    dftR<R1, +1>(r, im);
#pragma unroll
                    for (int b1 = 0; b1 < R1; ++b1) {
                        if (b1) cmul2<+1>(r[b1], im[b1], wr[b1], wi[b1]);
                        V v; v.x = r[b1]; v.y = im[b1];
                        tile[LY::idx_s64(q, b1, c)] = v;
                    }
                }
            } else if (ph == 4) dif_stage2<LY, T, G, +1>(tile, tw.B, tid, nthr);
            else if (g->in_place && stage != nullptr) dif_stage3<LY, T, G, +1>(tile, tid, nthr, ToStage{stage});
            else dif_stage3<LY, T, G, +1>(tile, tid, nthr, ToA{g, j, kt});
        }
    };
};

}  // namespace fftc
}  // namespace pm
