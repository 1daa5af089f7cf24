// pm_fftcore.cuh — in-shared-memory FFT building blocks for the hand-written slab transform
// (pm_fft.cu).  Everything here is __host__ __device__ and takes (tid, nthr) explicitly, so that the
// very same code can be stepped through sequentially on the CPU (tests/fft_host_harness.cu) — stages
// are separated by __syncthreads() on the device and touch thread-private positions in between.
//
// Transform length N = R1·8·8 with R1 ∈ {1, 2, 4, 8}.  Two flows over a tile of C independent lines:
//
//   DIT  (A, B, C):  input element a = a1 + R1·a2 + 8·R1·a3 at position 64·a1 + 8·a2 + a3
//                    (stage A gathers it straight from global memory through a Source functor),
//                    output natural.
//   DIF  (1, 2, 3):  input natural, output element b = b1 + R1·b2 + 8·R1·b3 at position
//                    64·b1 + 8·b2 + b3.
//
// so that   natural --DIT--> natural-order spectrum --(pointwise factor)--> DIF --> scattered store
// needs no reordering pass, and DIT stage C / DIF stage 1 touch the same positions (the x-solve keeps
// them in registers across the Green's-function multiply).
//
// Tile layouts (template policy L).  nat(p, c) is where the bulk copy from global memory leaves element
// p of line c; idx(p, c) is where the in-place stages keep it (same footprint, swizzled so that the
// stride-1, stride-8 and stride-64 accesses of the three stages are bank-conflict free):
//   ColSwz<C>         [position][C] with C·sizeof(complex) = 64 B; lanes run over c first.  Swizzle:
//                     position p sits in row p ^ bit3(p) ^ bit6(p).  Used for the y and x passes.
//   RowSwz<M,C,RAW>   [line][M]; lanes run over positions.  nat has row pitch RAW (the padded global
//                     rows), idx pitch M with the low three position bits XORed with bits 3-5 and 6-7.
//                     Used for the z pass.
//   ColLayout / RowLayout: unswizzled reference layouts (CPU tests).
#pragma once

#include <cuda_runtime.h>

#ifndef PM_HD
#ifdef __CUDACC__
#define PM_HD __host__ __device__ __forceinline__
#else
#define PM_HD inline
#endif
#endif

namespace pm {
namespace fftc {

template <typename T> struct Vec2;
template <> struct Vec2<double> { using type = double2; };
template <> struct Vec2<float> { using type = float2; };

PM_HD double fma_(double a, double b, double c) { return fma(a, b, c); }
PM_HD float fma_(float a, float b, float c) { return fmaf(a, b, c); }

// ---------------------------------------------------------------------------------------------
// small DFTs in registers, natural order in and out.  DIR = −1: e^{−2πi nk/R};  +1: e^{+2πi nk/R}
// ---------------------------------------------------------------------------------------------
template <int DIR, typename T>
PM_HD void dft2(T (&r)[2], T (&i)[2]) {
    const T ar = r[0] + r[1], ai = i[0] + i[1];
    r[1] = r[0] - r[1]; i[1] = i[0] - i[1];
    r[0] = ar; i[0] = ai;
}

template <int DIR, typename T>
PM_HD void dft4(T (&r)[4], T (&i)[4]) {
    const T s0r = r[0] + r[2], s0i = i[0] + i[2], d0r = r[0] - r[2], d0i = i[0] - i[2];
    const T s1r = r[1] + r[3], s1i = i[1] + i[3], d1r = r[1] - r[3], d1i = i[1] - i[3];
    // forward: −i·d1 = (d1i, −d1r);  inverse: +i·d1 = (−d1i, d1r)
    const T tr = DIR < 0 ? d1i : -d1i;
    const T ti = DIR < 0 ? -d1r : d1r;
    r[0] = s0r + s1r; i[0] = s0i + s1i;
    r[2] = s0r - s1r; i[2] = s0i - s1i;
    r[1] = d0r + tr;  i[1] = d0i + ti;
    r[3] = d0r - tr;  i[3] = d0i - ti;
}

template <int DIR, typename T>
PM_HD void dft8(T (&r)[8], T (&i)[8]) {
    const T h = (T)0.70710678118654752440;
    T a0r = r[0] + r[4], a0i = i[0] + i[4], a4r = r[0] - r[4], a4i = i[0] - i[4];
    T a1r = r[1] + r[5], a1i = i[1] + i[5], a5r = r[1] - r[5], a5i = i[1] - i[5];
    T a2r = r[2] + r[6], a2i = i[2] + i[6], a6r = r[2] - r[6], a6i = i[2] - i[6];
    T a3r = r[3] + r[7], a3i = i[3] + i[7], a7r = r[3] - r[7], a7i = i[3] - i[7];
    T t;
    if (DIR < 0) {
        t = (a5r + a5i) * h; a5i = (a5i - a5r) * h; a5r = t;        // ·(1−i)/√2
        t = a6i; a6i = -a6r; a6r = t;                               // ·(−i)
        t = (a7i - a7r) * h; a7i = (-a7r - a7i) * h; a7r = t;       // ·(−1−i)/√2
    } else {
        t = (a5r - a5i) * h; a5i = (a5r + a5i) * h; a5r = t;        // ·(1+i)/√2
        t = -a6i; a6i = a6r; a6r = t;                               // ·(+i)
        t = (-a7r - a7i) * h; a7i = (a7r - a7i) * h; a7r = t;       // ·(−1+i)/√2
    }
    T b0r = a0r + a2r, b0i = a0i + a2i, b2r = a0r - a2r, b2i = a0i - a2i;
    T b1r = a1r + a3r, b1i = a1i + a3i, b3r = a1r - a3r, b3i = a1i - a3i;
    T b4r = a4r + a6r, b4i = a4i + a6i, b6r = a4r - a6r, b6i = a4i - a6i;
    T b5r = a5r + a7r, b5i = a5i + a7i, b7r = a5r - a7r, b7i = a5i - a7i;
    if (DIR < 0) {
        t = b3i; b3i = -b3r; b3r = t;
        t = b7i; b7i = -b7r; b7r = t;
    } else {
        t = -b3i; b3i = b3r; b3r = t;
        t = -b7i; b7i = b7r; b7r = t;
    }
    r[0] = b0r + b1r; i[0] = b0i + b1i; r[4] = b0r - b1r; i[4] = b0i - b1i;
    r[2] = b2r + b3r; i[2] = b2i + b3i; r[6] = b2r - b3r; i[6] = b2i - b3i;
    r[1] = b4r + b5r; i[1] = b4i + b5i; r[5] = b4r - b5r; i[5] = b4i - b5i;
    r[3] = b6r + b7r; i[3] = b6i + b7i; r[7] = b6r - b7r; i[7] = b6i - b7i;
}

// 16 points as 4 × 4: n = 4·n1 + n2, k = k1 + 4·k2; 4-point transforms over n1, twiddles ω16^(n2·k1), 4-point transforms over n2
template <int DIR, typename T>
PM_HD void dft16(T (&r)[16], T (&i)[16]) {
    const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173;   // cos, sin of π/8
    const T h = (T)0.70710678118654752440;
    T yr[4][4], yi[4][4];      // [n2][k1]
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
        T ar[4], ai[4];
#pragma unroll
        for (int n1 = 0; n1 < 4; ++n1) { ar[n1] = r[4 * n1 + n2]; ai[n1] = i[4 * n1 + n2]; }
        dft4<DIR>(ar, ai);
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) { yr[n2][k1] = ar[k1]; yi[n2][k1] = ai[k1]; }
    }
    // ω16^m = (cos(mπ/8), ∓sin(mπ/8)) for DIR = ∓1; m = n2·k1
    auto tw = [&](T& xr, T& xi, T wr, T ws) {      // multiply by (wr, DIR < 0 ? −ws : +ws)
        const T wi = DIR < 0 ? -ws : ws;
        const T t = fma_(xr, wr, -(xi * wi));
        xi = fma_(xr, wi, xi * wr);
        xr = t;
    };
    tw(yr[1][1], yi[1][1], c1, s1);     // m = 1
    tw(yr[1][2], yi[1][2], h, h);       // 2
    tw(yr[1][3], yi[1][3], s1, c1);     // 3
    tw(yr[2][1], yi[2][1], h, h);       // 2
    tw(yr[2][2], yi[2][2], (T)0, (T)1); // 4
    tw(yr[2][3], yi[2][3], -h, h);      // 6
    tw(yr[3][1], yi[3][1], s1, c1);     // 3
    tw(yr[3][2], yi[3][2], -h, h);      // 6
    tw(yr[3][3], yi[3][3], -c1, -s1);   // 9: cos(9π/8) = −cos(π/8), sin(9π/8) = −sin(π/8)
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        T ar[4], ai[4];
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) { ar[n2] = yr[n2][k1]; ai[n2] = yi[n2][k1]; }
        dft4<DIR>(ar, ai);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) { r[k1 + 4 * k2] = ar[k2]; i[k1 + 4 * k2] = ai[k2]; }
    }
}

template <int R, int DIR, typename T>
PM_HD void dftR(T (&r)[R], T (&i)[R]) {
    if constexpr (R == 2) dft2<DIR>(r, i);
    else if constexpr (R == 4) dft4<DIR>(r, i);
    else if constexpr (R == 8) dft8<DIR>(r, i);
    else if constexpr (R == 16) dft16<DIR>(r, i);
    // R == 1: identity
}

// (r, i) ·= w (DIR < 0) or ·= conj(w) (DIR > 0); the table holds w = e^{−2πi m/NT}
template <int DIR, typename T, typename V>
PM_HD void cmul(T& r, T& i, const V w) {
    const T wi = DIR < 0 ? w.y : -w.y;
    const T t = fma_(r, w.x, -(i * wi));
    i = fma_(r, wi, i * w.x);
    r = t;
}

// Powers w⁰ … w^(R−1) of one table entry w = e^{−2πi·m/NT}.  The twiddles of a butterfly are the powers of ONE root
// (ω64^(a·b3), ωN^(a·q) for a = 0 … R−1), so a thread loads that root once and multiplies — the kernels are bound by
// the shared-memory/L1 data pipe (ncu r01: 42 % of all LDS were twiddle loads) and have FP64 issue slots to spare.
// Squarings for even powers, one product for odd ones: at most ⌈log₂R⌉ roundings deep (≈ 3 ulp for R = 8).
template <int R, typename T, typename V>
PM_HD void tw_powers(const V* wp, T (&pr)[R], T (&pi)[R]) {
    pr[0] = (T)1; pi[0] = (T)0;
    if constexpr (R == 1) return;      // (the table entry may not even exist then)
    const V w = *wp;
    if constexpr (R > 1) { pr[1] = w.x; pi[1] = w.y; }
#ifdef PM_TW_SERIAL
#pragma unroll
    for (int m = 2; m < R; ++m) {      // w^m = w^(m−1)·w: one product each, R − 2 roundings deep
        const T a = pr[m - 1], b = pi[m - 1];
        pr[m] = fma_(a, w.x, -(b * w.y));
        pi[m] = fma_(a, w.y, b * w.x);
    }
    return;
#endif
#pragma unroll
    for (int m = 2; m < R; ++m) {
        if ((m & 1) == 0) {
            const T a = pr[m / 2], b = pi[m / 2];
            pr[m] = (a - b) * (a + b);
            pi[m] = (a + a) * b;
        } else {
            const T a = pr[m - 1], b = pi[m - 1];
            pr[m] = fma_(a, w.x, -(b * w.y));
            pi[m] = fma_(a, w.y, b * w.x);
        }
    }
}
// (r, i) ·= (wr, wi) (DIR < 0) or ·= conj (DIR > 0)
template <int DIR, typename T>
PM_HD void cmul2(T& r, T& i, const T wr, const T wi_in) {
    const T wi = DIR < 0 ? wi_in : -wi_in;
    const T t = fma_(r, wr, -(i * wi));
    i = fma_(r, wi, i * wr);
    r = t;
}

// ---------------------------------------------------------------------------------------------
// layouts
// ---------------------------------------------------------------------------------------------
template <int C_>
struct ColLayout {
    static constexpr int C = C_;
    static PM_HD int idx(int p, int c) { return p * C + c; }
    // strided accessors used by the stages (k is a compile-time constant after unrolling):
    //   idx_s8 (p0, k, c) = idx(p0 + 8·k, c)   with p0 = 64·x + y, y < 8
    //   idx_s64(q,  k, c) = idx(64·k + q, c)   with q < 64
    //   idx_s1 (p0, k, c) = idx(p0 + k, c)     with p0 a multiple of 8, k < 8
    static PM_HD int idx_s8(int p0, int k, int c) { return idx(p0 + 8 * k, c); }
    static PM_HD int idx_s64(int q, int k, int c) { return idx(64 * k + q, c); }
    static PM_HD int idx_s1(int p0, int k, int c) { return idx(p0 + k, c); }
    // butterfly b of a stage with P butterflies per line  ->  (line, butterfly-in-line)
    template <int P> static PM_HD void decode(int b, int& c, int& u) { c = b % C; u = b / C; }
};

template <int N_, int C_>
struct RowLayout {
    static constexpr int C = C_;
    static constexpr int PITCH = N_ + N_ / 8 + N_ / 64 + 1;
    static PM_HD int idx(int p, int c) { return c * PITCH + p + (p >> 3) + (p >> 6); }
    // strided accessors used by the stages (k is a compile-time constant after unrolling):
    //   idx_s8 (p0, k, c) = idx(p0 + 8·k, c)   with p0 = 64·x + y, y < 8
    //   idx_s64(q,  k, c) = idx(64·k + q, c)   with q < 64
    //   idx_s1 (p0, k, c) = idx(p0 + k, c)     with p0 a multiple of 8, k < 8
    static PM_HD int idx_s8(int p0, int k, int c) { return idx(p0 + 8 * k, c); }
    static PM_HD int idx_s64(int q, int k, int c) { return idx(64 * k + q, c); }
    static PM_HD int idx_s1(int p0, int k, int c) { return idx(p0 + k, c); }
    template <int P> static PM_HD void decode(int b, int& c, int& u) { u = b % P; c = b / P; }
};

template <int C_>
struct ColSwz {
    static constexpr int C = C_;
    static PM_HD int swz(int p) { return p ^ ((p >> 3) & 1) ^ ((p >> 6) & 1); }
    static PM_HD int idx(int p, int c) { return swz(p) * C + c; }
    // The strided accessors (see ColLayout) written so that, with k a compile-time constant, every access is
    // one of TWO per-thread base addresses plus an immediate offset — the plain idx() costs two to three
    // integer instructions and an address register per access, and these kernels are issue-bound.
    //   p0 + 8k: bit 3 = k & 1 (p0's own bit 3 is clear), bit 6 = bit 6 of p0
    static PM_HD int idx_s8(int p0, int k, int c) { return ((p0 ^ (((p0 >> 6) ^ k) & 1)) + 8 * k) * C + c; }
    //   64k + q: bit 3 = bit 3 of q, bit 6 = k & 1
    static PM_HD int idx_s64(int q, int k, int c) { return ((q ^ (((q >> 3) ^ k) & 1)) + 64 * k) * C + c; }
    //   p0 + k: the swizzle bit s is a property of p0; k ^ s = k + s for even k, k − s for odd k
    static PM_HD int idx_s1(int p0, int k, int c) {
        const int s = ((p0 >> 3) ^ (p0 >> 6)) & 1;
        return (p0 + k + ((k & 1) ? -s : s)) * C + c;
    }
    static PM_HD int nat(int p, int c) { return p * C + c; }
    template <int P> static PM_HD void decode(int b, int& c, int& u) { c = b % C; u = b / C; }
};

template <int M_, int C_, int RAWPITCH_>
struct RowSwz {
    static constexpr int C = C_;
    static PM_HD int idx(int p, int c) { return c * M_ + ((p & ~7) | ((p ^ (p >> 3) ^ ((p >> 6) << 1)) & 7)); }
    // strided accessors used by the stages (k is a compile-time constant after unrolling):
    //   idx_s8 (p0, k, c) = idx(p0 + 8·k, c)   with p0 = 64·x + y, y < 8
    //   idx_s64(q,  k, c) = idx(64·k + q, c)   with q < 64
    //   idx_s1 (p0, k, c) = idx(p0 + k, c)     with p0 a multiple of 8, k < 8
    static PM_HD int idx_s8(int p0, int k, int c) { return idx(p0 + 8 * k, c); }
    static PM_HD int idx_s64(int q, int k, int c) { return idx(64 * k + q, c); }
    static PM_HD int idx_s1(int p0, int k, int c) { return idx(p0 + k, c); }
    static PM_HD int nat(int p, int c) { return c * RAWPITCH_ + p; }
    template <int P> static PM_HD void decode(int b, int& c, int& u) { u = b % P; c = b / P; }
};

// Twiddle tables of one grid size G (all e^{−2πi·…}), laid out so that the lanes of a warp read
// consecutive entries (row layout) or one broadcast entry per quarter-warp (column layout):
//   B[a·8 + b]   = ω64^(a·b)            a, b < 8        middle stages
//   C[a·64 + q]  = ωG^(a·q)             a < G/64, q < 64   first/last stage of a G-point transform; an
//                                                       M = G/2-point transform uses rows 2a (ωM = ωG²)
//   R[k]         = ωG^k                 k ≤ G/8         real<->complex packing; k up to G/4 by symmetry
template <typename V>
struct Twiddles {
    const V* B;
    const V* C;
    const V* R;
};
template <int G> constexpr int twiddle_entries() { return 64 + G + G / 8 + 1; }

// ω_{2M}^k for 0 ≤ k ≤ M/2 from R (entries k ≤ M/4):  ω^k = −i·conj(ω^(M/2−k))
template <int M, typename V>
PM_HD V tw_r(const V* R, int k) {
    if (k <= M / 4) return R[k];
    const V w = R[M / 2 - k];
    V o; o.x = -w.y; o.y = -w.x;
    return o;
}

// ---------------------------------------------------------------------------------------------
// DIT flow.  TWS = G/N: row stride into the C table (1 for a G-point transform, 2 for G/2 points).
// ---------------------------------------------------------------------------------------------
// Stage A: gathers element a = u + (N/8)·a3 (u = a1 + R1·a2) of line c through src(c, a),
// 8-point DFT over a3 -> b3, result at position 64·a1 + 8·a2 + b3 of `dst`.
template <class L, typename T, int N, int DIR, typename V, class Source>
PM_HD void dit_stageA(const Source& src, V* dst, int tid, int nthr) {
    constexpr int R1 = N / 64;
    constexpr int P = N / 8;
    for (int b = tid; b < P * L::C; b += nthr) {
        int c, u;
        L::template decode<P>(b, c, u);
        const int a1 = u % R1, a2 = u / R1;
        T r[8], i[8];
#pragma unroll
        for (int a3 = 0; a3 < 8; ++a3) {
            const V v = src(c, u + P * a3);
            r[a3] = v.x; i[a3] = v.y;
        }
        dft8<DIR>(r, i);
        const int p0 = 64 * a1 + 8 * a2;
#pragma unroll
        for (int b3 = 0; b3 < 8; ++b3) {
            V v; v.x = r[b3]; v.y = i[b3];
            dst[L::idx_s1(p0, b3, c)] = v;
        }
    }
}

// Stage A in place: the inputs already sit at positions 64·a1 + 8·a2 + a3 (c2r_pre puts them there).
template <class L, typename T, int N, int DIR, typename V>
PM_HD void dit_stageA_inplace(V* tile, int tid, int nthr) {
    constexpr int P = N / 8;
    for (int b = tid; b < P * L::C; b += nthr) {
        int c, u;
        L::template decode<P>(b, c, u);   // u = a2 + 8·a1: positions 8·u + a3
        T r[8], i[8];
#pragma unroll
        for (int a3 = 0; a3 < 8; ++a3) {
            const V v = tile[L::idx_s1(8 * u, a3, c)];
            r[a3] = v.x; i[a3] = v.y;
        }
        dft8<DIR>(r, i);
#pragma unroll
        for (int b3 = 0; b3 < 8; ++b3) {
            V v; v.x = r[b3]; v.y = i[b3];
            tile[L::idx_s1(8 * u, b3, c)] = v;
        }
    }
}

// Stage B: positions 64·a1 + 8·a2 + b3 over a2; input a2 times ω64^(a2·b3); DFT8 -> b2, in place.
template <class L, typename T, int N, int DIR, typename V>
PM_HD void dit_stageB(V* tile, const V* twB, int tid, int nthr) {
    constexpr int P = N / 8;
    for (int b = tid; b < P * L::C; b += nthr) {
        int c, u;
        L::template decode<P>(b, c, u);
        const int b3 = u & 7, a1 = u >> 3;
        const int p0 = 64 * a1 + b3;
        T r[8], i[8], wr[8], wi[8];
        tw_powers<8>(&twB[8 + b3], wr, wi);     // ω64^(a2·b3), a2 = 0 … 7
#pragma unroll
        for (int a2 = 0; a2 < 8; ++a2) {
            const V v = tile[L::idx_s8(p0, a2, c)];
            r[a2] = v.x; i[a2] = v.y;
            if (a2) cmul2<DIR>(r[a2], i[a2], wr[a2], wi[a2]);
        }
        dft8<DIR>(r, i);
#pragma unroll
        for (int b2 = 0; b2 < 8; ++b2) {
            V v; v.x = r[b2]; v.y = i[b2];
            tile[L::idx_s8(p0, b2, c)] = v;
        }
    }
}

// Stage C: positions 64·a1 + q over a1 (q = 8·b2 + b3); input a1 times ωN^(a1·q); DFT_R1 -> b1;
// result (natural index 64·b1 + q) handed to sink(c, index, re, im).
template <class L, typename T, int N, int TWS, int DIR, typename V, class Sink>
PM_HD void dit_stageC(const V* tile, const V* twC, int tid, int nthr, const Sink& sink) {
    constexpr int R1 = N / 64;
    for (int b = tid; b < 64 * L::C; b += nthr) {
        int c, q;
        L::template decode<64>(b, c, q);
        T r[R1], i[R1], wr[R1], wi[R1];
        tw_powers<R1>(&twC[TWS * 64 + q], wr, wi);     // ωN^(a1·q)
#pragma unroll
        for (int a1 = 0; a1 < R1; ++a1) {
            const V v = tile[L::idx_s64(q, a1, c)];
            r[a1] = v.x; i[a1] = v.y;
            if (a1) cmul2<DIR>(r[a1], i[a1], wr[a1], wi[a1]);
        }
        dftR<R1, DIR>(r, i);
#pragma unroll
        for (int b1 = 0; b1 < R1; ++b1) sink(c, 64 * b1 + q, r[b1], i[b1]);
    }
}

// ---------------------------------------------------------------------------------------------
// DIF flow
// ---------------------------------------------------------------------------------------------
// Stage 1 on values already in registers (natural index 64·a1 + q): DFT_R1 -> b1, times ωN^(b1·q),
// written to position 64·b1 + q.
template <class L, typename T, int N, int TWS, int DIR, typename V>
PM_HD void dif_stage1_regs(T (&r)[N / 64], T (&i)[N / 64], V* tile, const V* twC, int c, int q) {
    constexpr int R1 = N / 64;
    dftR<R1, DIR>(r, i);
    T wr[R1], wi[R1];
    tw_powers<R1>(&twC[TWS * 64 + q], wr, wi);
#pragma unroll
    for (int b1 = 0; b1 < R1; ++b1) {
        if (b1) cmul2<DIR>(r[b1], i[b1], wr[b1], wi[b1]);
        V v; v.x = r[b1]; v.y = i[b1];
        tile[L::idx_s64(q, b1, c)] = v;
    }
}

// Stage 2: positions 64·b1 + 8·a2 + a3 over a2: DFT8 -> b2, times ω64^(a3·b2), in place.
template <class L, typename T, int N, int DIR, typename V>
PM_HD void dif_stage2(V* tile, const V* twB, int tid, int nthr) {
    constexpr int P = N / 8;
    for (int b = tid; b < P * L::C; b += nthr) {
        int c, u;
        L::template decode<P>(b, c, u);
        const int a3 = u & 7, b1 = u >> 3;
        const int p0 = 64 * b1 + a3;
        T r[8], i[8];
#pragma unroll
        for (int a2 = 0; a2 < 8; ++a2) {
            const V v = tile[L::idx_s8(p0, a2, c)];
            r[a2] = v.x; i[a2] = v.y;
        }
        dft8<DIR>(r, i);
        T wr[8], wi[8];
        tw_powers<8>(&twB[8 + a3], wr, wi);     // ω64^(a3·b2)
#pragma unroll
        for (int b2 = 0; b2 < 8; ++b2) {
            if (b2) cmul2<DIR>(r[b2], i[b2], wr[b2], wi[b2]);
            V v; v.x = r[b2]; v.y = i[b2];
            tile[L::idx_s8(p0, b2, c)] = v;
        }
    }
}

// Stage 3: positions 8·u + a3 (u = b2 + 8·b1) over a3: DFT8 -> b3; element b1 + R1·b2 + 8·R1·b3
// handed to sink(c, element, re, im).
template <class L, typename T, int N, int DIR, typename V, class Sink>
PM_HD void dif_stage3(const V* tile, int tid, int nthr, const Sink& sink) {
    constexpr int R1 = N / 64;
    constexpr int P = N / 8;
    for (int b = tid; b < P * L::C; b += nthr) {
        int c, u;
        L::template decode<P>(b, c, u);
        const int b2 = u & 7, b1 = u >> 3;
        T r[8], i[8];
#pragma unroll
        for (int a3 = 0; a3 < 8; ++a3) {
            const V v = tile[L::idx_s1(8 * u, a3, c)];
            r[a3] = v.x; i[a3] = v.y;
        }
        dft8<DIR>(r, i);
        const int e0 = b1 + R1 * b2;
#pragma unroll
        for (int b3 = 0; b3 < 8; ++b3) sink(c, e0 + 8 * R1 * b3, r[b3], i[b3]);
    }
}

// ---------------------------------------------------------------------------------------------
// real <-> half-length complex packing (z pass).  M complex points per line, real length 2M = G.
// ---------------------------------------------------------------------------------------------
// After the forward M-point transform of z_n = x_2n + i·x_2n+1 (natural order in `tile`):
//   X_k = E + ω^k·O,   E = (Z_k + conj Z_{M−k})/2,   O = −i·(Z_k − conj Z_{M−k})/2,   ω = e^{−2πi/2M}
//   X_{M−k} = conj(E) − conj(ω^k·O)
// for k = 0 … M/2; sink(c, k, re, im) receives k = 0 … M−1 (X_M — the Nyquist mode, which the potential
// nullifies, mesh.py:3615-3622 — is delivered as zero at k = M).
template <class L, typename T, int M, typename V, class Sink>
PM_HD void r2c_post(const V* tile, const V* twR, int tid, int nthr, const Sink& sink) {
    constexpr int P = M / 2 + 1;
    for (int b = tid; b < P * L::C; b += nthr) {
        int c, k;
        L::template decode<P>(b, c, k);
        const int kp = (M - k) & (M - 1);
        const V zk = tile[L::idx(k, c)];
        const V zp = tile[L::idx(kp, c)];
        const T er = (T)0.5 * (zk.x + zp.x), ei = (T)0.5 * (zk.y - zp.y);
        // D = Z_k − conj Z_p = (zk.x − zp.x, zk.y + zp.y);  O = −i·D/2 = (D.y/2, −D.x/2)
        T orr = (T)0.5 * (zk.y + zp.y), oi = (T)-0.5 * (zk.x - zp.x);
        cmul<-1>(orr, oi, tw_r<M>(twR, k));
        sink(c, k, er + orr, ei + oi);
        if (k == 0) sink(c, M, (T)0, (T)0);
        else if (k != kp) sink(c, kp, er - orr, -ei + oi);
    }
}

// Before the inverse M-point transform:  Z_k = A + i·U,  Z_{M−k} = conj(A) + i·conj(U),
//   A = X_k + conj X_{M−k},  U = conj(ω^k)·(X_k − conj X_{M−k}),  with X_M := 0 and Im X_0 ignored.
// src(c, k) delivers X_k (k < M) from global memory; Z_k goes to its DIT input position of `tile`.
template <class L, typename T, int M, typename V, class Source>
PM_HD void c2r_pre(const Source& src, V* tile, const V* twR, int tid, int nthr) {
    constexpr int R1 = M / 64;
    constexpr int P = M / 2 + 1;
    for (int b = tid; b < P * L::C; b += nthr) {
        int c, k;
        L::template decode<P>(b, c, k);
        const int kp = (M - k) & (M - 1);
        V zk, zp;
        if (k == 0) {
            const T x0 = src(c, 0).x;
            zk.x = x0; zk.y = x0;
            zp = zk;
        } else {
            const V xk = src(c, k);
            const V xp = src(c, kp);
            const T ar = xk.x + xp.x, ai = xk.y - xp.y;
            T ur = xk.x - xp.x, ui = xk.y + xp.y;
            cmul<+1>(ur, ui, tw_r<M>(twR, k));
            zk.x = ar - ui; zk.y = ai + ur;      // A + i·U
            zp.x = ar + ui; zp.y = -ai + ur;     // conj(A) + i·conj(U)
        }
        // DIT input position of element a: a = a1 + R1·a2 + 8·R1·a3  ->  64·a1 + 8·a2 + a3
        {
            const int a1 = k % R1, a2 = (k / R1) & 7, a3 = k / (8 * R1);
            tile[L::idx(64 * a1 + 8 * a2 + a3, c)] = zk;
        }
        if (k != kp) {
            const int a1 = kp % R1, a2 = (kp / R1) & 7, a3 = kp / (8 * R1);
            tile[L::idx(64 * a1 + 8 * a2 + a3, c)] = zp;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// the real-transform pre/post-processing fused into the neighbouring stage of the z pass (no extra trip of the tile
// through shared memory, one block barrier less per tile)
// ---------------------------------------------------------------------------------------------
// One mirror pair of r2c_post: Z_k, Z_{M−k} -> X_k, X_{M−k}; needs k ≤ M/2, kp = (M − k) mod M.
template <typename T, int M, typename V, class Sink>
PM_HD void r2c_pair(const V* twR, const Sink& sink, int c, int k, T zkr, T zki, int kp, T zpr, T zpi) {
    const T er = (T)0.5 * (zkr + zpr), ei = (T)0.5 * (zki - zpi);
    T orr = (T)0.5 * (zki + zpi), oi = (T)-0.5 * (zkr - zpr);
    cmul<-1>(orr, oi, tw_r<M>(twR, k));
    sink(c, k, er + orr, ei + oi);
    if (k == 0) sink(c, M, (T)0, (T)0);
    else if (k != kp) sink(c, kp, er - orr, -ei + oi);
}
template <typename T, int M, typename V, class Sink>
PM_HD void r2c_pair_any(const V* twR, const Sink& sink, int c, int k, T zkr, T zki, int kp, T zpr, T zpi) {
    if (k <= kp) r2c_pair<T, M>(twR, sink, c, k, zkr, zki, kp, zpr, zpi);
    else r2c_pair<T, M>(twR, sink, c, kp, zpr, zpi, k, zkr, zki);
}

// Forward: DIT stage C of the M-point transform + r2c_post.  One item = row c and the column pair (qa, qb) = (m, 64 − m)
// for m = 1 … 31, (0, 32) for m = 0: its outputs Z[64·b1 + qa], Z[64·b1 + qb] (b1 < R1) hold both members of every mirror
// pair (k, M − k) they belong to, because M − (64·b1 + qa) = 64·(R1 − 1 − b1) + (64 − qa).
template <class L, typename T, int M, typename V, class Sink>
PM_HD void r2c_stageC_post(const V* tile, const V* twC, const V* twR, int tid, int nthr, const Sink& sink) {
    constexpr int R1 = M / 64;
    for (int b = tid; b < 32 * L::C; b += nthr) {
        const int m = b & 31, c = b >> 5;
        const int qa = m, qb = m ? 64 - m : 32;
        T ar[R1], ai[R1], br[R1], bi[R1];
        {
            T wr[R1], wi[R1];
            tw_powers<R1>(&twC[2 * 64 + qa], wr, wi);     // ωM^(a1·qa) = ωG^(2·a1·qa)
#pragma unroll
            for (int a1 = 0; a1 < R1; ++a1) {
                const V v = tile[L::idx_s64(qa, a1, c)];
                ar[a1] = v.x; ai[a1] = v.y;
                if (a1) cmul2<-1>(ar[a1], ai[a1], wr[a1], wi[a1]);
            }
            dftR<R1, -1>(ar, ai);
            tw_powers<R1>(&twC[2 * 64 + qb], wr, wi);
#pragma unroll
            for (int a1 = 0; a1 < R1; ++a1) {
                const V v = tile[L::idx_s64(qb, a1, c)];
                br[a1] = v.x; bi[a1] = v.y;
                if (a1) cmul2<-1>(br[a1], bi[a1], wr[a1], wi[a1]);
            }
            dftR<R1, -1>(br, bi);
        }
        if (m != 0) {
#pragma unroll
            for (int b1 = 0; b1 < R1; ++b1)
                r2c_pair_any<T, M>(twR, sink, c, 64 * b1 + qa, ar[b1], ai[b1], 64 * (R1 - 1 - b1) + qb, br[R1 - 1 - b1], bi[R1 - 1 - b1]);
        } else {
            // qa = 0: k = 64·b1 pairs with 64·(R1 − b1) (k = 0 and k = M/2 with themselves)
#pragma unroll
            for (int b1 = 0; b1 <= R1 / 2; ++b1) {
                const int bp = (R1 - b1) % R1;
                r2c_pair<T, M>(twR, sink, c, 64 * b1, ar[b1], ai[b1], 64 * bp, ar[bp], ai[bp]);
            }
            // qb = 32: k = 64·b1 + 32 pairs with 64·(R1 − 1 − b1) + 32
#pragma unroll
            for (int b1 = 0; b1 <= (R1 - 1) / 2; ++b1) {
                const int bp = R1 - 1 - b1;
                r2c_pair<T, M>(twR, sink, c, 64 * b1 + 32, br[b1], bi[b1], 64 * bp + 32, br[bp], bi[bp]);
            }
        }
    }
}

// One mirror pair of c2r_pre: X_k, X_{M−k} -> Z_k, Z_{M−k}; needs 0 < k ≤ M/2.
template <typename T, int M, typename V>
PM_HD void c2r_pair(const V* twR, int k, const V xk, const V xp, T& zkr, T& zki, T& zpr, T& zpi) {
    const T ar = xk.x + xp.x, ai = xk.y - xp.y;
    T ur = xk.x - xp.x, ui = xk.y + xp.y;
    cmul<+1>(ur, ui, tw_r<M>(twR, k));
    zkr = ar - ui; zki = ai + ur;      // A + i·U
    zpr = ar + ui; zpi = -ai + ur;     // conj(A) + i·conj(U)
}

// Inverse: c2r_pre + DIT stage A of the M-point transform.  One item = row c and the butterfly pair (ua, ub) = (m, P − m)
// for m = 1 … P/2 − 1, (0, P/2) for m = 0 (P = M/8 butterflies per row): element a = u + P·a3 of butterfly u needs X_a and
// X_{M−a}, and M − a = (P − u) + P·(7 − a3) is element 7 − a3 of butterfly P − u.  src(c, k) delivers X_k (k < M).
template <class L, typename T, int M, typename V, class Source>
PM_HD void c2r_pre_stageA(const Source& src, V* tile, const V* twR, int tid, int nthr) {
    constexpr int R1 = M / 64, P = M / 8, H = P / 2;
    for (int b = tid; b < H * L::C; b += nthr) {
        const int m = b % H, c = b / H;
        const int ua = m, ub = m ? P - m : H;
        V xa[8], xb[8];
#pragma unroll
        for (int a3 = 0; a3 < 8; ++a3) { xa[a3] = src(c, ua + P * a3); xb[a3] = src(c, ub + P * a3); }
        T zar[8], zai[8], zbr[8], zbi[8];
        if (m != 0) {
#pragma unroll
            for (int a3 = 0; a3 < 8; ++a3) {
                const int k = ua + P * a3;            // 0 < k < M, mirror M − k = ub + P·(7 − a3)
                if (k <= M / 2) c2r_pair<T, M>(twR, k, xa[a3], xb[7 - a3], zar[a3], zai[a3], zbr[7 - a3], zbi[7 - a3]);
                else c2r_pair<T, M>(twR, M - k, xb[7 - a3], xa[a3], zbr[7 - a3], zbi[7 - a3], zar[a3], zai[a3]);
            }
        } else {
            // ua = 0: k = P·a3; k = 0 alone (X_M := 0, Im X_0 ignored), k = M/2 with itself, a3 = 1 … 3 with 8 − a3
            zar[0] = xa[0].x; zai[0] = xa[0].x;
            T dr, di;
#pragma unroll
            for (int a3 = 1; a3 < 4; ++a3) c2r_pair<T, M>(twR, P * a3, xa[a3], xa[8 - a3], zar[a3], zai[a3], zar[8 - a3], zai[8 - a3]);
            c2r_pair<T, M>(twR, M / 2, xa[4], xa[4], zar[4], zai[4], dr, di);
            // ub = P/2: k = P/2 + P·a3 with P/2 + P·(7 − a3)
#pragma unroll
            for (int a3 = 0; a3 < 4; ++a3) c2r_pair<T, M>(twR, H + P * a3, xb[a3], xb[7 - a3], zbr[a3], zbi[a3], zbr[7 - a3], zbi[7 - a3]);
        }
        dft8<+1>(zar, zai);
        dft8<+1>(zbr, zbi);
        const int pa = 64 * (ua % R1) + 8 * (ua / R1), pb = 64 * (ub % R1) + 8 * (ub / R1);
#pragma unroll
        for (int b3 = 0; b3 < 8; ++b3) {
            V v; v.x = zar[b3]; v.y = zai[b3];
            tile[L::idx_s1(pa, b3, c)] = v;
            v.x = zbr[b3]; v.y = zbi[b3];
            tile[L::idx_s1(pb, b3, c)] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// in-place variants for tiles that a bulk copy left in natural order (L::nat) in the SAME buffer the
// stages work in (L::idx): everything is read into registers, the caller places a block barrier,
// then the results are written.
// ---------------------------------------------------------------------------------------------
// NBT = ceil(butterflies / NTHR) butterflies per thread, 16 registers each: rg[16·it + a3] real, rg[16·it + 8 + a3] imaginary
template <class L, typename T, int N, int NTHR, int NBT, typename V>
PM_HD void dit_stageA_load(const V* tile, int tid, T (&rg)[16 * NBT]) {
    constexpr int P = N / 8;
#pragma unroll
    for (int it = 0; it < NBT; ++it) {
        const int b = tid + it * NTHR;
        if (b < P * L::C) {
            int c, u;
            L::template decode<P>(b, c, u);
#pragma unroll
            for (int a3 = 0; a3 < 8; ++a3) {
                const V v = tile[L::nat(u + P * a3, c)];
                rg[16 * it + a3] = v.x; rg[16 * it + 8 + a3] = v.y;
            }
        }
    }
}

template <class L, typename T, int N, int DIR, int NTHR, int NBT, typename V>
PM_HD void dit_stageA_store(V* tile, int tid, T (&rg)[16 * NBT]) {
    constexpr int R1 = N / 64;
    constexpr int P = N / 8;
#pragma unroll
    for (int it = 0; it < NBT; ++it) {
        const int b = tid + it * NTHR;
        if (b < P * L::C) {
            int c, u;
            L::template decode<P>(b, c, u);
            const int a1 = u % R1, a2 = u / R1;
            T r[8], i[8];
#pragma unroll
            for (int a3 = 0; a3 < 8; ++a3) { r[a3] = rg[16 * it + a3]; i[a3] = rg[16 * it + 8 + a3]; }
            dft8<DIR>(r, i);
            const int p0 = 64 * a1 + 8 * a2;
#pragma unroll
            for (int b3 = 0; b3 < 8; ++b3) {
                V v; v.x = r[b3]; v.y = i[b3];
                tile[L::idx_s1(p0, b3, c)] = v;
            }
        }
    }
}

// c2r_pre in two halves: ITEMS = ceil((M/2+1)·C / NTHR) pairs (Z_k, Z_{M−k}) per thread in rg[4·ITEMS]
template <class L, typename T, int M, int NTHR, int ITEMS, typename V>
PM_HD void c2r_pre_load(const V* tile, const V* twR, int tid, T (&rg)[4 * ITEMS]) {
    constexpr int P = M / 2 + 1;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int b = tid + it * NTHR;
        if (b < P * L::C) {
            int c, k;
            L::template decode<P>(b, c, k);
            const int kp = (M - k) & (M - 1);
            V zk, zp;
            if (k == 0) {
                const T x0 = tile[L::nat(0, c)].x;
                zk.x = x0; zk.y = x0;
                zp = zk;
            } else {
                const V xk = tile[L::nat(k, c)];
                const V xp = tile[L::nat(kp, c)];
                const T ar = xk.x + xp.x, ai = xk.y - xp.y;
                T ur = xk.x - xp.x, ui = xk.y + xp.y;
                cmul<+1>(ur, ui, tw_r<M>(twR, k));
                zk.x = ar - ui; zk.y = ai + ur;      // A + i·U
                zp.x = ar + ui; zp.y = -ai + ur;     // conj(A) + i·conj(U)
            }
            rg[4 * it + 0] = zk.x; rg[4 * it + 1] = zk.y; rg[4 * it + 2] = zp.x; rg[4 * it + 3] = zp.y;
        }
    }
}

template <class L, typename T, int M, int NTHR, int ITEMS, typename V>
PM_HD void c2r_pre_store(V* tile, int tid, const T (&rg)[4 * ITEMS]) {
    constexpr int R1 = M / 64;
    constexpr int P = M / 2 + 1;
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int b = tid + it * NTHR;
        if (b < P * L::C) {
            int c, k;
            L::template decode<P>(b, c, k);
            const int kp = (M - k) & (M - 1);
            V zk, zp;
            zk.x = rg[4 * it + 0]; zk.y = rg[4 * it + 1]; zp.x = rg[4 * it + 2]; zp.y = rg[4 * it + 3];
            {
                const int a1 = k % R1, a2 = (k / R1) & 7, a3 = k / (8 * R1);
                tile[L::idx(64 * a1 + 8 * a2 + a3, c)] = zk;
            }
            if (k != kp) {
                const int a1 = kp % R1, a2 = (kp / R1) & 7, a3 = kp / (8 * R1);
                tile[L::idx(64 * a1 + 8 * a2 + a3, c)] = zp;
            }
        }
    }
}

}  // namespace fftc
}  // namespace pm
