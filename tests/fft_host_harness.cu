// CPU walk-through of the shared-memory FFT building blocks (concept_b200/csrc/pm_fftcore.cuh):
// the device stages are executed sequentially for tid = 0 … nthr−1 and compared with an O(N²)
// long-double DFT.  Built and run by tests/test_fftcore_host.py (no GPU needed).
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pm_fftops.cuh"

using namespace pm::fftc;

static const long double PI = 3.14159265358979323846264338327950288L;

// B | C | R tables of pm_fftcore.cuh for grid size G
template <typename V>
struct HostTw {
    std::vector<V> data;
    Twiddles<V> tw;
    explicit HostTw(int G) : data(64 + G + G / 8 + 1) {
        auto w = [&](long double num, long double den) {
            const long double a = -2.0L * PI * num / den;
            V v; v.x = (decltype(v.x))cosl(a); v.y = (decltype(v.y))sinl(a);
            return v;
        };
        for (int a = 0; a < 8; ++a) for (int b = 0; b < 8; ++b) data[a * 8 + b] = w(a * b, 64);
        for (int a = 0; a < G / 64; ++a) for (int q = 0; q < 64; ++q) data[64 + a * 64 + q] = w(a * q, G);
        for (int k = 0; k <= G / 8; ++k) data[64 + G + k] = w(k, G);
        tw.B = data.data(); tw.C = data.data() + 64; tw.R = data.data() + 64 + G;
    }
};

template <typename T> struct NatSource {   // natural-order host arrays [c][n]
    const std::vector<std::vector<T>>*re, *im;
    typename Vec2<T>::type operator()(int c, int n) const { typename Vec2<T>::type v; v.x = (*re)[c][n]; v.y = (*im)[c][n]; return v; }
};

template <typename T> struct Store {
    T* re; T* im; int N;   // out[c][index]
    void operator()(int c, int idx, T r, T i) const { re[c * N + idx] = r; im[c * N + idx] = i; }
};

template <typename T, class L> struct StoreTile {
    typename Vec2<T>::type* tile;
    void operator()(int c, int idx, T r, T i) const { typename Vec2<T>::type v; v.x = r; v.y = i; tile[L::idx(idx, c)] = v; }
};

static double frand() { return rand() / (double)RAND_MAX - 0.5; }

// reference DFT of line c (natural order), sign dir
template <typename T>
static void ref_dft(const std::vector<T>& xr, const std::vector<T>& xi, int N, int dir, std::vector<long double>& yr,
                    std::vector<long double>& yi) {
    yr.assign(N, 0); yi.assign(N, 0);
    for (int k = 0; k < N; ++k) {
        long double sr = 0, si = 0;
        for (int n = 0; n < N; ++n) {
            const long double a = dir * 2.0L * PI * ((long long)n * k % N) / N;
            const long double cr = cosl(a), ci = sinl(a);
            sr += xr[n] * cr - xi[n] * ci;
            si += xr[n] * ci + xi[n] * cr;
        }
        yr[k] = sr; yi[k] = si;
    }
}

template <typename T, int N, int C, bool ROW>
static double test_complex(int nthr) {
    using V = typename Vec2<T>::type;
    using L = typename std::conditional<ROW, RowLayout<N, C>, ColLayout<C>>::type;
    HostTw<V> ht(N);
    const Twiddles<V>& tw = ht.tw;
    std::vector<V> tile((size_t)(N + N / 8 + N / 64 + 1) * C + 16);
    std::vector<std::vector<T>> xr(C, std::vector<T>(N)), xi(C, std::vector<T>(N));
    for (int c = 0; c < C; ++c)
        for (int n = 0; n < N; ++n) { xr[c][n] = (T)frand(); xi[c][n] = (T)frand(); }
    NatSource<T> src{&xr, &xi};
    double err = 0;
    for (int dir = -1; dir <= 1; dir += 2) {
        std::vector<T> ore(C * N), oim(C * N);
        Store<T> st{ore.data(), oim.data(), N};
        // ---- DIT: natural -> natural
        for (int t = 0; t < nthr; ++t) {
            if (dir < 0) dit_stageA<L, T, N, -1>(src, tile.data(), t, nthr);
            else dit_stageA<L, T, N, +1>(src, tile.data(), t, nthr);
        }
        for (int t = 0; t < nthr; ++t) {
            if (dir < 0) dit_stageB<L, T, N, -1>(tile.data(), tw.B, t, nthr);
            else dit_stageB<L, T, N, +1>(tile.data(), tw.B, t, nthr);
        }
        for (int t = 0; t < nthr; ++t) {
            if (dir < 0) dit_stageC<L, T, N, 1, -1>(tile.data(), tw.C, t, nthr, st);
            else dit_stageC<L, T, N, 1, +1>(tile.data(), tw.C, t, nthr, st);
        }
        std::vector<long double> yr, yi;
        for (int c = 0; c < C; ++c) {
            ref_dft(xr[c], xi[c], N, dir, yr, yi);
            for (int k = 0; k < N; ++k) {
                err = fmax(err, fabs((double)(ore[c * N + k] - yr[k])));
                err = fmax(err, fabs((double)(oim[c * N + k] - yi[k])));
            }
        }
        // ---- DIF: natural (registers) -> scattered elements
        std::vector<T> dre(C * N), dim_(C * N);
        Store<T> st2{dre.data(), dim_.data(), N};
        constexpr int R1 = N / 64;
        for (int t = 0; t < nthr; ++t)
            for (int b = t; b < 64 * L::C; b += nthr) {
                int c, q;
                L::template decode<64>(b, c, q);
                T r[R1], i[R1];
                for (int a1 = 0; a1 < R1; ++a1) { r[a1] = xr[c][64 * a1 + q]; i[a1] = xi[c][64 * a1 + q]; }
                if (dir < 0) dif_stage1_regs<L, T, N, 1, -1>(r, i, tile.data(), tw.C, c, q);
                else dif_stage1_regs<L, T, N, 1, +1>(r, i, tile.data(), tw.C, c, q);
            }
        for (int t = 0; t < nthr; ++t) {
            if (dir < 0) dif_stage2<L, T, N, -1>(tile.data(), tw.B, t, nthr);
            else dif_stage2<L, T, N, +1>(tile.data(), tw.B, t, nthr);
        }
        for (int t = 0; t < nthr; ++t) {
            if (dir < 0) dif_stage3<L, T, N, -1>(tile.data(), t, nthr, st2);
            else dif_stage3<L, T, N, +1>(tile.data(), t, nthr, st2);
        }
        for (int c = 0; c < C; ++c) {
            ref_dft(xr[c], xi[c], N, dir, yr, yi);
            for (int k = 0; k < N; ++k) {
                err = fmax(err, fabs((double)(dre[c * N + k] - yr[k])));
                err = fmax(err, fabs((double)(dim_[c * N + k] - yi[k])));
            }
        }
    }
    return err;
}

// z pass: 2M reals per line <-> M+1 complex (Nyquist dropped)
template <typename T, int M, int C>
static double test_real(int nthr) {
    using V = typename Vec2<T>::type;
    using L = RowLayout<M, C>;
    HostTw<V> ht(2 * M);
    const Twiddles<V>& tw = ht.tw;
    std::vector<V> tile((size_t)L::PITCH * C + 16);
    std::vector<std::vector<T>> x(C, std::vector<T>(2 * M)), zr(C, std::vector<T>(M)), zi(C, std::vector<T>(M));
    for (int c = 0; c < C; ++c)
        for (int n = 0; n < M; ++n) {
            x[c][2 * n] = (T)frand(); x[c][2 * n + 1] = (T)frand();
            zr[c][n] = x[c][2 * n]; zi[c][n] = x[c][2 * n + 1];
        }
    NatSource<T> src{&zr, &zi};
    double err = 0;
    // forward
    StoreTile<T, L> stt{tile.data()};
    for (int t = 0; t < nthr; ++t) dit_stageA<L, T, M, -1>(src, tile.data(), t, nthr);
    for (int t = 0; t < nthr; ++t) dit_stageB<L, T, M, -1>(tile.data(), tw.B, t, nthr);
    for (int t = 0; t < nthr; ++t) dit_stageC<L, T, M, 2, -1>(tile.data(), tw.C, t, nthr, stt);
    std::vector<T> ore(C * (M + 1)), oim(C * (M + 1));
    Store<T> st{ore.data(), oim.data(), M + 1};
    for (int t = 0; t < nthr; ++t) r2c_post<L, T, M>(tile.data(), tw.R, t, nthr, st);
    std::vector<std::vector<long double>> Xr(C, std::vector<long double>(M + 1)), Xi(C, std::vector<long double>(M + 1));
    for (int c = 0; c < C; ++c) {
        for (int k = 0; k <= M; ++k) {
            long double sr = 0, si = 0;
            for (int n = 0; n < 2 * M; ++n) {
                const long double a = -2.0L * PI * ((long long)n * k % (2 * M)) / (2 * M);
                sr += x[c][n] * cosl(a); si += x[c][n] * sinl(a);
            }
            Xr[c][k] = sr; Xi[c][k] = si;
            if (k < M) {
                err = fmax(err, fabs((double)(ore[c * (M + 1) + k] - sr)));
                err = fmax(err, fabs((double)(oim[c * (M + 1) + k] - si)));
            } else {
                err = fmax(err, fabs((double)ore[c * (M + 1) + k]) + fabs((double)oim[c * (M + 1) + k]));   // delivered as zero
            }
        }
    }
    // inverse: feed the exact spectrum with X_M := 0
    std::vector<std::vector<T>> sr(C, std::vector<T>(M)), si(C, std::vector<T>(M));
    for (int c = 0; c < C; ++c)
        for (int k = 0; k < M; ++k) { sr[c][k] = (T)Xr[c][k]; si[c][k] = (T)Xi[c][k]; }
    NatSource<T> spec{&sr, &si};
    for (int t = 0; t < nthr; ++t) c2r_pre<L, T, M>(spec, tile.data(), tw.R, t, nthr);
    for (int t = 0; t < nthr; ++t) dit_stageA_inplace<L, T, M, +1>(tile.data(), t, nthr);
    for (int t = 0; t < nthr; ++t) dit_stageB<L, T, M, +1>(tile.data(), tw.B, t, nthr);
    std::vector<T> zre(C * M), zim(C * M);
    Store<T> stz{zre.data(), zim.data(), M};
    for (int t = 0; t < nthr; ++t) dit_stageC<L, T, M, 2, +1>(tile.data(), tw.C, t, nthr, stz);
    for (int c = 0; c < C; ++c)
        for (int n = 0; n < 2 * M; ++n) {
            // 2M·x_n minus the dropped Nyquist term X_M·(−1)^n
            const long double expect = 2.0L * M * x[c][n] - Xr[c][M] * ((n & 1) ? -1 : 1);
            const T got = (n & 1) ? zim[c * M + n / 2] : zre[c * M + n / 2];
            err = fmax(err, fabs((double)(got - expect)) / (2.0 * M));
        }
    // ---- the fused variants the kernels use: last stage + r2c_post, c2r_pre + first stage
    for (int t = 0; t < nthr; ++t) dit_stageA<L, T, M, -1>(src, tile.data(), t, nthr);
    for (int t = 0; t < nthr; ++t) dit_stageB<L, T, M, -1>(tile.data(), tw.B, t, nthr);
    std::fill(ore.begin(), ore.end(), (T)99); std::fill(oim.begin(), oim.end(), (T)99);
    for (int t = 0; t < nthr; ++t) r2c_stageC_post<L, T, M>(tile.data(), tw.C, tw.R, t, nthr, st);
    for (int c = 0; c < C; ++c)
        for (int k = 0; k <= M; ++k) {
            if (k < M) {
                err = fmax(err, fabs((double)(ore[c * (M + 1) + k] - Xr[c][k])));
                err = fmax(err, fabs((double)(oim[c * (M + 1) + k] - Xi[c][k])));
            } else {
                err = fmax(err, fabs((double)ore[c * (M + 1) + k]) + fabs((double)oim[c * (M + 1) + k]));
            }
        }
    for (int t = 0; t < nthr; ++t) c2r_pre_stageA<L, T, M>(spec, tile.data(), tw.R, t, nthr);
    for (int t = 0; t < nthr; ++t) dit_stageB<L, T, M, +1>(tile.data(), tw.B, t, nthr);
    std::fill(zre.begin(), zre.end(), (T)99); std::fill(zim.begin(), zim.end(), (T)99);
    for (int t = 0; t < nthr; ++t) dit_stageC<L, T, M, 2, +1>(tile.data(), tw.C, t, nthr, stz);
    for (int c = 0; c < C; ++c)
        for (int n = 0; n < 2 * M; ++n) {
            const long double expect = 2.0L * M * x[c][n] - Xr[c][M] * ((n & 1) ? -1 : 1);
            const T got = (n & 1) ? zim[c * M + n / 2] : zre[c * M + n / 2];
            err = fmax(err, fabs((double)(got - expect)) / (2.0 * M));
        }
    return err;
}

// ---- a whole fused solve on the CPU: real -> A -> B -> A -> real, exactly the device tile operations ----
template <typename T, int G>
static int solve3d(const char* fin, const char* fsep, const char* fout, double prefactor, int nranks) {
    constexpr int nthr = 256;
    using S = SlabFFT<T, G, nthr>;
    using V = typename S::V;
    const int nxl = G / nranks;
    std::vector<double> in((size_t)G * G * G), sep(G);
    FILE* f = fopen(fin, "rb");
    if (!f || fread(in.data(), sizeof(double), in.size(), f) != in.size()) return 2;
    fclose(f);
    f = fopen(fsep, "rb");
    if (!f || fread(sep.data(), sizeof(double), G, f) != (size_t)G) return 2;
    fclose(f);
    // per "rank": padded real slab (padding holds garbage), A and B
    std::vector<std::vector<T>> slab(nranks, std::vector<T>((size_t)nxl * G * S::Gp, (T)7));
    std::vector<std::vector<V>> A(nranks, std::vector<V>((size_t)nxl * S::M * G)), B(nranks, std::vector<V>((size_t)nxl * S::M * G));
    for (int i = 0; i < G; ++i)
        for (int j = 0; j < G; ++j)
            for (int k = 0; k < G; ++k)
                slab[i / nxl][((size_t)(i % nxl) * G + j) * S::Gp + k] = (T)in[((size_t)i * G + j) * G + k];
    HostTw<V> ht(G);
    std::vector<V> buf(S::kBufElems + 16);
    std::vector<T> regs((size_t)nthr * S::kRegs);
    auto run = [&](const auto& op, int nloads) {
        for (int k = 0; k < nloads && op.kBulk; ++k) {
            const TileLoad l = op.load(k);
            if (l.bytes % 16 || l.dst_bytes % 16 || ((uintptr_t)l.src) % 16) { printf("misaligned load\n"); exit(3); }
            memcpy(reinterpret_cast<char*>(buf.data()) + l.dst_bytes, l.src, l.bytes);
        }
        for (int ph = 0; ph < op.kPhases; ++ph)
            for (int t = 0; t < nthr; ++t)
                op.phase(ph, buf.data(), ht.tw, t, nthr, reinterpret_cast<T(&)[S::kRegs]>(regs[(size_t)t * S::kRegs]));
    };
    for (int r = 0; r < nranks; ++r)
        for (int p = 0; p < nxl; ++p) {
            T* plane = slab[r].data() + (size_t)p * G * S::Gp;
            V* a_plane = A[r].data() + (size_t)p * S::NKT * G * S::CY;
            for (int t = 0; t < S::kZTilesPerPlane; ++t) run(typename S::ZFwd{plane, a_plane, t * S::CZ, false}, 1);
            for (int kt = 0; kt < S::NKT; ++kt) run(typename S::YFwd{a_plane + (size_t)kt * G * S::CY, B[r].data(), p, kt, nxl}, 1);
        }
    typename S::XGeom xg;
    for (int r = 0; r < kMaxFftPeers; ++r) { xg.a[r] = r < nranks ? A[r].data() : nullptr; xg.b[r] = r < nranks ? B[r].data() : nullptr; }
    xg.nranks = nranks;
    xg.in_place = 0;
    xg.nxl_shift = 0;
    while ((1 << xg.nxl_shift) < nxl) ++xg.nxl_shift;
    xg.sep = sep.data();
    xg.prefactor = prefactor;
    for (int j = 0; j < G; ++j)
        for (int kt = 0; kt < S::NKT; ++kt) run(typename S::XSolve{&xg, j, kt}, nranks);
    for (int r = 0; r < nranks; ++r)
        for (int p = 0; p < nxl; ++p) {
            T* plane = slab[r].data() + (size_t)p * G * S::Gp;
            V* a_plane = A[r].data() + (size_t)p * S::NKT * G * S::CY;
            for (int kt = 0; kt < S::NKT; ++kt) run(typename S::YInv{a_plane + (size_t)kt * G * S::CY, plane, kt}, 1);
            for (int t = 0; t < S::kZTilesPerPlane; ++t) run(typename S::ZInv{plane, t * S::CZ}, 1);
        }
    std::vector<double> out((size_t)G * G * G);
    for (int i = 0; i < G; ++i)
        for (int j = 0; j < G; ++j)
            for (int k = 0; k < G; ++k)
                out[((size_t)i * G + j) * G + k] = (double)slab[i / nxl][((size_t)(i % nxl) * G + j) * S::Gp + k];
    f = fopen(fout, "wb");
    if (!f || fwrite(out.data(), sizeof(double), out.size(), f) != out.size()) return 2;
    fclose(f);
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 8 && std::string(argv[1]) == "solve3d") {
        // solve3d <f64|f32> <in.bin> <sep.bin> <out.bin> <prefactor> <nranks> [G = 128 | 256]
        const double pre = atof(argv[6]);
        const int nr = atoi(argv[7]);
        const int G = argc >= 9 ? atoi(argv[8]) : 128;
        const bool f64 = std::string(argv[2]) == "f64";
        if (G == 256) return f64 ? solve3d<double, 256>(argv[3], argv[4], argv[5], pre, nr) : solve3d<float, 256>(argv[3], argv[4], argv[5], pre, nr);
        return f64 ? solve3d<double, 128>(argv[3], argv[4], argv[5], pre, nr) : solve3d<float, 128>(argv[3], argv[4], argv[5], pre, nr);
    }
    int fails = 0;
    auto report = [&](const char* name, double err, double tol) {
        printf("%-40s max err %.3e %s\n", name, err, err < tol ? "ok" : "FAIL");
        if (!(err < tol)) ++fails;
    };
    report("col f64 N=1024 C=4 nthr=256", test_complex<double, 1024, 4, false>(256), 1e-12);
    report("col f64 N=512 C=8  nthr=512", test_complex<double, 512, 8, false>(512), 1e-12);
    report("col f64 N=256 C=8  nthr=512", test_complex<double, 256, 8, false>(512), 1e-12);
    report("col f64 N=128 C=8  nthr=512", test_complex<double, 128, 8, false>(512), 1e-12);
    report("col f64 N=64  C=8  nthr=512", test_complex<double, 64, 8, false>(512), 1e-12);
    report("col f32 N=512 C=16 nthr=512", test_complex<float, 512, 16, false>(512), 2e-4);
    report("col f32 N=128 C=16 nthr=256", test_complex<float, 128, 16, false>(256), 1e-4);
    report("row f64 N=256 C=16 nthr=512", test_complex<double, 256, 16, true>(512), 1e-12);
    report("row f64 N=64  C=16 nthr=512", test_complex<double, 64, 16, true>(512), 1e-12);
    report("real f64 M=256 C=16", test_real<double, 256, 16>(512), 1e-12);
    report("real f64 M=128 C=16", test_real<double, 128, 16>(512), 1e-12);
    report("real f64 M=64  C=16", test_real<double, 64, 16>(512), 1e-12);
    report("real f32 M=256 C=16", test_real<float, 256, 16>(512), 2e-4);
    report("real f64 M=512 C=8", test_real<double, 512, 8>(256), 1e-12);
    return fails;
}
