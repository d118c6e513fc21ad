// Host-side run of the any-length Poisson kernels (fen_b200/csrc/fft_any.cuh): the kernels are sequences of
// __host__ __device__ phases, so blocks and threads are loops here and the block barrier is the phase boundary.
// Every kernel is compared with direct O(n^2) sums of FFTW's definitions (long double accumulation), global indexing,
// pitches, scaling, spectral divide and ghost writes included.  Built and run by tests/test_host_logic.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../fen_b200/csrc/fft_any.cuh"
using namespace fen;

static unsigned g_seed = 2463534242u;
static double rnd() { g_seed = g_seed * 1664525u + 1013904223u; return (g_seed >> 8) / 16777216.0 - 0.5; }
static const long double PI_L = 3.141592653589793238462643383279502884L;

template <class K> static void launch(const typename K::Args& a, int gx, int gy, int nth, int half) {
    std::vector<double2> sm((size_t)2 * half, make_double2(NAN, NAN));
    const int nph = K::nphases(a);
    for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx)
            for (int ph = 0; ph < nph; ++ph)
                for (int tid = 0; tid < nth; ++tid) K::phase(ph, a, sm.data(), sm.data() + half, bx, by, tid, nth);
}

static std::vector<double2> table(int L) {
    std::vector<double2> t((size_t)(L > 0 ? L : 1));
    for (int m = 0; m < L; ++m) {
        long double ang = -2.0L * PI_L * m / L;
        t[m] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
    return t;
}
static std::vector<double2> half_phases(int n) {
    std::vector<double2> t((size_t)n);
    for (int k = 0; k < n; ++k) {
        long double ang = -PI_L * k / (2.0L * n);
        t[k] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
    return t;
}
// direct DFT of one strided line
static void dft(const double2* x, long long stride, int L, int dir, std::vector<double2>& out) {
    out.resize(L);
    std::vector<long double> c(L), s(L);
    for (int m = 0; m < L; ++m) { c[m] = cosl(2.0L * PI_L * m / L); s[m] = dir * sinl(2.0L * PI_L * m / L); }
    for (int k = 0; k < L; ++k) {
        long double re = 0, im = 0;
        for (int n = 0; n < L; ++n) {
            const int m = (int)((long long)n * k % L);
            re += x[stride * n].x * c[m] - x[stride * n].y * s[m];
            im += x[stride * n].x * s[m] + x[stride * n].y * c[m];
        }
        out[k] = make_double2((double)re, (double)im);
    }
}

static int g_fail = 0;
static void report(const char* what, int L, double err, double scale) {
    const double tol = 2e-13 * scale * (L > 4 ? log2((double)L) : 2.0);
    const bool ok = err <= tol;      // NaN fails
    printf("%-14s L=%-5d err=%.3e tol=%.3e %s\n", what, L, err, tol, ok ? "" : "FAIL");
    if (!ok) g_fail = 1;
}

// ---- strided lines, modes 0 / 1 / 2 -------------------------------------------------------------------------------
static void check_lines(int L, int mode, int every = 1) {
    const int PC = 8, nouter = 2, NL = any_lines_per_block(L);
    const long long sl = (long long)PC * nouter + 8, so = PC;       // a pitch with padding, outer index in between
    std::vector<double2> C((size_t)sl * L + 64), ref;
    for (auto& v : C) v = make_double2(rnd(), rnd());
    ref = C;
    std::vector<double2> tw = table(L);
    std::vector<double> lx(PC), lo(nouter + 3), ll(L);
    for (auto& v : lx) v = -fabs(rnd()) - 0.1;
    for (auto& v : lo) v = -fabs(rnd());
    for (auto& v : ll) v = -fabs(rnd());
    lx[0] = 0.0; lo[3] = 0.0; ll[0] = 0.0;                           // one exactly singular mode: kx = 0, outer 0 (o0 = 3)
    AnyLinesArgs a;
    a.P = any_plan(L); a.NL = NL; a.mode = mode; a.C = C.data(); a.sl = sl; a.so = so; a.o0 = 3; a.tw = tw.data();
    a.scale = mode == 0 ? 1.0 / 3.0 : 1.0; a.lx = lx.data(); a.lo = lo.data(); a.ll = ll.data(); a.norm = 7.0 * L;
    if (a.P.L != L) { printf("no plan for %d\n", L); g_fail = 1; return; }
    launch<AnyLines>(a, PC / NL, nouter, ANY_THREADS, NL * L);
    double err = 0, big = 0;
    std::vector<double2> X, Y, tmp(L);
    for (int o = 0; o < nouter; ++o)
        for (int kx = 0; kx < PC; ++kx) {
            if ((kx + 3 * o) % every) continue;       // long lines: the O(L^2) reference of a few columns only
            const double2* in = ref.data() + kx + so * o;
            dft(in, sl, L, mode == 1 ? +1 : -1, X);
            if (mode == 2) {
                for (int i = 0; i < L; ++i) {
                    const double lam = (lx[kx] + lo[3 + o]) + ll[i];
                    if (lam == 0.0) tmp[i] = make_double2(0, 0);
                    else tmp[i] = make_double2(X[i].x / a.norm / lam, X[i].y / a.norm / lam);
                }
                dft(tmp.data(), 1, L, +1, Y);
            } else {
                Y = X;
                for (auto& v : Y) { v.x *= a.scale; v.y *= a.scale; }
            }
            for (int i = 0; i < L; ++i) {
                const double2 got = C[kx + so * o + sl * i];
                err = fmax(err, fmax(fabs(got.x - Y[i].x), fabs(got.y - Y[i].y)));
                if (!(got.x == got.x)) err = NAN;
                big = fmax(big, fmax(fabs(Y[i].x), fabs(Y[i].y)));
            }
        }
    // nothing outside the lines was touched
    for (size_t e = 0; e < C.size(); ++e) {
        const long long i = (long long)e / sl, rem = (long long)e % sl;
        if (i < L && rem < (long long)PC * nouter) continue;
        if (C[e].x != ref[e].x || C[e].y != ref[e].y) { printf("lines: wrote outside at %zu\n", e); g_fail = 1; break; }
    }
    report(mode == 0 ? "lines fwd" : (mode == 1 ? "lines inv" : "lines solve"), L, err, big);
}

// ---- rows: r2c / c2r and the cosine transforms ---------------------------------------------------------------------
struct Rows {
    int nx, ny, nz, PC;
    AnyLayout lay;
    size_t elems;
    std::vector<double> f;
    std::vector<double2> C;
    Rows(int nx_, int ny_, int nz_, int PC_) : nx(nx_), ny(ny_), nz(nz_), PC(PC_) {
        lay.xoff = 16;
        const int px = (16 + nx + 1 + 15) / 16 * 16;
        lay.sy = px; lay.sz = (long long)px * (ny + 2);
        elems = (size_t)lay.sz * (nz + 2);
        f.assign(elems, 0.0);
        C.assign((size_t)PC * ny * nz, make_double2(0, 0));
    }
};

static void check_rows_c(int N) {
    const int ny = 3, nz = 2, PC = (N / 2 + 1 + 7) / 8 * 8, NR = any_lines_per_block(N);
    Rows R(N, ny, nz, PC);
    for (auto& v : R.f) v = rnd();
    std::vector<double> f0 = R.f;
    std::vector<double2> tw = table(N);
    AnyRowsArgs a;
    a.P = any_plan(N); a.NR = NR; a.inverse = 0; a.lay = R.lay; a.f = R.f.data(); a.C = R.C.data(); a.PC = PC; a.ny = ny;
    a.nrows = ny * nz; a.tw = tw.data(); a.twq = nullptr; a.scale = 1.0 / (double)(float)N;
    launch<AnyRowsC>(a, (a.nrows + NR - 1) / NR, 1, ANY_THREADS, NR * N);
    double err = 0, big = 0;
    std::vector<double2> line(N), X;
    for (int r = 0; r < a.nrows; ++r) {
        for (int i = 0; i < N; ++i) line[i] = make_double2(f0[R.lay.idx(1 + i, r % ny + 1, r / ny + 1)], 0.0);
        dft(line.data(), 1, N, -1, X);
        for (int k = 0; k < PC; ++k) {
            const double2 got = R.C[(size_t)PC * r + k];
            const double2 want = k <= N / 2 ? make_double2(X[k].x * a.scale, X[k].y * a.scale) : make_double2(0, 0);
            err = fmax(err, fmax(fabs(got.x - want.x), fabs(got.y - want.y)));
            if (!(got.x == got.x)) err = NAN;
            big = fmax(big, fabs(want.x));
        }
    }
    report("rows r2c", N, err, big);
    // c2r of that spectrum gives the rows back (times N * scale), ghosts included; the imaginary parts of the DC and
    // Nyquist coefficients are ignored, as FFTW's c2r ignores them
    for (int r = 0; r < a.nrows; ++r) { R.C[(size_t)PC * r].y = 0.37; if (N % 2 == 0) R.C[(size_t)PC * r + N / 2].y = -0.21; }
    for (auto& v : R.f) v = 777.0;
    a.inverse = 1; a.scale = 1.0;
    launch<AnyRowsC>(a, (a.nrows + NR - 1) / NR, 1, ANY_THREADS, NR * N);
    err = 0; big = 0;
    size_t touched = 0;
    for (int r = 0; r < a.nrows; ++r) {
        const int j = r % ny + 1, k = r / ny + 1;
        for (int i = 0; i <= N + 1; ++i) {
            const int src = i == 0 ? N : (i == N + 1 ? 1 : i);
            const double want = f0[R.lay.idx(src, j, k)] * N * (1.0 / (double)(float)N);
            const double got = R.f[R.lay.idx(i, j, k)];
            err = fmax(err, fabs(got - want));
            if (!(got == got)) err = NAN;
            big = fmax(big, fabs(want));
            ++touched;
        }
    }
    size_t changed = 0;
    for (auto& v : R.f) if (v != 777.0) ++changed;
    if (changed != touched) { printf("rows c2r: %zu elements written, %zu expected\n", changed, touched); g_fail = 1; }
    report("rows c2r", N, err, big);
}

static void dct2(const std::vector<double>& x, std::vector<double>& y) {      // REDFT10
    const int N = (int)x.size();
    y.resize(N);
    for (int k = 0; k < N; ++k) {
        long double s = 0;
        for (int n = 0; n < N; ++n) s += x[n] * cosl(PI_L * (n + 0.5L) * k / N);
        y[k] = (double)(2.0L * s);
    }
}
static void dct3(const std::vector<double>& x, std::vector<double>& y) {      // REDFT01
    const int N = (int)x.size();
    y.resize(N);
    for (int k = 0; k < N; ++k) {
        long double s = x[0];
        for (int n = 1; n < N; ++n) s += 2.0L * x[n] * cosl(PI_L * n * (k + 0.5L) / N);
        y[k] = (double)s;
    }
}

static void check_rows_dct(int N) {
    const int ny = 2, nz = 2, PC = (N + 7) / 8 * 8, NR = any_lines_per_block(N);
    Rows R(N, ny, nz, PC);
    for (auto& v : R.f) v = rnd();
    std::vector<double> f0 = R.f;
    std::vector<double2> tw = table(N), twq = half_phases(N);
    AnyRowsArgs a;
    a.P = any_plan(N); a.NR = NR; a.inverse = 0; a.lay = R.lay; a.f = R.f.data(); a.C = R.C.data(); a.PC = PC; a.ny = ny;
    a.nrows = ny * nz; a.tw = tw.data(); a.twq = twq.data(); a.scale = 0.5;
    launch<AnyRowsDct>(a, (a.nrows + NR - 1) / NR, 1, ANY_THREADS, NR * N);
    double err = 0, big = 0;
    std::vector<double> x(N), y;
    for (int r = 0; r < a.nrows; ++r) {
        for (int i = 0; i < N; ++i) x[i] = f0[R.lay.idx(1 + i, r % ny + 1, r / ny + 1)];
        dct2(x, y);
        for (int k = 0; k < N; ++k) {
            const double2 got = R.C[(size_t)PC * r + k];
            err = fmax(err, fmax(fabs(got.x - 0.5 * y[k]), fabs(got.y)));
            if (!(got.x == got.x)) err = NAN;
            big = fmax(big, fabs(y[k]));
        }
    }
    report("rows dct2", N, err, big);
    for (auto& v : R.C) v = make_double2(rnd(), 123.0);             // the imaginary parts must be ignored
    std::vector<double2> C0 = R.C;
    a.inverse = 1; a.scale = 0.25;
    launch<AnyRowsDct>(a, (a.nrows + NR - 1) / NR, 1, ANY_THREADS, NR * N);
    err = 0; big = 0;
    for (int r = 0; r < a.nrows; ++r) {
        for (int i = 0; i < N; ++i) x[i] = C0[(size_t)PC * r + i].x;
        dct3(x, y);
        for (int i = 0; i < N; ++i) {
            const double got = R.f[R.lay.idx(1 + i, r % ny + 1, r / ny + 1)];
            err = fmax(err, fabs(got - 0.25 * y[i]));
            if (!(got == got)) err = NAN;
            big = fmax(big, fabs(y[i]));
        }
    }
    report("rows dct3", N, err, big);
}

static void check_lines_dct(int L) {
    const int PC = 8, nouter = 2, NL = any_lines_per_block(L);
    const long long sl = (long long)PC * nouter, so = PC;
    std::vector<double2> C((size_t)sl * L), ref;
    for (auto& v : C) v = make_double2(rnd(), 55.0);
    ref = C;
    std::vector<double2> tw = table(L), twq = half_phases(L);
    AnyLinesDctArgs a;
    a.P = any_plan(L); a.NL = NL; a.C = C.data(); a.sl = sl; a.so = so; a.tw = tw.data(); a.twq = twq.data();
    for (int inverse = 0; inverse < 2; ++inverse) {
        C = ref;
        a.inverse = inverse; a.scale = inverse ? 1.0 : 1.0 / (2.0 * L);
        launch<AnyLinesDct>(a, PC / NL, nouter, ANY_THREADS, NL * L);
        double err = 0, big = 0;
        std::vector<double> x(L), y;
        for (int o = 0; o < nouter; ++o)
            for (int kx = 0; kx < PC; ++kx) {
                for (int i = 0; i < L; ++i) x[i] = ref[kx + so * o + sl * i].x;
                if (inverse) dct3(x, y); else dct2(x, y);
                for (int i = 0; i < L; ++i) {
                    const double2 got = C[kx + so * o + sl * i];
                    err = fmax(err, fmax(fabs(got.x - y[i] * a.scale), fabs(got.y)));
                    if (!(got.x == got.x)) err = NAN;
                    big = fmax(big, fabs(y[i] * a.scale));
                }
            }
        report(inverse ? "lines dct3" : "lines dct2", L, err, big);
    }
}

int main() {
    // plans: supported / rejected lengths
    if (!any_supported(96) || !any_supported(4608) || !any_supported(3072) || !any_supported(1) ||
        !any_supported(31 * 29) || !any_supported(2 * 61) || any_supported(67) || any_supported(2 * 6144) ||
        any_supported(0)) {
        printf("plan support table wrong\n");
        return 1;
    }
    { AnyPlan p = any_plan(4608); int prod = 1; for (int s = 0; s < p.nst; ++s) prod *= p.radix[s];
      if (prod != 4608) { printf("plan product %d\n", prod); return 1; } }
    const int small[] = {1, 2, 3, 4, 5, 6, 7, 9, 10, 12, 15, 16, 22, 37, 48, 58, 62, 96, 100, 106, 122, 124, 243, 1001};
    for (int L : small) { check_lines(L, 0); check_lines(L, 1); check_lines(L, 2); }
    check_lines(3072, 0, 5);
    check_lines(4608, 2, 7);
    check_lines(6144, 1, 11);
    const int rows[] = {2, 3, 5, 6, 10, 12, 15, 48, 96, 100, 243};
    for (int N : rows) { check_rows_c(N); check_rows_dct(N); check_lines_dct(N); }
    check_rows_c(1536);
    printf(g_fail ? "FAILED\n" : "OK\n");
    return g_fail;
}
