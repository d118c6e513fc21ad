// Host-side run of the 32-values-per-thread transforms of fen_b200/csrc/fft_wide.cuh: the in-register 16- and
// 32-point butterflies and the two stages of the 512-point transform (threads as loops, the barrier as the loop
// boundary) against O(n^2) DFT sums, both directions.  Built and run by tests/test_host_logic.py (no GPU needed).
#include <cmath>
#include <cstdio>
#include <vector>
#include "../../fen_b200/csrc/fft_wide.cuh"
using namespace fen;

static unsigned seed = 2026u;
static double rnd() { seed = seed * 1664525u + 1013904223u; return (seed >> 8) / 16777216.0 - 0.5; }

template <int N, int DIR> static double dft_err(const double2* in, const double2* out) {
    double emax = 0;
    for (int k = 0; k < N; ++k) {
        double re = 0, im = 0;
        for (int n = 0; n < N; ++n) {
            const double ang = DIR * 2.0 * M_PI * (double)((n * k) % N) / N;
            re += in[n].x * cos(ang) - in[n].y * sin(ang);
            im += in[n].x * sin(ang) + in[n].y * cos(ang);
        }
        emax = fmax(emax, fmax(fabs(out[k].x - re), fabs(out[k].y - im)));
    }
    return emax;
}

template <int DIR> static double check_small() {
    double2 a[16], b[16], c[32], d[32];
    for (int i = 0; i < 16; ++i) a[i] = b[i] = make_double2(rnd(), rnd());
    for (int i = 0; i < 32; ++i) c[i] = d[i] = make_double2(rnd(), rnd());
    bfly16<DIR>(b);
    bfly32<DIR>(d);
    return fmax(dft_err<16, DIR>(a, b), dft_err<32, DIR>(c, d));
}

template <int DIR> static double check_512(int IS, int NL) {
    constexpr int L = 512, T = 16;
    std::vector<double2> tw(L), in((size_t)L * NL), s((size_t)L * IS), regs((size_t)NL * T * 32), out((size_t)L * NL);
    for (int m = 0; m < L; ++m) tw[m] = make_double2(cos(-2.0 * M_PI * m / L), sin(-2.0 * M_PI * m / L));
    for (auto& x : in) x = make_double2(rnd(), rnd());
    auto R = [&](int line, int t) -> double2(&)[32] { return *reinterpret_cast<double2(*)[32]>(&regs[((size_t)line * T + t) * 32]); };
    for (int line = 0; line < NL; ++line)
        for (int t = 0; t < T; ++t)
            for (int m = 0; m < 32; ++m) R(line, t)[m] = in[(size_t)line * L + t + 16 * m];      // v[m] = x[t + 16 m]
    for (int line = 0; line < NL; ++line)
        for (int t = 0; t < T; ++t) wide512_stage1<DIR>(R(line, t), s.data(), IS, line, t);
    for (int line = 0; line < NL; ++line)
        for (int t = 0; t < T; ++t) wide512_stage2<DIR>(R(line, t), s.data(), IS, line, t, tw.data());
    for (int line = 0; line < NL; ++line)
        for (int t = 0; t < T; ++t)
            for (int m = 0; m < 32; ++m) out[(size_t)line * L + t + 16 * m] = R(line, t)[m];      // v[m] = X[t + 16 m]
    double emax = 0;
    for (int line = 0; line < NL; ++line) emax = fmax(emax, dft_err<L, DIR>(&in[(size_t)line * L], &out[(size_t)line * L]));
    return emax;
}

int main() {
    const double e1 = check_small<-1>(), e2 = check_small<+1>();
    const double e3 = check_512<-1>(8, 8), e4 = check_512<+1>(8, 8);
    printf("bfly16/32 fwd=%.3e inv=%.3e  fft512_wide fwd=%.3e inv=%.3e\n", e1, e2, e3, e4);
    return (e1 < 1e-13 && e2 < 1e-13 && e3 < 1e-11 && e4 < 1e-11) ? 0 : 1;
}
