// Host-side emulation of the shared-memory Stockham FFT index logic in fen_b200/csrc/fft_core.cuh:
// threads are loops, the barrier between stage_load and stage_store is the loop boundary.
// Built and run by tests/test_host_logic.py (no GPU needed).  Prints max abs error vs an O(L^2) DFT.
#include <cmath>
#include <cstdio>
#include <vector>
#include "../../fen_b200/csrc/fft_core.cuh"
using namespace fen;

template <int L, int R, int DIR>
static void run_stage(std::vector<double2>& s, int IS, int NL, int Ns, const double2* tw) {
    constexpr int T = FftPlan<L>::T;
    std::vector<double2> regs((size_t)NL * T * 8);
    for (int line = 0; line < NL; ++line)
        for (int t = 0; t < T; ++t) stage_load<L, R, DIR>(&regs[((size_t)line * T + t) * 8], s.data(), IS, line, t);
    for (int line = 0; line < NL; ++line)
        for (int t = 0; t < T; ++t)
            stage_store<L, R, DIR>(&regs[((size_t)line * T + t) * 8], s.data(), IS, line, t, Ns, tw);
}

template <int L, int DIR> static double check(int IS, int NL) {
    std::vector<double2> tw(L > 0 ? L : 1);
    for (int m = 0; m < L; ++m) tw[m] = make_double2(cos(-2.0 * M_PI * m / L), sin(-2.0 * M_PI * m / L));
    std::vector<double2> s((size_t)L * IS, make_double2(0, 0)), in((size_t)L * NL);
    unsigned seed = 12345u + L;
    for (int line = 0; line < NL; ++line)
        for (int i = 0; i < L; ++i) {
            seed = seed * 1664525u + 1013904223u; double a = (seed >> 8) / 16777216.0 - 0.5;
            seed = seed * 1664525u + 1013904223u; double b = (seed >> 8) / 16777216.0 - 0.5;
            in[(size_t)line * L + i] = make_double2(a, b);
            s[(size_t)i * IS + line] = make_double2(a, b);
        }
    int Ns = 1;
    if constexpr (L >= 8) {
        for (int st = 0; st < FftPlan<L>::N8; ++st) { run_stage<L, 8, DIR>(s, IS, NL, Ns, tw.data()); Ns *= 8; }
    }
    constexpr int REM = FftPlan<L>::REM;
    if constexpr (REM > 1) run_stage<L, REM, DIR>(s, IS, NL, Ns, tw.data());
    double emax = 0;
    for (int line = 0; line < NL; ++line)
        for (int k = 0; k < L; ++k) {
            double re = 0, im = 0;
            for (int n = 0; n < L; ++n) {
                double ang = DIR * 2.0 * M_PI * (double)((long long)n * k % L) / L;
                double2 x = in[(size_t)line * L + n];
                re += x.x * cos(ang) - x.y * sin(ang);
                im += x.x * sin(ang) + x.y * cos(ang);
            }
            double2 y = s[(size_t)k * IS + line];
            emax = fmax(emax, fmax(fabs(y.x - re), fabs(y.y - im)));
        }
    return emax;
}

template <int L> static int both() {
    double ef = check<L, -1>(8, 8), eb = check<L, +1>(9, 8);
    printf("L=%d fwd=%.3e inv=%.3e\n", L, ef, eb);
    return (ef < 1e-11 && eb < 1e-11) ? 0 : 1;
}

// register-to-register flow of fft_regs<L, DIR, true> (L >= 64): inputs and outputs live in the threads'
// registers, v[m] <-> element t + m * L / 8; barriers are the boundaries between the thread loops
template <int L, int DIR, bool PR = false> static double check_regs(int IS, int NL) {
    constexpr int T = FftPlan<L>::T, N8 = FftPlan<L>::N8, REM = FftPlan<L>::REM;
    constexpr int NST = N8 + (REM > 1 ? 1 : 0);
    std::vector<double2> tw(L);
    for (int m = 0; m < L; ++m) tw[m] = make_double2(cos(-2.0 * M_PI * m / L), sin(-2.0 * M_PI * m / L));
    std::vector<double2> s(PR ? (size_t)NL * IS : (size_t)L * IS, make_double2(0, 0)), in((size_t)L * NL);
    std::vector<double2> regs((size_t)NL * T * 8);
    unsigned seed = 777u + L;
    for (int line = 0; line < NL; ++line)
        for (int i = 0; i < L; ++i) {
            seed = seed * 1664525u + 1013904223u; double a = (seed >> 8) / 16777216.0 - 0.5;
            seed = seed * 1664525u + 1013904223u; double b = (seed >> 8) / 16777216.0 - 0.5;
            in[(size_t)line * L + i] = make_double2(a, b);
        }
    auto R = [&](int line, int t) { return &regs[((size_t)line * T + t) * 8]; };
    for (int line = 0; line < NL; ++line)
        for (int t = 0; t < T; ++t)
            for (int m = 0; m < 8; ++m) R(line, t)[m] = in[(size_t)line * L + t + m * (L / 8)];
    int Ns = 1;
    bool done = false;
    for (int st = 0; st < N8 && !done; ++st) {
        if (st > 0)
            for (int line = 0; line < NL; ++line)
                for (int t = 0; t < T; ++t) stage_load<L, 8, DIR, PR>(R(line, t), s.data(), IS, line, t);
        for (int line = 0; line < NL; ++line)
            for (int t = 0; t < T; ++t) {
                stage_compute<L, 8, DIR>(R(line, t), t, Ns, tw.data());
                if (st != NST - 1) stage_write<L, 8, PR>(R(line, t), s.data(), IS, line, t, Ns);
            }
        if (st == NST - 1) done = true;
        Ns *= 8;
    }
    if constexpr (REM > 1) {
        for (int line = 0; line < NL; ++line)
            for (int t = 0; t < T; ++t) stage_load<L, REM, DIR, PR>(R(line, t), s.data(), IS, line, t);
        for (int line = 0; line < NL; ++line)
            for (int t = 0; t < T; ++t) {
                stage_compute<L, REM, DIR>(R(line, t), t, Ns, tw.data());
                last_permute<REM>(R(line, t));
            }
    }
    double emax = 0;
    for (int line = 0; line < NL; ++line)
        for (int k = 0; k < L; ++k) {
            double re = 0, im = 0;
            for (int n = 0; n < L; ++n) {
                double ang = DIR * 2.0 * M_PI * (double)((long long)n * k % L) / L;
                double2 x = in[(size_t)line * L + n];
                re += x.x * cos(ang) - x.y * sin(ang);
                im += x.x * sin(ang) + x.y * cos(ang);
            }
            const int t = k % (L / 8), m = k / (L / 8);
            double2 y = R(line, t)[m];
            emax = fmax(emax, fmax(fabs(y.x - re), fabs(y.y - im)));
        }
    return emax;
}

// padded-row layout: 8 neighbouring threads of one line (an aligned group = one 128-byte shared-memory wavefront of
// 16-byte accesses) must hit 8 distinct 16-byte bank groups (position mod 8) in every stage read and write
template <int L, int R> static int conflicts_stage(int Ns, int IS) {
    constexpr int T = FftPlan<L>::T, NB = (L / R) / T;
    int bad = 0;
    for (int b = 0; b < NB; ++b)
        for (int r = 0; r < R; ++r)
            for (int t0 = 0; t0 < T; t0 += 8) {
                int seen_r = 0, seen_w = 0;
                for (int q = 0; q < 8; ++q) {
                    const int j = t0 + q + b * T, k = j & (Ns - 1), j0 = (j - k) * R + k;
                    seen_r |= 1 << (spos<true>(j + r * (L / R), IS, 0) & 7);
                    seen_w |= 1 << (spos<true>(j0 + r * Ns, IS, 0) & 7);
                }
                bad += (seen_r != 0xff) + (seen_w != 0xff);
            }
    return bad;
}
template <int L> static int conflicts() {
    constexpr int N8 = FftPlan<L>::N8, REM = FftPlan<L>::REM;
    const int IS = L + L / 8 + 1;
    int bad = 0, Ns = 1;
    for (int st = 0; st < N8; ++st) { bad += conflicts_stage<L, 8>(Ns, IS); Ns *= 8; }
    if constexpr (REM > 1) bad += conflicts_stage<L, REM>(Ns, IS);
    return bad;
}

template <int L> static int both_regs() {
    double ef = check_regs<L, -1>(8, 8), eb = check_regs<L, +1>(9, 4);
    double pf = check_regs<L, -1, true>(L + L / 8 + 1, 8), pb = check_regs<L, +1, true>(L + L / 8 + 1, 3);
    const int bc = conflicts<L>();
    printf("L=%d regs fwd=%.3e inv=%.3e  padded rows fwd=%.3e inv=%.3e bank-conflicting groups=%d\n", L, ef, eb, pf, pb, bc);
    return (ef < 1e-11 && eb < 1e-11 && pf < 1e-11 && pb < 1e-11 && bc == 0) ? 0 : 1;
}

int main() {
    int bad = 0;
    bad += both_regs<64>(); bad += both_regs<128>(); bad += both_regs<256>(); bad += both_regs<512>();
    bad += both_regs<1024>(); bad += both_regs<2048>();
    bad += both<2>(); bad += both<4>(); bad += both<8>(); bad += both<16>(); bad += both<32>();
    bad += both<64>(); bad += both<128>(); bad += both<256>(); bad += both<512>(); bad += both<1024>();
    bad += both<2048>();
    printf(bad ? "FAIL\n" : "OK\n");
    return bad;
}
