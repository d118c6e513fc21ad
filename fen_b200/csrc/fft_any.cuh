// fft_any.cuh -- the ANY-LENGTH path of the Poisson solver: mixed-radix Stockham transforms in shared memory for
// the grid sizes that are not powers of two (the reference has no size restriction: FFTW plans of any n,
// src/poisson.f90:148-151, 635-674; its own drivers use 96 x 96, 16 x 48, 3072 x 4608 --
// test/small_test/fsi/Pan_Eulerian/Pan.f90:33-34, test/small_test/io/test_MF.f90, test/large_test/
// startup_flow_cylinder/main.f90:37-38).  The power-of-two lengths keep the tuned register-path kernels of
// poisson.cu / fft_core.cuh; this path is the coverage path: every length whose prime factors are <= 61, up to
// ANY_MAX_L points, one rank.
//
// Structure.  Every kernel is a sequence of PHASES separated by block barriers:
//     load (global -> shared, with the r2c / Hermitian / DCT reordering of the variant)
//     one phase per Stockham stage (radix 4, 2, 3, 5, then the remaining primes up to 61), ping-pong between two shared buffers
//     [spectral divide, then the inverse stages]                                   (fused solve)
//     store (shared -> global, with the scaling / post-twiddle of the variant)
// A phase is a __host__ __device__ function of (arguments, shared buffers, block index, thread index, block size), so
// tests/cpu/test_fft_any.cu runs the SAME code on the CPU -- blocks and threads as loops, barriers as phase
// boundaries -- against direct O(n^2) DFT / DCT sums, global indexing, pitches and ghost writes included.
//
// Transform definitions (FFTW's, as the reference uses them): forward = sum x_n exp(-2 pi i n k / L), unnormalised;
// backward = the same with +i; REDFT10 = 2 sum x_n cos(pi (n + 1/2) k / L); REDFT01 = x_0 + 2 sum_{n>=1} x_n
// cos(pi n (k + 1/2) / L).  The cosine transforms go through one complex transform of the same length (Makhoul's
// reordering, as in poisson.cu's power-of-two kernels; valid for odd lengths too).
#pragma once
#include <cuda_runtime.h>

#include "fft_core.cuh"

namespace fen {

constexpr int ANY_MAX_L = 6144;          // two ping-pong lines of 16-byte elements: 2 * 6144 * 16 B = 192 KB of shared memory
constexpr int ANY_MAX_RADIX = 61;
constexpr int ANY_MAX_STAGES = 16;
constexpr int ANY_THREADS = 256;

struct AnyPlan {
    int L = 0, nst = 0;
    int radix[ANY_MAX_STAGES] = {};
};

// radices of a length: 4s first, then 2, 3, 5, then the remaining primes up to ANY_MAX_RADIX.  nst = 0 with L > 1
// means "not supported" (a larger prime factor, or too long).
inline AnyPlan any_plan(int L) {
    AnyPlan p;
    p.L = L;
    if (L < 1 || L > ANY_MAX_L) { p.L = 0; return p; }
    int n = L;
    while (n % 4 == 0 && p.nst < ANY_MAX_STAGES) { p.radix[p.nst++] = 4; n /= 4; }
    for (int r = 2; r <= ANY_MAX_RADIX && n > 1; ++r)
        while (n % r == 0 && p.nst < ANY_MAX_STAGES) { p.radix[p.nst++] = r; n /= r; }
    if (n != 1) { p.nst = 0; p.L = 0; }
    return p;
}
inline bool any_supported(int L) { return L >= 1 && any_plan(L).L == L; }

// lines (or rows) one block transforms at once: the largest of 8, 4, 2, 1 whose two buffers fit
inline int any_lines_per_block(int L) {
    for (int nl = 8; nl > 1; nl /= 2)
        if ((size_t)2 * nl * L * sizeof(double2) <= (size_t)2 * ANY_MAX_L * sizeof(double2)) return nl;
    return 1;
}
inline size_t any_smem_bytes(int L, int nl) { return (size_t)2 * nl * L * sizeof(double2); }

// ---- one Stockham stage -----------------------------------------------------------------------------------------
// Element i of line `line` lives at buf[i * NL + line].  Stage with radix r after the radices whose product is Ns:
// butterfly j (j < L / r), k = j mod Ns:  x_q = in[j + q L/r] * w_L^(q k L/(Ns r)),  y_m = sum_q x_q w_r^(q m),
// out[(j - k) r + k + m Ns] = y_m.  tw[n] = exp(-2 pi i n / L), n < L; dir = -1 forward, +1 backward (conjugated).
FEN_HD double2 any_tw(const double2* tw, int n, int dir) {
    const double2 w = tw[n];
    return dir < 0 ? w : make_double2(w.x, -w.y);
}
FEN_HD void any_stage(const AnyPlan& P, int st, int Ns, const double2* in, double2* out, int NL, int tid, int nthreads,
                      const double2* tw, int dir) {
    const int L = P.L, r = P.radix[st], Lr = L / r;
    const int line = tid % NL, t = tid / NL, T = nthreads / NL;
    if (t >= T) return;                              // nthreads is a multiple of NL in every launch; belt and braces
    const int step = L / (Ns * r);
    for (int j = t; j < Lr; j += T) {
        const int k = j % Ns;
        const int o0 = (j - k) * r + k;
        if (r == 2) {
            double2 a = in[j * NL + line], b = in[(j + Lr) * NL + line];
            if (Ns > 1) b = cmul(b, any_tw(tw, k * step, dir));
            out[o0 * NL + line] = cadd(a, b);
            out[(o0 + Ns) * NL + line] = csub(a, b);
        } else if (r == 4) {
            double2 a0 = in[j * NL + line], a1 = in[(j + Lr) * NL + line], a2 = in[(j + 2 * Lr) * NL + line],
                    a3 = in[(j + 3 * Lr) * NL + line];
            if (Ns > 1) {
                a1 = cmul(a1, any_tw(tw, k * step, dir));
                a2 = cmul(a2, any_tw(tw, 2 * k * step, dir));
                a3 = cmul(a3, any_tw(tw, 3 * k * step, dir));
            }
            if (dir < 0) bfly4<-1>(a0, a1, a2, a3);
            else bfly4<+1>(a0, a1, a2, a3);
            out[o0 * NL + line] = a0;
            out[(o0 + Ns) * NL + line] = a1;
            out[(o0 + 2 * Ns) * NL + line] = a2;
            out[(o0 + 3 * Ns) * NL + line] = a3;
        } else {
            double2 x[ANY_MAX_RADIX];
            for (int q = 0; q < r; ++q) {
                double2 v = in[(j + q * Lr) * NL + line];
                if (q > 0 && Ns > 1) v = cmul(v, any_tw(tw, q * k * step, dir));
                x[q] = v;
            }
            for (int m = 0; m < r; ++m) {
                double2 acc = x[0];
                for (int q = 1; q < r; ++q) acc = cadd(acc, cmul(x[q], any_tw(tw, ((q * m) % r) * Lr, dir)));
                out[(o0 + m * Ns) * NL + line] = acc;
            }
        }
    }
}

// Phases [ph0, ph0 + nst) are the stages of one transform that starts in buffer `first` (0 or 1); returns true when
// `ph` was one of them.  The result of the whole transform is in buffer (first + nst) & 1.
FEN_HD bool any_stage_phase(const AnyPlan& P, int ph, int ph0, int first, double2* s0, double2* s1, int NL, int tid,
                            int nthreads, const double2* tw, int dir) {
    const int st = ph - ph0;
    if (st < 0 || st >= P.nst) return false;
    int Ns = 1;
    for (int q = 0; q < st; ++q) Ns *= P.radix[q];
    const bool from0 = ((first + st) & 1) == 0;
    any_stage(P, st, Ns, from0 ? s0 : s1, from0 ? s1 : s0, NL, tid, nthreads, tw, dir);
    return true;
}

FEN_HD int any_dct_perm(int i, int N) { return (i & 1) ? N - 1 - (i >> 1) : (i >> 1); }

struct AnyLayout {                        // the part of fen::Layout the row kernels need
    int xoff;
    long long sy, sz;
    FEN_HD long long idx(int i, int j, int k) const { return (long long)(xoff - 1 + i) + sy * j + sz * k; }
};

// =================================================================================================================
// strided lines (y or z): NL consecutive columns per block; mode 0 forward, 1 backward, 2 forward + divide + backward
// =================================================================================================================
struct AnyLinesArgs {
    AnyPlan P;
    int NL, mode;
    double2* C;
    long long sl, so;                     // line element stride, outer stride (complex elements)
    int o0;
    const double2* tw;                    // exp(-2 pi i n / L), n < L
    double scale;
    const double* lx; const double* lo; const double* ll;
    double norm;
};
struct AnyLines {
    typedef AnyLinesArgs Args;
    FEN_HD static int nphases(const Args& a) { return a.mode == 2 ? 2 * a.P.nst + 3 : a.P.nst + 2; }
    FEN_HD static void phase(int ph, const Args& a, double2* s0, double2* s1, int bx, int by, int tid, int nth) {
        const int L = a.P.L, NL = a.NL, nst = a.P.nst;
        double2* base = a.C + (long long)bx * NL + a.so * by;
        if (ph == 0) {
            for (int e = tid; e < NL * L; e += nth) {
                const int line = e % NL, i = e / NL;
                s0[e] = base[line + a.sl * i];
            }
            return;
        }
        if (any_stage_phase(a.P, ph, 1, 0, s0, s1, NL, tid, nth, a.tw, a.mode == 1 ? +1 : -1)) return;
        double2* res = (nst & 1) ? s1 : s0;
        if (a.mode != 2) {
            for (int e = tid; e < NL * L; e += nth) {
                const int line = e % NL, i = e / NL;
                const double2 v = res[e];
                base[line + a.sl * i] = make_double2(v.x * a.scale, v.y * a.scale);
            }
            return;
        }
        if (ph == nst + 1) {
            // poisson.f90:992 then :998-1002 (pp: :458 then :463-467): divide by float(nx*ny*nz), then by the sum of
            // the modified wavenumbers, the exactly singular mode set to zero
            for (int e = tid; e < NL * L; e += nth) {
                const int line = e % NL, i = e / NL;
                double lam = a.lx[bx * NL + line];
                if (a.lo) lam = lam + a.lo[a.o0 + by];
                lam = lam + a.ll[i];
                double2 v = res[e];
                v.x = v.x / a.norm;
                v.y = v.y / a.norm;
                if (lam == 0.0) v = make_double2(0.0, 0.0);
                else { v.x = v.x / lam; v.y = v.y / lam; }
                res[e] = v;
            }
            return;
        }
        if (any_stage_phase(a.P, ph, nst + 2, nst & 1, s0, s1, NL, tid, nth, a.tw, +1)) return;
        for (int e = tid; e < NL * L; e += nth) {     // after 2 nst stages the data is back in s0
            const int line = e % NL, i = e / NL;
            base[line + a.sl * i] = s0[e];
        }
    }
};

// =================================================================================================================
// x direction, periodic: real rows <-> half spectrum (FFTW r2c / c2r through one complex transform of length nx)
// =================================================================================================================
struct AnyRowsArgs {
    AnyPlan P;                            // L = nx
    int NR, inverse;                      // rows per block
    AnyLayout lay;
    double* f;
    double2* C;
    int PC, ny, nrows;
    const double2* tw;                    // exp(-2 pi i n / nx)
    const double2* twq;                   // exp(-i pi k / (2 nx)): cosine transforms only
    double scale;
};
struct AnyRowsC {                         // r2c (inverse = 0) and c2r (inverse = 1)
    typedef AnyRowsArgs Args;
    FEN_HD static int nphases(const Args& a) { return a.P.nst + 2; }
    FEN_HD static void phase(int ph, const Args& a, double2* s0, double2* s1, int bx, int, int tid, int nth) {
        const int N = a.P.L, NR = a.NR, nst = a.P.nst;
        if (ph == 0) {
            for (int e = tid; e < NR * N; e += nth) {
                const int row = e % NR, i = e / NR;
                const int r = bx * NR + row;
                double2 z = make_double2(0.0, 0.0);
                if (r < a.nrows) {
                    if (!a.inverse) {
                        z.x = a.f[a.lay.idx(1 + i, r % a.ny + 1, r / a.ny + 1)];
                    } else {
                        // the half spectrum FFTW's c2r reads: X[k], k <= nx/2; the rest is its Hermitian image
                        const double2* X = a.C + (size_t)a.PC * r;
                        if (i <= N / 2) z = X[i];
                        else { z = X[N - i]; z.y = -z.y; }
                    }
                }
                s0[e] = z;
            }
            return;
        }
        if (any_stage_phase(a.P, ph, 1, 0, s0, s1, NR, tid, nth, a.tw, a.inverse ? +1 : -1)) return;
        const double2* res = (nst & 1) ? s1 : s0;
        for (int e = tid; e < NR * N; e += nth) {
            const int row = e % NR, i = e / NR;
            const int r = bx * NR + row;
            if (r >= a.nrows) continue;
            const double2 v = res[e];
            if (!a.inverse) {
                if (i <= N / 2) a.C[(size_t)a.PC * r + i] = make_double2(v.x * a.scale, v.y * a.scale);
            } else {
                double* frow = a.f + a.lay.idx(1, r % a.ny + 1, r / a.ny + 1);
                const double val = v.x * a.scale;
                frow[i] = val;
                if (i == 0) frow[N] = val;            // periodic x ghosts of the row (scalar.f90:257,276), as the
                if (i == N - 1) frow[-1] = val;       // power-of-two c2r pass writes them
            }
        }
    }
};
struct AnyRowsDct {                       // REDFT10 (inverse = 0) and REDFT01 (inverse = 1) of the rows, full-width C
    typedef AnyRowsArgs Args;
    FEN_HD static int nphases(const Args& a) { return a.P.nst + 2; }
    FEN_HD static void phase(int ph, const Args& a, double2* s0, double2* s1, int bx, int, int tid, int nth) {
        const int N = a.P.L, NR = a.NR, nst = a.P.nst;
        if (ph == 0) {
            for (int e = tid; e < NR * N; e += nth) {
                const int row = e % NR, i = e / NR;
                const int r = bx * NR + row;
                double2 z = make_double2(0.0, 0.0);
                int dst = a.inverse ? i : any_dct_perm(i, N);
                if (r < a.nrows) {
                    if (!a.inverse) {
                        z.x = a.f[a.lay.idx(1 + i, r % a.ny + 1, r / a.ny + 1)];
                    } else {
                        const double2* X = a.C + (size_t)a.PC * r;
                        const double xk = X[i].x, xm = i > 0 ? X[N - i].x : 0.0;
                        const double2 q = a.twq[i];
                        z = cmul(make_double2(xk, -xm), make_double2(q.x, -q.y));
                    }
                }
                s0[dst * NR + row] = z;
            }
            return;
        }
        if (any_stage_phase(a.P, ph, 1, 0, s0, s1, NR, tid, nth, a.tw, a.inverse ? +1 : -1)) return;
        const double2* res = (nst & 1) ? s1 : s0;
        for (int e = tid; e < NR * N; e += nth) {
            const int row = e % NR, i = e / NR;
            const int r = bx * NR + row;
            if (r >= a.nrows) continue;
            if (!a.inverse) {
                const double2 V = res[e], q = a.twq[i];
                a.C[(size_t)a.PC * r + i] = make_double2(2.0 * (V.x * q.x - V.y * q.y) * a.scale, 0.0);
            } else {
                a.f[a.lay.idx(1 + i, r % a.ny + 1, r / a.ny + 1)] = res[any_dct_perm(i, N) * NR + row].x * a.scale;
            }
        }
    }
};

// cosine transforms along a strided direction, in place on the real parts of C (nnn: the y direction)
struct AnyLinesDctArgs {
    AnyPlan P;
    int NL, inverse;
    double2* C;
    long long sl, so;
    const double2* tw;
    const double2* twq;
    double scale;
};
struct AnyLinesDct {
    typedef AnyLinesDctArgs Args;
    FEN_HD static int nphases(const Args& a) { return a.P.nst + 2; }
    FEN_HD static void phase(int ph, const Args& a, double2* s0, double2* s1, int bx, int by, int tid, int nth) {
        const int L = a.P.L, NL = a.NL, nst = a.P.nst;
        double2* base = a.C + (long long)bx * NL + a.so * by;
        if (ph == 0) {
            for (int e = tid; e < NL * L; e += nth) {
                const int line = e % NL, i = e / NL;
                if (!a.inverse) {
                    s0[any_dct_perm(i, L) * NL + line] = make_double2(base[line + a.sl * i].x, 0.0);
                } else {
                    const double xk = base[line + a.sl * i].x, xm = i > 0 ? base[line + a.sl * (L - i)].x : 0.0;
                    const double2 q = a.twq[i];
                    s0[e] = cmul(make_double2(xk, -xm), make_double2(q.x, -q.y));
                }
            }
            return;
        }
        if (any_stage_phase(a.P, ph, 1, 0, s0, s1, NL, tid, nth, a.tw, a.inverse ? +1 : -1)) return;
        const double2* res = (nst & 1) ? s1 : s0;
        for (int e = tid; e < NL * L; e += nth) {
            const int line = e % NL, i = e / NL;
            if (!a.inverse) {
                const double2 V = res[e], q = a.twq[i];
                base[line + a.sl * i] = make_double2(2.0 * (V.x * q.x - V.y * q.y) * a.scale, 0.0);
            } else {
                base[line + a.sl * i] = make_double2(res[any_dct_perm(i, L) * NL + line].x * a.scale, 0.0);
            }
        }
    }
};

// ---- the device wrapper: phases separated by block barriers ---------------------------------------------------------
#ifdef __CUDACC__
template <class K>
__global__ void __launch_bounds__(ANY_THREADS) k_any(const __grid_constant__ typename K::Args a, int half) {
    extern __shared__ double2 any_smem[];
    double2* s0 = any_smem;
    double2* s1 = any_smem + half;
    const int nph = K::nphases(a);
    for (int ph = 0; ph < nph; ++ph) {
        K::phase(ph, a, s0, s1, blockIdx.x, blockIdx.y, threadIdx.x, blockDim.x);
        __syncthreads();
    }
}
#endif

}  // namespace fen
