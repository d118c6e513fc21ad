// fft_core.cuh -- shared-memory Stockham FFT building blocks (fp64 complex, power-of-two lengths).
//
// Replaces FFTW's 1-D plans executed line by line in the reference (src/poisson.f90:148-151,
// 635-674; executes :967,977,987,1010,1020,1030): unnormalised transforms, forward sign -1.
//
// A block transforms NL lines of length L that live in shared memory as s[idx * IS + line]
// (IS >= NL, chosen so that both "8 threads = 8 lines, same idx" and "8 threads = 8 consecutive
// idx, same line" are bank-conflict free: IS = 8 for strided lines, 9 for contiguous rows).
// Each line is worked on by T = max(L/8, 1) threads; every stage is an autosort (Stockham)
// radix-R pass:  read R inputs at stride L/R -> twiddle -> radix-R butterfly -> write at
// stride Ns.  Reads and writes of one stage are separated by a block barrier, so the pass is
// done in place.
//
// The functions are __host__ __device__ so that tests/cpu/test_fft_core.cu can run the same
// index logic on the CPU (threads emulated by loops, barriers by phase boundaries).
#pragma once
#include <cuda_runtime.h>

#ifndef FEN_HD
#define FEN_HD __host__ __device__ __forceinline__
#endif

// FEN_STRIDED_TWP (build-time A/B switch, default on): the strided y / z passes also form w^3, w^5, w^6, w^7 of a
// radix-8 stage as products of three table entries instead of loading seven (within ~2 ulp of the table values).
// Measured (profiles/r02g_twp.json, r02g_slab_twp.json): 512-point passes 0.38 -> 0.37 ms, 1024-point passes of the
// 8-GPU slab 0.59 / 0.61 / 0.58 -> 0.56 / 0.55 / 0.54 ms.
#ifndef FEN_STRIDED_TWP
#define FEN_STRIDED_TWP 1
#endif

namespace fen {

constexpr bool kStridedTwp = FEN_STRIDED_TWP != 0;

FEN_HD double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
FEN_HD double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
FEN_HD double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
FEN_HD double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
// multiply by -i (DIR = -1, forward) or +i (DIR = +1, inverse)
template <int DIR> FEN_HD double2 rot90(double2 a) {
    return DIR < 0 ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x);
}
template <int DIR> FEN_HD double2 twid(double2 w) { return DIR < 0 ? w : cconj(w); }

// ---- radix butterflies (in registers) ------------------------------------------------------
template <int DIR> FEN_HD void bfly2(double2& a, double2& b) {
    double2 t = a;
    a = cadd(t, b);
    b = csub(t, b);
}
template <int DIR> FEN_HD void bfly4(double2& a0, double2& a1, double2& a2, double2& a3) {
    // X[k] = sum_n a_n w^(nk), w = exp(DIR * 2 pi i / 4)
    double2 s02 = cadd(a0, a2), d02 = csub(a0, a2);
    double2 s13 = cadd(a1, a3), d13 = rot90<DIR>(csub(a1, a3));
    a0 = cadd(s02, s13);
    a2 = csub(s02, s13);
    a1 = cadd(d02, d13);
    a3 = csub(d02, d13);
}
template <int DIR> FEN_HD void bfly8(double2 (&v)[8]) {
    const double h = 0.70710678118654752440;
    // first layer: pairs (n, n+4)
    double2 a0 = cadd(v[0], v[4]), b0 = csub(v[0], v[4]);
    double2 a1 = cadd(v[1], v[5]), b1 = csub(v[1], v[5]);
    double2 a2 = cadd(v[2], v[6]), b2 = csub(v[2], v[6]);
    double2 a3 = cadd(v[3], v[7]), b3 = csub(v[3], v[7]);
    // twiddles w8^n on the odd half: w8 = exp(DIR * i pi/4)
    // b1 *= (1 + DIR*i) h ; b2 *= DIR*i ; b3 *= (-1 + DIR*i) h
    b1 = DIR < 0 ? make_double2((b1.x + b1.y) * h, (b1.y - b1.x) * h)
                 : make_double2((b1.x - b1.y) * h, (b1.y + b1.x) * h);
    b2 = rot90<DIR>(b2);
    b3 = DIR < 0 ? make_double2((b3.y - b3.x) * h, -(b3.x + b3.y) * h)
                 : make_double2(-(b3.x + b3.y) * h, (b3.x - b3.y) * h);
    bfly4<DIR>(a0, a1, a2, a3);   // even outputs 0,2,4,6
    bfly4<DIR>(b0, b1, b2, b3);   // odd outputs 1,3,5,7
    v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
    v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
}
template <int R, int DIR> FEN_HD void bfly(double2 (&v)[R]) {
    if constexpr (R == 2) bfly2<DIR>(v[0], v[1]);
    if constexpr (R == 4) bfly4<DIR>(v[0], v[1], v[2], v[3]);
    if constexpr (R == 8) bfly8<DIR>(v);
}

// ---- plan: radices per length -----------------------------------------------------------------
template <int L> struct FftPlan {
    static constexpr int T = (L >= 8) ? L / 8 : 1;          // threads per line
    static constexpr int LOG2 = (L <= 1) ? 0 : 1 + FftPlan<L / 2>::LOG2;
    static constexpr int N8 = (L >= 8) ? LOG2 / 3 : 0;      // number of radix-8 stages
    static constexpr int REM = (L >= 8) ? (1 << (LOG2 - 3 * N8)) : L;   // trailing radix 1,2,4
};
template <> struct FftPlan<0> { static constexpr int LOG2 = 0; };

// Shared-memory position of element idx of line `line`.  Default: idx-major, s[idx * IS + line] (8 threads = 8
// lines of one idx -> one 128-byte wavefront).  PR ("padded rows"): line-major with one pad slot every 8 elements,
// s[line * IS + idx + idx / 8], for kernels whose warps own ONE line each (contiguous rows): the Stockham read
// (consecutive idx) and write (idx = 8 j + r, or consecutive) patterns of 8 neighbouring threads then both fall
// into 8 distinct 16-byte bank groups.
template <bool PR> FEN_HD int spos(int idx, int IS, int line) {
    return PR ? line * IS + idx + (idx >> 3) : idx * IS + line;
}

// One Stockham stage, split at the barrier: load phase then compute+store phase.
// v must hold 8 complex values (T threads x 8 = L elements).  For L < 8 only R = L values are used.
template <int L, int R, int DIR, bool PR = false>
FEN_HD void stage_load(double2* v, const double2* s, int IS, int line, int t) {
    constexpr int T = FftPlan<L>::T;
    constexpr int NB = (L / R) / T;     // butterflies per thread
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        int j = t + b * T;
#pragma unroll
        for (int r = 0; r < R; ++r) v[b * R + r] = s[spos<PR>(j + r * (L / R), IS, line)];
    }
}
// twiddle + butterfly of one stage, in place on v[b * R + r].
// TWP (radix 8 only): load w, w^2, w^4 and form w^3, w^5, w^6, w^7 as products (3 table loads instead of 7;
// the products are within ~2 ulp of the table values).
template <int L, int R, int DIR, bool TWP = false>
FEN_HD void stage_compute(double2* v, int t, int Ns, const double2* tw) {
    constexpr int T = FftPlan<L>::T;
    constexpr int NB = (L / R) / T;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        int j = t + b * T;
        int k = j & (Ns - 1);
        double2 w[R];
#pragma unroll
        for (int r = 0; r < R; ++r) w[r] = v[b * R + r];
        if (Ns > 1) {
            int step = k * (L / (Ns * R));          // twiddle index of r = 1 in the length-L table
            if constexpr (TWP && R == 8) {
#ifdef __CUDA_ARCH__
                const double2 c1 = twid<DIR>(__ldg(&tw[step])), c2 = twid<DIR>(__ldg(&tw[step * 2])),
                              c4 = twid<DIR>(__ldg(&tw[step * 4]));
#else
                const double2 c1 = twid<DIR>(tw[step]), c2 = twid<DIR>(tw[step * 2]), c4 = twid<DIR>(tw[step * 4]);
#endif
                const double2 c3 = cmul(c1, c2), c5 = cmul(c1, c4), c6 = cmul(c2, c4);
                const double2 c7 = cmul(c3, c4);
                w[1] = cmul(w[1], c1); w[2] = cmul(w[2], c2); w[3] = cmul(w[3], c3); w[4] = cmul(w[4], c4);
                w[5] = cmul(w[5], c5); w[6] = cmul(w[6], c6); w[7] = cmul(w[7], c7);
            } else {
#pragma unroll
                for (int r = 1; r < R; ++r) {
#ifdef __CUDA_ARCH__
                    double2 c = __ldg(&tw[step * r]);
#else
                    double2 c = tw[step * r];
#endif
                    w[r] = cmul(w[r], twid<DIR>(c));
                }
            }
        }
        bfly<R, DIR>(w);
#pragma unroll
        for (int r = 0; r < R; ++r) v[b * R + r] = w[r];
    }
}
// autosort scatter of one stage's results
template <int L, int R, bool PR = false>
FEN_HD void stage_write(const double2* v, double2* s, int IS, int line, int t, int Ns) {
    constexpr int T = FftPlan<L>::T;
    constexpr int NB = (L / R) / T;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        int j = t + b * T;
        int k = j & (Ns - 1);
        int j0 = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) s[spos<PR>(j0 + r * Ns, IS, line)] = v[b * R + r];
    }
}
template <int L, int R, int DIR>
FEN_HD void stage_store(double2* v, double2* s, int IS, int line, int t, int Ns, const double2* tw) {
    stage_compute<L, R, DIR>(v, t, Ns, tw);
    stage_write<L, R>(v, s, IS, line, t, Ns);
}
// After the LAST stage (Ns = L / R) the value in v[b * R + r] is output element t + (b + r * (8 / R)) * (L / 8)
// (for L >= 64, T = L / 8): bring the registers to "natural" order v[m] = X[t + m * L / 8], which is also the
// order a first radix-8 stage consumes -- so transforms chain through registers.
template <int R> FEN_HD void last_permute(double2* v) {
    if constexpr (R != 8) {
        double2 w[8];
#pragma unroll
        for (int b = 0; b < 8 / R; ++b)
#pragma unroll
            for (int r = 0; r < R; ++r) w[b + r * (8 / R)] = v[b * R + r];
#pragma unroll
        for (int m = 0; m < 8; ++m) v[m] = w[m];
    }
}

// Full in-place transform of the block's lines (device only: every thread of the block must call
// it, `active` false for threads that own no line so that barriers stay uniform).
#ifdef __CUDACC__
template <int L, int DIR>
__device__ __forceinline__ void fft_lines(double2* s, int IS, int line, int t, bool active,
                                          const double2* tw) {
    if constexpr (L <= 1) return;
    double2 v[8];
    int Ns = 1;
    if constexpr (L >= 8) {
#pragma unroll
        for (int st = 0; st < FftPlan<L>::N8; ++st) {
            if (active) stage_load<L, 8, DIR>(v, s, IS, line, t);
            __syncthreads();
            if (active) stage_store<L, 8, DIR>(v, s, IS, line, t, Ns, tw);
            __syncthreads();
            Ns *= 8;
        }
    }
    constexpr int REM = FftPlan<L>::REM;
    if constexpr (REM > 1) {
        if (active) stage_load<L, REM, DIR>(v, s, IS, line, t);
        __syncthreads();
        if (active) stage_store<L, REM, DIR>(v, s, IS, line, t, Ns, tw);
        __syncthreads();
    }
}

// Register-to-register transform for L >= 64 (T = L / 8 threads per line, 8 values per thread).
// In:  v[m] = x[t + m * L / 8].  Out (OUT_REG): v[m] = X[t + m * L / 8]; otherwise the result is left in shared
// memory (s[idx * IS + line], barrier done).  The first stage takes its inputs and the last stage leaves its outputs
// in registers, so a 512-point transform costs 4 shared-memory passes and 3 block barriers instead of 8 and 8.
// On return no thread still reads shared memory written before the call's last barrier, so the caller may start
// the next transform (e.g. the inverse of the fused solve) without another barrier.
// RSYNC: the barriers of the exchange only involve the T threads of one line -- valid when each line's shared-memory
// region is private to its T threads, which are whole warps or parts of one warp (thread = line * T + t, padded-row
// layout): __syncwarp when a line fits a warp, a named barrier (id 1 + line, T threads) otherwise.  The warps of a block
// then run through the transform independently instead of meeting at ~6 block-wide barriers.
template <int L, bool RSYNC> __device__ __forceinline__ void fft_sync(int line) {
    if constexpr (!RSYNC) {
        __syncthreads();
    } else if constexpr (FftPlan<L>::T <= 32) {
        __syncwarp();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(line + 1), "n"(FftPlan<L>::T) : "memory");
    }
}
template <int L, int DIR, bool OUT_REG, bool TWP = false, bool PR = false, bool RSYNC = false>
__device__ __forceinline__ void fft_regs(double2 (&v)[8], double2* s, int IS, int line, int t, const double2* tw) {
    static_assert(L >= 64, "fft_regs needs T = L / 8 >= 8 threads per line");
    constexpr int N8 = FftPlan<L>::N8, REM = FftPlan<L>::REM;
    constexpr int NST = N8 + (REM > 1 ? 1 : 0);
    int Ns = 1;
#pragma unroll
    for (int st = 0; st < N8; ++st) {
        if (st > 0) {
            stage_load<L, 8, DIR, PR>(v, s, IS, line, t);
            fft_sync<L, RSYNC>(line);
        }
        stage_compute<L, 8, DIR, TWP>(v, t, Ns, tw);
        if (OUT_REG && st == NST - 1) return;
        stage_write<L, 8, PR>(v, s, IS, line, t, Ns);
        fft_sync<L, RSYNC>(line);
        Ns *= 8;
    }
    if constexpr (REM > 1) {
        stage_load<L, REM, DIR, PR>(v, s, IS, line, t);
        fft_sync<L, RSYNC>(line);
        stage_compute<L, REM, DIR>(v, t, Ns, tw);
        if (OUT_REG) {
            last_permute<REM>(v);
        } else {
            stage_write<L, REM, PR>(v, s, IS, line, t, Ns);
            fft_sync<L, RSYNC>(line);
        }
    }
}
#endif

}  // namespace fen
