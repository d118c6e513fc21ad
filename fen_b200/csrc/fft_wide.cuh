// fft_wide.cuh -- 512-point transforms with THIRTY-TWO values per thread (radix 32 x 16): one shared-memory exchange
// per transform instead of two.
//
// Why: the fused z solve (forward c2c + spectral divide + inverse c2c of the last direction, poisson.f90:985-1012) runs two
// transforms per 64 KB tile.  With 8 values per thread (radix 8 x 8 x 8, fft_core.cuh) that is four exchanges = eight
// shared-memory passes per tile, and the kernel sits at 0.45 of HBM with the L1 / shared-memory data path 67 % busy, the
// fp64 pipe 35 %, DRAM 37 % and 32 warps per SM in lock-step through ~11 block-wide barriers (r01v capture).  With 32
// values per thread a 512-point line is 16 threads, a tile of 8 lines is 128 threads = 4 warps; the transform is a
// radix-32 stage (4 x 8, in registers), ONE exchange, and two radix-16 stages (4 x 4) per thread:
//   * shared-memory passes per tile: 4 instead of 8; barriers: 4 (of 4 warps) instead of ~11 (of 16 warps);
//   * every thread has 32 independent 16-byte loads in flight before the first butterfly (the whole 64 KB tile), and
//     three tiles are resident per SM (192 KB shared memory, 384 threads x ~168 registers).
// The arithmetic differs from the radix-8 path only in the order of the butterfly additions (both are exact DFT
// factorisations with table twiddles), i.e. by a few ulp; the parity bounds are the same.
#pragma once
#include "fft_core.cuh"

namespace fen {

// exp(DIR * 2 pi i m / N) for the small in-register twiddles
template <int DIR> FEN_HD double2 wconst(double c, double s) { return make_double2(c, DIR < 0 ? -s : s); }

// cos / sin(2 pi m / 32) as literals: the index is a compile-time constant after unrolling, so these fold away (a local
// constexpr table ended up on the stack)
FEN_HD constexpr double cos32(int m) {
    switch (m & 31) {
        case 0: return 1.0;
        case 1: case 31: return 0.98078528040323044913;
        case 2: case 30: return 0.92387953251128675613;
        case 3: case 29: return 0.83146961230254523708;
        case 4: case 28: return 0.70710678118654752440;
        case 5: case 27: return 0.55557023301960222474;
        case 6: case 26: return 0.38268343236508977173;
        case 7: case 25: return 0.19509032201612826785;
        case 8: case 24: return 0.0;
        case 9: case 23: return -0.19509032201612826785;
        case 10: case 22: return -0.38268343236508977173;
        case 11: case 21: return -0.55557023301960222474;
        case 12: case 20: return -0.70710678118654752440;
        case 13: case 19: return -0.83146961230254523708;
        case 14: case 18: return -0.92387953251128675613;
        case 15: case 17: return -0.98078528040323044913;
        default: return -1.0;      // 16
    }
}
FEN_HD constexpr double sin32(int m) { return cos32(m + 24); }      // sin(x) = cos(x - pi/2) = cos(2 pi (m - 8) / 32)

// 16-point DFT in registers, output in natural order: n = 4 n1 + n2, k = k1 + 4 k2
template <int DIR> FEN_HD void bfly16(double2 (&v)[16]) {
    // step 1: radix 4 over n1 for every n2  ->  A[n2][k1] left in v[4 k1 + n2]
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) bfly4<DIR>(v[n2], v[n2 + 4], v[n2 + 8], v[n2 + 12]);
    // step 2: twiddles W16^(n2 k1)
    const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;   // cos, sin(pi/8)
    const double h = 0.70710678118654752440;
    const double2 w1 = wconst<DIR>(c1, s1), w2 = wconst<DIR>(h, h), w3 = wconst<DIR>(s1, c1);
    const double2 w6 = wconst<DIR>(-h, h), w9 = wconst<DIR>(-c1, -s1);
    v[4 * 1 + 1] = cmul(v[4 * 1 + 1], w1); v[4 * 1 + 2] = cmul(v[4 * 1 + 2], w2); v[4 * 1 + 3] = cmul(v[4 * 1 + 3], w3);
    v[4 * 2 + 1] = cmul(v[4 * 2 + 1], w2); v[4 * 2 + 2] = rot90<DIR>(v[4 * 2 + 2]);  v[4 * 2 + 3] = cmul(v[4 * 2 + 3], w6);
    v[4 * 3 + 1] = cmul(v[4 * 3 + 1], w3); v[4 * 3 + 2] = cmul(v[4 * 3 + 2], w6); v[4 * 3 + 3] = cmul(v[4 * 3 + 3], w9);
    // step 3: radix 4 over n2 for every k1  ->  X[k1 + 4 k2] left in v[4 k1 + k2]
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) bfly4<DIR>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
    // natural order: X[k1 + 4 k2] -> w[k1 + 4 k2]
    double2 w[16];
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) w[k1 + 4 * k2] = v[4 * k1 + k2];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = w[q];
}

// 32-point DFT in registers, output in natural order: n = 8 n1 + n2 (n1 < 4, n2 < 8), k = k1 + 4 k2 (k1 < 4, k2 < 8)
template <int DIR> FEN_HD void bfly32(double2 (&v)[32]) {
    // step 1: radix 4 over n1 for every n2  ->  A[n2][k1] left in v[8 k1 + n2]
#pragma unroll
    for (int n2 = 0; n2 < 8; ++n2) bfly4<DIR>(v[n2], v[n2 + 8], v[n2 + 16], v[n2 + 24]);
    // step 2: twiddles W32^(n2 k1), n2 < 8, k1 < 4 (exponents 0 .. 21)
#pragma unroll
    for (int k1 = 1; k1 < 4; ++k1)
#pragma unroll
        for (int n2 = 1; n2 < 8; ++n2) {
            const int m = n2 * k1;
            if (m == 8) v[8 * k1 + n2] = rot90<DIR>(v[8 * k1 + n2]);
            else v[8 * k1 + n2] = cmul(v[8 * k1 + n2], wconst<DIR>(cos32(m), sin32(m)));
        }
    // step 3: radix 8 over n2 for every k1  ->  X[k1 + 4 k2] left in blk[k2]
    double2 w[32];
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
        double2 blk[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) blk[q] = v[8 * k1 + q];
        bfly8<DIR>(blk);
#pragma unroll
        for (int k2 = 0; k2 < 8; ++k2) w[k1 + 4 * k2] = blk[k2];
    }
#pragma unroll
    for (int q = 0; q < 32; ++q) v[q] = w[q];
}

// 512-point transform, 16 threads per line (t), 32 values per thread.  In: v[m] = x[t + 16 m].  Out: v[m] = X[t + 16 m].
// Shared memory s[idx * IS + line]; one exchange between the two stages (the functions are __host__ __device__ so that
// tests/cpu/test_fft_wide.cu runs the same code on the CPU, threads as loops, the barrier as the loop boundary).
// stage 1: radix 32, Ns = 1 (no twiddles); butterfly j = t writes X1[32 t + r]
template <int DIR> FEN_HD void wide512_stage1(double2 (&v)[32], double2* s, int IS, int line, int t) {
    bfly32<DIR>(v);
#pragma unroll
    for (int r = 0; r < 32; ++r) s[(32 * t + r) * IS + line] = v[r];
}
// stage 2: radix 16, Ns = 32; butterflies j = t + 16 b, b < 2: inputs j + 32 r, twiddles w^(j r), outputs j + 32 r
template <int DIR> FEN_HD void wide512_stage2(double2 (&v)[32], const double2* s, int IS, int line, int t,
                                              const double2* tw) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        const int j = t + 16 * b;
        double2 u[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) u[r] = s[(j + 32 * r) * IS + line];
        // w^r = exp(DIR 2 pi i j r / 512): w, w^2, w^4, w^8 from the table, the rest as products (fft_core.cuh: TWP)
#ifdef __CUDA_ARCH__
        const double2 c1 = twid<DIR>(__ldg(&tw[j])), c2 = twid<DIR>(__ldg(&tw[2 * j])), c4 = twid<DIR>(__ldg(&tw[4 * j])),
                      c8 = twid<DIR>(__ldg(&tw[8 * j]));
#else
        const double2 c1 = twid<DIR>(tw[j]), c2 = twid<DIR>(tw[2 * j]), c4 = twid<DIR>(tw[4 * j]), c8 = twid<DIR>(tw[8 * j]);
#endif
        const double2 c3 = cmul(c1, c2), c5 = cmul(c1, c4), c6 = cmul(c2, c4), c7 = cmul(c3, c4);
        u[1] = cmul(u[1], c1); u[2] = cmul(u[2], c2); u[3] = cmul(u[3], c3); u[4] = cmul(u[4], c4);
        u[5] = cmul(u[5], c5); u[6] = cmul(u[6], c6); u[7] = cmul(u[7], c7); u[8] = cmul(u[8], c8);
        u[9] = cmul(u[9], cmul(c1, c8)); u[10] = cmul(u[10], cmul(c2, c8)); u[11] = cmul(u[11], cmul(c3, c8));
        u[12] = cmul(u[12], cmul(c4, c8)); u[13] = cmul(u[13], cmul(c5, c8)); u[14] = cmul(u[14], cmul(c6, c8));
        u[15] = cmul(u[15], cmul(c7, c8));
        bfly16<DIR>(u);
        // output j + 32 r = t + 16 (b + 2 r): natural slot m = b + 2 r
#pragma unroll
        for (int r = 0; r < 16; ++r) v[b + 2 * r] = u[r];
    }
}

#ifdef __CUDACC__
template <int DIR, class Sync>
__device__ __forceinline__ void fft512_wide(double2 (&v)[32], double2* s, int IS, int line, int t, const double2* tw,
                                            Sync sync) {
    wide512_stage1<DIR>(v, s, IS, line, t);
    sync();
    wide512_stage2<DIR>(v, s, IS, line, t, tw);
    sync();           // the caller may overwrite the exchange buffer (the inverse transform's first stage)
}
#endif

}  // namespace fen
