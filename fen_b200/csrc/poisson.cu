// poisson.cu -- FEN's fast direct Poisson solver (src/poisson.f90) as hand-written sm_100a kernels.
//
// Reference pipeline (ppp, poisson.f90:941-1034): FFTW r2c along x line by line -> 2decomp
// transpose -> c2c along y -> transpose -> c2c along z -> scale, divide by the modified
// wavenumbers -> inverse chain.  ppn (:1038-1173) replaces the z transforms by a Thomas solve and
// removes the mean.  2-D variants pp (:416-505) / pn (:306-412) are the same with nz = 1.
//
// Here the spectral field lives in ONE half-spectrum work array
//     C[kx + PC * (j + ny * k)],  kx in [0, nx/2],  PC = roundup(nx/2 + 1, 8)  (128-byte rows)
// and every pass reads it once and writes it once:
//     k_fft_x_r2c     rows of the real field -> C          (x contiguous; 8 rows per block)
//     k_fft_lines     c2c along y or z, in place           (strided lines; 8 kx = 128 B per block,
//                                                            so the 2decomp x<->y / y<->z transposes
//                                                            of the reference disappear on one GPU)
//     k_fft_solve     forward c2c + spectral divide + inverse c2c along the last direction, fused
//     k_thomas_fwd/bwd  Thomas algorithm along the last direction, one thread per (kx, j) system,
//                     with the reference's operation order and no FMA contraction so that the
//                     exactly-zero pivot of the singular mode (hazard H5) is preserved
//     k_fft_x_c2r     C -> rows of the real field
// cuFFT is not used here: it is the cross-check (tests/test_gpu_parity.py::test_poisson_ppp_matches_cufft for
// correctness, bench.py's extra.cufft_crosscheck for speed: 3.85 ms against 2.30 ms at 512^3).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "fen_internal.cuh"
#include "fft_core.cuh"
#include "fft_any.cuh"
#include "slab_bulk.cuh"

namespace fen {

struct Poisson {
    char variant[4] = {0, 0, 0, 0};
    int nx = 0, ny = 0, nz = 0;     // global sizes
    int nzl = 0;                    // local z planes (slab)
    int nyl = 0;                    // local y lines in the z-pencil layout (multi-rank)
    int M = 0;                      // nx / 2
    int PC = 0;                     // complex row pitch
    double2* C = nullptr;           // [PC][ny][nzl]
    double2* Cz = nullptr;          // [PC][nyl][nz] (multi-rank only; == C on one rank)
    double2 *tw_x = nullptr, *twr_x = nullptr, *tw_y = nullptr, *tw_z = nullptr;
    double2* tw_xa = nullptr;       // exp(-2 pi i n / nx), n < nx: any-length path of a periodic x direction (fft_any.cuh)
    double2 *twq_x = nullptr, *twq_y = nullptr;   // exp(-i pi k / (2n)): DCT half-sample phases (Neumann directions)
    double *mwn_x = nullptr, *mwn_y = nullptr, *mwn_z = nullptr;
    double *ta = nullptr, *tb = nullptr, *tc = nullptr;   // tridiagonal a, b, c
    double* c1 = nullptr;           // Thomas c1 table, indexed like the line layout
    int tri_n = 0;
    bool multi = false;             // C / Cz live in the comm arena and the transposes are peer stores
    // multi-rank ppp / ppn with 64..1024-point lines: granule-blocked transposed layouts + bulk stores (slab_bulk.cuh)
    bool blocked = false;
    double2* Cr = nullptr;          // blocked path: row-layout output of the y inverse (input of the x c2r pass) and,
                                    // for ppn, the [g][jl][k][8] staging of the back substitution; local memory
    // chunked overlap of the blocked path (solve_blocked): the link-bound transposing kernels run on `aux` (forward)
    // / the main stream (backward) in nchunk pieces while the HBM-bound x pass / y inverse of the neighbouring piece
    // runs on the other stream
    cudaStream_t aux = nullptr;
    cudaEvent_t ev_piece[FEN_MAX_CHUNKS] = {};
    cudaEvent_t ev_join = nullptr;
    int nchunk = 1;
    double2* peerC[FEN_MAX_RANKS] = {};
    double2* peerCz[FEN_MAX_RANKS] = {};
};

// Destination of the fused transposes (C3 in SURVEY.md: transpose_y_to_z / z_to_y, poisson.f90:982,1015):
// element `idx` of the line a block has just transformed belongs to rank idx / blk; it is stored straight
// into that rank's array (mapped peer memory, comm.cu) at the position the next stage reads it from.
struct ScArgs {
    double2* peer[FEN_MAX_RANKS];
    int sh, mask;          // blk = 1 << sh; sh < 0: blk is not a power of two (row-copy transposes of any-length grids)
    long long dsl, dso;    // strides of (idx % blk) and of the outer index in the destination
    int o0;                // global offset of this rank's outer index in the destination
    int blk;
};
__device__ __forceinline__ double2* sc_dst(const ScArgs& q, int kx, int idx, int outer) {
    const int r = q.sh >= 0 ? idx >> q.sh : idx / q.blk;
    const int e = q.sh >= 0 ? idx & q.mask : idx - r * q.blk;
    return q.peer[r] + kx + q.dsl * e + q.dso * (q.o0 + outer);
}
static void sc_block(ScArgs& q, int blk) {            // host: the rank that owns idx is idx / blk
    q.blk = blk;
    q.sh = -1; q.mask = 0;
    if (blk > 0 && (blk & (blk - 1)) == 0) { q.sh = 0; while ((1 << q.sh) < blk) ++q.sh; q.mask = blk - 1; }
}

static inline double f32(long long n) { return (double)(float)n; }   // Fortran float(n), hazard H1

// =================================================================================================
// x direction: real <-> half-spectrum, rows contiguous
// =================================================================================================
struct XArgs {
    Layout L;
    double* f;            // real field (device layout)
    double2* C;
    int PC, ny, nrows;    // nrows = ny * nzl
    const double2* tw;    // exp(-2 pi i m / M), m < M
    const double2* twr;   // exp(-2 pi i k / N), k <= M
    double scale;         // applied to the r2c output (1/float(nx) in ppn/pn, else 1)
    int r0;               // first row of this launch (rows r0 .. nrows-1; 0 unless a caller launches the rows in pieces)
};

// Poisson right-hand side computed on the fly by the x pass (navier_stokes.f90:111-121 fused into
// poisson.f90:965-969): rhs = ((u-u_im)*id + (v-v_jm)*id + (w-w_km)*id) * rho / dt, uniform rho.
struct DivArgs {
    const double* u; const double* v; const double* w;
    double idelta, rho0, dt;
};
// rhs of the two cells (2*idx+1, 2*idx+2) of the row whose first interior element is c0
__device__ __forceinline__ double2 div_pair(const DivArgs& dv, const Layout& L, long long c) {
    const double id = dv.idelta;
    const double2 u2 = *reinterpret_cast<const double2*>(dv.u + c);
    const double um = dv.u[c - 1];
    const double2 v2 = *reinterpret_cast<const double2*>(dv.v + c);
    const double2 vm = *reinterpret_cast<const double2*>(dv.v + c - L.sy);
    const double2 w2 = *reinterpret_cast<const double2*>(dv.w + c);
    const double2 wm = *reinterpret_cast<const double2*>(dv.w + c - L.sz);
    double d0 = (u2.x - um) * id + (v2.x - vm.x) * id;      // fields.f90:144-147
    d0 = d0 + (w2.x - wm.x) * id;
    double d1 = (u2.y - u2.x) * id + (v2.y - vm.y) * id;
    d1 = d1 + (w2.y - wm.y) * id;
    return make_double2(d0 * dv.rho0 / dv.dt, d1 * dv.rho0 / dv.dt);   // navier_stokes.f90:118
}

constexpr int XR = 8;     // rows per block
constexpr int XIS = 9;    // smem idx stride (see fft_core.cuh)

template <int M, bool DIV>
__global__ void __launch_bounds__(XR* FftPlan<M>::T) k_fft_x_r2c(XArgs a, DivArgs dv) {
    extern __shared__ double2 s[];
    constexpr int T = FftPlan<M>::T;
    constexpr int NT = XR * T;
    const int tid = threadIdx.x;
    const int row0 = a.r0 + blockIdx.x * XR;
    // load: thread e -> (row, idx), consecutive threads read consecutive 16-byte pairs of a row
    for (int e = tid; e < XR * M; e += NT) {
        const int row = e / M, idx = e - row * M;
        const int r = row0 + row;
        double2 z = make_double2(0.0, 0.0);
        if (r < a.nrows) {
            const int j = r % a.ny, k = r / a.ny;
            const long long c0 = a.L.idx(1, j + 1, k + 1);
            if (DIV) z = div_pair(dv, a.L, c0 + 2 * idx);
            else z = reinterpret_cast<const double2*>(a.f + c0)[idx];
        }
        s[idx * XIS + row] = z;
    }
    __syncthreads();
    fft_lines<M, -1>(s, XIS, tid % XR, tid / XR, true, a.tw);
    // post-process pairs (k, M-k) and write X[0..M]:  X[k] = (Zk + conj Zm)/2 + w_k (Zk - conj Zm)/(2i)
    auto pair_out = [&](int row, int k) {
        const int r = row0 + row;
        if (r >= a.nrows) return;
        const int km = M - k;
        const double2 zk = s[(k % M) * XIS + row];
        const double2 zm = s[(km % M) * XIS + row];
        double2* dst = a.C + (size_t)a.PC * r;
        {
            const double2 E = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));
            const double2 O = make_double2(0.5 * (zk.y + zm.y), -0.5 * (zk.x - zm.x));
            const double2 w = __ldg(&a.twr[k]);
            const double2 X = cadd(E, cmul(w, O));
            dst[k] = make_double2(X.x * a.scale, X.y * a.scale);
        }
        if (km != k) {
            const double2 E = make_double2(0.5 * (zm.x + zk.x), 0.5 * (zm.y - zk.y));
            const double2 O = make_double2(0.5 * (zm.y + zk.y), -0.5 * (zm.x - zk.x));
            const double2 w = __ldg(&a.twr[km]);
            const double2 X = cadd(E, cmul(w, O));
            dst[km] = make_double2(X.x * a.scale, X.y * a.scale);
        }
    };
    if constexpr (M >= 16) {
        // NT == M threads: thread -> pair k = tid % (M/2) of rows (tid / (M/2)) + 2q; fully unrolled, no divisions;
        // the self-paired middle element k = M/2 of row q is done by thread q
        const int k = tid % (M / 2), rg = tid / (M / 2);
#pragma unroll
        for (int q = 0; q < XR / 2; ++q) pair_out(rg + 2 * q, k);
        if (tid < XR) pair_out(tid, M / 2);
    } else {
        constexpr int NP = M / 2 + 1;
        for (int e = tid; e < XR * NP; e += NT) pair_out(e / NP, e % NP);
    }
}

template <int M>
__global__ void __launch_bounds__(XR* FftPlan<M>::T) k_fft_x_c2r(XArgs a) {
    extern __shared__ double2 s[];
    constexpr int T = FftPlan<M>::T;
    constexpr int NT = XR * T;
    const int tid = threadIdx.x;
    const int row0 = a.r0 + blockIdx.x * XR;
    for (int e = tid; e < XR * (M + 1); e += NT) {
        const int row = e / (M + 1), k = e - row * (M + 1);
        const int r = row0 + row;
        double2 x = make_double2(0.0, 0.0);
        if (r < a.nrows) {
            x = a.C[(size_t)a.PC * r + k];
            if (k == 0 || k == M) x.y = 0.0;     // c2r ignores the imaginary part of DC / Nyquist
        }
        s[k * XIS + row] = x;
    }
    __syncthreads();
    constexpr int NP = M / 2 + 1;
    for (int e = tid; e < XR * NP; e += NT) {
        const int row = e / NP, k = e - row * NP;
        const int km = M - k;
        const double2 xk = s[k * XIS + row];
        const double2 xm = s[km * XIS + row];
        // Z[k] = (Xk + conj Xm) + i conj(w_k) (Xk - conj Xm),  w_k = exp(-2 pi i k / N)
        {
            const double2 E = make_double2(xk.x + xm.x, xk.y - xm.y);
            const double2 D = make_double2(xk.x - xm.x, xk.y + xm.y);
            const double2 w = cconj(__ldg(&a.twr[k]));
            const double2 O = cmul(w, D);
            s[k * XIS + row] = make_double2(E.x - O.y, E.y + O.x);
        }
        if (km != k && km < M) {
            const double2 E = make_double2(xm.x + xk.x, xm.y - xk.y);
            const double2 D = make_double2(xm.x - xk.x, xm.y + xk.y);
            const double2 w = cconj(__ldg(&a.twr[km]));
            const double2 O = cmul(w, D);
            s[km * XIS + row] = make_double2(E.x - O.y, E.y + O.x);
        }
    }
    __syncthreads();
    fft_lines<M, +1>(s, XIS, tid % XR, tid / XR, true, a.tw);
    for (int e = tid; e < XR * M; e += NT) {
        const int row = e / M, idx = e - row * M;
        const int r = row0 + row;
        if (r >= a.nrows) continue;
        const int j = r % a.ny, k = r / a.ny;
        double* frow = a.f + a.L.idx(1, j + 1, k + 1);
        const double2 val = s[idx * XIS + row];
        reinterpret_cast<double2*>(frow)[idx] = val;
        // periodic x ghosts of the row (scalar.f90:257,276): every variant built here is periodic in x
        if (idx == 0) frow[2 * M] = val.x;
        if (idx == M - 1) frow[-1] = val.y;
    }
}

// =================================================================================================
// strided directions (y, z): NL consecutive kx per block, lines of length Lf at stride sl
// =================================================================================================
struct LArgs {
    double2* C;
    long long sl, so;      // line element stride, outer stride (in complex elements)
    int o0;                // global index of outer element 0 (for lam_o)
    const double2* tw;
    double scale;          // multiply after the forward transform (1/float(ny) in ppn)
    // spectral divide (k_fft_solve only): lam = (lx[kx] + lo[o]) + ll[l]   (poisson.f90:998)
    const double* lx; const double* lo; const double* ll;
    double norm;           // float(nx*ny*nz)  (poisson.f90:992)
    int cx0;               // first kx group of this launch (0 unless a caller splits the kx range)
};

template <int Lf, int DIR, int NL, bool SC>
__global__ void __launch_bounds__(NL* FftPlan<Lf>::T) k_fft_lines(LArgs a, ScArgs q) {
    extern __shared__ double2 s[];
    constexpr int T = FftPlan<Lf>::T;
    const int tid = threadIdx.x;
    const int line = tid % NL, t = tid / NL;
    const int kx = (blockIdx.x + a.cx0) * NL + line;
    double2* base = a.C + kx + a.so * blockIdx.y;
    for (int idx = t; idx < Lf; idx += T) s[idx * NL + line] = base[a.sl * idx];
    __syncthreads();
    fft_lines<Lf, DIR>(s, NL, line, t, true, a.tw);
    const double sc = a.scale;
    for (int idx = t; idx < Lf; idx += T) {
        double2 v = s[idx * NL + line];
        v = make_double2(v.x * sc, v.y * sc);
        if (SC) *sc_dst(q, kx, idx, blockIdx.y) = v;
        else base[a.sl * idx] = v;
    }
}

template <int Lf, int NL, bool SC>
__global__ void __launch_bounds__(NL* FftPlan<Lf>::T) k_fft_solve(LArgs a, ScArgs q) {
    extern __shared__ double2 s[];
    constexpr int T = FftPlan<Lf>::T;
    const int tid = threadIdx.x;
    const int line = tid % NL, t = tid / NL;
    const int kx = (blockIdx.x + a.cx0) * NL + line;
    double2* base = a.C + kx + a.so * blockIdx.y;
    for (int idx = t; idx < Lf; idx += T) s[idx * NL + line] = base[a.sl * idx];
    __syncthreads();
    fft_lines<Lf, -1>(s, NL, line, t, true, a.tw);
    double lxo = a.lx[kx];
    if (a.lo) lxo = lxo + a.lo[a.o0 + blockIdx.y];
    for (int idx = t; idx < Lf; idx += T) {
        const double lam = lxo + a.ll[idx];
        double2 v = s[idx * NL + line];
        if (lam == 0.0) {                      // poisson.f90:998-999
            v = make_double2(0.0, 0.0);
        } else {                               // :992 then :1001
            v.x = (v.x / a.norm) / lam;
            v.y = (v.y / a.norm) / lam;
        }
        s[idx * NL + line] = v;
    }
    __syncthreads();
    fft_lines<Lf, +1>(s, NL, line, t, true, a.tw);
    for (int idx = t; idx < Lf; idx += T) {
        if (SC) *sc_dst(q, kx, idx, blockIdx.y) = s[idx * NL + line];
        else base[a.sl * idx] = s[idx * NL + line];
    }
}

// =================================================================================================
// Register-path kernels (transform length >= 64): the first radix-8 stage takes its operands straight from
// global memory and the last one stores straight back (fft_core.cuh: fft_regs), so a pass costs 4 shared
// memory sweeps and 3 barriers instead of 8 and 8, and every thread has 8 independent 16-byte loads in
// flight before the first butterfly.
// =================================================================================================

// ---- x passes with one warp per row ("padded rows", fft_core.cuh: spos<true>) --------------------------------
// Thread (row, t) owns elements t + m*T of its row, so every global access of a warp is one contiguous 512-byte
// run AND the transform runs on the register path; the line-major padded shared-memory layout keeps the Stockham
// exchanges bank-conflict free (proved by tests/cpu/test_fft_core.cu) and private to the row's threads.
// Row-private r2c: thread (row, t) computes / loads the elements t + m T of its row (one 512-byte run per warp), the
// transform runs on the register path with row-level barriers (RSYNC) and twiddle products (TWP), and the pair post-pass is done by the row's own threads -- thread t takes the pairs k = t + m T, m < 4
// (k < M/2), thread 0 also the self-paired k = M/2 -- so that no block-wide barrier remains.  With one warp per row
// the stage twiddles of a warp are 32 different table entries per load instruction (the x passes are bound by the
// L1 / shared-memory data path: 91 % in the r01v capture of the c2r pass), hence the products.
template <int M, bool DIV, bool TWP>
__global__ void __launch_bounds__(XR* FftPlan<M>::T, (XR * FftPlan<M>::T <= 1024) ? 1024 / (XR * FftPlan<M>::T) : 1)
k_fft_x_r2c_v(XArgs a, DivArgs dv) {
    extern __shared__ double2 s[];
    constexpr int T = FftPlan<M>::T;
    constexpr int RS = M + M / 8 + 1;
    const int tid = threadIdx.x;
    const int row = tid / T, t = tid % T;
    const int r = a.r0 + blockIdx.x * XR + row;
    const bool valid = r < a.nrows;
    double2 v[8];
    {
        const int j = valid ? r % a.ny : 0, k = valid ? r / a.ny : 0;
        const long long c0 = a.L.idx(1, j + 1, k + 1);
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            if (DIV) v[m] = div_pair(dv, a.L, c0 + 2 * (t + m * T));
            else v[m] = reinterpret_cast<const double2*>(a.f + c0)[t + m * T];
            if (!valid) v[m] = make_double2(0.0, 0.0);
        }
    }
    fft_regs<M, -1, false, TWP, true, true>(v, s, RS, row, t, a.tw);
    if (!valid) return;
    // pairs (k, M-k):  X[k] = (Zk + conj Zm)/2 + w_k (Zk - conj Zm)/(2i)
    double2* dst = a.C + (size_t)a.PC * r;
    auto pair_out = [&](int k) {
        const int km = M - k;
        const double2 zk = s[spos<true>(k % M, RS, row)];
        const double2 zm = s[spos<true>(km % M, RS, row)];
        {
            const double2 E = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));
            const double2 O = make_double2(0.5 * (zk.y + zm.y), -0.5 * (zk.x - zm.x));
            const double2 X = cadd(E, cmul(__ldg(&a.twr[k]), O));
            dst[k] = make_double2(X.x * a.scale, X.y * a.scale);
        }
        if (km != k) {
            const double2 E = make_double2(0.5 * (zm.x + zk.x), 0.5 * (zm.y - zk.y));
            const double2 O = make_double2(0.5 * (zm.y + zk.y), -0.5 * (zm.x - zk.x));
            const double2 X = cadd(E, cmul(__ldg(&a.twr[km]), O));
            dst[km] = make_double2(X.x * a.scale, X.y * a.scale);
        }
    };
#pragma unroll
    for (int m = 0; m < 4; ++m) pair_out(t + m * T);
    if (t == 0) pair_out(M / 2);
}

// Row-private c2r WITHOUT staging: the T threads of a row read X[k] (ascending) and X[M-k] (descending) of their own
// row straight from global memory -- both are whole 512-byte runs per warp, the second hits the lines the first just
// brought into L1 -- so the pair pre-pass needs no shared-memory pass and no barrier at all; the transform then runs
// with row-level barriers (RSYNC).  One shared-memory store and two loads per element fewer than a staged pre-pass.
template <int M, bool TWP>
__global__ void __launch_bounds__(XR* FftPlan<M>::T, (XR * FftPlan<M>::T <= 1024) ? 1024 / (XR * FftPlan<M>::T) : 1)
k_fft_x_c2r_d(XArgs a) {
    extern __shared__ double2 s[];
    constexpr int T = FftPlan<M>::T;
    constexpr int RS = M + M / 8 + 1;
    const int tid = threadIdx.x;
    const int row = tid / T, t = tid % T;
    const int r = a.r0 + blockIdx.x * XR + row;
    const bool valid = r < a.nrows;
    const double2* X = a.C + (size_t)a.PC * (valid ? r : a.r0);
    double2 v[8];
    {
        double2 xk[8], xm[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            xk[m] = X[t + m * T];
            xm[m] = X[M - (t + m * T)];
        }
        if (t == 0) { xk[0].y = 0.0; xm[0].y = 0.0; }            // c2r ignores the imaginary part of DC / Nyquist
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            // Z[k] = (Xk + conj Xm) + i conj(w_k) (Xk - conj Xm),  m = M - k, w_k = exp(-2 pi i k / N)
            const double2 E = make_double2(xk[m].x + xm[m].x, xk[m].y - xm[m].y);
            const double2 D = make_double2(xk[m].x - xm[m].x, xk[m].y + xm[m].y);
            const double2 O = cmul(cconj(__ldg(&a.twr[t + m * T])), D);
            v[m] = make_double2(E.x - O.y, E.y + O.x);
        }
    }
    fft_regs<M, +1, true, TWP, true, true>(v, s, RS, row, t, a.tw);
    if (!valid) return;
    const int j = r % a.ny, k3 = r / a.ny;
    double* frow = a.f + a.L.idx(1, j + 1, k3 + 1);
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int idx = t + m * T;
        reinterpret_cast<double2*>(frow)[idx] = v[m];
        if (idx == 0) frow[2 * M] = v[m].x;       // periodic x ghosts (scalar.f90:257,276)
        if (idx == M - 1) frow[-1] = v[m].y;
    }
}

template <int Lf, int DIR, int NL, bool SC>
__global__ void __launch_bounds__(NL* FftPlan<Lf>::T, (NL * FftPlan<Lf>::T <= 1024) ? 1024 / (NL * FftPlan<Lf>::T) : 1) k_fft_lines_r(LArgs a, ScArgs q) {
    extern __shared__ double2 s[];
    constexpr int T = FftPlan<Lf>::T;
    const int tid = threadIdx.x;
    const int line = tid % NL, t = tid / NL;
    const int kx = (blockIdx.x + a.cx0) * NL + line;
    double2* base = a.C + kx + a.so * blockIdx.y;
    double2 v[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = base[a.sl * (t + m * T)];
    fft_regs<Lf, DIR, true, kStridedTwp>(v, s, NL, line, t, a.tw);
    const double sc = a.scale;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int idx = t + m * T;
        const double2 o = make_double2(v[m].x * sc, v[m].y * sc);
        if (SC) *sc_dst(q, kx, idx, blockIdx.y) = o;
        else base[a.sl * idx] = o;
    }
}

template <int Lf, int NL, bool SC, bool TWP = kStridedTwp>
__global__ void __launch_bounds__(NL* FftPlan<Lf>::T, (NL * FftPlan<Lf>::T <= 1024) ? 1024 / (NL * FftPlan<Lf>::T) : 1) k_fft_solve_r(LArgs a, ScArgs q) {
    extern __shared__ double2 s[];
    constexpr int T = FftPlan<Lf>::T;
    const int tid = threadIdx.x;
    const int line = tid % NL, t = tid / NL;
    double2 v[8];
    {
        const double2* base = a.C + ((blockIdx.x + a.cx0) * NL + line) + a.so * blockIdx.y;
#pragma unroll
        for (int m = 0; m < 8; ++m) v[m] = base[a.sl * (t + m * T)];
    }
    fft_regs<Lf, -1, true, TWP>(v, s, NL, line, t, a.tw);
    // poisson.f90:992 then :998-1001.  norm = float(nx*ny*nz) is a power of two here (power-of-two transform
    // lengths), so x/norm == x*(1/norm) exactly; the division by lambda is one rounded reciprocal and a
    // multiply (<= 1 ulp from x/lambda, far inside the 1e-12 parity bound) instead of four fp64 divisions.
    {
        const int kx = (blockIdx.x + a.cx0) * NL + line;
        double lxo = __ldg(&a.lx[kx]);
        if (a.lo) lxo = lxo + __ldg(&a.lo[a.o0 + blockIdx.y]);
        const double inorm = 1.0 / a.norm;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const double lam = lxo + __ldg(&a.ll[t + m * T]);
            const double rl = lam == 0.0 ? 0.0 : inorm / lam;
            v[m].x *= rl;
            v[m].y *= rl;
        }
    }
    fft_regs<Lf, +1, true, TWP>(v, s, NL, line, t, a.tw);
    const int kx = (blockIdx.x + a.cx0) * NL + line;
    double2* base = a.C + kx + a.so * blockIdx.y;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const int idx = t + m * T;
        if (SC) *sc_dst(q, kx, idx, blockIdx.y) = v[m];
        else base[a.sl * idx] = v[m];
    }
}

// =================================================================================================
// Neumann directions: FFTW REDFT10 / REDFT01 (DCT-II / DCT-III, poisson.f90:272-275, :801-804, :888-909)
// through one complex transform of the same length (Makhoul's reordering):
//   DCT-II :  v[n] = x[2n], v[N-1-n] = x[2n+1];  V = FFT(v);  Y[k] = 2 Re(exp(-i pi k / 2N) V[k])
//   DCT-III:  W[k] = (X[k] - i X[N-k]) exp(+i pi k / 2N), X[N] = 0;  v = FFT^-1(W) (unnormalised);
//             y[2n] = Re v[n], y[2n+1] = Re v[N-1-n]
// (checked against scipy.fft.dct types 2 / 3 = FFTW's definitions in tests/).  The data of these variants is
// real; it is carried in the real part of a full-width complex work array C[i + PC*(j + ny*k)], PC =
// roundup(nx, 8), so that the y transforms and the Thomas kernels are the same ones the periodic variants use.
// =================================================================================================
struct DArgs {
    Layout L;
    double* f;
    double2* C;
    int PC, ny, nrows;
    const double2* tw;     // exp(-2 pi i m / N)
    const double2* twq;    // exp(-i pi k / (2N))
    double scale;
};
__device__ __forceinline__ int dct_perm(int i, int N) { return (i & 1) ? N - 1 - (i >> 1) : (i >> 1); }

template <int N, int DIR>
__global__ void __launch_bounds__(XR* FftPlan<N>::T) k_dct_x(DArgs a) {
    extern __shared__ double2 s[];
    constexpr int T = FftPlan<N>::T;
    constexpr int NT = XR * T;
    const int tid = threadIdx.x;
    const int row0 = blockIdx.x * XR;
    for (int e = tid; e < XR * N; e += NT) {
        const int row = e / N, i = e - row * N;
        const int r = row0 + row;
        double2 z = make_double2(0.0, 0.0);
        if (r < a.nrows) {
            if (DIR < 0) {
                const int j = r % a.ny, k = r / a.ny;
                z.x = a.f[a.L.idx(1 + i, j + 1, k + 1)];
                s[dct_perm(i, N) * XIS + row] = z;
                continue;
            }
            const double2* X = a.C + (size_t)a.PC * r;
            const double xk = X[i].x, xm = i > 0 ? X[N - i].x : 0.0;
            z = cmul(make_double2(xk, -xm), cconj(__ldg(&a.twq[i])));
        }
        s[(DIR < 0 ? dct_perm(i, N) : i) * XIS + row] = z;
    }
    __syncthreads();
    fft_lines<N, DIR>(s, XIS, tid % XR, tid / XR, true, a.tw);
    for (int e = tid; e < XR * N; e += NT) {
        const int row = e / N, i = e - row * N;
        const int r = row0 + row;
        if (r >= a.nrows) continue;
        if (DIR < 0) {
            const double2 V = s[i * XIS + row], q = __ldg(&a.twq[i]);
            a.C[(size_t)a.PC * r + i] = make_double2(2.0 * (V.x * q.x - V.y * q.y) * a.scale, 0.0);
        } else {
            const int j = r % a.ny, k = r / a.ny;
            a.f[a.L.idx(1 + i, j + 1, k + 1)] = s[dct_perm(i, N) * XIS + row].x * a.scale;
        }
    }
}

// DCT along a strided direction, in place on the real parts of C (nnn: y direction)
template <int Lf, int DIR, int NL>
__global__ void __launch_bounds__(NL* FftPlan<Lf>::T) k_dct_lines(LArgs a, const double2* twq) {
    extern __shared__ double2 s[];
    constexpr int T = FftPlan<Lf>::T;
    const int tid = threadIdx.x;
    const int line = tid % NL, t = tid / NL;
    const int kx = (blockIdx.x + a.cx0) * NL + line;
    double2* base = a.C + kx + a.so * blockIdx.y;
    for (int i = t; i < Lf; i += T) {
        if (DIR < 0) {
            s[dct_perm(i, Lf) * NL + line] = make_double2(base[a.sl * i].x, 0.0);
        } else {
            const double xk = base[a.sl * i].x, xm = i > 0 ? base[a.sl * (Lf - i)].x : 0.0;
            s[i * NL + line] = cmul(make_double2(xk, -xm), cconj(__ldg(&twq[i])));
        }
    }
    __syncthreads();
    fft_lines<Lf, DIR>(s, NL, line, t, true, a.tw);
    for (int i = t; i < Lf; i += T) {
        if (DIR < 0) {
            const double2 V = s[i * NL + line], q = __ldg(&twq[i]);
            base[a.sl * i] = make_double2(2.0 * (V.x * q.x - V.y * q.y) * a.scale, 0.0);
        } else {
            base[a.sl * i] = make_double2(s[dct_perm(i, Lf) * NL + line].x * a.scale, 0.0);
        }
    }
}

// =================================================================================================
// Thomas algorithm along the last direction (poisson.f90:1092-1135 3-D form, :346-385 2-D form)
// =================================================================================================
struct TArgs {
    double2* C;
    double* c1;            // table, same indexing as C
    long long sl, so;
    int n, npc, nouter, o0;
    const double* a; const double* b; const double* c;
    const double* lx; const double* lo;     // lo == nullptr in 2-D
    int form2d;
    int mean;              // subtract the mean of phi (pn, ppn): see k_thomas_bwd
    // blocked z-pencil layout (slab_bulk.cuh): thread e of granule g = blockIdx.y owns system kx = 8 g + e % 8,
    // jl = e / 8 at C + so * g + e (so = granule stride, sl = plane stride nyl * 8)
    int blocked = 0;
    int g0 = 0;            // first granule of this launch (blocked layout; chunked launches)
};
struct TSys {              // one thread's system
    long long off;         // offset of its first element in C / c1
    int kx, lo_idx;
    bool valid;
};
__device__ __forceinline__ TSys thomas_sys(const TArgs& g) {
    TSys s;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (!g.blocked) {
        s.kx = e; s.lo_idx = g.o0 + blockIdx.y;
        s.off = e + g.so * blockIdx.y;
        s.valid = e < g.npc;
    } else {
        s.kx = (g.g0 + blockIdx.y) * 8 + (e & 7); s.lo_idx = g.o0 + (e >> 3);
        s.off = g.so * (g.g0 + blockIdx.y) + e;
        s.valid = e < g.nouter * 8;
    }
    return s;
}

// pivot term shared by both sweeps: 3-D: ((b + lx) + lo) - a*c1prev ; 2-D uses its own groupings
__device__ __forceinline__ double piv3(double b, double lx, double lo, double a, double c1p) {
    return __dsub_rn(__dadd_rn(__dadd_rn(b, lx), lo), __dmul_rn(a, c1p));
}

// c1 table (depends only on the grid): computed once at init with the reference's arithmetic
__global__ void k_thomas_c1(TArgs g) {
    const TSys sy = thomas_sys(g);
    if (!sy.valid) return;
    double* c1 = g.c1 + sy.off;
    const double lx = g.lx[sy.kx];
    const double lo = g.lo ? g.lo[sy.lo_idx] : 0.0;
    double c1p;
    if (g.form2d) c1p = __ddiv_rn(g.c[0], __dadd_rn(g.b[0], lx));                       // :350
    else c1p = __dmul_rn(g.c[0], __ddiv_rn(1.0, __dadd_rn(__dadd_rn(g.b[0], lx), lo)));  // :1096-1097
    c1[0] = c1p;
    for (int l = 1; l < g.n - 1; ++l) {
        if (g.form2d)   // c(j)/(b(j) - a(j)*c1(j-1) + mwn_x(i))  :357
            c1p = __ddiv_rn(g.c[l], __dadd_rn(__dsub_rn(g.b[l], __dmul_rn(g.a[l], c1p)), lx));
        else            // :1105-1106
            c1p = __dmul_rn(g.c[l], __ddiv_rn(1.0, piv3(g.b[l], lx, lo, g.a[l], c1p)));
        c1[g.sl * l] = c1p;
    }
    if (g.n > 1) c1[g.sl * (g.n - 1)] = 0.0;
}

__device__ __forceinline__ double2 c_scale(double2 v, double f) {
    return make_double2(__dmul_rn(v.x, f), __dmul_rn(v.y, f));
}
__device__ __forceinline__ double2 c_div(double2 v, double f) {
    return make_double2(__ddiv_rn(v.x, f), __ddiv_rn(v.y, f));
}
__device__ __forceinline__ double2 c_sub_ad(double2 r, double a, double2 d) {   // r - a*d
    return make_double2(__dsub_rn(r.x, __dmul_rn(a, d.x)), __dsub_rn(r.y, __dmul_rn(a, d.y)));
}

__global__ void __launch_bounds__(128) k_thomas_fwd(TArgs g) {
    const TSys sy = thomas_sys(g);
    if (!sy.valid) return;
    double2* C = g.C + sy.off;
    const double* c1t = g.c1 + sy.off;
    const double lx = g.lx[sy.kx];
    const double lo = g.lo ? g.lo[sy.lo_idx] : 0.0;
    const int n = g.n;
    double2 d;
    {
        const double2 r = C[0];
        if (g.form2d) d = c_div(r, __dadd_rn(g.b[0], lx));                                    // :351
        else d = c_scale(r, __ddiv_rn(1.0, __dadd_rn(__dadd_rn(g.b[0], lx), lo)));            // :1098
        C[0] = d;
    }
    constexpr int U = 8;
    for (int l0 = 1; l0 < n - 1; l0 += U) {
        double2 r[U];
        double cp[U];
#pragma unroll
        for (int q = 0; q < U; ++q)
            if (l0 + q < n - 1) {
                r[q] = C[g.sl * (l0 + q)];
                cp[q] = c1t[g.sl * (l0 + q - 1)];
            }
#pragma unroll
        for (int q = 0; q < U; ++q) {
            const int l = l0 + q;
            if (l < n - 1) {
                const double a = g.a[l];
                if (g.form2d)   // (rhs - a*d1)/(b + mwn_x - a*c1)   :358
                    d = c_div(c_sub_ad(r[q], a, d),
                              __dsub_rn(__dadd_rn(g.b[l], lx), __dmul_rn(a, cp[q])));
                else            // (rhs - a*d1)*factor               :1105-1107
                    d = c_scale(c_sub_ad(r[q], a, d), __ddiv_rn(1.0, piv3(g.b[l], lx, lo, a, cp[q])));
                C[g.sl * l] = d;
            }
        }
    }
    if (n > 1) {   // last row, exact-zero pivot guard  :1112-1121 / :362-371
        const int l = n - 1;
        const double a = g.a[l];
        const double c1p = c1t[g.sl * (l - 1)];
        const double fr = g.form2d ? __dsub_rn(__dadd_rn(g.b[l], lx), __dmul_rn(a, c1p))
                                   : piv3(g.b[l], lx, lo, a, c1p);
        const double2 r = C[g.sl * l];
        if (fr != 0.0) d = c_div(c_sub_ad(r, a, d), fr);
        else d = make_double2(0.0, 0.0);
        C[g.sl * l] = d;
    }
}

// Back substitution.  Mean removal (poisson.f90:1159-1171, :398-410): the mean of phi over the domain
// equals the average along the last direction of the (kx, ky) = (0, 0) spectral line, so the one thread
// that owns that line subtracts it there -- O(n) work instead of two sweeps over the real field
// (hazard H4).
__global__ void __launch_bounds__(128) k_thomas_bwd(TArgs g) {
    const TSys sy = thomas_sys(g);
    if (!sy.valid) return;
    const double2* C = g.C + sy.off;
    const double* c1t = g.c1 + sy.off;
    double2* O = g.C + sy.off;
    const long long osl = g.sl;
    const int n = g.n;
    const bool mean_line = g.mean && sy.kx == 0 && sy.lo_idx == 0;
    double2 x = C[g.sl * (n - 1)];                                  // :1124-1128
    double acc = x.x;
    O[osl * (n - 1)] = x;
    constexpr int U = 8;
    for (int l0 = n - 2; l0 >= 0; l0 -= U) {
        double2 d[U];
        double cc[U];
#pragma unroll
        for (int r = 0; r < U; ++r)
            if (l0 - r >= 0) {
                d[r] = C[g.sl * (l0 - r)];
                cc[r] = c1t[g.sl * (l0 - r)];
            }
#pragma unroll
        for (int r = 0; r < U; ++r)
            if (l0 - r >= 0) {                                      // x = d1 - c1*x(k+1)  :1132
                x = make_double2(__dsub_rn(d[r].x, __dmul_rn(cc[r], x.x)),
                                 __dsub_rn(d[r].y, __dmul_rn(cc[r], x.y)));
                acc += x.x;
                O[osl * (l0 - r)] = x;
            }
    }
    if (mean_line) {
        const double mean = acc / (double)n;
        for (int l = 0; l < n; ++l) {
            double2 v = O[osl * l];
            v.x -= mean;
            O[osl * l] = v;
        }
    }
}

// ---- Thomas for FEW, LONG systems (2-D pn / nn: only nx/2+1 systems of ny unknowns) -------------------------------
// k_thomas_fwd/bwd give each system one thread; with ~1000 systems of 4096 unknowns that is 9 blocks whose threads
// wait a full DRAM round trip every 8 rows (measured: 2.6 + 1.3 ms at 2048 x 4096, the bulk of the two-phase step).
// Here a block is ONE warp owning 32 neighbouring kx (one 512-byte run of C and one 256-byte run of c1 per row), and
// every thread streams its own column through a private shared-memory ring with cp.async, LPR - LPG rows (~86 KB per
// warp) in flight, so that ~33 warps keep ~3 MB of loads outstanding.  Threads only read what they copied
// themselves, so cp.async.wait_group is the only synchronisation.  The dependent chain of the forward sweep is cut
// from (mul, sub, div) to (mul, sub, mul) by taking the reciprocal pivot from the c1 table: 1/den_l = c1_l / c_l.
// That rounds differently from the reference's division by at most 1-2 ulp per row; the last row -- the one whose
// pivot is exactly zero for the singular mode (hazard H5) -- keeps the reference's expression and its zero test.
constexpr int LPG = 16, LPS = 8, LPR = LPG * LPS;
struct LpSmem {
    double2 c[LPR][32];
    double k[LPR][32];
};

__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* sdst, const void* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// BWD false: forward elimination (rows 1 .. n-2 pipelined, rows 0 and n-1 as in k_thomas_fwd);
// BWD true : back substitution (rows n-2 .. 0) + mean removal on the (0, 0) line.
// A warp is alone on its SM, so nothing hides a stall: full groups of LPG rows take a branch-free path that first
// pulls the whole group from shared memory into registers, then runs the dependent chain (2 fp64 operations per row
// backward, 3 forward), then stores; only the last, partial group goes row by row.
template <bool BWD>
__global__ void __launch_bounds__(32) k_thomas_lp(TArgs g, double cinv, double amid) {
    extern __shared__ __align__(16) unsigned char lp_raw[];
    LpSmem& s = *reinterpret_cast<LpSmem*>(lp_raw);
    const int lane = threadIdx.x;
    const bool active = blockIdx.x * 32 + lane < g.npc;
    const int kx = active ? blockIdx.x * 32 + lane : g.npc - 1;     // idle lanes shadow the last column, never store
    double2* C = g.C + kx;
    const double* c1t = g.c1 + kx;
    const long long sl = g.sl;
    const int n = g.n;
    const int m = BWD ? n - 1 : n - 2;                               // rows that go through the ring
    const int ng = m > 0 ? (m + LPG - 1) / LPG : 0;
    const long long step = BWD ? -sl : sl;                           // row p lives at (first + p * step)
    const long long first = BWD ? sl * (n - 2) : sl;
    auto issue = [&](int gi) {
        if (gi < ng) {
            const int p0 = gi * LPG;
            const double2* gc = C + first + step * p0;
            const double* gk = c1t + first + step * p0;
            if (p0 + LPG <= m) {
#pragma unroll
                for (int q = 0; q < LPG; ++q) {
                    cp_async16(&s.c[(p0 + q) % LPR][lane], gc + step * q);
                    cp_async8(&s.k[(p0 + q) % LPR][lane], gk + step * q);
                }
            } else {
                for (int q = 0; p0 + q < m; ++q) {
                    cp_async16(&s.c[(p0 + q) % LPR][lane], gc + step * q);
                    cp_async8(&s.k[(p0 + q) % LPR][lane], gk + step * q);
                }
            }
        }
        cp_async_commit();
    };
    for (int gi = 0; gi < LPS - 1; ++gi) issue(gi);
    const double lx = g.lx[kx];
    double2 d;
    double acc = 0.0;
    if (!BWD) {
        d = c_div(C[0], __dadd_rn(g.b[0], lx));                      // poisson.f90:351
        if (active) C[0] = d;
    } else {
        d = C[sl * (n - 1)];                                          // :374 x(n) = d1(n)
        acc = d.x;
    }
    for (int gi = 0; gi < ng; ++gi) {
        issue(gi + LPS - 1);
        cp_async_wait<LPS - 1>();
        const int p0 = gi * LPG;
        double2* out = C + first + step * p0;
        if (p0 + LPG <= m) {
            double2 r[LPG];
            double k[LPG];
#pragma unroll
            for (int q = 0; q < LPG; ++q) {
                r[q] = s.c[(p0 + q) % LPR][lane];
                k[q] = s.k[(p0 + q) % LPR][lane];
                if (!BWD) k[q] = __dmul_rn(k[q], cinv);              // 1/den_j = c1_j / c_j, off the dependent chain
            }
#pragma unroll
            for (int q = 0; q < LPG; ++q) {
                if (!BWD) {       // d1(j) = (rhs - a d1(j-1)) / den_j                                     :358
                    d = c_scale(c_sub_ad(r[q], amid, d), k[q]);
                } else {          // x(j) = d1(j) - c1(j) x(j+1)                                           :376
                    d = make_double2(__dsub_rn(r[q].x, __dmul_rn(k[q], d.x)), __dsub_rn(r[q].y, __dmul_rn(k[q], d.y)));
                    acc += d.x;
                }
                r[q] = d;
            }
            if (active) {
#pragma unroll
                for (int q = 0; q < LPG; ++q) out[step * q] = r[q];
            }
        } else {
            for (int q = 0; p0 + q < m; ++q) {
                const double2 r = s.c[(p0 + q) % LPR][lane];
                const double k = s.k[(p0 + q) % LPR][lane];
                if (!BWD) {
                    d = c_scale(c_sub_ad(r, amid, d), __dmul_rn(k, cinv));
                } else {
                    d = make_double2(__dsub_rn(r.x, __dmul_rn(k, d.x)), __dsub_rn(r.y, __dmul_rn(k, d.y)));
                    acc += d.x;
                }
                if (active) out[step * q] = d;
            }
        }
    }
    cp_async_wait<0>();
    if (!BWD) {
        if (n > 1) {   // last row with the reference's own pivot and exact-zero guard, :362-371
            const int l = n - 1;
            const double a = g.a[l];
            const double fr = __dsub_rn(__dadd_rn(g.b[l], lx), __dmul_rn(a, c1t[sl * (l - 1)]));
            const double2 r = C[sl * l];
            if (fr != 0.0) d = c_div(c_sub_ad(r, a, d), fr);
            else d = make_double2(0.0, 0.0);
            if (active) C[sl * l] = d;
        }
    } else if (g.mean && blockIdx.x == 0) {
        // mean(phi) = average of the (0, 0) spectral line (see k_thomas_bwd); the whole warp subtracts it
        __syncwarp();
        const double mean = __shfl_sync(0xffffffffu, acc, 0) / (double)n;
        double* C0 = reinterpret_cast<double*>(g.C);
        for (int l0 = lane; l0 < n; l0 += 32 * 8) {
            double v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (l0 + 32 * q < n) v[q] = C0[2 * sl * (l0 + 32 * q)];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (l0 + 32 * q < n) C0[2 * sl * (l0 + 32 * q)] = v[q] - mean;
        }
    }
}

// z-pencil -> y-slab transpose (transpose_z_to_y, poisson.f90:1138) after the Thomas back substitution, as its own
// kernel.  The back substitution walks z from the top plane down in every system at once, so when its stores go
// straight to the owning ranks (k_thomas_bwd<true>) ALL ranks write to the SAME destination at any moment: the
// receiver's NVLink ingress is shared by P-1 senders (measured at 8 GPUs: 261 GB/s per GPU against 654 GB/s for the
// forward transpose, whose FFT epilogue spreads its stores over all destinations).  Here the substitution stays in
// local HBM and this kernel copies rows of PC complex values with consecutive blocks cycling over the destination
// ranks, starting one past the sender -- every link of the switch is busy all the time, for one extra local read of
// the spectral array (16 B/cell against the 8 x slower link).
__global__ void __launch_bounds__(256) k_a2a_scatter(const double2* __restrict__ src, long long s_idx,
                                                     long long s_outer, int PC, int P, int rank, int blk, ScArgs q) {
    // rows (idx, outer) of PC complex values; idx is the transposed index, owned by rank idx / blk.  The same kernel
    // does transpose_y_to_z (idx = j, blk = ny / P) for the variants whose y pass has no fused epilogue.
    const int bx = blockIdx.x;                       // 0 .. P * blk - 1
    const int dest = (rank + 1 + bx % P) % P;
    const int idx = dest * blk + bx / P;
    const int o = blockIdx.y;
    const double2* s = src + s_idx * idx + s_outer * o;
    double2* dst = sc_dst(q, 0, idx, o);
    for (int kx = threadIdx.x; kx < PC; kx += blockDim.x) dst[kx] = s[kx];
}

// =================================================================================================
// host side
// =================================================================================================
static bool pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }
static bool blocked_len(int n) { return pow2(n) && n >= 64 && n <= 1024; }   // line lengths of the blocked slab path

template <typename T> static int upload(T** dptr, const std::vector<T>& h) {
    FEN_CUDA(cudaMalloc(dptr, h.size() * sizeof(T)));
    FEN_CUDA(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return FEN_OK;
}

static std::vector<double2> twiddles(int n, int count, int denom) {
    // exp(-2 pi i m / denom), m < count, evaluated in long double
    std::vector<double2> t((size_t)std::max(count, 1));
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int m = 0; m < count; ++m) {
        long double ang = -two_pi * (long double)m / (long double)denom;
        t[m] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
    (void)n;
    return t;
}

static std::vector<double> mwn(int n, double delta, int pad) {
    // modified wavenumbers 2(cos(2 pi (i-1)/float(n)) - 1)/delta**2   (poisson.f90:627-629)
    const double pi = std::acos(-1.0);      // global.f90:14
    std::vector<double> m((size_t)std::max(n, pad), 1.0);   // padding entries: harmless non-zero
    for (int i = 0; i < n; ++i) m[i] = 2.0 * (std::cos(2.0 * pi * (double)i / f32(n)) - 1.0) / (delta * delta);
    return m;
}

static std::vector<double> mwn_neumann(int n, double delta, int pad) {
    // 2(cos(pi (i-1)/float(n)) - 1)/delta**2   (poisson.f90:265, :794, :881, :899)
    const double pi = std::acos(-1.0);
    std::vector<double> m((size_t)std::max(n, pad), 1.0);
    for (int i = 0; i < n; ++i) m[i] = 2.0 * (std::cos(1.0 * pi * (double)i / f32(n)) - 1.0) / (delta * delta);
    return m;
}

static std::vector<double2> half_phases(int n) {
    // exp(-i pi k / (2n)), k < n, evaluated in long double
    std::vector<double2> t((size_t)n);
    const long double pi = 3.141592653589793238462643383279502884L;
    for (int k = 0; k < n; ++k) {
        long double ang = -pi * (long double)k / (2.0L * (long double)n);
        t[k] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
    return t;
}

void poisson_destroy(fen_ctx* c) {
    Poisson* p = c->ps;
    if (!p) return;
    // a captured step bakes this object's device buffers into its kernel arguments: a replay after the buffers are
    // gone (and a new Poisson object at the same heap address) would run on freed memory
    step_graphs_clear(c);
    if (!p->multi) {
        if (p->Cz && p->Cz != p->C) cudaFree(p->Cz);
        if (p->C) cudaFree(p->C);
    }
    if (p->Cr) cudaFree(p->Cr);
    if (p->aux) { cudaStreamSynchronize(p->aux); cudaStreamDestroy(p->aux); }
    for (int q = 0; q < FEN_MAX_CHUNKS; ++q) if (p->ev_piece[q]) cudaEventDestroy(p->ev_piece[q]);
    if (p->ev_join) cudaEventDestroy(p->ev_join);
    for (void* q : {(void*)p->tw_x, (void*)p->twr_x, (void*)p->tw_y, (void*)p->tw_z, (void*)p->twq_x, (void*)p->twq_y,
                    (void*)p->mwn_x, (void*)p->mwn_y, (void*)p->mwn_z, (void*)p->ta, (void*)p->tb,
                    (void*)p->tc, (void*)p->c1, (void*)p->tw_xa})
        if (q) cudaFree(q);
    delete p;
    c->ps = nullptr;
}

const char* poisson_variant(fen_ctx* c) { return c->ps ? c->ps->variant : ""; }

// ---- any-length path (fft_any.cuh): lengths the tuned kernels do not cover -------------------------------------
// tuned kernels: powers of two up to `maxn`; everything else any_supported() accepts goes through k_any<...>
static bool tuned_len(int n, int maxn) { return pow2(n) && n <= maxn; }

template <class K> static int launch_any(fen_ctx* c, const char* name, const typename K::Args& a, dim3 grid, int L, int nl) {
    FEN_ONCE_PER_DEVICE(c) {
        FEN_CUDA(cudaFuncSetAttribute(k_any<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)any_smem_bytes(ANY_MAX_L, 1)));
    }
    FEN_LAUNCH(c, name, k_any<K><<<grid, ANY_THREADS, any_smem_bytes(L, nl), c->stream>>>(a, nl * L));
    FEN_CUDA(cudaGetLastError());
    return FEN_OK;
}
static AnyLayout any_layout(const Layout& L) {
    AnyLayout l;
    l.xoff = L.xoff; l.sy = L.sy; l.sz = L.sz;
    return l;
}
static int any_rows(fen_ctx* c, const char* name, bool dct, int nx, const Layout& L, double* f, double2* C, int PC, int ny,
                    int nrows, const double2* tw, const double2* twq, double scale, bool inverse) {
    AnyRowsArgs a;
    a.P = any_plan(nx);
    if (a.P.L != nx || !tw) return set_error(FEN_ERR_UNSUPPORTED, "x transform length %d not supported", nx);
    a.NR = any_lines_per_block(nx); a.inverse = inverse ? 1 : 0; a.lay = any_layout(L); a.f = f; a.C = C; a.PC = PC;
    a.ny = ny; a.nrows = nrows; a.tw = tw; a.twq = twq; a.scale = scale;
    const dim3 grid((nrows + a.NR - 1) / a.NR);
    return dct ? launch_any<AnyRowsDct>(c, name, a, grid, nx, a.NR) : launch_any<AnyRowsC>(c, name, a, grid, nx, a.NR);
}

template <int M> static int set_smem_x() {
    const int bytes = (M + 1) * XIS * (int)sizeof(double2);
    if (bytes > 48 * 1024) {
        FEN_CUDA(cudaFuncSetAttribute(k_fft_x_r2c<M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        FEN_CUDA(cudaFuncSetAttribute(k_fft_x_r2c<M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        FEN_CUDA(cudaFuncSetAttribute(k_fft_x_c2r<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        if constexpr (M >= 64) {
            FEN_CUDA(cudaFuncSetAttribute((k_fft_x_r2c_v<M, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            FEN_CUDA(cudaFuncSetAttribute((k_fft_x_r2c_v<M, true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            FEN_CUDA(cudaFuncSetAttribute((k_fft_x_c2r_d<M, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        }
    }
    return FEN_OK;
}

// The x passes, rows contiguous.  M >= 64 (x lengths >= 128): the row-private kernels -- one warp (two at M = 512) owns
// a row from its coalesced global loads to its coalesced stores, row-level barriers only, stage twiddles as products of
// three table entries.  History (512^3, one B200; profiles/r01f_variants.txt, r02e/f/g_*.json): the x passes are bound
// by the L1 / shared-memory data path (91 % busy in the r01v capture of the c2r pass), not by HBM or the barriers:
//   c2r  staged block-wide 0.72 ms -> warp per row, staged 0.60 -> row-level barriers 0.57 -> no staging (X[k] and
//        X[M-k] straight from global) 0.55 -> twiddle products 0.47 ms (0.71 of HBM);  M = 512: 0.67 -> 0.52 ms
//   r2c  staged block-wide 0.87 ms -> row-private + twiddle products 0.86;              M = 512: 0.97 -> 0.90 ms
// The register path with eight rows interleaved in a warp (64-byte gathers per row: 1.69 ms), a persistent prefetching
// c2r (0.63 ms) and the fully staged c2r were measured and deleted.
// FEN_X_R2C=1 / FEN_X_C2R=1 select the block-wide staged kernels (the ones every M < 64 uses) for cross-checks.
static int x_variant(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

template <int M> static int launch_x(fen_ctx* c, const XArgs& a, bool fwd, const DivArgs* dv) {
    constexpr int T = FftPlan<M>::T;
    const int bytes = (M + 1) * XIS * (int)sizeof(double2);
    FEN_ONCE_PER_DEVICE(c) FEN_TRY(set_smem_x<M>());
    static const int vr2c = x_variant("FEN_X_R2C", 6), vc2r = x_variant("FEN_X_C2R", 7);
    dim3 grid((a.nrows - a.r0 + XR - 1) / XR), block(XR * T);
    DivArgs none{};
    bool done = false;
    if constexpr (M >= 64) {
        if (!fwd && vc2r != 1) {
            FEN_LAUNCH(c, "fft_x_c2r", k_fft_x_c2r_d<M, true><<<grid, block, bytes, c->stream>>>(a));
            done = true;
        }
        if (fwd && vr2c != 1) {
            if (dv) FEN_LAUNCH(c, "fft_x_r2c_div", k_fft_x_r2c_v<M, true, true><<<grid, block, bytes, c->stream>>>(a, *dv));
            else FEN_LAUNCH(c, "fft_x_r2c", k_fft_x_r2c_v<M, false, true><<<grid, block, bytes, c->stream>>>(a, none));
            done = true;
        }
    }
    if (!done) {
        if (fwd && dv) FEN_LAUNCH(c, "fft_x_r2c_div", k_fft_x_r2c<M, true><<<grid, block, bytes, c->stream>>>(a, *dv));
        else if (fwd) FEN_LAUNCH(c, "fft_x_r2c", k_fft_x_r2c<M, false><<<grid, block, bytes, c->stream>>>(a, none));
        else FEN_LAUNCH(c, "fft_x_c2r", k_fft_x_c2r<M><<<grid, block, bytes, c->stream>>>(a));
    }
    FEN_CUDA(cudaGetLastError());
    return FEN_OK;
}

static int dispatch_x(fen_ctx* c, int M, const XArgs& a, bool fwd, const DivArgs* dv = nullptr) {
    if (!tuned_len(a.L.nx, 2048)) {
        if (dv) return set_error(FEN_ERR_STATE, "the fused right-hand side needs a power-of-two x length");
        return any_rows(c, fwd ? "fft_x_r2c_any" : "fft_x_c2r_any", false, a.L.nx, a.L, a.f, a.C, a.PC, a.ny, a.nrows,
                        c->ps->tw_xa, nullptr, a.scale, !fwd);
    }
    switch (M) {
#define FEN_CASE(m) case m: return launch_x<m>(c, a, fwd, dv);
        FEN_CASE(1) FEN_CASE(2) FEN_CASE(4) FEN_CASE(8) FEN_CASE(16) FEN_CASE(32) FEN_CASE(64)
        FEN_CASE(128) FEN_CASE(256) FEN_CASE(512) FEN_CASE(1024)
#undef FEN_CASE
    }
    return set_error(FEN_ERR_UNSUPPORTED, "x FFT length %d not supported (power of two, 2..2048)", 2 * M);
}

// mode 0: forward, 1: inverse, 2: fused solve; sc != nullptr: scatter the result to the owning ranks
template <int Lf, int NL> static int launch_lines(fen_ctx* c, const LArgs& a, int mode, int nchunks, int nouter,
                                                  const ScArgs* sc) {
    constexpr int T = FftPlan<Lf>::T;
    const int bytes = Lf * NL * (int)sizeof(double2);
    FEN_ONCE_PER_DEVICE(c) {
        if (bytes > 48 * 1024) {
            if constexpr (Lf >= 64) {
                FEN_CUDA(cudaFuncSetAttribute(k_fft_lines_r<Lf, -1, NL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
                FEN_CUDA(cudaFuncSetAttribute(k_fft_lines_r<Lf, -1, NL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
                FEN_CUDA(cudaFuncSetAttribute(k_fft_lines_r<Lf, +1, NL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
                FEN_CUDA(cudaFuncSetAttribute(k_fft_solve_r<Lf, NL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
                FEN_CUDA(cudaFuncSetAttribute(k_fft_solve_r<Lf, NL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            } else {
                FEN_CUDA(cudaFuncSetAttribute(k_fft_lines<Lf, -1, NL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
                FEN_CUDA(cudaFuncSetAttribute(k_fft_lines<Lf, -1, NL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
                FEN_CUDA(cudaFuncSetAttribute(k_fft_lines<Lf, +1, NL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
                FEN_CUDA(cudaFuncSetAttribute(k_fft_solve<Lf, NL, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
                FEN_CUDA(cudaFuncSetAttribute(k_fft_solve<Lf, NL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            }
        }
    }
    dim3 grid(nchunks, nouter), block(NL * T);
    ScArgs none;
    memset(&none, 0, sizeof(none));
    if constexpr (Lf >= 64) {
        if (mode == 0 && !sc) FEN_LAUNCH(c, "fft_lines_fwd", k_fft_lines_r<Lf, -1, NL, false><<<grid, block, bytes, c->stream>>>(a, none));
        if (mode == 0 && sc) FEN_LAUNCH(c, "fft_lines_fwd_a2a", k_fft_lines_r<Lf, -1, NL, true><<<grid, block, bytes, c->stream>>>(a, *sc));
        if (mode == 1) FEN_LAUNCH(c, "fft_lines_inv", k_fft_lines_r<Lf, +1, NL, false><<<grid, block, bytes, c->stream>>>(a, none));
        if (mode == 2 && !sc) FEN_LAUNCH(c, "fft_solve", k_fft_solve_r<Lf, NL, false><<<grid, block, bytes, c->stream>>>(a, none));
        if (mode == 2 && sc) FEN_LAUNCH(c, "fft_solve_a2a", k_fft_solve_r<Lf, NL, true><<<grid, block, bytes, c->stream>>>(a, *sc));
    } else {
        if (mode == 0 && !sc) FEN_LAUNCH(c, "fft_lines_fwd", k_fft_lines<Lf, -1, NL, false><<<grid, block, bytes, c->stream>>>(a, none));
        if (mode == 0 && sc) FEN_LAUNCH(c, "fft_lines_fwd_a2a", k_fft_lines<Lf, -1, NL, true><<<grid, block, bytes, c->stream>>>(a, *sc));
        if (mode == 1) FEN_LAUNCH(c, "fft_lines_inv", k_fft_lines<Lf, +1, NL, false><<<grid, block, bytes, c->stream>>>(a, none));
        if (mode == 2 && !sc) FEN_LAUNCH(c, "fft_solve", k_fft_solve<Lf, NL, false><<<grid, block, bytes, c->stream>>>(a, none));
        if (mode == 2 && sc) FEN_LAUNCH(c, "fft_solve_a2a", k_fft_solve<Lf, NL, true><<<grid, block, bytes, c->stream>>>(a, *sc));
    }
    FEN_CUDA(cudaGetLastError());
    return FEN_OK;
}

static int dispatch_lines(fen_ctx* c, int Lf, const LArgs& a, int mode, int PC, int nouter, const ScArgs* sc = nullptr) {
    if (!tuned_len(Lf, 2048)) {
        if (sc || a.cx0) return set_error(FEN_ERR_UNSUPPORTED, "any-length transforms (%d) run on one rank", Lf);
        AnyLinesArgs q;
        q.P = any_plan(Lf);
        if (q.P.L != Lf) return set_error(FEN_ERR_UNSUPPORTED, "FFT length %d not supported", Lf);
        q.NL = any_lines_per_block(Lf); q.mode = mode; q.C = a.C; q.sl = a.sl; q.so = a.so; q.o0 = a.o0; q.tw = a.tw;
        q.scale = a.scale; q.lx = a.lx; q.lo = a.lo; q.ll = a.ll; q.norm = a.norm;
        return launch_any<AnyLines>(c, mode == 0 ? "fft_lines_fwd_any" : (mode == 1 ? "fft_lines_inv_any" : "fft_solve_any"),
                                    q, dim3(PC / q.NL, nouter), Lf, q.NL);
    }
    // 1024-point lines: 8 columns per block would need 1024 threads and the whole register file (one block per SM);
    // 4 columns give two 512-thread blocks per SM -- measured 0.59 vs 0.62 ms on a 1024x1024x128 slab
    if (Lf == 1024) {
        LArgs b = a;
        b.cx0 = a.cx0 * 2;       // cx0 counts blocks of NL columns
        return launch_lines<1024, 4>(c, b, mode, PC / 4, nouter, sc);
    }
    switch (Lf) {
#define FEN_CASE(l) case l: return launch_lines<l, 8>(c, a, mode, PC / 8, nouter, sc);
        FEN_CASE(1) FEN_CASE(2) FEN_CASE(4) FEN_CASE(8) FEN_CASE(16) FEN_CASE(32) FEN_CASE(64)
        FEN_CASE(128) FEN_CASE(256) FEN_CASE(512)
#undef FEN_CASE
        case 2048: return launch_lines<2048, 4>(c, a, mode, PC / 4, nouter, sc);
    }
    return set_error(FEN_ERR_UNSUPPORTED, "FFT length %d not supported (power of two, 1..2048)", Lf);
}

template <int N> static int launch_dct_x(fen_ctx* c, const DArgs& a, bool fwd) {
    constexpr int T = FftPlan<N>::T;
    const int bytes = N * XIS * (int)sizeof(double2);
    FEN_ONCE_PER_DEVICE(c) {
        if (bytes > 48 * 1024) {
            FEN_CUDA(cudaFuncSetAttribute(k_dct_x<N, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            FEN_CUDA(cudaFuncSetAttribute(k_dct_x<N, +1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        }
    }
    dim3 grid((a.nrows + XR - 1) / XR), block(XR * T);
    if (fwd) FEN_LAUNCH(c, "dct_x_fwd", k_dct_x<N, -1><<<grid, block, bytes, c->stream>>>(a));
    else FEN_LAUNCH(c, "dct_x_inv", k_dct_x<N, +1><<<grid, block, bytes, c->stream>>>(a));
    FEN_CUDA(cudaGetLastError());
    return FEN_OK;
}
static int dispatch_dct_x(fen_ctx* c, int N, const DArgs& a, bool fwd) {
    if (!tuned_len(N, 1024))
        return any_rows(c, fwd ? "dct_x_fwd_any" : "dct_x_inv_any", true, N, a.L, a.f, a.C, a.PC, a.ny, a.nrows, a.tw,
                        a.twq, a.scale, !fwd);
    switch (N) {
#define FEN_CASE(m) case m: return launch_dct_x<m>(c, a, fwd);
        FEN_CASE(2) FEN_CASE(4) FEN_CASE(8) FEN_CASE(16) FEN_CASE(32) FEN_CASE(64) FEN_CASE(128) FEN_CASE(256)
        FEN_CASE(512) FEN_CASE(1024)
#undef FEN_CASE
    }
    return set_error(FEN_ERR_UNSUPPORTED, "DCT length %d not supported (power of two, 2..1024)", N);
}
template <int Lf> static int launch_dct_lines(fen_ctx* c, const LArgs& a, const double2* twq, bool fwd, int nchunks,
                                              int nouter) {
    constexpr int T = FftPlan<Lf>::T;
    const int bytes = Lf * 8 * (int)sizeof(double2);
    FEN_ONCE_PER_DEVICE(c) {
        if (bytes > 48 * 1024) {
            FEN_CUDA(cudaFuncSetAttribute(k_dct_lines<Lf, -1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            FEN_CUDA(cudaFuncSetAttribute(k_dct_lines<Lf, +1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        }
    }
    dim3 grid(nchunks, nouter), block(8 * T);
    if (fwd) FEN_LAUNCH(c, "dct_lines_fwd", k_dct_lines<Lf, -1, 8><<<grid, block, bytes, c->stream>>>(a, twq));
    else FEN_LAUNCH(c, "dct_lines_inv", k_dct_lines<Lf, +1, 8><<<grid, block, bytes, c->stream>>>(a, twq));
    FEN_CUDA(cudaGetLastError());
    return FEN_OK;
}
static int dispatch_dct_lines(fen_ctx* c, int Lf, const LArgs& a, const double2* twq, bool fwd, int PC, int nouter) {
    if (!tuned_len(Lf, 1024)) {
        AnyLinesDctArgs q;
        q.P = any_plan(Lf);
        if (q.P.L != Lf) return set_error(FEN_ERR_UNSUPPORTED, "DCT length %d not supported", Lf);
        q.NL = any_lines_per_block(Lf); q.inverse = fwd ? 0 : 1; q.C = a.C; q.sl = a.sl; q.so = a.so; q.tw = a.tw;
        q.twq = twq; q.scale = a.scale;
        return launch_any<AnyLinesDct>(c, fwd ? "dct_lines_fwd_any" : "dct_lines_inv_any", q, dim3(PC / q.NL, nouter), Lf,
                                       q.NL);
    }
    switch (Lf) {
#define FEN_CASE(l) case l: return launch_dct_lines<l>(c, a, twq, fwd, PC / 8, nouter);
        FEN_CASE(2) FEN_CASE(4) FEN_CASE(8) FEN_CASE(16) FEN_CASE(32) FEN_CASE(64) FEN_CASE(128) FEN_CASE(256)
        FEN_CASE(512) FEN_CASE(1024)
#undef FEN_CASE
    }
    return set_error(FEN_ERR_UNSUPPORTED, "DCT length %d not supported (power of two, 2..1024)", Lf);
}

static int log2i(int n) { int s = 0; while ((1 << s) < n) ++s; return s; }

static int poisson_build(fen_ctx* c);
int poisson_init(fen_ctx* c) {
    if (c->ps) poisson_destroy(c);
    step_graphs_clear(c);
    const int r = poisson_build(c);
    if (r != FEN_OK) poisson_destroy(c);      // never leave a half-built solver behind (null twiddles, C == nullptr)
    return r;
}
static int poisson_build(fen_ctx* c) {
    const fen_grid_desc& g = c->g;
    // variant from periodic_bc only (poisson.f90:68-111)
    bool per[3];
    per[0] = g.bc[0] == FEN_BC_PERIODIC && g.bc[1] == FEN_BC_PERIODIC;
    per[1] = g.bc[2] == FEN_BC_PERIODIC && g.bc[3] == FEN_BC_PERIODIC;
    per[2] = g.ndim == 3 ? (g.bc[4] == FEN_BC_PERIODIC && g.bc[5] == FEN_BC_PERIODIC) : true;
    const char* var = nullptr;
    if (g.ndim == 3) {
        if (per[0] && per[1] && per[2]) var = "ppp";
        else if (per[0] && per[1] && !per[2]) var = "ppn";
        else if (!per[0] && per[1] && !per[2]) var = "npn";
        else if (!per[0] && !per[1] && !per[2]) var = "nnn";
    } else {
        if (per[0] && per[1]) var = "pp";
        else if (per[0] && !per[1]) var = "pn";
        else if (!per[0] && !per[1]) var = "nn";
    }
    if (!var)
        return set_error(FEN_ERR_UNSUPPORTED, "Unable to find the proper poisson solver with the selected "
                                              "boundary conditions");          // poisson.f90:91-95
    const bool dctx = var[0] == 'n', dcty = g.ndim == 3 && var[1] == 'n';
    // the last direction of the *n variants is solved by the Thomas algorithm: any length (the reference has no
    // restriction either); FFT / DCT directions are powers of two up to 2048
    const bool thomas_last = var[g.ndim - 1] == 'n';
    const bool y_fft = !(g.ndim == 2 && thomas_last), z_fft = g.ndim == 3 && !thomas_last;
    // Transformed directions: powers of two up to 2048 (cosine transforms: 1024) run the tuned register-path kernels,
    // on any number of ranks; every other length whose prime factors are <= 61, up to 6144 points, runs the
    // any-length kernels of fft_any.cuh on one rank (the reference's FFTW plans take any n, poisson.f90:148-151)
    const bool multi_rank = g.nranks > 1 && g.ndim == 3;
    const bool tx = tuned_len(g.nx, dctx ? 1024 : 2048), ty = !y_fft || tuned_len(g.ny, dcty ? 1024 : 2048),
               tz = !z_fft || tuned_len(g.nz, 2048);
    if (g.nx < 2 || (!tx && !any_supported(g.nx)) || (!ty && !any_supported(g.ny)) || (!tz && !any_supported(g.nz)))
        return set_error(FEN_ERR_UNSUPPORTED, "transform sizes must be products of primes <= %d, 2..%d points "
                                              "(got %d %d %d)", ANY_MAX_RADIX, ANY_MAX_L, g.nx, g.ny, g.nz);
    // several ranks: the tuned power-of-two kernels ship their lines themselves (fused epilogues / bulk stores); any
    // other supported length runs its passes in place and the y <-> z transposes as staggered row copies
    // (k_a2a_scatter) -- periodic x only (the cosine-transform variants on slabs stay powers of two)
    if (multi_rank && !(tx && ty && tz) && dctx)
        return set_error(FEN_ERR_UNSUPPORTED, "on several ranks the cosine-transform variants need powers of two up to "
                                              "1024; got %d %d %d", g.nx, g.ny, g.nz);
    Poisson* p = new Poisson();
    c->ps = p;
    snprintf(p->variant, sizeof(p->variant), "%s", var);
    p->nx = g.nx; p->ny = g.ny; p->nz = g.nz; p->nzl = c->L.nzl;
    p->M = g.nx / 2;
    p->PC = spectral_pitch_grid(g);
    p->nyl = g.ny;
    if (g.nranks > 1 && g.ndim == 3) {
        // the spectral arrays live in the comm arena so that the peers can store into them
        if (g.ny % g.nranks || g.nz % g.nranks)
            return set_error(FEN_ERR_UNSUPPORTED, "slab transposes need ny and nz divisible by the number of ranks "
                                                  "(got %d ranks, ny %d, nz %d)", g.nranks, g.ny, g.nz);
        p->multi = true;                      // before any arena pointer is stored: poisson_destroy must not free them
        FEN_TRY(comm_spectral(c, p->peerC, p->peerCz));
        p->nyl = g.ny / g.nranks;
        p->C = p->peerC[g.rank];
        p->Cz = p->peerCz[g.rank];
        // ppp / ppn with 64..1024-point transform lines: blocked transposed layouts + bulk stores (slab_bulk.cuh);
        // FEN_SLAB_BULK=0 keeps the register-store epilogues (A/B switch)
        static const bool bulk_off = getenv("FEN_SLAB_BULK") && atoi(getenv("FEN_SLAB_BULK")) == 0;
        const bool is_ppp = !strcmp(var, "ppp"), is_ppn = !strcmp(var, "ppn");
        p->blocked = !bulk_off && (is_ppp || is_ppn) && blocked_len(g.ny) && (is_ppn || blocked_len(g.nz));
        if (p->blocked) {
            const size_t nC = (size_t)p->PC * g.ny * p->nzl;
            FEN_CUDA(cudaMalloc(&p->Cr, nC * sizeof(double2)));
            FEN_CUDA(cudaMemsetAsync(p->Cr, 0, nC * sizeof(double2), c->stream));
            // FEN_SLAB_CHUNKS = pieces of the overlapped transposes (1 = no overlap, one stream).  Default: 4 for ppp from
            // 4 ranks on (8 GPUs: 9.26 -> 9.07 ms/step, 4 GPUs: 8.81 -> 8.59, profiles/r02i_*, r02n_*), 1 otherwise: on 2
            // ranks the transposing kernels are as much HBM- as link-bound (7.82 vs 7.82 - 7.96 ms, r02h_*), and the
            // Thomas path of ppn loses with pieces (channel at 8 GPUs: 5.71 ms unchunked, 5.98 in 4 pieces, r02d / r02o)
            const char* e = getenv("FEN_SLAB_CHUNKS");
            p->nchunk = std::max(1, std::min(FEN_MAX_CHUNKS, e ? atoi(e) : (g.nranks >= 4 && is_ppp ? 4 : 1)));
            p->nchunk = std::min(p->nchunk, std::min(p->nzl, p->PC / 8));
            if (p->nchunk > 1) {
                int lo = 0, hi = 0;
                FEN_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
                // the link-bound pieces get the higher priority: their blocks mostly wait, the HBM-bound pass fills in
                FEN_CUDA(cudaStreamCreateWithPriority(&p->aux, cudaStreamNonBlocking, hi));
            }
            for (int q = 0; q < FEN_MAX_CHUNKS; ++q)
                FEN_CUDA(cudaEventCreateWithFlags(&p->ev_piece[q], cudaEventDisableTiming));
            FEN_CUDA(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
        }
    } else {
        const size_t nC = (size_t)p->PC * g.ny * p->nzl;
        FEN_CUDA(cudaMalloc(&p->C, nC * sizeof(double2)));
        FEN_CUDA(cudaMemsetAsync(p->C, 0, nC * sizeof(double2), c->stream));
        p->Cz = p->C;
    }
    const double d = g.delta;
    if (dctx) {
        FEN_TRY(upload(&p->tw_x, twiddles(g.nx, g.nx, g.nx)));
        FEN_TRY(upload(&p->twq_x, half_phases(g.nx)));
        FEN_TRY(upload(&p->mwn_x, mwn_neumann(g.nx, d, p->PC)));
    } else {
        FEN_TRY(upload(&p->tw_x, twiddles(p->M, p->M, p->M)));
        FEN_TRY(upload(&p->twr_x, twiddles(g.nx, p->M + 1, g.nx)));
        if (!tx) FEN_TRY(upload(&p->tw_xa, twiddles(g.nx, g.nx, g.nx)));
        FEN_TRY(upload(&p->mwn_x, mwn(g.nx, d, p->PC)));
    }
    const bool tri_y = !strcmp(var, "pn") || !strcmp(var, "nn");
    const bool tri_z = !strcmp(var, "ppn") || !strcmp(var, "npn") || !strcmp(var, "nnn");
    if (!tri_y) {
        FEN_TRY(upload(&p->tw_y, twiddles(g.ny, g.ny, g.ny)));
        if (dcty) {
            FEN_TRY(upload(&p->twq_y, half_phases(g.ny)));
            FEN_TRY(upload(&p->mwn_y, mwn_neumann(g.ny, d, 0)));
        } else {
            FEN_TRY(upload(&p->mwn_y, mwn(g.ny, d, 0)));
        }
    }
    if (!strcmp(var, "ppp")) {
        FEN_TRY(upload(&p->tw_z, twiddles(g.nz, g.nz, g.nz)));
        FEN_TRY(upload(&p->mwn_z, mwn(g.nz, d, 0)));
    }
    if (tri_y || tri_z) {
        // poisson.f90:219-232 / :744-757
        const int n = tri_y ? g.ny : g.nz;
        const int lo_face = tri_y ? 2 : 4;
        p->tri_n = n;
        std::vector<double> a((size_t)n, 1.0 / (d * d)), b((size_t)n, -2.0 / (d * d)), cc((size_t)n, 1.0 / (d * d));
        b[0] = b[0] + a[0];
        if (g.bc[lo_face] == FEN_BC_INFLOW && g.bc[lo_face + 1] == FEN_BC_OUTFLOW) b[n - 1] = b[n - 1] - cc[n - 1];
        else b[n - 1] = b[n - 1] + cc[n - 1];
        a[0] = 0.0;
        cc[n - 1] = 0.0;
        FEN_TRY(upload(&p->ta, a));
        FEN_TRY(upload(&p->tb, b));
        FEN_TRY(upload(&p->tc, cc));
        TArgs t;
        t.C = nullptr; t.mean = 0;
        t.n = n; t.npc = p->PC; t.a = p->ta; t.b = p->tb; t.c = p->tc; t.lx = p->mwn_x;
        if (tri_y) {
            t.sl = p->PC; t.so = 0; t.nouter = 1; t.o0 = 0; t.lo = nullptr; t.form2d = 1;
            FEN_CUDA(cudaMalloc(&p->c1, (size_t)p->PC * n * sizeof(double)));
        } else {
            t.sl = (long long)p->PC * p->nyl; t.so = p->PC; t.nouter = p->nyl; t.o0 = g.rank * p->nyl;
            t.lo = p->mwn_y; t.form2d = 0;
            FEN_CUDA(cudaMalloc(&p->c1, (size_t)p->PC * p->nyl * n * sizeof(double)));
            if (p->blocked) { t.blocked = 1; t.sl = (long long)p->nyl * 8; t.so = (long long)n * p->nyl * 8; }
        }
        t.c1 = p->c1;
        dim3 grid((p->PC + 127) / 128, t.nouter), block(128);
        if (p->blocked) grid = dim3((p->nyl * 8 + 127) / 128, p->PC / 8);
        FEN_LAUNCH(c, "thomas_c1", k_thomas_c1<<<grid, block, 0, c->stream>>>(t));
        FEN_CUDA(cudaGetLastError());
    }
    return FEN_OK;
}

// the x pass can compute the right-hand side div(v*) rho/dt itself (uniform rho, 3-D, register-path lengths)
bool poisson_can_fuse_rhs(fen_ctx* c) {
    static const bool off = getenv("FEN_NO_FUSED_RHS") != nullptr;     // tuning switch
    return !off && c->ps && c->ps->variant[0] == 'p' && c->g.ndim == 3 && c->uniform_props && tuned_len(c->g.nx, 2048);
}

// Thomas along y of a 2-D problem (pn, nn): few long systems -> the warp-per-32-columns streaming kernels.
// FEN_THOMAS_LP=0 selects the thread-per-system kernels (tuning / cross-check switch).
static int thomas_2d(fen_ctx* c, const TArgs& t) {
    static const bool lp = !getenv("FEN_THOMAS_LP") || atoi(getenv("FEN_THOMAS_LP")) != 0;
    if (!lp) {
        dim3 grid((t.npc + 127) / 128, 1), block(128);
        ScArgs none;
        memset(&none, 0, sizeof(none));
        FEN_LAUNCH(c, "thomas_fwd", k_thomas_fwd<<<grid, block, 0, c->stream>>>(t));
        FEN_LAUNCH(c, "thomas_bwd", k_thomas_bwd<<<grid, block, 0, c->stream>>>(t));
        FEN_CUDA(cudaGetLastError());
        return FEN_OK;
    }
    FEN_ONCE_PER_DEVICE(c) {
        FEN_CUDA(cudaFuncSetAttribute(k_thomas_lp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LpSmem)));
        FEN_CUDA(cudaFuncSetAttribute(k_thomas_lp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LpSmem)));
    }
    const double d = c->g.delta;
    const double cinv = 1.0 / (1.0 / (d * d));      // 1 / c_j, c_j = 1/delta**2 for every pipelined row (poisson.f90:219-232)
    dim3 grid((t.npc + 31) / 32), block(32);
    const double amid = 1.0 / (d * d);              // a_j of every row but the first (:219-232)
    FEN_LAUNCH(c, "thomas_fwd", k_thomas_lp<false><<<grid, block, sizeof(LpSmem), c->stream>>>(t, cinv, amid));
    FEN_LAUNCH(c, "thomas_bwd", k_thomas_lp<true><<<grid, block, sizeof(LpSmem), c->stream>>>(t, cinv, amid));
    FEN_CUDA(cudaGetLastError());
    return FEN_OK;
}

// ---- blocked slab path (slab_bulk.cuh): ppp / ppn on several ranks, 64..1024-point transform lines ----------------
template <int Lf> static int launch_bs(fen_ctx* c, int what, const BAddr& in, const BAddr& out, const double2* tw,
                                       double scale, const SolveArgs* sa, const BulkDst* d, dim3 grid, cudaStream_t st) {
    const int bytes = Lf * 8 * (int)sizeof(double2);
    FEN_ONCE_PER_DEVICE(c) {
        FEN_CUDA(cudaFuncSetAttribute(k_fft_lines_bs<Lf, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        FEN_CUDA(cudaFuncSetAttribute(k_fft_solve_bs<Lf>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        FEN_CUDA(cudaFuncSetAttribute(k_fft_lines_io<Lf, +1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    }
    if (what == 0) FEN_LAUNCH(c, "fft_lines_fwd_a2a", k_fft_lines_bs<Lf, -1><<<grid, Lf, bytes, st>>>(in, tw, scale, *d));
    if (what == 1) FEN_LAUNCH(c, "fft_lines_inv", k_fft_lines_io<Lf, +1><<<grid, Lf, bytes, st>>>(in, out, tw, scale));
    if (what == 2) FEN_LAUNCH(c, "fft_solve_a2a", k_fft_solve_bs<Lf><<<grid, Lf, bytes, st>>>(in, *sa, *d));
    FEN_CUDA(cudaGetLastError());
    return FEN_OK;
}
static int dispatch_bs(fen_ctx* c, int Lf, int what, const BAddr& in, const BAddr& out, const double2* tw, double scale,
                       const SolveArgs* sa, const BulkDst* d, dim3 grid, cudaStream_t st) {
    switch (Lf) {
#define FEN_CASE(l) case l: return launch_bs<l>(c, what, in, out, tw, scale, sa, d, grid, st);
        FEN_CASE(64) FEN_CASE(128) FEN_CASE(256) FEN_CASE(512) FEN_CASE(1024)
#undef FEN_CASE
    }
    return set_error(FEN_ERR_STATE, "blocked slab path: line length %d", Lf);
}

// x r2c -> y forward -> (bulk stores) -> z solve or Thomas -> (bulk stores) -> y inverse into the row array Cr.
//
// The two transposing stages are bound by the link (941 MB out per GPU at 1024^3 on 8: >= 1.2 ms each at the 770 GB/s a
// peer copy reaches), their blocks spend most of their life waiting for bulk stores to drain, and the passes either
// side of them are bound by local HBM.  So the work is cut in nchunk pieces and pipelined over two streams:
//   forward   piece q = z planes:  x r2c(q) on the main stream, y forward + stores(q) on `aux` behind it -- the x pass
//             of piece q+1 runs while piece q's stores travel;
//   backward  piece q = granules:  z solve (or Thomas + row shipping) + stores(q) on the main stream, then "piece q is
//             out" is raised on every peer; `aux` waits until every peer has raised it and runs the y inverse of those
//             granules while piece q+1 is solved and shipped.
// With per-kernel profiling on (bench.py's kernel table) everything runs on the main stream, piece by piece, so that
// the event brackets mean what they say; the timed region of the bench runs overlapped.
//
// Measured and not kept (round 2, records under profiles/): (i) moving the pieces with the COPY ENGINES instead -- the
// kernels fill a local send buffer laid out like the receivers' regions, one pitched cudaMemcpy2DAsync per peer and piece
// on separate copy streams, no SM involved -- reaches 373 / 438 GB/s at 8 GPUs against 675 / 649 GB/s for the bulk stores
// and makes the step 11.5 ms instead of 9.3 (r02l_bench_n8_copy_engines*.json; 495 GB/s to a single peer at 2 GPUs,
// r02k_*.json); (ii) capping the persistent transposing kernels below ~96 SMs' worth of blocks (r02i_*.json).
static int solve_blocked(fen_ctx* c, Poisson* p, bool ppp, XArgs xa, bool fuse_rhs, const DivArgs* dv) {
    const fen_grid_desc& g = c->g;
    const int NG = p->PC / 8, P = g.nranks;
    const long long ny = g.ny, nz = g.nz, nyl = p->nyl, nzl = p->nzl;
    cudaStream_t S = c->stream, T = (c->profiling || !p->aux) ? c->stream : p->aux;
    const bool two = T != S;
    // per-kernel profiling: one piece, one stream -- the kernel table then shows whole passes (and the transposes' GB/s
    // figures are those of the whole transfer, without the tails of four short launches)
    const int nq = c->profiling ? 1 : p->nchunk;
    // grid cap of the persistent transposing kernels while they share the GPU with the pass on the other stream
    // (FEN_SLAB_SMS = SMs' worth of their blocks, default 96.  Measured at 8 GPUs, 1024-point tiles, profiles/r02i_*.json:
    // 9.26 ms/step unchunked, 9.07 / 10.13 / 11.14 ms with 4 pieces at caps of 96 / 64 / 48 SMs -- a block ships a tile
    // every ~10 us, so the link needs about a hundred of them)
    static const int cap_sms = getenv("FEN_SLAB_SMS") ? std::max(1, atoi(getenv("FEN_SLAB_SMS"))) : 96;
    auto cap = [&](int Lf, long long ntiles) {
        const long long per_sm = Lf >= 1024 ? 1 : 1024 / Lf;
        return (unsigned)std::min<long long>(ntiles, two ? cap_sms * per_sm : ntiles);
    };
    BAddr none{nullptr, 0, 0, 0, 0, 0};
    BulkDst df;                                                           // -> Cz[((g*nz + k)*nyl + jl)*8 + kxi]
    memset(&df, 0, sizeof(df));
    for (int r = 0; r < P; ++r) df.peer[r] = p->peerCz[r];
    df.gs = nz * nyl * 8; df.os = nyl * 8; df.o0 = g.rank * (int)nzl; df.blk = (int)nyl; df.P = P; df.rank = g.rank;
    const double sy = ppp ? 1.0 : 1.0 / f32(g.ny);                        // poisson.f90:1087
    // ---- forward: pieces of z planes ----
    for (int q = 0; q < nq; ++q) {
        const int z0 = (int)(nzl * q / nq), z1 = (int)(nzl * (q + 1) / nq);
        if (z1 <= z0) continue;
        XArgs xq = xa;
        xq.r0 = z0 * (int)ny; xq.nrows = z1 * (int)ny;
        FEN_TRY(dispatch_x(c, p->M, xq, true, fuse_rhs ? dv : nullptr)); // :965-969 (+ :111-121 when fused)
        if (two) {
            FEN_CUDA(cudaEventRecord(p->ev_piece[q], S));
            FEN_CUDA(cudaStreamWaitEvent(T, p->ev_piece[q], 0));
        }
        BAddr rows_in{p->C, 8, (long long)p->PC * ny, p->PC, 0, z0};      // C[kx + PC*(j + ny*zl)]: lines over j
        df.ng = NG; df.no = z1 - z0;
        FEN_TRY(dispatch_bs(c, g.ny, 0, rows_in, none, p->tw_y, sy, nullptr, &df,
                            dim3(cap(g.ny, (long long)NG * (z1 - z0))), T));
    }
    if (two) {
        FEN_CUDA(cudaEventRecord(p->ev_join, T));
        FEN_CUDA(cudaStreamWaitEvent(S, p->ev_join, 0));
    }
    FEN_TRY(comm_transpose_fwd(c));                                       // transpose_y_to_z (:982 / :1090)
    // ---- backward: pieces of granules ----
    BulkDst db;                                                           // -> Cy[((g*ny + j)*nzl + zl)*8 + kxi]
    memset(&db, 0, sizeof(db));
    for (int r = 0; r < P; ++r) db.peer[r] = p->peerC[r];
    db.gs = ny * nzl * 8; db.os = nzl * 8; db.o0 = g.rank * (int)nyl; db.blk = (int)nzl; db.P = P; db.rank = g.rank;
    for (int q = 0; q < nq; ++q) {
        const int g0 = NG * q / nq, g1 = NG * (q + 1) / nq;
        if (g1 <= g0) continue;
        if (ppp) {
            BAddr zin{p->Cz, nz * nyl * 8, 8, nyl * 8, g0, 0};            // lines over k
            SolveArgs sa;
            sa.tw = p->tw_z; sa.lx = p->mwn_x; sa.lo = p->mwn_y; sa.ll = p->mwn_z;
            sa.norm = f32((long long)g.nx * g.ny * g.nz); sa.ow0 = g.rank * (int)nyl;
            db.ng = g1 - g0; db.no = (int)nyl;
            FEN_TRY(dispatch_bs(c, g.nz, 2, zin, none, nullptr, 1.0, &sa, &db,
                                dim3(cap(g.nz, (long long)(g1 - g0) * nyl)), S));
        } else {
            TArgs t;
            t.C = p->Cz; t.c1 = p->c1; t.sl = nyl * 8; t.so = nz * nyl * 8; t.n = g.nz; t.npc = p->PC;
            t.nouter = (int)nyl; t.o0 = g.rank * (int)nyl;
            t.a = p->ta; t.b = p->tb; t.c = p->tc; t.lx = p->mwn_x; t.lo = p->mwn_y; t.form2d = 0; t.mean = 1;
            t.blocked = 1; t.g0 = g0;
            dim3 grid(((unsigned)nyl * 8 + 127) / 128, g1 - g0), block(128);
            FEN_LAUNCH(c, "thomas_fwd", k_thomas_fwd<<<grid, block, 0, S>>>(t));
            FEN_LAUNCH(c, "thomas_bwd", k_thomas_bwd<<<grid, block, 0, S>>>(t));
            const int bytes = (int)nzl * 8 * (int)sizeof(double2);
            FEN_ONCE_PER_DEVICE(c)
                FEN_CUDA(cudaFuncSetAttribute(k_bulk_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
            FEN_LAUNCH(c, "a2a_scatter", k_bulk_rows<<<dim3((unsigned)(P * nyl), g1 - g0), 256, bytes, S>>>(
                                             p->Cz, nz * nyl * 8, nyl * 8, 8, db, (int)nyl, g0));
            FEN_CUDA(cudaGetLastError());
        }
        FEN_TRY(comm_chunk_signal(c, q, S));                              // my piece q is out ...
        if (two) {
            FEN_CUDA(cudaEventRecord(p->ev_piece[q], S));
            FEN_CUDA(cudaStreamWaitEvent(T, p->ev_piece[q], 0));
        }
        FEN_TRY(comm_chunk_wait(c, q, T));                                // ... and everybody's has arrived (:1015 / :1138)
        BAddr yin{p->C, ny * nzl * 8, 8, nzl * 8, g0, 0};                 // Cy lives in C's memory: lines over j
        BAddr rows_out{p->Cr, 8, (long long)p->PC * ny, p->PC, g0, 0};
        FEN_TRY(dispatch_bs(c, g.ny, 1, yin, rows_out, p->tw_y, 1.0, nullptr, nullptr, dim3(g1 - g0, (unsigned)nzl), T));
    }
    if (two) {
        FEN_CUDA(cudaEventRecord(p->ev_join, T));
        FEN_CUDA(cudaStreamWaitEvent(S, p->ev_join, 0));
    }
    return FEN_OK;
}

int poisson_solve(fen_ctx* c, double* f, bool fuse_rhs, double dt) {
    Poisson* p = c->ps;
    if (!p) return set_error(FEN_ERR_STATE, "solve_poisson before init_poisson_solver");
    DivArgs dv{};
    if (fuse_rhs) {
        if (!poisson_can_fuse_rhs(c)) return set_error(FEN_ERR_STATE, "fused Poisson right-hand side not available");
        Field *u, *v, *w;
        FEN_TRY(field_check(c, FEN_VX, &u));
        FEN_TRY(field_check(c, FEN_VY, &v));
        FEN_TRY(field_check(c, FEN_VZ, &w));
        dv.u = u->d; dv.v = v->d; dv.w = w->d;
        dv.idelta = 1.0 / c->g.delta;
        dv.rho0 = c->rho_uniform;
        dv.dt = dt;
    }
    const fen_grid_desc& g = c->g;
    const bool ppp = !strcmp(p->variant, "ppp"), ppn = !strcmp(p->variant, "ppn");
    const bool pp = !strcmp(p->variant, "pp"), pn = !strcmp(p->variant, "pn");
    const bool multi = g.nranks > 1 && g.ndim == 3;
    if (p->variant[0] == 'n') {
        // Neumann in x: nn (poisson.f90:509-596), npn (:1177-1312), nnn (:1316-1451)
        if (fuse_rhs) return set_error(FEN_ERR_STATE, "fused Poisson right-hand side needs a periodic x direction");
        const bool nn = !strcmp(p->variant, "nn"), npn = !strcmp(p->variant, "npn");
        DArgs da;
        da.L = c->L; da.f = f; da.C = p->C; da.PC = p->PC; da.ny = g.ny; da.nrows = g.ny * p->nzl;
        da.tw = p->tw_x; da.twq = p->twq_x;
        da.scale = nn ? 1.0 : 1.0 / (double)(2 * g.nx);                  // :1213, :1352 (real(nx*2, dp))
        FEN_TRY(dispatch_dct_x(c, g.nx, da, true));
        TArgs t;
        t.C = p->C; t.c1 = p->c1; t.npc = p->PC; t.a = p->ta; t.b = p->tb; t.c = p->tc; t.lx = p->mwn_x;
        t.mean = 0;                                                      // npn / nnn compute the mean but do not
        ScArgs none;                                                     // subtract it (:1310, :1449); nn has none
        memset(&none, 0, sizeof(none));
        LArgs la;
        la.C = p->C; la.sl = p->PC; la.so = (long long)p->PC * g.ny; la.tw = p->tw_y; la.o0 = 0; la.cx0 = 0;
        la.lx = nullptr; la.lo = nullptr; la.ll = nullptr; la.norm = 1.0;
        // slabs: the y <-> z transposes of these variants (poisson.f90:1229, :1279 / :1368, :1418) are separate
        // staggered row copies to the owning ranks (k_a2a_scatter); built for coverage, not fused into the passes
        ScArgs sf, sb;
        memset(&sf, 0, sizeof(sf));
        memset(&sb, 0, sizeof(sb));
        if (multi) {
            for (int r = 0; r < g.nranks; ++r) { sf.peer[r] = p->peerCz[r]; sb.peer[r] = p->peerC[r]; }
            sc_block(sf, p->nyl);
            sf.dsl = p->PC; sf.dso = (long long)p->PC * p->nyl; sf.o0 = g.rank * p->nzl;
            sc_block(sb, p->nzl);
            sb.dsl = (long long)p->PC * g.ny; sb.dso = p->PC; sb.o0 = g.rank * p->nyl;
        }
        if (nn) {
            t.sl = p->PC; t.so = 0; t.n = g.ny; t.nouter = 1; t.o0 = 0; t.lo = nullptr; t.form2d = 1;
        } else {
            la.scale = npn ? 1.0 / (double)g.ny : 1.0 / (double)(2 * g.ny);    // :1226, :1365
            if (npn) FEN_TRY(dispatch_lines(c, g.ny, la, 0, p->PC, p->nzl));   // full c2c of real data == r2c
            else FEN_TRY(dispatch_dct_lines(c, g.ny, la, p->twq_y, true, p->PC, p->nzl));
            if (multi) {
                FEN_LAUNCH(c, "a2a_scatter", k_a2a_scatter<<<dim3(g.ny, p->nzl), 256, 0, c->stream>>>(
                                                 p->C, (long long)p->PC, (long long)p->PC * g.ny, p->PC, g.nranks,
                                                 g.rank, p->nyl, sf));
                FEN_CUDA(cudaGetLastError());
                FEN_TRY(comm_transpose_fwd(c));
            }
            t.C = p->Cz;
            t.sl = (long long)p->PC * p->nyl; t.so = p->PC; t.n = g.nz; t.nouter = p->nyl;
            t.o0 = multi ? g.rank * p->nyl : 0;
            t.lo = p->mwn_y; t.form2d = 0;
        }
        if (nn) {
            FEN_TRY(thomas_2d(c, t));
        } else {
            dim3 tgrid((p->PC + 127) / 128, t.nouter), tblock(128);
            FEN_LAUNCH(c, "thomas_fwd", k_thomas_fwd<<<tgrid, tblock, 0, c->stream>>>(t));
            FEN_LAUNCH(c, "thomas_bwd", k_thomas_bwd<<<tgrid, tblock, 0, c->stream>>>(t));
            if (multi)
                FEN_LAUNCH(c, "a2a_scatter", k_a2a_scatter<<<dim3(g.nz, p->nyl), 256, 0, c->stream>>>(
                                                 p->Cz, (long long)p->PC * p->nyl, (long long)p->PC, p->PC, g.nranks,
                                                 g.rank, p->nzl, sb));
            FEN_CUDA(cudaGetLastError());
            if (multi) FEN_TRY(comm_transpose_bwd(c));
        }
        if (!nn) {
            la.scale = 1.0;
            if (npn) FEN_TRY(dispatch_lines(c, g.ny, la, 1, p->PC, p->nzl));
            else FEN_TRY(dispatch_dct_lines(c, g.ny, la, p->twq_y, false, p->PC, p->nzl));
        }
        da.scale = nn ? 1.0 / f32(2LL * g.nx) : 1.0;                     // :594 phi/float(nx*2)
        FEN_TRY(dispatch_dct_x(c, g.nx, da, false));
        return FEN_OK;
    }
    XArgs xa;
    xa.L = c->L; xa.f = f; xa.C = p->C; xa.PC = p->PC; xa.ny = g.ny; xa.nrows = g.ny * p->nzl;
    xa.tw = p->tw_x; xa.twr = p->twr_x;
    xa.scale = (ppn || pn) ? 1.0 / f32(g.nx) : 1.0;           // poisson.f90:1074, :341
    xa.r0 = 0;
    if (!p->blocked) FEN_TRY(dispatch_x(c, p->M, xa, true, fuse_rhs ? &dv : nullptr));   // blocked: in pieces, below

    LArgs la;
    la.lx = p->mwn_x; la.lo = nullptr; la.ll = nullptr; la.norm = 1.0; la.o0 = 0; la.cx0 = 0;
    if (pp) {
        // forward y + divide + inverse y fused (poisson.f90:451-478)
        la.C = p->C; la.sl = p->PC; la.so = 0; la.tw = p->tw_y; la.scale = 1.0;
        la.ll = p->mwn_y; la.norm = f32((long long)g.nx * g.ny);
        FEN_TRY(dispatch_lines(c, g.ny, la, 2, p->PC, 1));
    } else if (pn) {
        TArgs t;
        t.C = p->C; t.c1 = p->c1; t.sl = p->PC; t.so = 0; t.n = g.ny; t.npc = p->PC; t.nouter = 1; t.o0 = 0;
        t.a = p->ta; t.b = p->tb; t.c = p->tc; t.lx = p->mwn_x; t.lo = nullptr; t.form2d = 1; t.mean = 1;
        FEN_TRY(thomas_2d(c, t));
    } else {
        // y forward (poisson.f90:975-979 / :1080-1087); multi-rank: the result is stored straight into the
        // z-pencil arrays of the owning ranks (transpose_y_to_z, :982 / :1090)
        ScArgs sf, sb;
        memset(&sf, 0, sizeof(sf));
        memset(&sb, 0, sizeof(sb));
        if (multi) {
            for (int r = 0; r < g.nranks; ++r) { sf.peer[r] = p->peerCz[r]; sb.peer[r] = p->peerC[r]; }
            sc_block(sf, p->nyl);
            sf.dsl = p->PC; sf.dso = (long long)p->PC * p->nyl; sf.o0 = g.rank * p->nzl;
            sc_block(sb, p->nzl);
            sb.dsl = (long long)p->PC * g.ny; sb.dso = p->PC; sb.o0 = g.rank * p->nyl;
        }
        if (p->blocked) {
            FEN_TRY(solve_blocked(c, p, ppp, xa, fuse_rhs, &dv));
            xa.C = p->Cr;                                      // the y inverse left the rows there
        } else {
        // the fused epilogues need power-of-two lines and blocks; any other length (the any-length kernels of
        // fft_any.cuh, or an odd number of ranks) transforms in place and ships whole rows with k_a2a_scatter
        const bool rowcopy = multi && (!tuned_len(g.ny, 2048) || (ppp && !tuned_len(g.nz, 2048)) || sf.sh < 0 || sb.sh < 0);
        la.C = p->C; la.sl = p->PC; la.so = (long long)p->PC * g.ny; la.tw = p->tw_y;
        la.scale = ppn ? 1.0 / f32(g.ny) : 1.0;
        FEN_TRY(dispatch_lines(c, g.ny, la, 0, p->PC, p->nzl, multi && !rowcopy ? &sf : nullptr));
        if (rowcopy) {
            FEN_LAUNCH(c, "a2a_scatter", k_a2a_scatter<<<dim3(g.ny, p->nzl), 256, 0, c->stream>>>(
                                             p->C, (long long)p->PC, (long long)p->PC * g.ny, p->PC, g.nranks, g.rank,
                                             p->nyl, sf));
            FEN_CUDA(cudaGetLastError());
        }
        if (multi) FEN_TRY(comm_transpose_fwd(c));
        double2* Z = p->Cz;
        const long long slz = (long long)p->PC * p->nyl;
        if (ppp) {
            la.C = Z; la.sl = slz; la.so = p->PC; la.tw = p->tw_z; la.scale = 1.0; la.o0 = multi ? g.rank * p->nyl : 0;
            la.lo = p->mwn_y; la.ll = p->mwn_z; la.norm = f32((long long)g.nx * g.ny * g.nz);
            FEN_TRY(dispatch_lines(c, g.nz, la, 2, p->PC, p->nyl, multi && !rowcopy ? &sb : nullptr));
            if (rowcopy)
                FEN_LAUNCH(c, "a2a_scatter", k_a2a_scatter<<<dim3(g.nz, p->nyl), 256, 0, c->stream>>>(
                                                 Z, slz, (long long)p->PC, p->PC, g.nranks, g.rank, p->nzl, sb));
        } else {
            TArgs t;
            t.C = Z; t.c1 = p->c1; t.sl = slz; t.so = p->PC; t.n = g.nz; t.npc = p->PC; t.nouter = p->nyl;
            t.o0 = multi ? g.rank * p->nyl : 0;
            t.a = p->ta; t.b = p->tb; t.c = p->tc; t.lx = p->mwn_x; t.lo = p->mwn_y; t.form2d = 0; t.mean = 1;
            dim3 grid((p->PC + 127) / 128, p->nyl), block(128);
            FEN_LAUNCH(c, "thomas_fwd", k_thomas_fwd<<<grid, block, 0, c->stream>>>(t));
            // the back substitution stays in local HBM and a staggered row copy ships it (see k_a2a_scatter)
            FEN_LAUNCH(c, "thomas_bwd", k_thomas_bwd<<<grid, block, 0, c->stream>>>(t));
            if (multi)
                FEN_LAUNCH(c, "a2a_scatter", k_a2a_scatter<<<dim3(g.nz, p->nyl), 256, 0, c->stream>>>(
                                                 Z, slz, (long long)p->PC, p->PC, g.nranks, g.rank, p->nzl, sb));
        }
        FEN_CUDA(cudaGetLastError());
        if (multi) FEN_TRY(comm_transpose_bwd(c));             // transpose_z_to_y (:1015 / :1138)
        // y inverse (:1018-1022 / :1141-1145)
        la.C = p->C; la.sl = p->PC; la.so = (long long)p->PC * g.ny; la.tw = p->tw_y; la.scale = 1.0;
        la.lo = nullptr; la.ll = nullptr;
        FEN_TRY(dispatch_lines(c, g.ny, la, 1, p->PC, p->nzl));
        }
    }
    FEN_CUDA(cudaGetLastError());
    xa.scale = 1.0;
    FEN_TRY(dispatch_x(c, p->M, xa, false));                   // :1028-1032
    return FEN_OK;
}

}  // namespace fen
