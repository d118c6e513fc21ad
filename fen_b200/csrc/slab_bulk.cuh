// slab_bulk.cuh -- the y <-> z transposes of the slab decomposition (transpose_y_to_z / transpose_z_to_y of
// 2decomp&FFT, src/poisson.f90:982,1015 ppp / :1090,1138 ppn) as BULK asynchronous stores into peer memory.
//
// Round 1 stored every transformed element straight from registers into the owning rank's array: 128-byte (64-byte
// for 1024-point lines) st.global runs scattered over all peers, 423 / 385 GB/s at 8 GPUs, below NCCL's all-to-all.
// The link wants long contiguous writes.  What a block owns after its transform is a tile of Lf line elements x 8
// consecutive kx (one 128-byte "granule" per element); rank r must receive the elements idx in [r*blk, (r+1)*blk).
// So the arrays a transpose WRITES are laid out granule-blocked, with the transposed index next to the granule:
//
//     z-pencil   Cz[ ((g * nz + k ) * nyl + jl) * 8 + kxi ]     g = kx / 8, kxi = kx % 8, k global, jl local
//     y-slab     Cy[ ((g * ny + j ) * nzl + zl) * 8 + kxi ]     j global, zl local
//
// In both the part of a tile that goes to one rank is ONE contiguous run of blk * 128 bytes (16 KB at 1024^3 on 8
// GPUs), both in the tile as it sits in shared memory after the transform ([idx][8]) and in the destination: the
// epilogue is one cp.async.bulk.global.shared::cta per destination rank, issued by eight different threads in
// rotated order (rank + 1 first).  The READ side of the next stage pays instead -- its lines are 128-byte granules at
// a stride of blk * 128 bytes -- which local HBM serves at ~85 % of peak (the single-GPU passes read the same way).
//
// k_fft_lines_bs   y forward:   row layout C[kx + PC*(j + ny*zl)]  ->  transform over j  ->  peers' Cz
// k_fft_solve_bs   z solve:     Cz (lines over k)  ->  forward, divide, inverse           ->  peers' Cy
// k_fft_lines_io   y inverse:   Cy (lines over j)  ->  transform                          ->  row layout (local)
// k_bulk_rows      Thomas path: the back substitution's solution in Cz, gathered run by run    ->  peers' Cy
#pragma once
#include "fen_internal.cuh"
#include "fft_core.cuh"

namespace fen {

// element (g, o, idx, line) of an array:  p + gs*g + os*o + is*idx + line.  A launch covers the granules g0 + blockIdx.x
// and the outer indices ob + blockIdx.y (the chunked, overlapped form of the solve launches pieces of the range)
struct BAddr {
    double2* p;
    long long gs, os, is;
    int g0, ob;
};
// destination of a tile: rank r receives idx in [r*blk, (r+1)*blk) as one run at peer[r] + gs*g + os*(o0 + o).
// The transposing kernels are PERSISTENT: gridDim.x blocks walk the ng x no tiles of a launch (tile = granule fastest).
// Their blocks spend most of their life waiting for the link, and a block of 1024 threads with 128 KB of shared memory
// owns its SM; capping the grid (a few dozen blocks keep the link full) leaves the other SMs to the HBM-bound pass of
// the neighbouring piece that runs beside it on the second stream (poisson.cu: solve_blocked).
struct BulkDst {
    double2* peer[FEN_MAX_RANKS];
    long long gs, os;
    int o0, blk, P, rank;
    int ng, no;            // tiles of this launch: granules x outer indices
};

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // writes complete (not only the source reads)
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void bulk_commit_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the source has been read: the tile may be reused
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// the tile sits in shared memory as s[idx * 8 + line]; threads 0 .. P-1 each ship one destination's run and wait until
// the copy engine has read it; the barrier at the end releases the tile for the block's next one
__device__ __forceinline__ void bulk_scatter_tile(const double2* s, const BulkDst& d, int g, int o, int tid) {
    fence_async_smem();
    __syncthreads();
    if (tid < d.P) {
        const int r = (d.rank + 1 + tid) % d.P;                  // own rank last
        double2* dst = d.peer[r] + d.gs * g + d.os * (d.o0 + o);
        bulk_store(dst, s + (size_t)r * d.blk * 8, (unsigned)(d.blk * 8 * sizeof(double2)));
        bulk_commit_wait_read();
    }
    __syncthreads();
}

template <int Lf, int DIR>
__global__ void __launch_bounds__(Lf, (Lf <= 512) ? 1024 / Lf : 1)
k_fft_lines_bs(BAddr in, const double2* tw, double scale, BulkDst d) {
    extern __shared__ __align__(128) double2 s[];
    constexpr int T = Lf / 8;
    const int tid = threadIdx.x;
    const int line = tid & 7, t = tid >> 3;
    for (int tile = blockIdx.x; tile < d.ng * d.no; tile += gridDim.x) {
        const int g = in.g0 + tile % d.ng, o = in.ob + tile / d.ng;
        const double2* base = in.p + in.gs * g + in.os * o + line;
        double2 v[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) v[m] = base[in.is * (t + m * T)];
        fft_regs<Lf, DIR, true, kStridedTwp>(v, s, 8, line, t, tw);
#pragma unroll
        for (int m = 0; m < 8; ++m) s[(t + m * T) * 8 + line] = make_double2(v[m].x * scale, v[m].y * scale);
        bulk_scatter_tile(s, d, g, o, tid);
    }
    if (tid < d.P) bulk_wait_all();          // every write of this block has landed before the block retires
}

struct SolveArgs {
    const double2* tw;
    const double* lx; const double* lo; const double* ll;   // lam = (lx[kx] + lo[ow0 + o]) + ll[idx]  (poisson.f90:998)
    double norm;                                            // float(nx*ny*nz)  (:992)
    int ow0;
};

template <int Lf>
__global__ void __launch_bounds__(Lf, (Lf <= 512) ? 1024 / Lf : 1)
k_fft_solve_bs(BAddr in, SolveArgs a, BulkDst d) {
    extern __shared__ __align__(128) double2 s[];
    constexpr int T = Lf / 8;
    const int tid = threadIdx.x;
    const int line = tid & 7, t = tid >> 3;
    const double inorm = 1.0 / a.norm;
    for (int tile = blockIdx.x; tile < d.ng * d.no; tile += gridDim.x) {
        const int g = in.g0 + tile % d.ng, o = in.ob + tile / d.ng;
        double2 v[8];
        {
            const double2* base = in.p + in.gs * g + in.os * o + line;
#pragma unroll
            for (int m = 0; m < 8; ++m) v[m] = base[in.is * (t + m * T)];
        }
        fft_regs<Lf, -1, true, kStridedTwp>(v, s, 8, line, t, a.tw);
        {   // poisson.f90:992 then :998-1001; see k_fft_solve_r for the rounding argument (power-of-two norm)
            double lxo = __ldg(&a.lx[g * 8 + line]);
            if (a.lo) lxo = lxo + __ldg(&a.lo[a.ow0 + o]);
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const double lam = lxo + __ldg(&a.ll[t + m * T]);
                const double rl = lam == 0.0 ? 0.0 : inorm / lam;
                v[m].x *= rl;
                v[m].y *= rl;
            }
        }
        fft_regs<Lf, +1, true, kStridedTwp>(v, s, 8, line, t, a.tw);
#pragma unroll
        for (int m = 0; m < 8; ++m) s[(t + m * T) * 8 + line] = v[m];
        bulk_scatter_tile(s, d, g, o, tid);
    }
    if (tid < d.P) bulk_wait_all();
}

// transform with separate source and destination arrays (y inverse: blocked y-slab in, row layout out)
template <int Lf, int DIR>
__global__ void __launch_bounds__(Lf, (Lf <= 512) ? 1024 / Lf : 1)
k_fft_lines_io(BAddr in, BAddr out, const double2* tw, double scale) {
    extern __shared__ __align__(128) double2 s[];
    constexpr int T = Lf / 8;
    const int tid = threadIdx.x;
    const int line = tid & 7, t = tid >> 3;
    const int g = in.g0 + blockIdx.x, o = in.ob + blockIdx.y;
    const double2* base = in.p + in.gs * g + in.os * o + line;
    double2 v[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = base[in.is * (t + m * T)];
    fft_regs<Lf, DIR, true, kStridedTwp>(v, s, 8, line, t, tw);
    double2* ob = out.p + out.gs * g + out.os * o + line;
#pragma unroll
    for (int m = 0; m < 8; ++m) ob[out.is * (t + m * T)] = make_double2(v[m].x * scale, v[m].y * scale);
}

// Thomas path (ppn): the back substitution leaves the solution in place in Cz ([g][k][jl][8]).  One block ships the
// part of one (g, jl) column that belongs to rank r -- the nzl granules k in [r*nzl, (r+1)*nzl), 128 bytes each at a
// stride of nyl * 128 bytes -- by gathering it into shared memory (every warp instruction reads four whole granules)
// and issuing one bulk store of nzl * 128 bytes to the run it occupies in the destination Cy.  Consecutive blocks
// cycle over the destination ranks, starting one past the sender.
__global__ void __launch_bounds__(256) k_bulk_rows(const double2* __restrict__ src, long long s_gs, long long s_ks,
                                                   long long s_os, BulkDst d, int nouter, int g0) {
    extern __shared__ __align__(128) double2 s[];
    const int r = (d.rank + 1 + blockIdx.x % d.P) % d.P;
    const int o = blockIdx.x / d.P;                   // jl
    const int g = g0 + blockIdx.y;
    if (o >= nouter) return;
    const int n = d.blk * 8;
    const double2* sp = src + s_gs * g + s_os * o + s_ks * ((long long)r * d.blk);
    for (int e = threadIdx.x; e < n; e += blockDim.x) s[e] = sp[s_ks * (e >> 3) + (e & 7)];
    fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        bulk_store(d.peer[r] + d.gs * g + d.os * (d.o0 + o), s, (unsigned)(n * sizeof(double2)));
        bulk_commit_wait();
    }
}

}  // namespace fen
