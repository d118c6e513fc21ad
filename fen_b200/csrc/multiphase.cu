// multiphase.cu -- the two-phase path of FEN on the device: MTHINC volume of fluid (src/volume_of_fluid.f90),
// material properties (src/multiphase.f90) and the `#ifdef MF` branches of the fractional step
// (src/navier_stokes.f90:80-96, 113, 174-184, 405-501, 526-531, 553-558, 655-661, 724; src/solver.f90:87-98).
// 2-D only, as in the reference.  The per-cell arithmetic lives in vof_math.cuh.
//
// One advect_vof + material update + predictor costs, per cell (fields are read through L1/L2, so every field
// leaves HBM once per kernel):
//   k_vof_recon   x2 : 1 R (vof, 3x3)                      + 7 W (norm.x/y, l.x/y, curv, h, d)       =  64 B each
//   k_vof_sweep   x2 : 7 R (vof, norm, l, d, face velocity) + 1 W  (+2 R in the final sweep)         =  64 / 80 B
//   k_mf_props       : 3 R (vof, p, p_o)                    + 3 W (rho, mu, p_hat)                   =  48 B
//   k_mf_pred        : 10 R (u, v, p, p_hat, rho, mu, vof, curv, dv_o x2) + 4 W (u*, v*, dv_o x2)    = 112 B
//   k_mf_rhs         : 2 R + 1 W                                                                     =  24 B
//   k_mf_corr        : 4 R (phi, u*, v*, p) + 4 W (u, v, p, p_o)                                     =  64 B
// The reference makes ~60 whole-array passes for the same work (every `RHS%x%f = ...` statement is one).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <vector>

#include "fen_internal.cuh"
#include "vof_math.cuh"

namespace fen {

constexpr int MTX = 64, MTY = 4;

static dim3 mf_grid(const Layout& L) { return dim3((L.nx + MTX - 1) / MTX, (L.ny + MTY - 1) / MTY, 1); }

// Periodic x ghosts written by the PRODUCING kernel (scalar.f90:257,276: f(0,j) = f(nx,j), f(nx+1,j) = f(1,j)): when
// every field a kernel writes is periodic in x, its edge threads store the two ghost cells of their row as well and the
// ghost update skips its x pass (ghost.cu: x_done) -- seven tiny column-strided launches less per two-phase step.  The
// y pass still copies whole rows, x ghosts included, so corners come out in the reference's order.
__device__ __forceinline__ void put_x(double* f, long long c, int i, int nx, double val, int xper) {
    f[c] = val;
    if (xper) {
        if (i == 1) f[c + nx] = val;
        if (i == nx) f[c - nx] = val;
    }
}
static bool x_periodic(fen_ctx* c, const int* ids, int n) {
    static const bool off = getenv("FEN_MF_XGHOST") && atoi(getenv("FEN_MF_XGHOST")) == 0;   // cross-check switch
    if (off) return false;
    for (int q = 0; q < n; ++q) {
        const Field& f = c->fields[ids[q]];
        if (f.bc_type[FEN_LEFT] != FEN_PERIODIC || f.bc_type[FEN_RIGHT] != FEN_PERIODIC) return false;
    }
    return true;
}

// ---- reconstruction: compute_norm + the cell loop of get_h_from_vof -----------------------------------------------
struct ReconArgs {
    Layout L;
    const double* vof;
    double *nx, *ny, *lx, *ly, *curv, *h, *d;
    double delta, idelta, idelta2, beta, cut;
    int quadratic;
    int xper;                                 // the kernel also writes the periodic x ghosts of its seven fields
};

__global__ void __launch_bounds__(MTX* MTY, 4) k_vof_recon(ReconArgs a) {
    const int i = blockIdx.x * MTX + threadIdx.x + 1;
    const int j = blockIdx.y * MTY + threadIdx.y + 1;
    if (i > a.L.nx || j > a.L.ny) return;
    const long long c = a.L.idx(i, j, 1);
    const long long sy = a.L.sy;
    const double* __restrict__ f = a.vof;
    double s[3][3];
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int q = 0; q < 3; ++q) s[b][q] = f[c + (b - 1) * sy + (q - 1)];
    const VofRecon r = vof_norm(s, a.delta, a.idelta, a.idelta2, a.quadratic != 0);
    double h, d;
    vof_h_d(s[1][1], r.nx, r.ny, r.lx, r.ly, a.beta, a.cut, h, d);
    const int nx = a.L.nx;
    put_x(a.nx, c, i, nx, r.nx, a.xper); put_x(a.ny, c, i, nx, r.ny, a.xper); put_x(a.lx, c, i, nx, r.lx, a.xper);
    put_x(a.ly, c, i, nx, r.ly, a.xper); put_x(a.curv, c, i, nx, r.curv, a.xper);
    put_x(a.h, c, i, nx, h, a.xper); put_x(a.d, c, i, nx, d, a.xper);
}

// The same reconstruction with the corner normals of a 64 x 4 tile computed once in shared memory: the reference
// evaluates every corner from the four cells around it with identical operands, so sharing it changes no bit and
// cuts the square roots / divisions of the interface band (and of the 1 +- eps gas phase) from 5 + 10 to about
// 2.3 + 4.5 per cell.
__global__ void __launch_bounds__(VT_N, 4) k_vof_recon_tile(ReconArgs a) {
    __shared__ VofTile T;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * VT_X + tx;
    const int i0 = blockIdx.x * VT_X, j0 = blockIdx.y * VT_Y;     // the tile's low halo cell (global indices)
    vof_tile_load(T, tid, a.vof + a.L.idx(i0, j0, 1), a.L.sy, min(VT_FW, a.L.nx + 2 - i0), min(VT_FH, a.L.ny + 2 - j0));
    __syncthreads();
    vof_tile_corners(T, tid, a.idelta);
    __syncthreads();
    const int i = i0 + 1 + tx, j = j0 + 1 + ty;
    if (i > a.L.nx || j > a.L.ny) return;
    double v00;
    const VofRecon r = vof_tile_cell(T, tx, ty, a.delta, a.idelta2, a.quadratic != 0, v00);
    double h, d;
    vof_h_d(v00, r.nx, r.ny, r.lx, r.ly, a.beta, a.cut, h, d);
    const long long c = a.L.idx(i, j, 1);
    const int nx = a.L.nx;
    put_x(a.nx, c, i, nx, r.nx, a.xper); put_x(a.ny, c, i, nx, r.ny, a.xper); put_x(a.lx, c, i, nx, r.lx, a.xper);
    put_x(a.ly, c, i, nx, r.ly, a.xper); put_x(a.curv, c, i, nx, r.curv, a.xper);
    put_x(a.h, c, i, nx, h, a.xper); put_x(a.d, c, i, nx, d, a.xper);
}

// ---- directional split sweeps of advect_vof (volume_of_fluid.f90:457-538) ----------------------------------------
struct SweepArgs {
    Layout L;
    const double* src;                        // the reconstructed field: vof (first sweep) or vof1 (second)
    const double *nx, *ny, *lx, *ly, *d;
    const double *u, *v;                      // v%x, v%y
    double* out;
    double dt, delta, beta, cut;
    int x_first;                              // value of x_first when advect_vof was entered
    int xper;                                 // the kernel also writes the periodic x ghosts of `out`
};

// DIR 1: x, 2: y.  FINAL: second sweep fused with the update to time n+1 (:492-498 / :532-538).
// Every face flux is evaluated ONCE (the reference's compute_flux evaluates it from both sides): the flux through the
// + face of a cell is the flux through the - face of its neighbour, same upwind cell, same arguments, same bits.
//   x sweep: the - face flux of cell i is the + face flux lane - 1 of the warp computed (shuffle); lane 0 evaluates its own;
//   y sweep: a thread marches SWEEP_RY consecutive rows and carries the + face flux of row j as the - face flux of row j+1.
constexpr int SWEEP_RY = 8;
template <int DIR, bool FINAL>
__global__ void __launch_bounds__(MTX* MTY) k_vof_sweep(SweepArgs a) {
    const int i = blockIdx.x * MTX + threadIdx.x + 1;
    const long long s = DIR == 1 ? 1 : a.L.sy;
    const double* __restrict__ vel = DIR == 1 ? a.u : a.v;
    auto face = [&](long long c, double uf) {                // flux through the + face of cell c, face velocity uf
        const long long cu = uf >= 0.0 ? c : c + s;          // upwind cell (:567-581)
        const double vf = a.src[cu];
        // outside the interface band the flux is vof times the swept volume and the reconstruction is never looked at
        // (vof_flux returns before it touches it): do not load it either -- five of the seven fields a sweep reads, for
        // nearly every cell of the domain (the r02u capture: 474 MB read per sweep = all seven fields everywhere)
        if (vf <= a.cut || vf >= (1.0 - a.cut)) return vof_flux(DIR, uf, a.dt, a.delta, a.beta, a.cut, vf, 0.0, 0.0, 0.0, 0.0, 0.0);
        return vof_flux(DIR, uf, a.dt, a.delta, a.beta, a.cut, vf, a.nx[cu], a.ny[cu], a.lx[cu], a.ly[cu], a.d[cu]);
    };
    auto finish = [&](long long c, double up, double um, double fp, double fm) {
        const double s0 = a.src[c];
        const double val = (s0 - (fp - fm) / a.delta) / (1.0 - a.dt * (up - um) / a.delta);
        if (!FINAL) {
            put_x(a.out, c, i, a.L.nx, val, a.xper);
            return;
        }
        // here src = vof1 and val = vof2
        const double dux = a.u[c] - a.u[c - 1];
        const double dvy = a.v[c] - a.v[c - a.L.sy];
        if (a.x_first) put_x(a.out, c, i, a.L.nx, val - a.dt * (s0 * dux / a.delta + val * dvy / a.delta), a.xper);
        else put_x(a.out, c, i, a.L.nx, val - a.dt * (val * dux / a.delta + s0 * dvy / a.delta), a.xper);
    };
    if (DIR == 1) {
        const int j = blockIdx.y * MTY + threadIdx.y + 1;
        const bool valid = i <= a.L.nx && j <= a.L.ny;
        const long long c = a.L.idx(valid ? i : 1, valid ? j : 1, 1);
        const double up = vel[c], um = vel[c - s];
        const double fp = valid ? face(c, up) : 0.0;
        double fm = __shfl_up_sync(0xffffffffu, fp, 1);      // lane - 1 holds cell i - 1 of the same row (MTX % 32 == 0)
        if ((threadIdx.x & 31) == 0 && valid) fm = face(c - s, um);
        if (valid) finish(c, up, um, fp, fm);
    } else {
        if (i > a.L.nx) return;
        const int j0 = (blockIdx.y * MTY + threadIdx.y) * SWEEP_RY + 1;
        if (j0 > a.L.ny) return;
        long long c = a.L.idx(i, j0, 1);
        double um = vel[c - s];
        double fm = face(c - s, um);
        for (int r = 0; r < SWEEP_RY && j0 + r <= a.L.ny; ++r, c += s) {
            const double up = vel[c];
            const double fp = face(c, up);
            finish(c, up, um, fp, fm);
            um = up;
            fm = fp;
        }
    }
}

// ---- get_vof_from_distance (:676-718): tanh profile + 2x2 Gauss quadrature from host-evaluated distances ----------
// dist: [5][ny][nx] = distance at (xm,ym), (xp,ym), (xm,yp), (xp,yp), (x,y)
__global__ void __launch_bounds__(MTX* MTY) k_vof_from_dist(Layout L, const double* dist, double* vof, double* h,
                                                           double beta, double delta) {
    const int i = blockIdx.x * MTX + threadIdx.x + 1;
    const int j = blockIdx.y * MTY + threadIdx.y + 1;
    if (i > L.nx || j > L.ny) return;
    const size_t n = (size_t)L.nx * L.ny, o = (size_t)(j - 1) * L.nx + (i - 1);
    auto T = [&](int q) { return 0.5 * (1.0 + tanh(beta * dist[q * n + o] / delta)); };
    const long long c = L.idx(i, j, 1);
    vof[c] = 0.5 * (0.5 * (T(0) + T(1)) + 0.5 * (T(2) + T(3)));
    h[c] = T(4);
}

// ---- update_material_properties (multiphase.f90:121-137) and p_hat (navier_stokes.f90:88-92) ---------------------
struct PropsArgs {
    Layout L;
    const double* vof; double* rho; double* mu;
    double rho_0, rho_1, mu_0, mu_1;
    const double* p; const double* p_o; double* p_hat;     // p_hat == nullptr: material properties only
    int ccfl; double dt, dt_o;
    int xper;                                 // the kernel also writes the periodic x ghosts of rho, mu, p_hat
};
__global__ void __launch_bounds__(MTX* MTY) k_mf_props(PropsArgs a) {
    const int i = blockIdx.x * MTX + threadIdx.x + 1;
    const int j = blockIdx.y * MTY + threadIdx.y + 1;
    if (i > a.L.nx || j > a.L.ny) return;
    const long long c = a.L.idx(i, j, 1);
    const double f = a.vof[c];
    put_x(a.rho, c, i, a.L.nx, a.rho_1 * f + a.rho_0 * (1.0 - f), a.xper);
    put_x(a.mu, c, i, a.L.nx, a.mu_1 * f + a.mu_0 * (1.0 - f), a.xper);
    if (a.p_hat) {
        const double p = a.p[c], po = a.p_o[c];
        put_x(a.p_hat, c, i, a.L.nx, a.ccfl ? po + (a.dt + a.dt_o) * (p - po) / a.dt_o : 2.0 * p - po, a.xper);
    }
}

// ---- predictor ---------------------------------------------------------------------------------------------------
struct MfPredArgs {
    Layout L;
    const double *u, *v, *p, *p_hat, *rho, *mu, *vof, *curv;
    const double *sx, *sy;                    // may be null (S == 0)
    double *dvox, *dvoy, *un, *vn;
    MfPrm k;
    int xper;                                 // the kernel also writes the periodic x ghosts of un, vn
};
__global__ void __launch_bounds__(MTX* MTY, 4) k_mf_pred(MfPredArgs a) {
    const int i = blockIdx.x * MTX + threadIdx.x + 1;
    const int j = blockIdx.y * MTY + threadIdx.y + 1;
    if (i > a.L.nx || j > a.L.ny) return;
    const long long c = a.L.idx(i, j, 1);
    const long long sy = a.L.sy;
    MfCell q;
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const long long o = c + (b - 1) * sy + (e - 1);
            q.u[b][e] = a.u[o];
            q.v[b][e] = a.v[o];
            q.m[b][e] = a.mu[o];
        }
    q.rho0 = a.rho[c]; q.rhoip = a.rho[c + 1]; q.rhojp = a.rho[c + sy];
    q.p0 = a.p[c]; q.pip = a.p[c + 1]; q.pjp = a.p[c + sy];
    q.h0 = a.p_hat[c]; q.hip = a.p_hat[c + 1]; q.hjp = a.p_hat[c + sy];
    q.c0 = a.curv[c]; q.cip = a.curv[c + 1]; q.cjp = a.curv[c + sy];
    q.f0 = a.vof[c]; q.fip = a.vof[c + 1]; q.fjp = a.vof[c + sy];
    q.sx = a.sx ? a.sx[c] : 0.0;
    q.sy = a.sy ? a.sy[c] : 0.0;
    q.dvox = a.dvox[c]; q.dvoy = a.dvoy[c];
    double un, vn, dvx, dvy;
    mf_predict_cell(q, a.k, un, vn, dvx, dvy);
    put_x(a.un, c, i, a.L.nx, un, a.xper); put_x(a.vn, c, i, a.L.nx, vn, a.xper);
    a.dvox[c] = dvx; a.dvoy[c] = dvy;               // dv_o = dv (:201-202)
}

// phi = div(v) * rhomin / dt (navier_stokes.f90:111-113)
__global__ void __launch_bounds__(MTX* MTY) k_mf_rhs(Layout L, const double* u, const double* v, double* phi,
                                                    double id, double rhomin, double dt) {
    const int i = blockIdx.x * MTX + threadIdx.x + 1;
    const int j = blockIdx.y * MTY + threadIdx.y + 1;
    if (i > L.nx || j > L.ny) return;
    const long long c = L.idx(i, j, 1);
    const double d = (u[c] - u[c - 1]) * id + (v[c] - v[c - L.sy]) * id;       // fields.f90:144-145
    phi[c] = d * rhomin / dt;
}

// correct_velocity_field (MF branch, :526-528) + update_pressure (:557, :561)
__global__ void __launch_bounds__(MTX* MTY) k_mf_corr(Layout L, double* u, double* v, double* p, double* p_o,
                                                     const double* phi, double id, double dt, double irhomin, int xper) {
    const int i = blockIdx.x * MTX + threadIdx.x + 1;
    const int j = blockIdx.y * MTY + threadIdx.y + 1;
    if (i > L.nx || j > L.ny) return;
    const long long c = L.idx(i, j, 1);
    const double f0 = phi[c];
    put_x(u, c, i, L.nx, u[c] - ((phi[c + 1] - f0) * id) * dt * irhomin, xper);
    put_x(v, c, i, L.nx, v[c] - ((phi[c + L.sy] - f0) * id) * dt * irhomin, xper);
    const double p0 = p[c];
    put_x(p_o, c, i, L.nx, p0, xper);
    put_x(p, c, i, L.nx, p0 + f0, xper);
}

// ---- host side ---------------------------------------------------------------------------------------------------
static int mf_need(fen_ctx* c, bool ns) {
    if (!c) return set_error(FEN_ERR_ARG, "null context");
    cudaSetDevice(c->device);
    if (!c->mf || !c->mf->vof_fields) return set_error(FEN_ERR_STATE, "allocate_vof_fields has not been called");
    if (ns && !c->mf->ns) return set_error(FEN_ERR_STATE, "init_solver_mf has not been called");
    return FEN_OK;
}

static int fptr(fen_ctx* c, int id, double** out) {
    Field* f;
    FEN_TRY(field_check(c, id, &f));
    *out = f->d;
    return FEN_OK;
}

// Periodic -> 0, Wall -> 2 on the four faces (volume_of_fluid.f90:68-198, multiphase.f90:60-100); anything else
// prints the reference's error and keeps the default
static void wire_mf_bc(fen_ctx* c, const int* ids, int n) {
    static const char* face_name[4] = {"left", "right", "bottom", "top"};
    for (int face = 0; face < 4; ++face) {
        const int s = c->g.bc[face];
        int t;
        if (s == FEN_BC_PERIODIC) t = FEN_PERIODIC;
        else if (s == FEN_BC_WALL) t = FEN_NEUMANN;
        else { fprintf(stderr, "ERROR: wrong bc on %s boundary\n", face_name[face]); continue; }
        for (int q = 0; q < n; ++q) c->fields[ids[q]].bc_type[face] = t;
    }
}

static int recon(fen_ctx* c, int src_id) {
    // get_h_from_vof (:228-303): compute_norm (+ ghosts of curv, norm, l), then h and d (+ ghosts)
    const fen_mf_params& m = c->mf->prm;
    ReconArgs a;
    a.L = c->L;
    double* src;
    FEN_TRY(fptr(c, src_id, &src));
    a.vof = src;
    FEN_TRY(fptr(c, FEN_NORMX, &a.nx)); FEN_TRY(fptr(c, FEN_NORMY, &a.ny));
    FEN_TRY(fptr(c, FEN_LX, &a.lx)); FEN_TRY(fptr(c, FEN_LY, &a.ly));
    FEN_TRY(fptr(c, FEN_CURV, &a.curv)); FEN_TRY(fptr(c, FEN_H, &a.h)); FEN_TRY(fptr(c, FEN_D, &a.d));
    a.delta = c->g.delta;
    a.idelta = 1.0 / a.delta;
    a.idelta2 = 1.0 / (a.delta * a.delta);
    a.beta = m.beta; a.cut = m.cut; a.quadratic = m.quadratic;
    const int ids[7] = {FEN_CURV, FEN_NORMX, FEN_NORMY, FEN_LX, FEN_LY, FEN_H, FEN_D};
    const bool xp = x_periodic(c, ids, 7);
    a.xper = xp ? 1 : 0;
    // FEN_VOF_TILE=0: one thread evaluates all four corners of its cell (the first version; cross-check switch)
    static const bool tile = !getenv("FEN_VOF_TILE") || atoi(getenv("FEN_VOF_TILE")) != 0;
    static_assert(VT_X == MTX && VT_Y == MTY, "the tiled reconstruction uses the 64 x 4 blocks of mf_grid");
    if (tile) FEN_LAUNCH(c, "vof_recon", k_vof_recon_tile<<<mf_grid(c->L), dim3(MTX, MTY), 0, c->stream>>>(a));
    else FEN_LAUNCH(c, "vof_recon", k_vof_recon<<<mf_grid(c->L), dim3(MTX, MTY), 0, c->stream>>>(a));
    FEN_CUDA(cudaGetLastError());
    // curv, norm, l (:392-394) and h, d (:299-300): seven fields, one launch per direction
    return ghost_update_list(c, ids, 7, xp);
}

static int sweep(fen_ctx* c, int dir, bool final, int src_id, int out_id, int vx, double dt, bool x_first) {
    const fen_mf_params& m = c->mf->prm;
    SweepArgs a;
    a.L = c->L;
    double *src, *nx, *ny, *lx, *ly, *d, *u, *v;
    FEN_TRY(fptr(c, src_id, &src));
    FEN_TRY(fptr(c, FEN_NORMX, &nx)); FEN_TRY(fptr(c, FEN_NORMY, &ny));
    FEN_TRY(fptr(c, FEN_LX, &lx)); FEN_TRY(fptr(c, FEN_LY, &ly)); FEN_TRY(fptr(c, FEN_D, &d));
    FEN_TRY(fptr(c, vx, &u)); FEN_TRY(fptr(c, vx + 1, &v));
    FEN_TRY(fptr(c, out_id, &a.out));
    a.src = src; a.nx = nx; a.ny = ny; a.lx = lx; a.ly = ly; a.d = d; a.u = u; a.v = v;
    a.dt = dt; a.delta = c->g.delta; a.beta = m.beta; a.cut = m.cut; a.x_first = x_first ? 1 : 0;
    a.xper = x_periodic(c, &out_id, 1) ? 1 : 0;       // the caller's ghost update of out_id skips its x pass then
    const dim3 g = mf_grid(c->L), b(MTX, MTY);
    const dim3 gy(g.x, (c->L.ny + MTY * SWEEP_RY - 1) / (MTY * SWEEP_RY), 1);   // y sweep: SWEEP_RY rows per thread
    if (dir == 1 && !final) FEN_LAUNCH(c, "vof_sweep", k_vof_sweep<1, false><<<g, b, 0, c->stream>>>(a));
    if (dir == 2 && !final) FEN_LAUNCH(c, "vof_sweep", k_vof_sweep<2, false><<<gy, b, 0, c->stream>>>(a));
    if (dir == 1 && final) FEN_LAUNCH(c, "vof_sweep", k_vof_sweep<1, true><<<g, b, 0, c->stream>>>(a));
    if (dir == 2 && final) FEN_LAUNCH(c, "vof_sweep", k_vof_sweep<2, true><<<gy, b, 0, c->stream>>>(a));
    FEN_CUDA(cudaGetLastError());
    return FEN_OK;
}

// advect_vof(v, dt), volume_of_fluid.f90:434-554.  vof1 is the hidden field FEN_VOF1, freshly "allocated" on every
// call (default boundary types); vof2 never exists in memory (it is a per-cell temporary of the final sweep) and the
// result is written straight into vof's own buffer, so no pointer changes hands.
static int advect_vof(fen_ctx* c, int vx, double dt) {
    fen_mf_params& m = c->mf->prm;
    Field* v0;
    FEN_TRY(field_check(c, vx, &v0));
    if (v0->gl < 1) return set_error(FEN_ERR_ARG, "advect_vof: the velocity needs ghost nodes");
    const bool xf = m.x_first != 0;
    FEN_TRY(recon(c, FEN_VOF));                                            // :456
    {   // call vof1%allocate(vof%G, 1), :448: default (periodic) boundary types, zero values
        Field& t = c->fields[FEN_VOF1];
        for (int q = 0; q < 6; ++q) { t.bc_type[q] = FEN_PERIODIC; t.bc_mode[q] = BC_ZERO; t.bc_value[q] = 0.0; }
    }
    FEN_TRY(sweep(c, xf ? 1 : 2, false, FEN_VOF, FEN_VOF1, vx, dt, xf));   // :458-466 / :502-506
    {   // `vof = vof1` (:470, :510) is a derived-type assignment: vof takes vof1's boundary types too (hazard H13)
        Field& f = c->fields[FEN_VOF];
        const Field& t = c->fields[FEN_VOF1];
        for (int q = 0; q < 6; ++q) {
            f.bc_type[q] = t.bc_type[q]; f.bc_mode[q] = BC_ZERO; f.bc_value[q] = 0.0;
        }
    }
    {
        const int id1 = FEN_VOF1;                                          // :471 (the data of vof is vof1's)
        FEN_TRY(ghost_update(c, FEN_VOF1, 1, x_periodic(c, &id1, 1)));
    }
    FEN_TRY(recon(c, FEN_VOF1));                                           // :472
    FEN_TRY(sweep(c, xf ? 2 : 1, true, FEN_VOF1, FEN_VOF, vx, dt, xf));    // :478-498 / :516-538
    m.x_first = xf ? 0 : 1;                                                // :500 / :540
    const int id0 = FEN_VOF;
    return ghost_update(c, FEN_VOF, 1, x_periodic(c, &id0, 1));            // :546
}

static int material_properties(fen_ctx* c, bool with_phat, double dt) {
    const fen_mf_params& m = c->mf->prm;
    PropsArgs a;
    a.L = c->L;
    double *vof, *p = nullptr, *po = nullptr;
    FEN_TRY(fptr(c, FEN_VOF, &vof));
    a.vof = vof;
    c->uniform_props = false;                     // rho and mu are genuine fields from here on (hazard H11)
    FEN_TRY(fptr(c, FEN_RHO, &a.rho));
    FEN_TRY(fptr(c, FEN_MU, &a.mu));
    a.rho_0 = m.rho_0; a.rho_1 = m.rho_1; a.mu_0 = m.mu_0; a.mu_1 = m.mu_1;
    a.p_hat = nullptr;
    a.ccfl = 0; a.dt = dt; a.dt_o = c->prm.dt_o;
    if (with_phat) {
        FEN_TRY(fptr(c, FEN_P, &p)); FEN_TRY(fptr(c, FEN_PO, &po)); FEN_TRY(fptr(c, FEN_PHAT, &a.p_hat));
        a.ccfl = c->prm.constant_CFL ? 1 : 0;
    }
    a.p = p; a.p_o = po;
    const int ids[3] = {FEN_RHO, FEN_MU, FEN_PHAT};                        // multiphase.f90:134-135, navier_stokes.f90:95
    const bool xp = x_periodic(c, ids, with_phat ? 3 : 2);
    a.xper = xp ? 1 : 0;
    FEN_LAUNCH(c, "mf_props", k_mf_props<<<mf_grid(c->L), dim3(MTX, MTY), 0, c->stream>>>(a));
    FEN_CUDA(cudaGetLastError());
    return ghost_update_list(c, ids, with_phat ? 3 : 2, xp);
}

int mf_step_front(fen_ctx* c, double dt) {
    FEN_TRY(advect_vof(c, FEN_VX, dt));                                    // navier_stokes.f90:82
    return material_properties(c, true, dt);                               // :85-95
}

int mf_predict(fen_ctx* c, double dt) {
    const fen_mf_params& m = c->mf->prm;
    MfPredArgs a;
    a.L = c->L;
    double *u, *v, *p, *ph, *rho, *mu, *vof, *curv;
    FEN_TRY(fptr(c, FEN_VX, &u)); FEN_TRY(fptr(c, FEN_VY, &v)); FEN_TRY(fptr(c, FEN_P, &p));
    FEN_TRY(fptr(c, FEN_PHAT, &ph)); FEN_TRY(fptr(c, FEN_RHO, &rho)); FEN_TRY(fptr(c, FEN_MU, &mu));
    FEN_TRY(fptr(c, FEN_VOF, &vof)); FEN_TRY(fptr(c, FEN_CURV, &curv));
    a.u = u; a.v = v; a.p = p; a.p_hat = ph; a.rho = rho; a.mu = mu; a.vof = vof; a.curv = curv;
    a.sx = a.sy = nullptr;
    if (c->has_source) {
        double *sx, *sy;
        FEN_TRY(fptr(c, FEN_SX, &sx)); FEN_TRY(fptr(c, FEN_SY, &sy));
        a.sx = sx; a.sy = sy;
    }
    FEN_TRY(fptr(c, FEN_DVOX, &a.dvox)); FEN_TRY(fptr(c, FEN_DVOY, &a.dvoy));
    if (!c->vnew[0] || !c->vnew[1]) return set_error(FEN_ERR_STATE, "init_solver has not allocated the predictor buffers");
    a.un = c->vnew[0]; a.vn = c->vnew[1];
    a.k.id = 1.0 / c->g.delta;
    a.k.dt = dt;
    a.k.A = 1.0 + 0.5 * dt / c->prm.dt_o;             // :157
    a.k.B = -0.5 * dt / c->prm.dt_o;                  // :158
    a.k.g0 = c->prm.g[0]; a.k.g1 = c->prm.g[1];
    a.k.sigma = m.sigma; a.k.irhomin = m.irhomin;
    a.k.has_source = c->has_source ? 1 : 0;
    const int vids[2] = {FEN_VX, FEN_VY};
    const bool xp = x_periodic(c, vids, 2);
    a.xper = xp ? 1 : 0;
    FEN_LAUNCH(c, "mf_pred", k_mf_pred<<<mf_grid(c->L), dim3(MTX, MTY), 0, c->stream>>>(a));
    FEN_CUDA(cudaGetLastError());
    for (int q = 0; q < 2; ++q) std::swap(c->fields[FEN_VX + q].d, c->vnew[q]);
    return ghost_update(c, FEN_VX, 2, xp);            // :208
}

int mf_poisson_rhs(fen_ctx* c, double dt) {
    double *u, *v, *phi;
    FEN_TRY(fptr(c, FEN_VX, &u)); FEN_TRY(fptr(c, FEN_VY, &v)); FEN_TRY(fptr(c, FEN_PHI, &phi));
    FEN_LAUNCH(c, "poisson_rhs", k_mf_rhs<<<mf_grid(c->L), dim3(MTX, MTY), 0, c->stream>>>(
                                     c->L, u, v, phi, 1.0 / c->g.delta, c->mf->prm.rhomin, dt));
    FEN_CUDA(cudaGetLastError());
    return FEN_OK;
}

int mf_correct(fen_ctx* c, double dt) {
    double *u, *v, *p, *po, *phi;
    FEN_TRY(fptr(c, FEN_VX, &u)); FEN_TRY(fptr(c, FEN_VY, &v)); FEN_TRY(fptr(c, FEN_P, &p));
    FEN_TRY(fptr(c, FEN_PO, &po)); FEN_TRY(fptr(c, FEN_PHI, &phi));
    // v (:544), p (:564) and p_o (p_o%f = p%f copies p's ghosts too, :557; same boundary types as p)
    const int ids[4] = {FEN_VX, FEN_VY, FEN_PO, FEN_P};
    const bool xp = x_periodic(c, ids, 4);
    FEN_LAUNCH(c, "corr", k_mf_corr<<<mf_grid(c->L), dim3(MTX, MTY), 0, c->stream>>>(
                              c->L, u, v, p, po, phi, 1.0 / c->g.delta, dt, c->mf->prm.irhomin, xp ? 1 : 0));
    FEN_CUDA(cudaGetLastError());
    return ghost_update_list(c, ids, 4, xp);
}

int mf_set_timestep(fen_ctx* c, double U, double* dt) {
    // set_timestep with -DMF, navier_stokes.f90:641, 655-664 (2-D)
    fen_mf_params& m = c->mf->prm;
    fen_ns_params& p = c->prm;
    const double d = c->g.delta;
    p.dt_conv = p.CFL * d / U;
    p.dt_visc = 0.125 * d * d * std::min(m.rho_0 / m.mu_0, m.rho_1 / m.mu_1);
    *dt = std::min(p.dt_conv, p.dt_visc);
    if (m.sigma > 0.0) {
        const double pi = acos(-1.0);                                        // global.f90:14
        m.dt_surf = sqrt(0.5 * (m.rho_0 + m.rho_1) * (d * d * d) / (pi * m.sigma + 1.0e-16));
        *dt = std::min(*dt, m.dt_surf);
    }
    p.dt_o = *dt;
    return FEN_OK;
}

void mf_destroy(fen_ctx* c) {
    if (!c->mf) return;
    for (int id = FEN_VOF; id <= FEN_VOF1; ++id) { free_field(c->fields[id]); c->fields[id].exists = false; }
    delete c->mf;
    c->mf = nullptr;
}

}  // namespace fen

using namespace fen;

extern "C" {

int fen_gpu_mf_get_params(fen_ctx* c, fen_mf_params* p) {
    if (!c || !p) return set_error(FEN_ERR_ARG, "null argument");
    if (!c->mf) {
        c->mf = new Multiphase();
        fen_mf_params& m = c->mf->prm;
        m.rho_0 = m.rho_1 = m.mu_0 = m.mu_1 = 1.0;                  // multiphase.f90:18
        m.sigma = 0.0;                                             // multiphase.f90:21
        m.beta = 1.0; m.cut = 1.0e-8; m.quadratic = 1; m.x_first = 1;   // volume_of_fluid.f90:24-37
        m.dt_surf = std::numeric_limits<double>::infinity();        // uninitialised in the reference (hazard H16)
        m.rhomin = m.irhomin = 1.0;
    }
    *p = c->mf->prm;
    return FEN_OK;
}

int fen_gpu_mf_set_params(fen_ctx* c, const fen_mf_params* p) {
    fen_mf_params cur;
    if (!p) return set_error(FEN_ERR_ARG, "null argument");
    FEN_TRY(fen_gpu_mf_get_params(c, &cur));
    c->mf->prm = *p;
    return FEN_OK;
}

int fen_gpu_allocate_vof_fields(fen_ctx* c) {
    fen_mf_params cur;
    FEN_TRY(fen_gpu_mf_get_params(c, &cur));          // creates the module state with the reference's defaults
    if (c->g.ndim != 2)
        return set_error(FEN_ERR_UNSUPPORTED, "the volume-of-fluid solver is 2-D only, as in the reference "
                                              "(volume_of_fluid.f90:307-396 has no z terms)");
    FEN_CUDA(cudaSetDevice(c->device));
    // vof, h, d, curv: cell centred; norm, l: vectors (components tagged x, y), all with one ghost layer (:60-65)
    const int ids[9] = {FEN_VOF, FEN_H, FEN_D, FEN_CURV, FEN_NORMX, FEN_NORMY, FEN_LX, FEN_LY, FEN_VOF1};
    const int loc[9] = {FEN_LOC_C, FEN_LOC_C, FEN_LOC_C, FEN_LOC_C, FEN_LOC_X, FEN_LOC_Y, FEN_LOC_X, FEN_LOC_Y, FEN_LOC_C};
    for (int q = 0; q < 9; ++q) {
        init_field(c, ids[q], 1, loc[q]);
        FEN_TRY(field_alloc(c, c->fields[ids[q]]));
    }
    wire_mf_bc(c, ids, 8);                            // vof1 keeps the defaults
    c->mf->vof_fields = true;
    return FEN_OK;
}

int fen_gpu_destroy_vof(fen_ctx* c) {
    if (!c) return set_error(FEN_ERR_ARG, "null context");
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    mf_destroy(c);
    return FEN_OK;
}

int fen_gpu_get_vof_from_distance(fen_ctx* c, fen_distance_fn fn, void* user, double x0, double y0) {
    FEN_TRY(mf_need(c, false));
    if (!fn) {
        fprintf(stderr, "ERROR: distance function not defined.\n");       // volume_of_fluid.f90:688
        return set_error(FEN_ERR_ARG, "distance function not defined");
    }
    const Layout& L = c->L;
    const double delta = c->g.delta, beta = c->mf->prm.beta;
    const size_t n = (size_t)L.nx * L.ny;
    std::vector<double> dist(5 * n);
    for (int j = 1; j <= L.ny; ++j) {
        const double y = y0 + (j - 0.5) * delta;                          // grid.f90:155-164
        const double yp = y + delta * (VOF_RP - 0.5), ym = y + delta * (VOF_RM - 0.5);
        for (int i = 1; i <= L.nx; ++i) {
            const double x = x0 + (i - 0.5) * delta;
            const double xp = x + delta * (VOF_RP - 0.5), xm = x + delta * (VOF_RM - 0.5);
            const size_t o = (size_t)(j - 1) * L.nx + (i - 1);
            dist[0 * n + o] = fn(user, xm, ym);
            dist[1 * n + o] = fn(user, xp, ym);
            dist[2 * n + o] = fn(user, xm, yp);
            dist[3 * n + o] = fn(user, xp, yp);
            dist[4 * n + o] = fn(user, x, y);
        }
    }
    double* d_dist = nullptr;
    FEN_CUDA(cudaMalloc(&d_dist, 5 * n * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(d_dist, dist.data(), 5 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    double *vof = nullptr, *h = nullptr;
    int r = e == cudaSuccess ? FEN_OK : set_error(FEN_ERR_CUDA, "copy of the distances failed: %s", cudaGetErrorString(e));
    if (r == FEN_OK) r = fptr(c, FEN_VOF, &vof);
    if (r == FEN_OK) r = fptr(c, FEN_H, &h);
    if (r == FEN_OK) {
        FEN_LAUNCH(c, "vof_from_dist", k_vof_from_dist<<<mf_grid(L), dim3(MTX, MTY), 0, c->stream>>>(L, d_dist, vof, h, beta, delta));
        if (cudaGetLastError() != cudaSuccess) r = set_error(FEN_ERR_CUDA, "k_vof_from_dist launch failed");
    }
    cudaStreamSynchronize(c->stream);
    cudaFree(d_dist);
    if (r != FEN_OK) return r;
    FEN_TRY(ghost_update(c, FEN_VOF, 1));                                  // :714-715
    return ghost_update(c, FEN_H, 1);
}

int fen_gpu_get_h_from_vof(fen_ctx* c) {
    FEN_TRY(mf_need(c, false));
    return recon(c, FEN_VOF);
}

int fen_gpu_advect_vof(fen_ctx* c, int vector_x, double dt) {
    FEN_TRY(mf_need(c, false));
    return advect_vof(c, vector_x, dt);
}

int fen_gpu_check_vof_integral(fen_ctx* c, double* i1, double* i2) {
    FEN_TRY(mf_need(c, false));
    if (!i1 || !i2) return set_error(FEN_ERR_ARG, "null argument");
    double* vof;
    FEN_TRY(fptr(c, FEN_VOF, &vof));
    FEN_TRY(reduce_field(c, vof, 1, c->d_red));
    FEN_TRY(fetch_red(c, 1));
    FEN_CUDA(cudaStreamSynchronize(c->stream));
    const double d = c->g.delta, ncell = (double)c->L.nx * (double)c->L.ny;
    const double s = c->h_red[0];
    *i1 = s * d * d * d;                               // :750 (delta**3 although the case is 2-D)
    *i2 = (ncell - s) * d * d * d;                     // sum(1 - vof), :744
    return FEN_OK;
}

int fen_gpu_update_material_properties(fen_ctx* c) {
    FEN_TRY(mf_need(c, false));
    if (!c->solver_init) return set_error(FEN_ERR_STATE, "init_solver has not been called");
    return material_properties(c, false, 0.0);
}

int fen_gpu_init_solver_mf(fen_ctx* c, fen_distance_fn fn, void* user, double x0, double y0) {
    if (!c) return set_error(FEN_ERR_ARG, "null context");
    if (c->g.ndim != 2)
        return set_error(FEN_ERR_UNSUPPORTED, "the two-phase solver is 2-D only, as in the reference");
    fen_mf_params keep;
    FEN_TRY(fen_gpu_mf_get_params(c, &keep));          // rho_0 ... sigma, beta set by the driver before init_solver
    FEN_TRY(fen_gpu_init_solver(c));                   // solver.f90:58-61
    FEN_TRY(fen_gpu_allocate_vof_fields(c));           // :87
    // allocate_multiphase_fields, multiphase.f90:46-117 (grad_p_hat never exists here: the predictor differences p_hat)
    const int ids[2] = {FEN_PHAT, FEN_PO};
    for (int id : ids) {
        init_field(c, id, 1, FEN_LOC_C);
        FEN_TRY(field_alloc(c, c->fields[id]));
    }
    wire_mf_bc(c, ids, 2);
    if (fn) FEN_TRY(fen_gpu_get_vof_from_distance(c, fn, user, x0, y0));    // :89
    else fprintf(stderr, "ERROR: distance function not defined.\n");
    FEN_TRY(material_properties(c, false, 0.0));       // :90
    fen_mf_params& m = c->mf->prm;
    m.rhomin = std::min(m.rho_0, m.rho_1);             // :92
    m.irhomin = 1.0 / m.rhomin;                        // :93
    c->mf->ns = true;
    return FEN_OK;
}

}  // extern "C"
