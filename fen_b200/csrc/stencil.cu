// stencil.cu -- the staggered-grid finite-difference kernels of the fractional step.
//
//   k_pred   : predicted_velocity_field (src/navier_stokes.f90:140-213) with everything it calls
//              fused into one pass: center_to_face (fields.f90:175), add_advection (:261-353),
//              add_diffusion const-mu branch (:384-404 -> fields.f90:298-343), body force (:245-255),
//              gradient(p) (fields.f90:31), RHS assembly (:169-172), v += dt*RHS (:187-198),
//              dv_o = dv (:201-205).  Reads u,v,w,p,dv_o once, writes u*,v*,w*,dv_o once
//              (104 B/cell in the uniform-property case).
//   k_rhs    : divergence(v, phi); phi = phi*rho/dt      (fields.f90:120-153, navier_stokes.f90:111-121)
//   k_corr   : gradient(phi); v -= grad*dt/rhof; p += phi (navier_stokes.f90:505-541, 550-561)
//   k_check  : maxdiv (signed max) and max(|u|+|v|+|w|)   (navier_stokes.f90:570-619)
// plus the stand-alone field operators of src/fields.f90 used by the reference's unit tests.
//
// All kernels map threadIdx.x to the x (fastest) index so every warp reads/writes contiguous
// 256-byte runs; a block owns a (TX x TY) column of cells and marches KZ planes in z so that the
// three z-planes a stencil needs stay in L1 while the 126 MB L2 holds the neighbouring tiles.
#include <cstdlib>

#include "fen_internal.cuh"
#include "tma.cuh"

namespace fen {

constexpr int TX = 64, TY = 4, KZ = 16;

struct StArgs {
    Layout L;
    const double* u; const double* v; const double* w; const double* p;
    const double* rho; const double* mu;           // fields (general path) or nullptr
    const double* sx; const double* sy_; const double* sz_;
    double* dvox; double* dvoy; double* dvoz;
    double* un; double* vn; double* wn;
    double rho0, mu0;                               // uniform values
    double idelta, idelta2, dt, A, B, g0, g1, g2;
    int xper;                                       // x periodic: the kernel also writes the x ghosts of its rows
};

__device__ __forceinline__ double sq(double x) { return x * x; }

// The 27 velocity values the explicit terms of one cell read.
struct Sten {
    double u0, uip, uim, ujp, ujm, ukp, ukm, uimjp, uimkp;
    double v0, vip, vim, vjp, vjm, vkp, vkm, vipjm, vjmkp;
    double w0, wip, wim, wjp, wjm, wkp, wkm, wipkm, wjpkm;
};

// advection (navier_stokes.f90:297-347) + diffusion (fields.f90:326-337, navier_stokes.f90:394-397) from the
// stencil values; the one place where this arithmetic lives.  UNIT: rhof == 1 exactly, x/1 == x, so the
// divisions are dropped without changing a bit.
template <bool D3, bool ADV_ONLY, bool UNIT>
__device__ __forceinline__ void explicit_from(const Sten& s, double id, double id2, double mu, double rfx, double rfy,
                                              double rfz, double& dvx, double& dvy, double& dvz) {
    double uuip = 0.25 * sq(s.uip + s.u0);
    double uuim = 0.25 * sq(s.uim + s.u0);
    double uvjp = (s.ujp + s.u0) * (s.vip + s.v0) * 0.25;
    double uvjm = (s.u0 + s.ujm) * (s.vipjm + s.vjm) * 0.25;
    dvx = 0.0 - (uuip - uuim) * id - (uvjp - uvjm) * id;
    double vuip = (s.vip + s.v0) * (s.ujp + s.u0) * 0.25;
    double vuim = (s.v0 + s.vim) * (s.uimjp + s.uim) * 0.25;
    double vvjp = 0.25 * sq(s.vjp + s.v0);
    double vvjm = 0.25 * sq(s.vjm + s.v0);
    dvy = 0.0 - (vuip - vuim) * id - (vvjp - vvjm) * id;
    dvz = 0.0;
    if (D3) {
        double uwkp = (s.ukp + s.u0) * (s.wip + s.w0) * 0.25;
        double uwkm = (s.u0 + s.ukm) * (s.wipkm + s.wkm) * 0.25;
        dvx = dvx - (uwkp - uwkm) * id;
        double vwkp = (s.vkp + s.v0) * (s.wjp + s.w0) * 0.25;
        double vwkm = (s.v0 + s.vkm) * (s.wjpkm + s.wkm) * 0.25;
        dvy = dvy - (vwkp - vwkm) * id;
        double wuip = (s.w0 + s.wip) * (s.u0 + s.ukp) * 0.25;
        double wuim = (s.w0 + s.wim) * (s.uim + s.uimkp) * 0.25;
        double wvjp = (s.w0 + s.wjp) * (s.v0 + s.vkp) * 0.25;
        double wvjm = (s.w0 + s.wjm) * (s.vjm + s.vjmkp) * 0.25;
        double wwkp = (s.w0 + s.wkp) * (s.w0 + s.wkp) * 0.25;
        double wwkm = (s.w0 + s.wkm) * (s.w0 + s.wkm) * 0.25;
        dvz = 0.0 - (wuip - wuim) * id - (wvjp - wvjm) * id - (wwkp - wwkm) * id;
    }
    if (ADV_ONLY) return;
    double lx = ((s.uip - 2.0 * s.u0 + s.uim) + (s.ujp - 2.0 * s.u0 + s.ujm)) * id2;
    double ly = ((s.vip - 2.0 * s.v0 + s.vim) + (s.vjp - 2.0 * s.v0 + s.vjm)) * id2;
    if (D3) {
        lx = lx + (s.ukp - 2.0 * s.u0 + s.ukm) * id2;
        ly = ly + (s.vkp - 2.0 * s.v0 + s.vkm) * id2;
        double lz = ((s.wip - 2.0 * s.w0 + s.wim) + (s.wjp - 2.0 * s.w0 + s.wjm) + (s.wkp - 2.0 * s.w0 + s.wkm)) * id2;
        dvz = dvz + (UNIT ? mu * lz : mu * lz / rfz);
    }
    dvx = dvx + (UNIT ? mu * lx : mu * lx / rfx);
    dvy = dvy + (UNIT ? mu * ly : mu * ly / rfy);
}

// explicit terms dv (advection + diffusion [+ source]) at one cell, values read from global memory
// (general-property / 2-D / stand-alone operator path); also returns rhof.
template <bool D3, bool GEN, bool ADV_ONLY>
__device__ __forceinline__ void explicit_terms(const StArgs& a, long long c, double& dvx, double& dvy,
                                               double& dvz, double& rfx, double& rfy, double& rfz) {
    const long long sy = a.L.sy, sz = a.L.sz;
    const double* __restrict__ u = a.u;
    const double* __restrict__ v = a.v;
    const double* __restrict__ w = a.w;
    Sten s;
    s.u0 = u[c]; s.uip = u[c + 1]; s.uim = u[c - 1]; s.ujp = u[c + sy]; s.ujm = u[c - sy];
    s.v0 = v[c]; s.vip = v[c + 1]; s.vim = v[c - 1]; s.vjp = v[c + sy]; s.vjm = v[c - sy];
    s.uimjp = u[c - 1 + sy]; s.vipjm = v[c + 1 - sy];
    s.ukp = s.ukm = s.vkp = s.vkm = s.uimkp = s.vjmkp = 0.0;
    s.w0 = s.wip = s.wim = s.wjp = s.wjm = s.wkp = s.wkm = s.wipkm = s.wjpkm = 0.0;
    if (D3) {
        s.ukp = u[c + sz]; s.ukm = u[c - sz]; s.vkp = v[c + sz]; s.vkm = v[c - sz];
        s.w0 = w[c]; s.wip = w[c + 1]; s.wim = w[c - 1]; s.wjp = w[c + sy]; s.wjm = w[c - sy];
        s.wkp = w[c + sz]; s.wkm = w[c - sz];
        s.wipkm = w[c + 1 - sz]; s.wjpkm = w[c + sy - sz];
        s.uimkp = u[c - 1 + sz]; s.vjmkp = v[c - sy + sz];
    }
    // face densities, fields.f90:197-200
    double mu = a.mu0;
    rfx = rfy = rfz = 1.0;
    if (!ADV_ONLY) {
        if (GEN) {
            const double* __restrict__ rho = a.rho;
            const double r0 = rho[c];
            rfx = 0.5 * (rho[c + 1] + r0);
            rfy = 0.5 * (rho[c + sy] + r0);
            rfz = D3 ? 0.5 * (rho[c + sz] + r0) : 1.0;
            mu = a.mu[c];
        } else {
            rfx = rfy = rfz = 0.5 * (a.rho0 + a.rho0);
        }
    }
    explicit_from<D3, ADV_ONLY, false>(s, a.idelta, a.idelta2, mu, rfx, rfy, rfz, dvx, dvy, dvz);
    if (ADV_ONLY) return;
    // body force, navier_stokes.f90:248-251
    if (GEN && a.sx) {
        dvx = dvx + a.sx[c] / rfx;
        dvy = dvy + a.sy_[c] / rfy;
        if (D3) dvz = dvz + a.sz_[c] / rfz;
    }
}

template <bool D3, bool GEN>
__global__ void __launch_bounds__(TX* TY) k_pred(StArgs a) {
    const int i = blockIdx.x * TX + threadIdx.x + 1;
    const int j = blockIdx.y * TY + threadIdx.y + 1;
    if (i > a.L.nx || j > a.L.ny) return;
    const int kb = blockIdx.z * KZ + 1;
    const int ke = min(kb + KZ - 1, a.L.nzl);
    const double id = a.idelta, dt = a.dt;
    const double* __restrict__ p = a.p;
    for (int k = kb; k <= ke; ++k) {
        const long long c = a.L.idx(i, j, k);
        double dvx, dvy, dvz, rfx, rfy, rfz;
        explicit_terms<D3, GEN, false>(a, c, dvx, dvy, dvz, rfx, rfy, rfz);
        const double p0 = p[c];
        // RHS = -grad_p/rhof + A*dv + B*dv_o + g  (navier_stokes.f90:169-172)
        double gx = (p[c + 1] - p0) * id;
        double gy = (p[c + a.L.sy] - p0) * id;
        double rx = -gx / rfx + a.A * dvx + a.B * a.dvox[c] + a.g0;
        double ry = -gy / rfy + a.A * dvy + a.B * a.dvoy[c] + a.g1;
        a.un[c] = a.u[c] + dt * rx;
        a.vn[c] = a.v[c] + dt * ry;
        a.dvox[c] = dvx;
        a.dvoy[c] = dvy;
        if (D3) {
            double gz = (p[c + a.L.sz] - p0) * id;
            double rz = -gz / rfz + a.A * dvz + a.B * a.dvoz[c] + a.g2;
            a.wn[c] = a.w[c] + dt * rz;
            a.dvoz[c] = dvz;
        }
    }
}

// ---- 3-D uniform-property predictor with TMA-staged tiles --------------------------------------------
// A block owns a PX x PY column of cells and marches PKZ planes in z.  For every plane the copy engine
// drops the (PX+4) x (PY+2) box (tile + halo; the box must start on a 16-byte boundary -- measured with
// scripts/probes/tma_probe.cu: an odd fp64 start coordinate is an illegal instruction -- so it starts two
// cells left of the tile) of u, v and w into a 4-deep shared-memory ring
// (cp.async.bulk.tensor, one elected thread, mbarrier completion), two planes ahead of the arithmetic, so
// the 27 stencil values of a cell are shared-memory reads and each of u, v, w leaves HBM once.  p and dv_o
// are plain coalesced loads, prefetched one plane ahead in registers.
constexpr int PX = 64, PY = 8, PKZ = 32;
constexpr int PTW = PX + 4, PTH = PY + 2;
constexpr int PTILE = PTW * PTH;                                  // doubles per box
constexpr int PTILE_B = (PTILE * 8 + 127) / 128 * 128;            // bytes per box slot (128-byte aligned for TMA)
constexpr int PRED_SMEM = 12 * PTILE_B + 64;

template <bool UNIT, bool XG>
__global__ void __launch_bounds__(PX* PY, 2)
k_pred_tma(const __grid_constant__ CUtensorMap mu_, const __grid_constant__ CUtensorMap mv_,
           const __grid_constant__ CUtensorMap mw_, StArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 12 * PTILE_B);
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * PX + tx;
    const int i = blockIdx.x * PX + tx + 1;
    const int j = blockIdx.y * PY + ty + 1;
    const int kb = blockIdx.z * PKZ + 1;
    const int ke = min(kb + PKZ - 1, a.L.nzl);
    const int x0 = a.L.xoff - 2 + blockIdx.x * PX;     // box origin: cell (i0 - 2, j0 - 1), an even element index
    const int y0 = blockIdx.y * PY;
    if (tid == 0) {
        for (int q = 0; q < 4; ++q) mbar_init(bar + q, 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto slot = [&](int q, int f) { return reinterpret_cast<double*>(smem + (size_t)((q & 3) * 3 + f) * PTILE_B); };
    auto issue = [&](int plane) {          // plane: local z index kb-1 .. ke+1
        const int q = plane - (kb - 1);
        uint64_t* b = bar + (q & 3);
        mbar_expect_tx(b, 3u * PTILE * 8u);
        tma_load_3d(slot(q, 0), &mu_, x0, y0, plane, b);
        tma_load_3d(slot(q, 1), &mv_, x0, y0, plane, b);
        tma_load_3d(slot(q, 2), &mw_, x0, y0, plane, b);
    };
    if (tid == 0) {
        issue(kb - 1);
        issue(kb);
        issue(kb + 1);
    }
    const bool active = i <= a.L.nx && j <= a.L.ny;
    const long long sy = a.L.sy, sz = a.L.sz;
    long long c = a.L.idx(active ? i : 1, active ? j : 1, kb);
    const double* __restrict__ p = a.p;
    const double id = a.idelta, id2 = a.idelta2, dt = a.dt;
    const double rf = 0.5 * (a.rho0 + a.rho0);
    // own-column values prefetched one plane ahead
    double p0 = p[c], pkp = p[c + sz];
    double pip = p[c + 1], pjp = p[c + sy];
    double dox = a.dvox[c], doy = a.dvoy[c], doz = a.dvoz[c];
    mbar_wait(bar + 0, 0);
    mbar_wait(bar + 1, 0);
    const int o = (ty + 1) * PTW + tx + 2;             // my cell inside a box
    const int goff = !XG ? 0 : (i == 1 ? a.L.nx : (i == a.L.nx ? -a.L.nx : 0));   // my periodic image, if any
    for (int k = kb; k <= ke; ++k) {
        const int q = k - (kb - 1);                    // ring position of the centre plane
        mbar_wait(bar + ((q + 1) & 3), ((q + 1) >> 2) & 1);
        __syncthreads();                               // everybody is done with plane k-2: its slot is free
        if (tid == 0 && k + 2 <= ke + 1) issue(k + 2);
        // next plane's streamed values
        double n_pkp = 0.0, n_pip = 0.0, n_pjp = 0.0, n_dox = 0.0, n_doy = 0.0, n_doz = 0.0;
        if (k < ke) {
            const long long cn = c + sz;
            n_pkp = p[cn + sz]; n_pip = p[cn + 1]; n_pjp = p[cn + sy];
            n_dox = a.dvox[cn]; n_doy = a.dvoy[cn]; n_doz = a.dvoz[cn];
        }
        const double* um = slot(q - 1, 0); const double* uc = slot(q, 0); const double* up = slot(q + 1, 0);
        const double* vm = slot(q - 1, 1); const double* vc = slot(q, 1); const double* vp = slot(q + 1, 1);
        const double* wm = slot(q - 1, 2); const double* wc = slot(q, 2); const double* wp = slot(q + 1, 2);
        Sten s;
        s.u0 = uc[o]; s.uip = uc[o + 1]; s.uim = uc[o - 1]; s.ujp = uc[o + PTW]; s.ujm = uc[o - PTW];
        s.uimjp = uc[o - 1 + PTW]; s.ukp = up[o]; s.ukm = um[o]; s.uimkp = up[o - 1];
        s.v0 = vc[o]; s.vip = vc[o + 1]; s.vim = vc[o - 1]; s.vjp = vc[o + PTW]; s.vjm = vc[o - PTW];
        s.vipjm = vc[o + 1 - PTW]; s.vkp = vp[o]; s.vkm = vm[o]; s.vjmkp = vp[o - PTW];
        s.w0 = wc[o]; s.wip = wc[o + 1]; s.wim = wc[o - 1]; s.wjp = wc[o + PTW]; s.wjm = wc[o - PTW];
        s.wkp = wp[o]; s.wkm = wm[o]; s.wipkm = wm[o + 1]; s.wjpkm = wm[o + PTW];
        double dvx, dvy, dvz;
        explicit_from<true, false, UNIT>(s, id, id2, a.mu0, rf, rf, rf, dvx, dvy, dvz);
        // RHS = -grad_p/rhof + A*dv + B*dv_o + g  (navier_stokes.f90:169-172); v += dt*RHS (:187-198)
        const double gx = (pip - p0) * id, gy = (pjp - p0) * id, gz = (pkp - p0) * id;
        const double rx = (UNIT ? -gx : -gx / rf) + a.A * dvx + a.B * dox + a.g0;
        const double ry = (UNIT ? -gy : -gy / rf) + a.A * dvy + a.B * doy + a.g1;
        const double rz = (UNIT ? -gz : -gz / rf) + a.A * dvz + a.B * doz + a.g2;
        if (active) {
            a.un[c] = s.u0 + dt * rx;
            a.vn[c] = s.v0 + dt * ry;
            a.wn[c] = s.w0 + dt * rz;
            a.dvox[c] = dvx;                           // dv_o = dv (:201-205)
            a.dvoy[c] = dvy;
            a.dvoz[c] = dvz;
            if (XG && goff != 0) {                     // periodic x ghosts (scalar.f90:257,276)
                a.un[c + goff] = s.u0 + dt * rx;
                a.vn[c + goff] = s.v0 + dt * ry;
                a.wn[c + goff] = s.w0 + dt * rz;
            }
        }
        p0 = pkp; pkp = n_pkp; pip = n_pip; pjp = n_pjp;
        dox = n_dox; doy = n_doy; doz = n_doz;
        c += sz;
    }
}

// compute_explicit_terms / add_advection as stand-alone operators (reference tests call them)
template <bool D3, bool GEN, bool ADV_ONLY>
__global__ void __launch_bounds__(TX* TY) k_explicit(StArgs a) {
    const int i = blockIdx.x * TX + threadIdx.x + 1;
    const int j = blockIdx.y * TY + threadIdx.y + 1;
    if (i > a.L.nx || j > a.L.ny) return;
    const int kb = blockIdx.z * KZ + 1;
    const int ke = min(kb + KZ - 1, a.L.nzl);
    for (int k = kb; k <= ke; ++k) {
        const long long c = a.L.idx(i, j, k);
        double dvx, dvy, dvz, rfx, rfy, rfz;
        explicit_terms<D3, GEN, ADV_ONLY>(a, c, dvx, dvy, dvz, rfx, rfy, rfz);
        if (ADV_ONLY) {   // add_advection accumulates into RHS (navier_stokes.f90:305)
            a.un[c] += dvx;
            a.vn[c] += dvy;
            if (D3) a.wn[c] += dvz;
        } else {
            a.un[c] = dvx;
            a.vn[c] = dvy;
            if (D3) a.wn[c] = dvz;
        }
    }
}

struct RhsArgs {
    Layout L;
    const double* u; const double* v; const double* w;
    const double* rho; double rho0;
    double* phi;
    double idelta, dt;
    int scale;     // 1: phi = div * rho / dt ; 0: plain divergence
};

template <bool D3>
__device__ __forceinline__ double div_at(const Layout& L, const double* __restrict__ u,
                                         const double* __restrict__ v, const double* __restrict__ w,
                                         long long c, double id) {
    // fields.f90:144-147
    double d = (u[c] - u[c - 1]) * id + (v[c] - v[c - L.sy]) * id;
    if (D3) d = d + (w[c] - w[c - L.sz]) * id;
    return d;
}

template <bool D3>
__global__ void __launch_bounds__(TX* TY) k_rhs(RhsArgs a) {
    const int i = blockIdx.x * TX + threadIdx.x + 1;
    const int j = blockIdx.y * TY + threadIdx.y + 1;
    if (i > a.L.nx || j > a.L.ny) return;
    const int kb = blockIdx.z * KZ + 1;
    const int ke = min(kb + KZ - 1, a.L.nzl);
    for (int k = kb; k <= ke; ++k) {
        const long long c = a.L.idx(i, j, k);
        double d = div_at<D3>(a.L, a.u, a.v, a.w, c, a.idelta);
        if (a.scale) {
            double r = a.rho ? a.rho[c] : a.rho0;
            d = d * r / a.dt;      // navier_stokes.f90:118
        }
        a.phi[c] = d;
    }
}

struct CorrArgs {
    Layout L;
    double* u; double* v; double* w; double* p;
    const double* phi;
    const double* rho; double rho0;
    double idelta, dt;
};

template <bool D3>
__global__ void __launch_bounds__(TX* TY) k_corr(CorrArgs a) {
    const int i = blockIdx.x * TX + threadIdx.x + 1;
    const int j = blockIdx.y * TY + threadIdx.y + 1;
    if (i > a.L.nx || j > a.L.ny) return;
    const int kb = blockIdx.z * KZ + 1;
    const int ke = min(kb + KZ - 1, a.L.nzl);
    const double* __restrict__ phi = a.phi;
    const double id = a.idelta, dt = a.dt;
    for (int k = kb; k <= ke; ++k) {
        const long long c = a.L.idx(i, j, k);
        const double f0 = phi[c];
        double rfx, rfy, rfz;
        if (a.rho) {
            const double r0 = a.rho[c];
            rfx = 0.5 * (a.rho[c + 1] + r0);
            rfy = 0.5 * (a.rho[c + a.L.sy] + r0);
            rfz = D3 ? 0.5 * (a.rho[c + a.L.sz] + r0) : 1.0;
        } else {
            rfx = rfy = rfz = 0.5 * (a.rho0 + a.rho0);
        }
        // v = v - grad_p*dt/rhof   (navier_stokes.f90:533-536)
        a.u[c] = a.u[c] - ((phi[c + 1] - f0) * id) * dt / rfx;
        a.v[c] = a.v[c] - ((phi[c + a.L.sy] - f0) * id) * dt / rfy;
        if (D3) a.w[c] = a.w[c] - ((phi[c + a.L.sz] - f0) * id) * dt / rfz;
        a.p[c] = a.p[c] + f0;      // navier_stokes.f90:561 (ghosts are refreshed right after)
    }
}

// ---- 3-D uniform-property correction + pressure update + checks, TMA-staged ---------------------------
// correct_velocity_field (navier_stokes.f90:505-546), update_pressure (:550-566) and checks (:570-619) in one
// pass: reads phi, u*, v*, w*, p once and writes u, v, w, p once (72 B/cell); the divergence and velocity maxima
// of `checks` come out of the same registers, so the 24 B/cell re-read of u, v, w disappears.
// The divergence of the corrected field at (i,j,k) needs u(i-1), v(j-1), w(k-1): u(i-1) and v(j-1) are
// recomputed from the staged tiles with the same rounded operations the owning thread uses (corr_val, no FMA
// contraction), so they are bit-identical to the stored values; w(k-1) is the thread's own previous plane.
// Valid when x and y are periodic (ghost = periodic image, which the formula reproduces from the ghosts of u*
// and phi); z may be periodic, a rank boundary, or a wall with a zero/uniform normal velocity (w is then the
// wall value on the boundary faces, scalar.f90:355,377-378).  Output goes to the ping-pong buffers because
// neighbouring blocks still read u* from the halo of their boxes.
constexpr int CKZ = 32;
constexpr int CORR_SMEM = 16 * PTILE_B + 64;

struct CorrTArgs {
    Layout L;
    const double* p; double* pn;                   // p is updated in place (own cell only)
    double* un; double* vn; double* wn;
    double rho0, idelta, dt;
    int xper;                                      // also write the periodic x ghosts of u, v, w, p
    int wall_lo, wall_hi;                          // this rank owns the z wall on that side
    double wlo, whi;                               // wall-normal velocity there
    double* partial;                               // [2 * nblocks] (maxdiv, maxvel)
};

// v - ((phi_hi - phi_lo) * id) * dt / rhof   (navier_stokes.f90:533-536), every operation rounded once
template <bool UNIT>
__device__ __forceinline__ double corr_val(double vstar, double phi_hi, double phi_lo, double id, double dt, double rf) {
    double g = __dmul_rn(__dmul_rn(__dsub_rn(phi_hi, phi_lo), id), dt);
    if (!UNIT) g = __ddiv_rn(g, rf);
    return __dsub_rn(vstar, g);
}

template <bool UNIT>
__global__ void __launch_bounds__(PX* PY, 2)
k_corr_tma(const __grid_constant__ CUtensorMap mf_, const __grid_constant__ CUtensorMap mu_,
           const __grid_constant__ CUtensorMap mv_, const __grid_constant__ CUtensorMap mw_, CorrTArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16 * PTILE_B);
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * PX + tx;
    const int i = blockIdx.x * PX + tx + 1;
    const int j = blockIdx.y * PY + ty + 1;
    const int kb = blockIdx.z * CKZ + 1;
    const int ke = min(kb + CKZ - 1, a.L.nzl);
    const int x0 = a.L.xoff - 2 + blockIdx.x * PX;     // box origin: cell (i0 - 2, j0 - 1)
    const int y0 = blockIdx.y * PY;
    if (tid == 0) {
        for (int q = 0; q < 4; ++q) mbar_init(bar + q, 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto slot = [&](int q, int f) { return reinterpret_cast<double*>(smem + (size_t)((q & 3) * 4 + f) * PTILE_B); };
    // plane set z in [kb-1, ke+1]: phi always; w* for z <= ke; u*, v* for kb <= z <= ke
    auto issue = [&](int z) {
        const int q = z - (kb - 1);
        uint64_t* b = bar + (q & 3);
        const bool mid = z >= kb && z <= ke;
        mbar_expect_tx(b, (uint32_t)((mid ? 4 : (z < kb ? 2 : 1)) * PTILE * 8));
        tma_load_3d(slot(q, 0), &mf_, x0, y0, z, b);
        if (z <= ke) tma_load_3d(slot(q, 3), &mw_, x0, y0, z, b);
        if (mid) {
            tma_load_3d(slot(q, 1), &mu_, x0, y0, z, b);
            tma_load_3d(slot(q, 2), &mv_, x0, y0, z, b);
        }
    };
    if (tid == 0) {
        issue(kb - 1);
        issue(kb);
        issue(kb + 1);
        if (kb + 2 <= ke + 1) issue(kb + 2);
    }
    const bool active = i <= a.L.nx && j <= a.L.ny;
    const long long sz = a.L.sz;
    long long c = a.L.idx(active ? i : 1, active ? j : 1, kb);
    const double id = a.idelta, dt = a.dt;
    const double rf = 0.5 * (a.rho0 + a.rho0);
    const int o = (ty + 1) * PTW + tx + 2;             // my cell inside a box
    double p0 = a.p[c];
    mbar_wait(bar + 0, 0);
    mbar_wait(bar + 1, 0);
    // w(k-1) of the first plane: recomputed, or the wall value
    double wkm;
    {
        const double* fm = slot(0, 0); const double* fc = slot(1, 0); const double* wm = slot(0, 3);
        wkm = corr_val<UNIT>(wm[o], fc[o], fm[o], id, dt, rf);
        if (a.wall_lo && kb == 1) wkm = a.wlo;
    }
    double md = -1.0e300, mv = 0.0;
    for (int k = kb; k <= ke; ++k) {
        const int q = k - (kb - 1);                    // ring position of plane k
        mbar_wait(bar + ((q + 1) & 3), ((q + 1) >> 2) & 1);
        __syncthreads();                               // everybody is done with plane k-1: its slot is free
        if (tid == 0 && k + 3 <= ke + 1) issue(k + 3);
        double n_p = 0.0;
        if (k < ke) n_p = a.p[c + sz];
        const double* fc = slot(q, 0); const double* fp = slot(q + 1, 0);
        const double* us = slot(q, 1); const double* vs = slot(q, 2); const double* ws = slot(q, 3);
        const double f0 = fc[o];
        const double un = corr_val<UNIT>(us[o], fc[o + 1], f0, id, dt, rf);
        const double vn = corr_val<UNIT>(vs[o], fc[o + PTW], f0, id, dt, rf);
        double wn = corr_val<UNIT>(ws[o], fp[o], f0, id, dt, rf);
        if (a.wall_hi && k == a.L.nzl) wn = a.whi;     // scalar.f90:377: the last interior face is the wall
        const double uim = corr_val<UNIT>(us[o - 1], f0, fc[o - 1], id, dt, rf);
        const double vjm = corr_val<UNIT>(vs[o - PTW], f0, fc[o - PTW], id, dt, rf);
        if (active) {
            a.un[c] = un;
            a.vn[c] = vn;
            a.wn[c] = wn;
            a.pn[c] = p0 + f0;                         // navier_stokes.f90:561
            if (a.xper) {                              // periodic x ghosts (scalar.f90:257,276)
                if (i == 1) { const long long gh = c + a.L.nx; a.un[gh] = un; a.vn[gh] = vn; a.wn[gh] = wn; a.pn[gh] = p0 + f0; }
                if (i == a.L.nx) { const long long gh = c - a.L.nx; a.un[gh] = un; a.vn[gh] = vn; a.wn[gh] = wn; a.pn[gh] = p0 + f0; }
            }
            // checks (:587-617): divergence as fields.f90:144-147, signed max (H6)
            double d = (un - uim) * id + (vn - vjm) * id;
            d = d + (wn - wkm) * id;
            md = fmax(md, d);
            mv = fmax(mv, fabs(un) + fabs(vn) + fabs(wn));
        }
        wkm = wn;
        p0 = n_p;
        c += sz;
    }
    __shared__ double s0[PX * PY / 32], s1[PX * PY / 32];
    for (int o2 = 16; o2 > 0; o2 >>= 1) {
        md = fmax(md, __shfl_xor_sync(0xffffffffu, md, o2));
        mv = fmax(mv, __shfl_xor_sync(0xffffffffu, mv, o2));
    }
    if ((tid & 31) == 0) { s0[tid >> 5] = md; s1[tid >> 5] = mv; }
    __syncthreads();
    if (tid == 0) {
        for (int q = 1; q < PX * PY / 32; ++q) { md = fmax(md, s0[q]); mv = fmax(mv, s1[q]); }
        const long long b = blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z);
        a.partial[2 * b] = md;
        a.partial[2 * b + 1] = mv;
    }
}

// ---- reductions --------------------------------------------------------------------------------
__device__ __forceinline__ double warp_max(double x) {
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}
__device__ __forceinline__ double warp_sum(double x) {
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

struct CheckArgs {
    Layout L;
    const double* u; const double* v; const double* w;
    double idelta;
    double* partial;    // [2 * nblocks]
};

template <bool D3>
__global__ void __launch_bounds__(TX* TY) k_check(CheckArgs a) {
    const int i = blockIdx.x * TX + threadIdx.x + 1;
    double md = -1.0e300, mv = 0.0;
    // rows j, j + gridDim.y * TY, ...: a 2-D grid has one plane, so its launch covers y with few blocks that stride over
    // the rows (one cell per thread left 32 768 partial pairs to a one-block final pass: 0.13 ms of a 1.8 ms step);
    // max is exact in any order, so the result is the same bits
    for (int j = blockIdx.y * TY + threadIdx.y + 1; i <= a.L.nx && j <= a.L.ny; j += gridDim.y * TY) {
        const int kb = blockIdx.z * KZ + 1;
        const int ke = min(kb + KZ - 1, a.L.nzl);
        for (int k = kb; k <= ke; ++k) {
            const long long c = a.L.idx(i, j, k);
            md = fmax(md, div_at<D3>(a.L, a.u, a.v, a.w, c, a.idelta));   // signed max (H6)
            double vel = fabs(a.u[c]) + fabs(a.v[c]);
            if (D3) vel += fabs(a.w[c]);
            mv = fmax(mv, vel);
        }
    }
    __shared__ double s0[TX * TY / 32], s1[TX * TY / 32];
    const int tid = threadIdx.y * TX + threadIdx.x;
    md = warp_max(md);
    mv = warp_max(mv);
    if ((tid & 31) == 0) { s0[tid >> 5] = md; s1[tid >> 5] = mv; }
    __syncthreads();
    if (tid == 0) {
        for (int q = 1; q < TX * TY / 32; ++q) { md = fmax(md, s0[q]); mv = fmax(mv, s1[q]); }
        const long long b = blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z);
        a.partial[2 * b] = md;
        a.partial[2 * b + 1] = mv;
    }
}

// final pass over per-block partials: out[q] = reduce(partial[q + nq*b])
template <int OP>
__global__ void k_reduce_final(const double* partial, long long nb, int nq, double* out) {
    __shared__ double s[32];
    for (int q = 0; q < nq; ++q) {
        double x = OP == 0 ? -1.0e300 : 0.0;
        for (long long b = threadIdx.x; b < nb; b += blockDim.x) {
            double y = partial[nq * b + q];
            x = OP == 0 ? fmax(x, y) : x + y;
        }
        x = OP == 0 ? warp_max(x) : warp_sum(x);
        if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int r = 1; r < (int)blockDim.x / 32; ++r) x = OP == 0 ? fmax(x, s[r]) : x + s[r];
            out[q] = x;
        }
        __syncthreads();
    }
}

template <int OP>
__global__ void __launch_bounds__(TX* TY) k_reduce_field(Layout L, const double* f, double* partial) {
    const int i = blockIdx.x * TX + threadIdx.x + 1;
    const int j = blockIdx.y * TY + threadIdx.y + 1;
    double x = OP == 0 ? -1.0e300 : 0.0;
    double mn = 1.0e300;
    if (i <= L.nx && j <= L.ny) {
        const int kb = blockIdx.z * KZ + 1;
        const int ke = min(kb + KZ - 1, L.nzl);
        for (int k = kb; k <= ke; ++k) {
            double y = f[L.idx(i, j, k)];
            x = OP == 0 ? fmax(x, y) : x + y;
            mn = fmin(mn, y);
        }
    }
    __shared__ double s0[TX * TY / 32], s1[TX * TY / 32];
    const int tid = threadIdx.y * TX + threadIdx.x;
    x = OP == 0 ? warp_max(x) : warp_sum(x);
    mn = -warp_max(-mn);
    if ((tid & 31) == 0) { s0[tid >> 5] = x; s1[tid >> 5] = -mn; }
    __syncthreads();
    if (tid == 0) {
        double m2 = -mn;
        for (int q = 1; q < TX * TY / 32; ++q) {
            x = OP == 0 ? fmax(x, s0[q]) : x + s0[q];
            m2 = fmax(m2, s1[q]);
        }
        const long long b = blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z);
        partial[2 * b] = x;
        partial[2 * b + 1] = m2;     // max of (-f): -(min f), reduced with max when OP == 0
    }
}

// ---- generic field operators (fields.f90) ----------------------------------------------------------
struct OpArgs {
    Layout L;
    const double* a0; const double* a1; const double* a2;
    double* o0; double* o1; double* o2;
    double idelta, idelta2;
    long long back;        // MODE 4: offset of the low-side neighbour in the averaging direction
};
// MODE 0 gradient of scalar a0 -> (o0,o1,o2); 1 laplacian of vector; 2 center_to_face; 3 laplacian of scalar a0 -> o0
// (fields.f90:256-294); 4 face_to_center a0 -> o0 (:210-252); 5 curl (a0,a1,a2) -> (o0,o1,o2), 2-D: -> o0 (:347-392)
template <bool D3, int MODE>
__global__ void __launch_bounds__(TX* TY) k_op(OpArgs a) {
    const int i = blockIdx.x * TX + threadIdx.x + 1;
    const int j = blockIdx.y * TY + threadIdx.y + 1;
    if (i > a.L.nx || j > a.L.ny) return;
    const int kb = blockIdx.z * KZ + 1;
    const int ke = min(kb + KZ - 1, a.L.nzl);
    const long long sy = a.L.sy, sz = a.L.sz;
    for (int k = kb; k <= ke; ++k) {
        const long long c = a.L.idx(i, j, k);
        if (MODE == 0) {
            const double s0 = a.a0[c];
            a.o0[c] = (a.a0[c + 1] - s0) * a.idelta;
            a.o1[c] = (a.a0[c + sy] - s0) * a.idelta;
            if (D3) a.o2[c] = (a.a0[c + sz] - s0) * a.idelta;
        } else if (MODE == 2) {
            const double s0 = a.a0[c];
            a.o0[c] = 0.5 * (a.a0[c + 1] + s0);
            a.o1[c] = 0.5 * (a.a0[c + sy] + s0);
            if (D3) a.o2[c] = 0.5 * (a.a0[c + sz] + s0);
        } else if (MODE == 3) {
            const double* f = a.a0;
            const double f0 = f[c];
            const double lx = f[c + 1] - 2.0 * f0 + f[c - 1];
            const double ly = f[c + sy] - 2.0 * f0 + f[c - sy];
            double r = (lx + ly) * a.idelta2;                                   // fields.f90:282-283
            if (D3) r = r + (f[c + sz] - 2.0 * f0 + f[c - sz]) * a.idelta2;     // :285-286
            a.o0[c] = r;
        } else if (MODE == 4) {
            a.o0[c] = 0.5 * (a.a0[c] + a.a0[c - a.back]);                       // fields.f90:225, :235, :245
        } else if (MODE == 5) {
            if (D3) {                                                           // fields.f90:374-379
                a.o0[c] = (a.a2[c + sy] - a.a2[c]) * a.idelta - (a.a1[c + sz] - a.a1[c]) * a.idelta;
                a.o1[c] = (a.a0[c + sz] - a.a0[c]) * a.idelta - (a.a2[c + 1] - a.a2[c]) * a.idelta;
                a.o2[c] = (a.a1[c + 1] - a.a1[c]) * a.idelta - (a.a0[c + sy] - a.a0[c]) * a.idelta;
            } else {                                                            // :383-384: z component in curl_v%x
                a.o0[c] = (a.a1[c + 1] - a.a1[c]) * a.idelta - (a.a0[c + sy] - a.a0[c]) * a.idelta;
            }
        } else {
            const double* in[3] = {a.a0, a.a1, a.a2};
            double* out[3] = {a.o0, a.o1, a.o2};
            for (int m = 0; m < (D3 ? 3 : 2); ++m) {
                const double* f = in[m];
                const double f0 = f[c];
                double lx = f[c + 1] - 2.0 * f0 + f[c - 1];
                double ly = f[c + sy] - 2.0 * f0 + f[c - sy];
                double r;
                if (D3) {
                    double lz = f[c + sz] - 2.0 * f0 + f[c - sz];
                    r = (m == 2) ? (lx + ly + lz) * a.idelta2 : (lx + ly) * a.idelta2 + lz * a.idelta2;
                } else {
                    r = (lx + ly) * a.idelta2;
                }
                out[m][c] = r;
            }
        }
    }
}

static dim3 st_grid(const Layout& L) {
    return dim3((L.nx + TX - 1) / TX, (L.ny + TY - 1) / TY, (L.nzl + KZ - 1) / KZ);
}
static long long st_blocks(const Layout& L) {
    dim3 g = st_grid(L);
    return (long long)g.x * g.y * g.z;
}

static int fill_props(fen_ctx* c, const double** rho, const double** mu, double* rho0, double* mu0) {
    *rho0 = c->rho_uniform;
    *mu0 = c->mu_uniform;
    *rho = nullptr;
    *mu = nullptr;
    if (!c->uniform_props) {
        Field *fr, *fm;
        FEN_TRY(field_check(c, FEN_RHO, &fr));
        FEN_TRY(field_check(c, FEN_MU, &fm));
        *rho = fr->d;
        *mu = fm->d;
    }
    return FEN_OK;
}

static int fill_st(fen_ctx* c, StArgs& a) {
    const bool d3 = c->g.ndim == 3;
    Field *u, *v, *w = nullptr, *p;
    FEN_TRY(field_check(c, FEN_VX, &u));
    FEN_TRY(field_check(c, FEN_VY, &v));
    if (d3) FEN_TRY(field_check(c, FEN_VZ, &w));
    FEN_TRY(field_check(c, FEN_P, &p));
    a.L = c->L;
    a.u = u->d; a.v = v->d; a.w = w ? w->d : nullptr; a.p = p->d;
    FEN_TRY(fill_props(c, &a.rho, &a.mu, &a.rho0, &a.mu0));
    a.sx = a.sy_ = a.sz_ = nullptr;
    if (c->has_source) {
        Field *sx, *sy, *sz = nullptr;
        FEN_TRY(field_check(c, FEN_SX, &sx));
        FEN_TRY(field_check(c, FEN_SY, &sy));
        if (d3) FEN_TRY(field_check(c, FEN_SZ, &sz));
        a.sx = sx->d; a.sy_ = sy->d; a.sz_ = sz ? sz->d : nullptr;
    }
    a.idelta = 1.0 / c->g.delta;
    a.idelta2 = 1.0 / (c->g.delta * c->g.delta);     // fields.f90:313: 1/delta**2
    return FEN_OK;
}

static bool general_path(fen_ctx* c) { return !c->uniform_props || c->has_source; }

// the general kernels read rho/mu as fields: materialise them if only the source term is general
static int ensure_general_fields(fen_ctx* c, StArgs& a) {
    if (a.rho) return FEN_OK;
    Field *fr, *fm;
    FEN_TRY(field_check(c, FEN_RHO, &fr));
    FEN_TRY(field_check(c, FEN_MU, &fm));
    a.rho = fr->d;
    a.mu = fm->d;
    return FEN_OK;
}

int ns_predict(fen_ctx* c, double dt) {
    const bool d3 = c->g.ndim == 3;
    StArgs a;
    FEN_TRY(fill_st(c, a));
    Field *dx, *dy, *dz = nullptr;
    FEN_TRY(field_check(c, FEN_DVOX, &dx));
    FEN_TRY(field_check(c, FEN_DVOY, &dy));
    if (d3) FEN_TRY(field_check(c, FEN_DVOZ, &dz));
    a.dvox = dx->d; a.dvoy = dy->d; a.dvoz = dz ? dz->d : nullptr;
    for (int m = 0; m < (d3 ? 3 : 2); ++m)
        if (!c->vnew[m]) return set_error(FEN_ERR_STATE, "init_solver has not allocated the predictor buffers");
    a.un = c->vnew[0]; a.vn = c->vnew[1]; a.wn = c->vnew[2];
    a.dt = dt;
    a.A = 1.0 + 0.5 * dt / c->prm.dt_o;      // navier_stokes.f90:157
    a.B = -0.5 * dt / c->prm.dt_o;           // navier_stokes.f90:158
    a.g0 = c->prm.g[0]; a.g1 = c->prm.g[1]; a.g2 = c->prm.g[2];
    const bool gen = general_path(c);
    if (gen) FEN_TRY(ensure_general_fields(c, a));
    dim3 grid = st_grid(c->L), block(TX, TY);
    bool x_done = false;
    a.xper = 0;
    if (d3 && !gen) {
        x_done = true;
        for (int q = 0; q < 3; ++q)
            x_done = x_done && c->fields[FEN_VX + q].bc_type[FEN_LEFT] == FEN_PERIODIC &&
                     c->fields[FEN_VX + q].bc_type[FEN_RIGHT] == FEN_PERIODIC;
        a.xper = x_done ? 1 : 0;
        // TMA-staged kernel: tensor maps of the three velocity buffers (cached per buffer)
        CUtensorMap* m[3];
        for (int q = 0; q < 3; ++q) FEN_TRY(field_tmap(c, q == 0 ? a.u : (q == 1 ? a.v : a.w), PTW, PTH, &m[q]));
        FEN_ONCE_PER_DEVICE(c) {
            FEN_CUDA(cudaFuncSetAttribute(k_pred_tma<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PRED_SMEM));
            FEN_CUDA(cudaFuncSetAttribute(k_pred_tma<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PRED_SMEM));
            FEN_CUDA(cudaFuncSetAttribute(k_pred_tma<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PRED_SMEM));
            FEN_CUDA(cudaFuncSetAttribute(k_pred_tma<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PRED_SMEM));
        }
        static const bool xg_env = !getenv("FEN_PRED_XG") || atoi(getenv("FEN_PRED_XG")) != 0;
        x_done = x_done && xg_env;
        dim3 pg((c->L.nx + PX - 1) / PX, (c->L.ny + PY - 1) / PY, (c->L.nzl + PKZ - 1) / PKZ), pb(PX, PY);
        const bool unit = a.rho0 == 1.0;
        if (unit && x_done) FEN_LAUNCH(c, "pred", k_pred_tma<true, true><<<pg, pb, PRED_SMEM, c->stream>>>(*m[0], *m[1], *m[2], a));
        else if (unit) FEN_LAUNCH(c, "pred", k_pred_tma<true, false><<<pg, pb, PRED_SMEM, c->stream>>>(*m[0], *m[1], *m[2], a));
        else if (x_done) FEN_LAUNCH(c, "pred", k_pred_tma<false, true><<<pg, pb, PRED_SMEM, c->stream>>>(*m[0], *m[1], *m[2], a));
        else FEN_LAUNCH(c, "pred", k_pred_tma<false, false><<<pg, pb, PRED_SMEM, c->stream>>>(*m[0], *m[1], *m[2], a));
    } else if (d3) {
        FEN_LAUNCH(c, "pred", k_pred<true, true><<<grid, block, 0, c->stream>>>(a));
    } else {
        if (gen) FEN_LAUNCH(c, "pred", k_pred<false, true><<<grid, block, 0, c->stream>>>(a));
        else FEN_LAUNCH(c, "pred", k_pred<false, false><<<grid, block, 0, c->stream>>>(a));
    }
    FEN_CUDA(cudaGetLastError());
    // v now lives in the freshly written buffers; the old ones become the next scratch
    for (int m = 0; m < (d3 ? 3 : 2); ++m) std::swap(c->fields[FEN_VX + m].d, c->vnew[m]);
    return ghost_update(c, FEN_VX, d3 ? 3 : 2, x_done);      // navier_stokes.f90:208
}

int op_explicit_terms(fen_ctx* c, int rhs_x, bool advection_only) {
    const bool d3 = c->g.ndim == 3;
    StArgs a;
    FEN_TRY(fill_st(c, a));
    Field* o[3] = {nullptr, nullptr, nullptr};
    for (int m = 0; m < (d3 ? 3 : 2); ++m) FEN_TRY(field_check(c, rhs_x + m, &o[m]));
    a.un = o[0]->d; a.vn = o[1]->d; a.wn = o[2] ? o[2]->d : nullptr;
    a.dvox = a.dvoy = a.dvoz = nullptr;
    a.dt = a.A = a.B = a.g0 = a.g1 = a.g2 = 0.0;
    a.xper = 0;
    const bool gen = general_path(c) && !advection_only;
    if (gen) FEN_TRY(ensure_general_fields(c, a));
    dim3 grid = st_grid(c->L), block(TX, TY);
#define FEN_EXPL(D3, GEN, ADV) FEN_LAUNCH(c, "explicit", k_explicit<D3, GEN, ADV><<<grid, block, 0, c->stream>>>(a))
    if (advection_only) { if (d3) FEN_EXPL(true, false, true); else FEN_EXPL(false, false, true); }
    else if (gen) { if (d3) FEN_EXPL(true, true, false); else FEN_EXPL(false, true, false); }
    else { if (d3) FEN_EXPL(true, false, false); else FEN_EXPL(false, false, false); }
#undef FEN_EXPL
    FEN_CUDA(cudaGetLastError());
    return FEN_OK;
}

static int launch_rhs(fen_ctx* c, int vx, int s, bool scale, double dt) {
    const bool d3 = c->g.ndim == 3;
    Field *u, *v, *w = nullptr, *o;
    FEN_TRY(field_check(c, vx, &u));
    FEN_TRY(field_check(c, vx + 1, &v));
    if (d3) FEN_TRY(field_check(c, vx + 2, &w));
    FEN_TRY(field_check(c, s, &o));
    RhsArgs a;
    a.L = c->L;
    a.u = u->d; a.v = v->d; a.w = w ? w->d : nullptr;
    const double* mu;
    double mu0;
    FEN_TRY(fill_props(c, &a.rho, &mu, &a.rho0, &mu0));
    a.phi = o->d;
    a.idelta = 1.0 / c->g.delta;
    a.dt = dt;
    a.scale = scale ? 1 : 0;
    dim3 grid = st_grid(c->L), block(TX, TY);
    if (d3) FEN_LAUNCH(c, scale ? "poisson_rhs" : "divergence", k_rhs<true><<<grid, block, 0, c->stream>>>(a));
    else FEN_LAUNCH(c, scale ? "poisson_rhs" : "divergence", k_rhs<false><<<grid, block, 0, c->stream>>>(a));
    FEN_CUDA(cudaGetLastError());
    return FEN_OK;
}

int ns_poisson_rhs(fen_ctx* c, double dt) { return launch_rhs(c, FEN_VX, FEN_PHI, true, dt); }
int op_divergence(fen_ctx* c, int vx, int s) { return launch_rhs(c, vx, s, false, 1.0); }

// the fused correction + checks kernel applies (see k_corr_tma): 3-D, uniform properties, x and y periodic for
// every field it touches, and in z either periodic / rank boundary or a wall with a zero or uniform w
static bool corr_fused_ok(fen_ctx* c) {
    if (c->g.ndim != 3 || !c->uniform_props) return false;
    const int ids[5] = {FEN_VX, FEN_VY, FEN_VZ, FEN_P, FEN_PHI};
    for (int id : ids)
        for (int face = 0; face < 4; ++face)
            if (c->fields[id].bc_type[face] != FEN_PERIODIC) return false;
    const Field& w = c->fields[FEN_VZ];
    for (int face = FEN_FRONT; face <= FEN_BACK; ++face) {
        const int t = w.bc_type[face];
        if (t == FEN_PERIODIC || t == FEN_HALO) continue;
        if (t == FEN_DIRICHLET && w.bc_mode[face] != BC_PLANE) continue;
        return false;
    }
    return true;
}

int ns_correct(fen_ctx* c, double dt, bool* checks_done) {
    const bool d3 = c->g.ndim == 3;
    if (checks_done) *checks_done = false;
    Field *u, *v, *w = nullptr, *p, *phi;
    FEN_TRY(field_check(c, FEN_VX, &u));
    FEN_TRY(field_check(c, FEN_VY, &v));
    if (d3) FEN_TRY(field_check(c, FEN_VZ, &w));
    FEN_TRY(field_check(c, FEN_P, &p));
    FEN_TRY(field_check(c, FEN_PHI, &phi));
    if (corr_fused_ok(c) && c->vnew[0] && c->vnew[1] && c->vnew[2]) {
        CUtensorMap* m[4];
        const double* src[4] = {phi->d, u->d, v->d, w->d};
        for (int q = 0; q < 4; ++q) FEN_TRY(field_tmap(c, src[q], PTW, PTH, &m[q]));
        FEN_ONCE_PER_DEVICE(c) {
            FEN_CUDA(cudaFuncSetAttribute(k_corr_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CORR_SMEM));
            FEN_CUDA(cudaFuncSetAttribute(k_corr_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CORR_SMEM));
        }
        FEN_TRY(ensure_red(c));
        CorrTArgs a;
        a.L = c->L;
        a.p = p->d; a.pn = p->d;
        a.un = c->vnew[0]; a.vn = c->vnew[1]; a.wn = c->vnew[2];
        a.rho0 = c->rho_uniform;
        a.idelta = 1.0 / c->g.delta;
        a.dt = dt;
        a.xper = 1;                                        // corr_fused_ok: x is periodic for u, v, w, p
        a.wall_lo = w->bc_type[FEN_FRONT] == FEN_DIRICHLET;
        a.wall_hi = w->bc_type[FEN_BACK] == FEN_DIRICHLET;
        a.wlo = w->bc_mode[FEN_FRONT] == BC_UNIFORM ? w->bc_value[FEN_FRONT] : 0.0;
        a.whi = w->bc_mode[FEN_BACK] == BC_UNIFORM ? w->bc_value[FEN_BACK] : 0.0;
        a.partial = c->d_red + 16;
        dim3 cg((c->L.nx + PX - 1) / PX, (c->L.ny + PY - 1) / PY, (c->L.nzl + CKZ - 1) / CKZ), cb(PX, PY);
        if (a.rho0 == 1.0)
            FEN_LAUNCH(c, "corr_check", k_corr_tma<true><<<cg, cb, CORR_SMEM, c->stream>>>(*m[0], *m[1], *m[2], *m[3], a));
        else
            FEN_LAUNCH(c, "corr_check", k_corr_tma<false><<<cg, cb, CORR_SMEM, c->stream>>>(*m[0], *m[1], *m[2], *m[3], a));
        FEN_CUDA(cudaGetLastError());
        for (int q = 0; q < 3; ++q) std::swap(c->fields[FEN_VX + q].d, c->vnew[q]);
        FEN_TRY(ghost_update(c, FEN_VX, 3, true));         // navier_stokes.f90:544
        FEN_TRY(ghost_update(c, FEN_P, 1, true));          // navier_stokes.f90:564
        if (checks_done) {
            const long long nb = (long long)cg.x * cg.y * cg.z;
            FEN_LAUNCH(c, "reduce", k_reduce_final<0><<<1, 256, 0, c->stream>>>(c->d_red + 16, nb, 2, c->d_red));
            FEN_CUDA(cudaGetLastError());
            if (c->g.nranks > 1) FEN_TRY(comm_allreduce(c, c->d_red, 2, 0));   // navier_stokes.f90:614, scalar.f90:194
            *checks_done = true;
        }
        return FEN_OK;
    }
    CorrArgs a;
    a.L = c->L;
    a.u = u->d; a.v = v->d; a.w = w ? w->d : nullptr; a.p = p->d; a.phi = phi->d;
    const double* mu;
    double mu0;
    FEN_TRY(fill_props(c, &a.rho, &mu, &a.rho0, &mu0));
    a.idelta = 1.0 / c->g.delta;
    a.dt = dt;
    dim3 grid = st_grid(c->L), block(TX, TY);
    if (d3) FEN_LAUNCH(c, "corr", k_corr<true><<<grid, block, 0, c->stream>>>(a));
    else FEN_LAUNCH(c, "corr", k_corr<false><<<grid, block, 0, c->stream>>>(a));
    FEN_CUDA(cudaGetLastError());
    FEN_TRY(ghost_update(c, FEN_VX, d3 ? 3 : 2));     // navier_stokes.f90:544
    return ghost_update(c, FEN_P, 1);                  // navier_stokes.f90:564
}

int ensure_red(fen_ctx* c) {
    const long long nb = st_blocks(c->L);
    if (c->d_red && c->red_blocks >= nb) return FEN_OK;
    if (c->d_red) cudaFree(c->d_red);
    c->red_blocks = (int)nb;
    FEN_CUDA(cudaMalloc(&c->d_red, (size_t)(2 * nb + 16) * sizeof(double)));
    if (!c->h_red) FEN_CUDA(cudaMallocHost(&c->h_red, 16 * sizeof(double)));
    return FEN_OK;
}

int ns_checks_launch(fen_ctx* c, double dt) {
    const bool d3 = c->g.ndim == 3;
    (void)dt;
    Field *u, *v, *w = nullptr;
    FEN_TRY(field_check(c, FEN_VX, &u));
    FEN_TRY(field_check(c, FEN_VY, &v));
    if (d3) FEN_TRY(field_check(c, FEN_VZ, &w));
    FEN_TRY(ensure_red(c));
    CheckArgs a;
    a.L = c->L;
    a.u = u->d; a.v = v->d; a.w = w ? w->d : nullptr;
    a.idelta = 1.0 / c->g.delta;
    a.partial = c->d_red + 16;
    dim3 grid = st_grid(c->L), block(TX, TY);
    if (!d3) grid.y = std::min<unsigned>(grid.y, 128);        // 2-D: the blocks stride over the rows (k_check)
    const long long nblocks = (long long)grid.x * grid.y * grid.z;
    if (d3) FEN_LAUNCH(c, "check", k_check<true><<<grid, block, 0, c->stream>>>(a));
    else FEN_LAUNCH(c, "check", k_check<false><<<grid, block, 0, c->stream>>>(a));
    FEN_LAUNCH(c, "reduce", k_reduce_final<0><<<1, 256, 0, c->stream>>>(c->d_red + 16, nblocks, 2, c->d_red));
    FEN_CUDA(cudaGetLastError());
    if (c->g.nranks > 1) FEN_TRY(comm_allreduce(c, c->d_red, 2, 0));   // navier_stokes.f90:614, scalar.f90:194
    return FEN_OK;
}

// reduce the interior of f: d_out[0] = max or sum, d_out[1] = max(-f) = -(min f)
int reduce_field(fen_ctx* c, const double* f, int op, double* d_out) {
    FEN_TRY(ensure_red(c));
    dim3 grid = st_grid(c->L), block(TX, TY);
    if (op == 0) {
        FEN_LAUNCH(c, "reduce", k_reduce_field<0><<<grid, block, 0, c->stream>>>(c->L, f, c->d_red + 16));
        FEN_LAUNCH(c, "reduce", k_reduce_final<0><<<1, 256, 0, c->stream>>>(c->d_red + 16, st_blocks(c->L), 2, d_out));
    } else {
        FEN_LAUNCH(c, "reduce", k_reduce_field<1><<<grid, block, 0, c->stream>>>(c->L, f, c->d_red + 16));
        // sum for slot 0; slot 1 (max of -f) is not meaningful under a sum-final, callers ignore it
        FEN_LAUNCH(c, "reduce", k_reduce_final<1><<<1, 256, 0, c->stream>>>(c->d_red + 16, st_blocks(c->L), 2, d_out));
    }
    FEN_CUDA(cudaGetLastError());
    return FEN_OK;
}

int field_is_uniform(fen_ctx* c, const double* f, bool* uniform, double* value) {
    FEN_TRY(ensure_red(c));
    FEN_TRY(reduce_field(c, f, 0, c->d_red));
    if (c->g.nranks > 1) FEN_TRY(comm_allreduce(c, c->d_red, 2, 0));
    FEN_CUDA(cudaMemcpyAsync(c->h_red, c->d_red, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    FEN_CUDA(cudaStreamSynchronize(c->stream));
    *uniform = (c->h_red[0] == -c->h_red[1]);
    *value = c->h_red[0];
    return FEN_OK;
}

static int launch_op(fen_ctx* c, int mode, int in0, int nin, int out0) {
    const bool d3 = c->g.ndim == 3;
    const int nc = d3 ? 3 : 2;
    OpArgs a;
    a.L = c->L;
    const double* in[3] = {nullptr, nullptr, nullptr};
    double* out[3] = {nullptr, nullptr, nullptr};
    for (int m = 0; m < nin; ++m) {
        Field* f;
        FEN_TRY(field_check(c, in0 + m, &f));
        if (f->gl < 1) return set_error(FEN_ERR_ARG, "operator input field %d needs ghost nodes", in0 + m);
        in[m] = f->d;
    }
    for (int m = 0; m < nc; ++m) {
        Field* f;
        FEN_TRY(field_check(c, out0 + m, &f));
        out[m] = f->d;
    }
    a.a0 = in[0]; a.a1 = in[1]; a.a2 = in[2];
    a.o0 = out[0]; a.o1 = out[1]; a.o2 = out[2];
    a.idelta = 1.0 / c->g.delta;
    a.idelta2 = 1.0 / (c->g.delta * c->g.delta);
    dim3 grid = st_grid(c->L), block(TX, TY);
#define FEN_OP(D3, M, NAME) FEN_LAUNCH(c, NAME, k_op<D3, M><<<grid, block, 0, c->stream>>>(a))
    if (mode == 0) { if (d3) FEN_OP(true, 0, "gradient"); else FEN_OP(false, 0, "gradient"); }
    if (mode == 1) { if (d3) FEN_OP(true, 1, "laplacian"); else FEN_OP(false, 1, "laplacian"); }
    if (mode == 2) { if (d3) FEN_OP(true, 2, "center_to_face"); else FEN_OP(false, 2, "center_to_face"); }
#undef FEN_OP
    FEN_CUDA(cudaGetLastError());
    return FEN_OK;
}

// scalar -> scalar and vector -> vector operators of fields_mod that the time step itself does not use, kept for the
// callers either side of it (the reference's own fields test calls laplacian(s, lap_s), its cavity driver curl(v, omega))
static int launch_op2(fen_ctx* c, int mode, int in0, int nin, int out0, int nout, int dir) {
    const bool d3 = c->g.ndim == 3;
    OpArgs a;
    a.L = c->L;
    const double* in[3] = {nullptr, nullptr, nullptr};
    double* out[3] = {nullptr, nullptr, nullptr};
    for (int m = 0; m < nin; ++m) {
        Field* f;
        FEN_TRY(field_check(c, in0 + m, &f));
        if (f->gl < 1) return set_error(FEN_ERR_ARG, "operator input field %d needs ghost nodes", in0 + m);
        in[m] = f->d;
    }
    for (int m = 0; m < nout; ++m) {
        Field* f;
        FEN_TRY(field_check(c, out0 + m, &f));
        for (int q = 0; q < nin; ++q)
            if (f->d == in[q]) return set_error(FEN_ERR_ARG, "operator output field %d is also an input", out0 + m);
        out[m] = f->d;
    }
    a.a0 = in[0]; a.a1 = in[1]; a.a2 = in[2];
    a.o0 = out[0]; a.o1 = out[1]; a.o2 = out[2];
    a.idelta = 1.0 / c->g.delta;
    a.idelta2 = 1.0 / (c->g.delta * c->g.delta);
    a.back = dir == 0 ? 1 : (dir == 1 ? c->L.sy : c->L.sz);
    dim3 grid = st_grid(c->L), block(TX, TY);
#define FEN_OP(D3, M, NAME) FEN_LAUNCH(c, NAME, k_op<D3, M><<<grid, block, 0, c->stream>>>(a))
    if (mode == 3) { if (d3) FEN_OP(true, 3, "laplacian_s"); else FEN_OP(false, 3, "laplacian_s"); }
    if (mode == 4) { if (d3) FEN_OP(true, 4, "face_to_center"); else FEN_OP(false, 4, "face_to_center"); }
    if (mode == 5) { if (d3) FEN_OP(true, 5, "curl"); else FEN_OP(false, 5, "curl"); }
#undef FEN_OP
    FEN_CUDA(cudaGetLastError());
    return FEN_OK;
}
int op_laplacian_scalar(fen_ctx* c, int s, int o) { return launch_op2(c, 3, s, 1, o, 1, 0); }
int op_face_to_center(fen_ctx* c, int sf, int sc, int dir) {
    if (dir < 0 || dir > 2 || (dir == 2 && c->g.ndim == 2))
        return set_error(FEN_ERR_ARG, "face_to_center: direction %d (0 = x, 1 = y, 2 = z in 3-D)", dir);
    return launch_op2(c, 4, sf, 1, sc, 1, dir);
}
int op_curl(fen_ctx* c, int vx, int ox) {
    const bool d3 = c->g.ndim == 3;
    return launch_op2(c, 5, vx, d3 ? 3 : 2, ox, d3 ? 3 : 1, 0);
}

int op_gradient(fen_ctx* c, int s, int vx) { return launch_op(c, 0, s, 1, vx); }
int op_laplacian(fen_ctx* c, int vx, int ox) { return launch_op(c, 1, vx, c->g.ndim == 3 ? 3 : 2, ox); }
int op_center_to_face(fen_ctx* c, int s, int vx) { return launch_op(c, 2, s, 1, vx); }

}  // namespace fen
