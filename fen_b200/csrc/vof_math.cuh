// vof_math.cuh -- per-cell arithmetic of FEN's MTHINC volume-of-fluid method (src/volume_of_fluid.f90, Ii et al.
// JCP 2012) and of the two-phase momentum terms (src/navier_stokes.f90, -DMF branches), written once as
// __host__ __device__ functions in the reference's operation order.  The kernels of multiphase.cu call them on the
// device; tests/cpu/vof_math_host.cpp calls the same functions from host loops so that the transcription can be checked
// on a machine without a GPU (tests/test_vof_math_host.py).
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define FEN_HD __host__ __device__ __forceinline__
#else
#define FEN_HD inline
#endif

namespace fen {

constexpr double VOF_SMALL = 1.0e-14;                             // global.f90:16
// two-point Gauss quadrature on [0, 1], volume_of_fluid.f90:33-34: 0.5*(1 +- 1/sqrt(3))
constexpr double VOF_RP = 0.78867513459481287;
constexpr double VOF_RM = 0.21132486540518708;

struct VofRecon {
    double nx, ny, lx, ly, curv;
};

// One corner of compute_norm (volume_of_fluid.f90:327-356): the Youngs gradient at the corner shared by the cells
// f00 = vof(i, j), f10 = vof(i+1, j), f01 = vof(i, j+1), f11 = vof(i+1, j+1), and its normalisation (:357-360).
// The reference evaluates every corner from the four cells around it with the same operands in the same order, so
// the value is bitwise shared (the tiled kernel computes it once).  A zero gradient gives 0 / sqrt(small) = 0 exactly.
FEN_HD void vof_corner(double f00, double f10, double f01, double f11, double idelta, double& mx, double& my,
                       double& nx, double& ny) {
    mx = 0.5 * (f10 + f11 - f00 - f01) * idelta;
    my = 0.5 * (f01 + f11 - f00 - f10) * idelta;
    if (mx == 0.0 && my == 0.0) {
        nx = 0.0;
        ny = 0.0;
        return;
    }
    const double r = sqrt(mx * mx + my * my + VOF_SMALL);
    nx = mx / r;
    ny = my / r;
}

// cell-centre part of compute_norm from the four corners (:338, :353, :361-373).  Corner order as in the reference:
// 0 = (i-1/2, j-1/2), 1 = (i-1/2, j+1/2), 2 = (i+1/2, j+1/2), 3 = (i+1/2, j-1/2).
FEN_HD VofRecon vof_norm_from_corners(const double mx[4], const double my[4], const double nx[4], const double ny[4],
                                      double delta, double idelta2, bool quadratic) {
    VofRecon o;
    const double mxc = 0.25 * (mx[0] + mx[1] + mx[2] + mx[3]);
    const double myc = 0.25 * (my[0] + my[1] + my[2] + my[3]);
    if (mxc == 0.0 && myc == 0.0) {
        o.nx = 0.0;
        o.ny = 0.0;
    } else {
        const double rc = sqrt(mxc * mxc + myc * myc + VOF_SMALL);
        o.nx = mxc / rc;
        o.ny = myc / rc;
    }
    if (quadratic) {
        o.lx = 0.5 * delta * (nx[3] + nx[2] - nx[1] - nx[0]);
        o.ly = 0.5 * delta * (ny[1] + ny[2] - ny[0] - ny[3]);
    } else {
        o.lx = 0.0;
        o.ly = 0.0;
    }
    o.curv = -(o.lx + o.ly) * idelta2;
    return o;
}

// compute_norm, volume_of_fluid.f90:307-396.  f[b][a] = vof(i + a - 1, j + b - 1).
FEN_HD VofRecon vof_norm(const double f[3][3], double delta, double idelta, double idelta2, bool quadratic) {
    double mx[4], my[4], nx[4], ny[4];
    vof_corner(f[0][0], f[0][1], f[1][0], f[1][1], idelta, mx[0], my[0], nx[0], ny[0]);      // i-1/2, j-1/2
    vof_corner(f[1][0], f[1][1], f[2][0], f[2][1], idelta, mx[1], my[1], nx[1], ny[1]);      // i-1/2, j+1/2
    vof_corner(f[1][1], f[1][2], f[2][1], f[2][2], idelta, mx[2], my[2], nx[2], ny[2]);      // i+1/2, j+1/2
    vof_corner(f[0][1], f[0][2], f[1][1], f[1][2], idelta, mx[3], my[3], nx[3], ny[3]);      // i+1/2, j-1/2
    return vof_norm_from_corners(mx, my, nx, ny, delta, idelta2, quadratic);
}

// ---- the same reconstruction for a 64 x 4 tile of cells with the corners computed once -------------------------
// Shared-memory tile: F = vof of the tile and its one-cell halo (66 x 6), then the 65 x 5 corners.  The three
// phases are separated by block barriers in the kernel (multiphase.cu: k_vof_recon_tile); tests/cpu/vof_math_host.cpp
// runs them from loops over tid, so the index logic is checked on a machine without a GPU.
constexpr int VT_X = 64, VT_Y = 4, VT_N = VT_X * VT_Y;
constexpr int VT_FW = VT_X + 2, VT_FH = VT_Y + 2, VT_CW = VT_X + 1, VT_CH = VT_Y + 1;
struct VofTile {
    double F[VT_FH][VT_FW];
    double cmx[VT_CH][VT_CW], cmy[VT_CH][VT_CW], cnx[VT_CH][VT_CW], cny[VT_CH][VT_CW];
};
// phase 1: f points at vof(i0, j0), the tile's low halo corner; wa x hb values are inside the array
FEN_HD void vof_tile_load(VofTile& T, int tid, const double* f, long long sy, int wa, int hb) {
    for (int e = tid; e < VT_FW * VT_FH; e += VT_N) {
        const int b = e / VT_FW, a = e - b * VT_FW;
        T.F[b][a] = (a < wa && b < hb) ? f[a + sy * b] : 0.0;
    }
}
// phase 2: corner (a, b) sits between tile cells (a, b), (a+1, b), (a, b+1), (a+1, b+1)
FEN_HD void vof_tile_corners(VofTile& T, int tid, double idelta) {
    for (int e = tid; e < VT_CW * VT_CH; e += VT_N) {
        const int b = e / VT_CW, a = e - b * VT_CW;
        vof_corner(T.F[b][a], T.F[b][a + 1], T.F[b + 1][a], T.F[b + 1][a + 1], idelta, T.cmx[b][a], T.cmy[b][a],
                   T.cnx[b][a], T.cny[b][a]);
    }
}
// phase 3: the cell (tx, ty) of the tile = tile cell (tx + 1, ty + 1)
FEN_HD VofRecon vof_tile_cell(const VofTile& T, int tx, int ty, double delta, double idelta2, bool quadratic,
                              double& vof00) {
    const int a = tx + 1, b = ty + 1;
    const int ca[4] = {a - 1, a - 1, a, a}, cb[4] = {b - 1, b, b, b - 1};
    double mx[4], my[4], nx[4], ny[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        mx[c] = T.cmx[cb[c]][ca[c]]; my[c] = T.cmy[cb[c]][ca[c]];
        nx[c] = T.cnx[cb[c]][ca[c]]; ny[c] = T.cny[cb[c]][ca[c]];
    }
    vof00 = T.F[b][a];
    return vof_norm_from_corners(mx, my, nx, ny, delta, idelta2, quadratic);
}

// coefficients of the quadratic surface, Eq. 12 of Ii et al. (volume_of_fluid.f90:254-266, 610-634)
struct VofSurf {
    double cx, cy, a10, a01, a20, a02;
    bool xdom;
};
FEN_HD VofSurf vof_surf(double nx, double ny, double lx, double ly) {
    VofSurf s;
    s.xdom = fabs(nx) == fmax(fabs(nx), fabs(ny));
    s.cx = s.xdom ? 0.0 : 1.0;
    s.cy = s.xdom ? 1.0 : 0.0;
    s.a10 = nx - 0.5 * s.cx * lx;
    s.a01 = ny - 0.5 * s.cy * ly;
    s.a20 = 0.5 * s.cx * lx;
    s.a02 = 0.5 * s.cy * ly;
    return s;
}
// P(x, y), volume_of_fluid.f90:400-411
FEN_HD double vof_P(const VofSurf& s, double x, double y) {
    return s.cx * s.a20 * (x * x) + s.cy * s.a02 * (y * y) + s.a10 * x + s.a01 * y;
}

// the cell body of get_h_from_vof, volume_of_fluid.f90:245-293
FEN_HD void vof_h_d(double vof, double nx, double ny, double lx, double ly, double beta, double cut, double& h,
                    double& d) {
    if (vof <= cut || vof >= (1.0 - cut)) {
        h = vof;
        d = 0.0;
        return;
    }
    const VofSurf s = vof_surf(nx, ny, lx, ly);
    const double A = (1.0 - s.cx) * exp(2.0 * beta * s.a10) + (1.0 - s.cy) * exp(2.0 * beta * s.a01);
    const double Bp = (1.0 - s.cx) * exp(2.0 * beta * vof_P(s, 0.0, VOF_RP)) +
                      (1.0 - s.cy) * exp(2.0 * beta * vof_P(s, VOF_RP, 0.0));
    const double Bm = (1.0 - s.cx) * exp(2.0 * beta * vof_P(s, 0.0, VOF_RM)) +
                      (1.0 - s.cy) * exp(2.0 * beta * vof_P(s, VOF_RM, 0.0));
    const double Q = (1.0 - s.cx) * exp(2.0 * beta * s.a10 * (2.0 * vof - 1.0)) +
                     (1.0 - s.cy) * exp(2.0 * beta * s.a01 * (2.0 * vof - 1.0));
    const double aa = A * Bm * Bp * (A - Q);
    const double bb = A * (Bp + Bm) * (1.0 - Q);
    const double cc = 1.0 - A * Q;
    // solve_quadratic, :415-430
    const double disc = sqrt(bb * bb - 4.0 * aa * cc);
    const double x1 = (-bb + disc) / (2.0 * aa);
    const double x2 = (-bb - disc) / (2.0 * aa);
    const double root = fmax(x1, x2);
    d = log(root) / (2.0 * beta);
    h = 0.5 * (1.0 + tanh(beta * (vof_P(s, 0.5, 0.5) + d)));
}

// An_Int, volume_of_fluid.f90:660-672
FEN_HD double vof_an_int(const VofSurf& s, double beta, double a, double b, double xa, double xb, double ya, double yb,
                         double c, double d0) {
    return 0.5 * (b - a + 1.0 / (c * beta) * log(cosh(beta * (vof_P(s, xb, yb) + d0)) /
                                                   cosh(beta * (vof_P(s, xa, ya) + d0))));
}

// compute_flux, volume_of_fluid.f90:558-642: flux through the face whose normal velocity is u; (vof, nx, ...) are the
// values of the UPWIND cell (cell of the face when u >= 0, the next one otherwise).  dir 1 = x, 2 = y.
FEN_HD double vof_flux(int dir, double u, double dt, double delta, double beta, double cut, double vof, double nx,
                       double ny, double lx, double ly, double d) {
    double a, b, sgn;
    if (u >= 0.0) { a = 1.0 - dt * u / delta; b = 1.0; sgn = 1.0; }
    else { a = 0.0; b = -dt * u / delta; sgn = -1.0; }
    double xa, xb, ya, yb;
    if (dir == 1) { xa = a; xb = b; ya = 0.0; yb = 1.0; }
    else { xa = 0.0; xb = 1.0; ya = a; yb = b; }
    if (vof <= cut || vof >= (1.0 - cut)) return sgn * delta * vof * (xb - xa) * (yb - ya);
    const VofSurf s = vof_surf(nx, ny, lx, ly);
    if (s.xdom) {
        // analytical integration in x, numerical in y (Gauss points rm*(ya+yb), rp*(ya+yb): as written, :620-623)
        const double qm = vof_an_int(s, beta, xa, xb, xa, xb, VOF_RM * (ya + yb), VOF_RM * (yb + ya), s.a10, d);
        const double qp = vof_an_int(s, beta, xa, xb, xa, xb, VOF_RP * (ya + yb), VOF_RP * (yb + ya), s.a10, d);
        return sgn * delta * (0.5 * (qm + qp) * (yb - ya));
    }
    const double qm = vof_an_int(s, beta, ya, yb, VOF_RM * (xa + xb), VOF_RM * (xa + xb), ya, yb, s.a01, d);
    const double qp = vof_an_int(s, beta, ya, yb, VOF_RP * (xa + xb), VOF_RP * (xa + xb), ya, yb, s.a01, d);
    return sgn * delta * (0.5 * (qm + qp) * (xb - xa));
}

// variable-viscosity stress divergence of add_diffusion, navier_stokes.f90:420-447 (2-D).
// m[b][a] = mu(i+a-1, j+b-1), u[b][a] = v%x, v[b][a] = v%y on the same 3x3 patch.
FEN_HD void mf_stress_div(const double m[3][3], const double u[3][3], const double v[3][3], double id, double& dx,
                          double& dy) {
    const double tauxxip = 2.0 * m[1][2] * (u[1][2] - u[1][1]) * id;
    const double tauxxim = 2.0 * m[1][1] * (u[1][1] - u[1][0]) * id;
    const double dtauxxdx = (tauxxip - tauxxim) * id;
    const double tauxyjp = 0.25 * (m[1][1] + m[1][2] + m[2][1] + m[2][2]) *
                           ((u[2][1] - u[1][1]) * id + (v[1][2] - v[1][1]) * id);
    const double tauxyjm = 0.25 * (m[0][1] + m[0][2] + m[1][1] + m[1][2]) *
                           ((u[1][1] - u[0][1]) * id + (v[0][2] - v[0][1]) * id);
    const double dtauxydy = (tauxyjp - tauxyjm) * id;
    dx = dtauxxdx + dtauxydy;
    const double tauyxip = tauxyjp;
    const double tauyxim = 0.25 * (m[1][0] + m[1][1] + m[2][0] + m[2][1]) *
                           ((u[2][0] - u[1][0]) * id + (v[1][1] - v[1][0]) * id);
    const double dtauyxdx = (tauyxip - tauyxim) * id;
    const double tauyyjp = 2.0 * m[2][1] * (v[2][1] - v[1][1]) * id;
    const double tauyyjm = 2.0 * m[1][1] * (v[1][1] - v[0][1]) * id;
    const double dtauyydy = (tauyyjp - tauyyjm) * id;
    dy = dtauyxdx + dtauyydy;
}

// divergence-form centred advection of add_advection, navier_stokes.f90:297-327 (2-D), starting from RHS = 0.
// u[b][a] = v%x(i+a-1, j+b-1), v[b][a] = v%y(...).
FEN_HD void mf_advection(const double u[3][3], const double v[3][3], double id, double& ax, double& ay) {
    const double u0 = u[1][1], v0 = v[1][1];
    const double uuip = 0.25 * ((u[1][2] + u0) * (u[1][2] + u0));
    const double uuim = 0.25 * ((u[1][0] + u0) * (u[1][0] + u0));
    const double uvjp = (u[2][1] + u0) * (v[1][2] + v0) * 0.25;
    const double uvjm = (u0 + u[0][1]) * (v[0][2] + v[0][1]) * 0.25;
    ax = 0.0 - (uuip - uuim) * id - (uvjp - uvjm) * id;
    const double vuip = (v[1][2] + v0) * (u[2][1] + u0) * 0.25;
    const double vuim = (v0 + v[1][0]) * (u[2][0] + u[1][0]) * 0.25;
    const double vvjp = 0.25 * ((v[2][1] + v0) * (v[2][1] + v0));
    const double vvjm = 0.25 * ((v[0][1] + v0) * (v[0][1] + v0));
    ay = 0.0 - (vuip - vuim) * id - (vvjp - vvjm) * id;
}

// everything predicted_velocity_field does for one cell in the two-phase build (navier_stokes.f90:140-213 with
// compute_explicit_terms :217-257: advection, stress divergence :405-452, surface tension :458-501, body force; the
// pressure splitting of Dodd & Ferrante :174-184).  Index 0 = the cell, ip / jp = its +x / +y neighbour.
struct MfCell {
    double u[3][3], v[3][3], m[3][3];
    double rho0, rhoip, rhojp;          // rho
    double p0, pip, pjp;                // p
    double h0, hip, hjp;                // p_hat
    double c0, cip, cjp;                // curv
    double f0, fip, fjp;                // vof
    double sx, sy;                      // S
    double dvox, dvoy;                  // dv_o
};
struct MfPrm {
    double id, dt, A, B, g0, g1, sigma, irhomin;
    int has_source;                     // S was written by the caller (it is identically zero otherwise)
};
FEN_HD void mf_predict_cell(const MfCell& q, const MfPrm& k, double& un, double& vn, double& dvx, double& dvy) {
    const double rfx = 0.5 * (q.rhoip + q.rho0);                                  // center_to_face, fields.f90:197-198
    const double rfy = 0.5 * (q.rhojp + q.rho0);
    mf_advection(q.u, q.v, k.id, dvx, dvy);
    double sdx, sdy;
    mf_stress_div(q.m, q.u, q.v, k.id, sdx, sdy);
    dvx = dvx + sdx / rfx;                                                        // :430
    dvy = dvy + sdy / rfy;                                                        // :442
    if (k.sigma != 0.0) {             // the term is an exact zero otherwise
        dvx = dvx + k.sigma * 0.5 * (q.cip + q.c0) * (q.fip - q.f0) * k.id / rfx;    // :486-487
        dvy = dvy + k.sigma * 0.5 * (q.cjp + q.c0) * (q.fjp - q.f0) * k.id / rfy;    // :488-489
    }
    if (k.has_source) {
        dvx = dvx + q.sx / rfx;                                                   // :248-249
        dvy = dvy + q.sy / rfy;
    }
    const double gx = (q.pip - q.p0) * k.id, gy = (q.pjp - q.p0) * k.id;          // gradient(p), fields.f90:55-56
    const double hx = (q.hip - q.h0) * k.id, hy = (q.hjp - q.h0) * k.id;          // gradient(p_hat)
    double rx = -gx / rfx + k.A * dvx + k.B * q.dvox + k.g0;                      // :169
    double ry = -gy / rfy + k.A * dvy + k.B * q.dvoy + k.g1;                      // :170
    rx = rx + gx / rfx - k.irhomin * gx - (1.0 / rfx - k.irhomin) * hx;           // :176-177
    ry = ry + gy / rfy - k.irhomin * gy - (1.0 / rfy - k.irhomin) * hy;           // :178-179
    un = q.u[1][1] + k.dt * rx;                                                   // :190
    vn = q.v[1][1] + k.dt * ry;
}

}  // namespace fen
