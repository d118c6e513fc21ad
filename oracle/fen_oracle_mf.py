"""CPU restatement (numpy, fp64) of FEN's two-phase path: MTHINC volume of fluid + one-fluid Navier-Stokes.

TEST INFRASTRUCTURE ONLY (see oracle/fen_oracle.py): nothing under ``fen_b200/`` may import this module.

Follows, in the reference's own operation order (paths relative to /root/reference):
  * src/volume_of_fluid.f90  -- allocate_vof_fields :54-224, get_h_from_vof :228-303, compute_norm :307-396,
    P :400-411, solve_quadratic :415-430, advect_vof :434-554, compute_flux :558-642, Num_Int :646-656,
    An_Int :660-672, get_vof_from_distance :676-718, check_vof_integral :722-754
  * src/multiphase.f90       -- allocate_multiphase_fields :46-117, update_material_properties :121-137
  * src/navier_stokes.f90    -- the ``#ifdef MF`` branches: :80-96 (interface advection, material properties,
    p_hat), :113 (phi*rhomin/dt), :174-184 (pressure splitting of Dodd & Ferrante), :405-452 (variable-viscosity
    stress divergence, 2-D only), :458-501 (CSF surface tension), :526-531 (correction with 1/rhomin),
    :553-558 (p_o = p), :655-661 and :724 (time-step limits)
  * src/solver.f90:87-98     -- init_solver wiring for MF

The reference's VoF and variable-viscosity code is 2-D only (no z terms exist in compute_norm, compute_flux or the
stress divergence), so this oracle is 2-D only.

Parity pinning.  The reference ships two known-answer data sets for this path (the second, the rising-bubble
benchmark curves of rising_bubble/com_ref.txt, is checked in tests/test_oracle_mf.py too).  The first is
test/small_test/multiphase/capillary_wave/prosperetti.csv (Prosperetti's analytic capillary-wave amplitude), and its
post-processing (capillary_wave/postpro.py:92-101) measures the error of the maximum interface amplitude against it;
tests/test_oracle_mf.py replays that case and holds the error to the N^-1 guide line the script draws (0.4/N).  The
other reference tests of this path are plot-only (reversed vortex, Zalesak, reconstruction, viscous decay, rising
bubble shapes): their properties -- phase-volume conservation, return to the initial shape after flow reversal,
a flat interface under gravity staying at rest -- are asserted in the same test file.

Hazards reproduced on purpose (they change results):
  H13  ``vof = vof1`` (volume_of_fluid.f90:470,510) is an intrinsic derived-type assignment: it copies vof1's freshly
       allocated boundary-condition types too, so after the first call of advect_vof the vof field has the DEFAULT
       (periodic) boundary types on every face, whatever allocate_vof_fields had wired.
  H14  the curvature, normals and d left behind by advect_vof come from the INTERMEDIATE field vof1 (the second
       get_h_from_vof call), and the surface tension of the following predictor uses that curvature with the final vof.
  H15  the Gauss points of the numerically integrated direction are rm*(a+b), rp*(a+b) (not a + r*(b-a)).
  H16  dt_surf is only assigned when sigma > 0 (navier_stokes.f90:657-660) but update_timestep always takes
       min(dt_conv, dt_visc, dt_surf) (:724): with sigma = 0 it is an uninitialised module variable.  Here it is +inf.
"""
from __future__ import annotations

import math

import numpy as np

from . import fen_oracle as fo
from .fen_oracle import Grid, Scalar, Vector, PI

SMALL = 1.0e-14                                   # global.f90:16
GRAVITY = 9.80665                                 # global.f90:15
RP = 0.5 * (1.0 + 1.0 / math.sqrt(3.0))           # volume_of_fluid.f90:33
RM = 0.5 * (1.0 - 1.0 / math.sqrt(3.0))           # volume_of_fluid.f90:34


def _wire(G: Grid, fields, wall_type: int = 2) -> None:
    """Periodic -> 0, Wall -> 2 on the four faces (volume_of_fluid.f90:68-198, multiphase.f90:60-100)."""
    for n, face in enumerate(fo.FACES[:4]):
        s = G.boundary_conditions[n]
        if s == "Periodic":
            t = 0
        elif s == "Wall":
            t = wall_type
        else:
            continue                # prints 'ERROR: wrong bc on ... boundary' and keeps the default
        for f in fields:
            f.bc_type[face] = t


class VoF:
    """Module state + procedures of ``volume_of_fluid_mod`` (src/volume_of_fluid.f90)."""

    def __init__(self, G: Grid):
        if G.ndim != 2:
            raise ValueError("the reference's VoF solver is 2-D only")
        self.G = G
        self.beta = 1.0                    # :24
        self.quadratic = True              # :27
        self.x_first = True                # :30
        self.cut = 1.0e-8                  # :37
        self.distance = None               # :46
        # allocate_vof_fields :54-224
        self.vof = Scalar(G, 1, "c", "vof")
        self.h = Scalar(G, 1, "c", "h")
        self.d = Scalar(G, 1, "c", "d")
        self.curv = Scalar(G, 1, "c", "curv")
        self.norm = Vector(G, 1, "norm")
        self.l = Vector(G, 1, "l")
        _wire(G, [self.vof, self.h, self.d, self.curv, self.norm.x, self.norm.y, self.l.x, self.l.y])

    # ---- reconstruction ------------------------------------------------------------------------------------
    def compute_norm(self) -> None:
        """volume_of_fluid.f90:307-396 (Youngs normals at the four corners, quadratic-surface curvature)."""
        G = self.G
        delta = G.delta
        idelta = 1.0 / delta
        idelta2 = 1.0 / delta ** 2
        F = self.vof.sh
        f00 = F()
        mx1 = 0.5 * (F(0, -1) + f00 - F(-1, -1) - F(-1, 0)) * idelta
        mx2 = 0.5 * (f00 + F(0, 1) - F(-1, 0) - F(-1, 1)) * idelta
        mx3 = 0.5 * (F(1, 0) + F(1, 1) - f00 - F(0, 1)) * idelta
        mx4 = 0.5 * (F(1, -1) + F(1, 0) - F(0, -1) - f00) * idelta
        mxc = 0.25 * (mx1 + mx2 + mx3 + mx4)
        my1 = 0.5 * (F(-1, 0) + f00 - F(-1, -1) - F(0, -1)) * idelta
        my2 = 0.5 * (F(-1, 1) + F(0, 1) - F(-1, 0) - f00) * idelta
        my3 = 0.5 * (F(0, 1) + F(1, 1) - f00 - F(1, 0)) * idelta
        my4 = 0.5 * (f00 + F(1, 0) - F(0, -1) - F(1, -1)) * idelta
        myc = 0.25 * (my1 + my2 + my3 + my4)
        nx, ny = [], []
        for mx, my in ((mx1, my1), (mx2, my2), (mx3, my3), (mx4, my4)):
            r = np.sqrt(mx ** 2 + my ** 2 + SMALL)
            nx.append(mx / r)
            ny.append(my / r)
        rc = np.sqrt(mxc ** 2 + myc ** 2 + SMALL)
        self.norm.x.I[...] = mxc / rc
        self.norm.y.I[...] = myc / rc
        if self.quadratic:
            self.l.x.I[...] = 0.5 * delta * (nx[3] + nx[2] - nx[1] - nx[0])
            self.l.y.I[...] = 0.5 * delta * (ny[1] + ny[2] - ny[0] - ny[3])
        else:
            self.l.x.I[...] = 0.0
            self.l.y.I[...] = 0.0
        self.curv.I[...] = -(self.l.x.I + self.l.y.I) * idelta2
        self.curv.update_ghost_nodes()
        self.norm.update_ghost_nodes()
        self.l.update_ghost_nodes()

    @staticmethod
    def _coeffs(nx, ny, lx, ly):
        """cx, cy and Eq. 12 coefficients (volume_of_fluid.f90:254-266, 610-634)."""
        xdom = np.abs(nx) == np.maximum(np.abs(nx), np.abs(ny))
        cx = np.where(xdom, 0.0, 1.0)
        cy = np.where(xdom, 1.0, 0.0)
        a10 = nx - 0.5 * cx * lx
        a01 = ny - 0.5 * cy * ly
        a20 = 0.5 * cx * lx
        a02 = 0.5 * cy * ly
        return xdom, cx, cy, a10, a01, a20, a02

    @staticmethod
    def _P(cx, cy, a10, a01, a20, a02, x, y):
        """volume_of_fluid.f90:400-411."""
        return cx * a20 * x ** 2 + cy * a02 * y ** 2 + a10 * x + a01 * y

    def get_h_from_vof(self) -> None:
        """volume_of_fluid.f90:228-303."""
        self.compute_norm()
        beta = self.beta
        vof = self.vof.sh()
        full = (vof <= self.cut) | (vof >= (1.0 - self.cut))
        nx, ny, lx, ly = self.norm.x.sh(), self.norm.y.sh(), self.l.x.sh(), self.l.y.sh()
        _, cx, cy, a10, a01, a20, a02 = self._coeffs(nx, ny, lx, ly)
        P = lambda x, y: self._P(cx, cy, a10, a01, a20, a02, x, y)   # noqa: E731
        with np.errstate(all="ignore"):
            A = (1.0 - cx) * np.exp(2.0 * beta * a10) + (1.0 - cy) * np.exp(2.0 * beta * a01)
            Bp = (1.0 - cx) * np.exp(2.0 * beta * P(0.0, RP)) + (1.0 - cy) * np.exp(2.0 * beta * P(RP, 0.0))
            Bm = (1.0 - cx) * np.exp(2.0 * beta * P(0.0, RM)) + (1.0 - cy) * np.exp(2.0 * beta * P(RM, 0.0))
            Q = (1.0 - cx) * np.exp(2.0 * beta * a10 * (2.0 * vof - 1.0)) + \
                (1.0 - cy) * np.exp(2.0 * beta * a01 * (2.0 * vof - 1.0))
            aa = A * Bm * Bp * (A - Q)
            bb = A * (Bp + Bm) * (1.0 - Q)
            cc = 1.0 - A * Q
            # solve_quadratic :415-430
            disc = np.sqrt(bb ** 2 - 4.0 * aa * cc)
            x1 = (-bb + disc) / (2.0 * aa)
            x2 = (-bb - disc) / (2.0 * aa)
            root = np.maximum(x1, x2)
            dd = np.log(root) / (2.0 * beta)
            hh = 0.5 * (1.0 + np.tanh(beta * (P(0.5, 0.5) + dd)))
        self.h.I[...] = np.where(full, vof, hh)
        self.d.I[...] = np.where(full, 0.0, dd)
        self.h.update_ghost_nodes()
        self.d.update_ghost_nodes()

    # ---- advection -----------------------------------------------------------------------------------------
    def _flux(self, direction: int, u_face: np.ndarray, dt: float) -> np.ndarray:
        """compute_flux (volume_of_fluid.f90:558-642) at every face 0..N of ``direction`` for the interior
        transverse range.  ``u_face`` has the face index 0..N along ``direction``."""
        G = self.G
        delta = G.delta
        beta = self.beta
        n = G.Nx if direction == 1 else G.Ny
        pos = u_face >= 0.0

        def upwind(s: Scalar):
            f = s.f[:, :, s.gl]
            if direction == 1:
                lo, hi = f[0:n + 1, 1:G.Ny + 1], f[1:n + 2, 1:G.Ny + 1]
            else:
                lo, hi = f[1:G.Nx + 1, 0:n + 1], f[1:G.Nx + 1, 1:n + 2]
            return np.where(pos, lo, hi)

        vof, nx, ny, lx, ly, dd = (upwind(s) for s in (self.vof, self.norm.x, self.norm.y, self.l.x, self.l.y, self.d))
        a = np.where(pos, 1.0 - dt * u_face / delta, 0.0)
        b = np.where(pos, 1.0, -dt * u_face / delta)
        sgn = np.where(pos, 1.0, -1.0)
        if direction == 1:
            xa, xb, ya, yb = a, b, 0.0, 1.0
        else:
            xa, xb, ya, yb = 0.0, 1.0, a, b
        full = (vof <= self.cut) | (vof >= (1.0 - self.cut))
        f_full = sgn * delta * vof * (xb - xa) * (yb - ya)
        xdom, cx, cy, a10, a01, a20, a02 = self._coeffs(nx, ny, lx, ly)
        P = lambda x, y: self._P(cx, cy, a10, a01, a20, a02, x, y)   # noqa: E731

        def an_int(a_, b_, xa_, xb_, ya_, yb_, c_, d0):            # :660-672
            return 0.5 * (b_ - a_ + 1.0 / (c_ * beta) * np.log(np.cosh(beta * (P(xb_, yb_) + d0)) /
                                                                np.cosh(beta * (P(xa_, ya_) + d0))))

        def num_int(a_, b_, qrm, qrp):                             # :646-656
            return 0.5 * (qrm + qrp) * (b_ - a_)

        with np.errstate(all="ignore"):
            fx = sgn * delta * num_int(ya, yb,
                                       an_int(xa, xb, xa, xb, RM * (ya + yb), RM * (yb + ya), a10, dd),
                                       an_int(xa, xb, xa, xb, RP * (ya + yb), RP * (yb + ya), a10, dd))
            fy = sgn * delta * num_int(xa, xb,
                                       an_int(ya, yb, RM * (xa + xb), RM * (xa + xb), ya, yb, a01, dd),
                                       an_int(ya, yb, RP * (xa + xb), RP * (xa + xb), ya, yb, a01, dd))
        return np.where(full, f_full, np.where(xdom, fx, fy))

    def _sweep(self, direction: int, v: Vector, dt: float, src: np.ndarray) -> np.ndarray:
        """One directional split step (volume_of_fluid.f90:457-466, 478-487); ``src`` is the interior of the field
        being advanced, fluxes come from the current ``self.vof`` reconstruction."""
        G = self.G
        delta = G.delta
        if direction == 1:
            uf = v.x.f[0:G.Nx + 1, 1:G.Ny + 1, v.x.gl]
            F = self._flux(1, uf, dt)
            fp, fm = F[1:, :], F[:-1, :]
            du = uf[1:, :] - uf[:-1, :]
        else:
            uf = v.y.f[1:G.Nx + 1, 0:G.Ny + 1, v.y.gl]
            F = self._flux(2, uf, dt)
            fp, fm = F[:, 1:], F[:, :-1]
            du = uf[:, 1:] - uf[:, :-1]
        return (src - (fp - fm) / delta) / (1.0 - dt * du / delta)

    def advect_vof(self, v: Vector, dt: float) -> None:
        """volume_of_fluid.f90:434-554."""
        G = self.G
        delta = G.delta
        vof1 = Scalar(G, 1, "c")                                     # :448-449
        vof2 = Scalar(G, 1, "c")
        self.get_h_from_vof()                                        # :456
        d1, d2 = (1, 2) if self.x_first else (2, 1)
        vof1.I[..., 0] = self._sweep(d1, v, dt, self.vof.I[..., 0])
        # vof = vof1 (:470, :510): derived-type assignment, boundary types included (hazard H13)
        self.vof.f[...] = vof1.f
        self.vof.bc_type = dict(vof1.bc_type)
        self.vof.update_ghost_nodes()
        self.get_h_from_vof()
        vof2.I[..., 0] = self._sweep(d2, v, dt, vof1.I[..., 0])
        dux = (v.x.sh() - v.x.sh(-1, 0, 0))
        dvy = (v.y.sh() - v.y.sh(0, -1, 0))
        if self.x_first:                                             # :492-498
            self.vof.I[...] = vof2.I - dt * (vof1.I * dux / delta + vof2.I * dvy / delta)
        else:                                                        # :532-538
            self.vof.I[...] = vof2.I - dt * (vof2.I * dux / delta + vof1.I * dvy / delta)
        self.x_first = not self.x_first
        self.vof.update_ghost_nodes()                                # :546

    # ---- initialisation / diagnostics ------------------------------------------------------------------------
    def get_vof_from_distance(self) -> None:
        """volume_of_fluid.f90:676-718; ``self.distance(x, y)`` must accept numpy arrays."""
        if self.distance is None:
            raise ValueError("ERROR: distance function not defined.")
        G = self.G
        delta, beta = G.delta, self.beta
        x = G.x[1:G.Nx + 1][:, None]
        y = G.y[1:G.Ny + 1][None, :]
        xp = x + delta * (RP - 0.5)
        xm = x + delta * (RM - 0.5)
        yp = y + delta * (RP - 0.5)
        ym = y + delta * (RM - 0.5)
        D = self.distance
        T = lambda xx, yy: 0.5 * (1.0 + np.tanh(beta * D(xx + 0.0 * yy, yy + 0.0 * xx) / delta))   # noqa: E731
        self.vof.I[..., 0] = 0.5 * (0.5 * (T(xm, ym) + T(xp, ym)) + 0.5 * (T(xm, yp) + T(xp, yp)))
        self.h.I[..., 0] = T(x, y)
        self.vof.update_ghost_nodes()
        self.h.update_ghost_nodes()

    def check_vof_integral(self):
        """volume_of_fluid.f90:722-754 (note delta**3 although the case is 2-D)."""
        d3 = self.G.delta * self.G.delta * self.G.delta
        return float(fo._seq_sum(self.vof.I)) * d3, float(fo._seq_sum(1.0 - self.vof.I)) * d3


class MultiphaseNavierStokes(fo.NavierStokes):
    """navier_stokes_mod compiled with -DMF, plus multiphase_mod and the MF part of init_solver."""

    def __init__(self, G: Grid, rho_0=1.0, rho_1=1.0, mu_0=1.0, mu_1=1.0, sigma=0.0, distance=None, beta=1.0):
        super().__init__(G, 1.0, 1.0)                                # density = viscosity = 1 (navier_stokes.f90:18)
        self.rho_0, self.rho_1, self.mu_0, self.mu_1, self.sigma = (float(rho_0), float(rho_1), float(mu_0),
                                                                    float(mu_1), float(sigma))
        self.constant_viscosity = False                              # solver.f90:83
        self.dt_surf = math.inf                                      # hazard H16
        self.vf = VoF(G)                                             # solver.f90:87
        self.vf.beta = float(beta)
        self.vf.distance = distance
        # allocate_multiphase_fields, multiphase.f90:46-117
        self.p_hat = Scalar(G, 1, "c", "p_hat")
        self.p_o = Scalar(G, 1, "c", "p_o")
        self.grad_p_hat = Vector(G, 0, "grad_p_hat")
        _wire(G, [self.p_hat, self.p_o])
        self.vf.get_vof_from_distance()                              # solver.f90:89
        self.update_material_properties()                            # solver.f90:90
        self.rhomin = min(self.rho_0, self.rho_1)                    # solver.f90:92
        self.irhomin = 1.0 / self.rhomin

    @property
    def vof(self):
        return self.vf.vof

    def update_material_properties(self) -> None:
        """multiphase.f90:121-137."""
        vof = self.vf.vof.f
        self.rho.f[...] = self.rho_1 * vof + self.rho_0 * (1.0 - vof)
        self.mu.f[...] = self.mu_1 * vof + self.mu_0 * (1.0 - vof)
        self.rho.update_ghost_nodes()
        self.mu.update_ghost_nodes()

    def set_timestep(self, U: float) -> float:
        """navier_stokes.f90:623-666 with MF (2-D)."""
        d = self.G.delta
        self.dt_conv = self.CFL * d / U
        self.dt_visc = 0.125 * d * d * min(self.rho_0 / self.mu_0, self.rho_1 / self.mu_1)   # :655
        dt = min(self.dt_conv, self.dt_visc)
        if self.sigma > 0.0:
            self.dt_surf = math.sqrt(0.5 * (self.rho_0 + self.rho_1) * d ** 3 / (PI * self.sigma + 1.0e-16))
            dt = min(dt, self.dt_surf)
        self.dt_o = dt
        return dt

    def update_timestep(self, dt: float) -> float:
        """navier_stokes.f90:670-730 with MF."""
        self.dt_o = dt
        max_vel = self._max_vel()
        self.dt_conv = self.CFL * self.G.delta / max_vel if max_vel > 0.0 else 1.0
        dt = min(self.dt_conv, self.dt_visc, self.dt_surf)           # :724
        if dt > 1.1 * self.dt_o:
            dt = 1.1 * self.dt_o
        return dt

    def add_diffusion(self, RHS: Vector) -> None:
        """navier_stokes.f90:405-452: div(2 mu D)/rhof, 2-D."""
        idelta = 1.0 / self.G.delta
        M, U, V = self.mu.sh, self.v.x.sh, self.v.y.sh
        tauxxip = 2.0 * M(1, 0) * (U(1, 0) - U()) * idelta
        tauxxim = 2.0 * M() * (U() - U(-1, 0)) * idelta
        dtauxxdx = (tauxxip - tauxxim) * idelta
        tauxyjp = 0.25 * (M() + M(1, 0) + M(0, 1) + M(1, 1)) * ((U(0, 1) - U()) * idelta + (V(1, 0) - V()) * idelta)
        tauxyjm = 0.25 * (M(0, -1) + M(1, -1) + M() + M(1, 0)) * \
            ((U() - U(0, -1)) * idelta + (V(1, -1) - V(0, -1)) * idelta)
        dtauxydy = (tauxyjp - tauxyjm) * idelta
        RHS.x.I[...] = RHS.x.I + (dtauxxdx + dtauxydy) / self.rhof.x.I
        tauyxip = tauxyjp
        tauyxim = 0.25 * (M(-1, 0) + M() + M(-1, 1) + M(0, 1)) * \
            ((U(-1, 1) - U(-1, 0)) * idelta + (V() - V(-1, 0)) * idelta)
        dtauyxdx = (tauyxip - tauyxim) * idelta
        tauyyjp = 2.0 * M(0, 1) * (V(0, 1) - V()) * idelta
        tauyyjm = 2.0 * M() * (V() - V(0, -1)) * idelta
        dtauyydy = (tauyyjp - tauyyjm) * idelta
        RHS.y.I[...] = RHS.y.I + (dtauyxdx + dtauyydy) / self.rhof.y.I

    def add_surface_tension(self, RHS: Vector) -> None:
        """navier_stokes.f90:458-501 (CSF)."""
        idelta = 1.0 / self.G.delta
        C, F = self.vf.curv.sh, self.vf.vof.sh
        RHS.x.I[...] = RHS.x.I + self.sigma * 0.5 * (C(1, 0) + C()) * (F(1, 0) - F()) * idelta / self.rhof.x.I
        RHS.y.I[...] = RHS.y.I + self.sigma * 0.5 * (C(0, 1) + C()) * (F(0, 1) - F()) * idelta / self.rhof.y.I

    def compute_explicit_terms(self, RHS: Vector) -> None:
        """navier_stokes.f90:217-257 with MF."""
        for comp in RHS.comps:
            comp.f[...] = 0.0
        self.add_advection(RHS)
        self.add_diffusion(RHS)
        self.add_surface_tension(RHS)                                # :241-243
        for comp, s, rf in zip(RHS.comps, self.S.comps, self.rhof.comps):
            comp.I[...] = comp.I + s.I / rf.I

    def predicted_velocity_field(self, dt: float) -> None:
        """navier_stokes.f90:140-213 with the MF pressure splitting (:174-184)."""
        A = 1.0 + 0.5 * dt / self.dt_o
        B = -0.5 * dt / self.dt_o
        fo.center_to_face(self.rho, self.rhof)
        self.compute_explicit_terms(self.dv)
        fo.gradient(self.p, self.grad_p)
        fo.gradient(self.p_hat, self.grad_p_hat)                     # :175
        for d, (vc, gp, gph, rf, dv, dvo) in enumerate(zip(self.v.comps, self.grad_p.comps, self.grad_p_hat.comps,
                                                           self.rhof.comps, self.dv.comps, self.dv_o.comps)):
            RHS = -gp.I / rf.I + A * dv.I + B * dvo.I + self.g[d]    # :169-170
            RHS = RHS + gp.I / rf.I - self.irhomin * gp.I - (1.0 / rf.I - self.irhomin) * gph.I   # :176-179
            vc.I[...] = vc.I + dt * RHS
        for dv, dvo in zip(self.dv.comps, self.dv_o.comps):
            dvo.f[...] = dv.f
        self.v.update_ghost_nodes()

    def correct_velocity_field(self, dt: float) -> None:
        """navier_stokes.f90:505-546, MF branch :526-531."""
        fo.gradient(self.phi, self.grad_p)
        for vc, gp in zip(self.v.comps, self.grad_p.comps):
            vc.I[...] = vc.I - gp.I * dt * self.irhomin
        self.v.update_ghost_nodes()

    def update_pressure(self) -> None:
        """navier_stokes.f90:550-566 with MF: p_o = p first."""
        self.p_o.f[...] = self.p.f                                   # :557
        super().update_pressure()

    def navier_stokes_solver(self, step: int, dt: float) -> float:
        """navier_stokes.f90:50-136 with MF."""
        if self.constant_CFL:
            dt = self.update_timestep(dt)
        self.vf.advect_vof(self.v, dt)                               # :82
        self.update_material_properties()                            # :85
        if self.constant_CFL:                                        # :88-92
            self.p_hat.f[...] = self.p_o.f + (dt + self.dt_o) * (self.p.f - self.p_o.f) / self.dt_o
        else:
            self.p_hat.f[...] = 2.0 * self.p.f - self.p_o.f
        self.p_hat.update_ghost_nodes()                              # :95
        self.predicted_velocity_field(dt)
        fo.divergence(self.v, self.phi)
        self.phi.f[...] = self.phi.f * self.rhomin / dt              # :113
        self.poisson.solve(self.phi)
        self.phi.update_ghost_nodes()
        self.correct_velocity_field(dt)
        self.update_pressure()
        self.checks(dt)
        return dt

    advance_solution = navier_stokes_solver
