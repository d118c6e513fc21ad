"""ctypes front end of the plain-C two-phase restatement (oracle/fen_oracle_mf_c.c): 2-D, x periodic, walls in y.

TEST INFRASTRUCTURE ONLY -- see the header of fen_oracle_mf_c.c.  The state is loaded from / compared with a
``fen_oracle_mf.MultiphaseNavierStokes`` object (``from_oracle`` / ``get``)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import build_c

_lib = None
(P, PHI, RHO, MU, U, V, VOF, H, D, CURV, NORMX, NORMY, LX, LY, PHAT, PO, DVOX, DVOY) = range(18)


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build_c.build_mf())
        d, i, p = C.c_double, C.c_int, C.c_void_p
        lib.fomf_create.restype = p
        lib.fomf_create.argtypes = [i, i, d, d, d, d, d, d, d]
        lib.fomf_destroy.argtypes = [p]
        lib.fomf_field_size.restype = C.c_long
        lib.fomf_field_size.argtypes = [p]
        lib.fomf_set_field.argtypes = [p, i, p]
        lib.fomf_get_field.argtypes = [p, i, p]
        lib.fomf_set_params.argtypes = [p, d, d, d]
        lib.fomf_set_vof_state.argtypes = [p, i, i]
        lib.fomf_set_wall_velocity.argtypes = [p, d, d]
        lib.fomf_x_first.argtypes = [p]
        lib.fomf_vof_bc_y.argtypes = [p]
        lib.fomf_step.argtypes = [p, d]
        lib.fomf_advect_vof.argtypes = [p, d]
        lib.fomf_update_material_properties.argtypes = [p]
        lib.fomf_poisson_solve.argtypes = [p, p]
        lib.fomf_maxdiv.restype = d
        lib.fomf_maxdiv.argtypes = [p]
        lib.fomf_maxcfl.restype = d
        lib.fomf_maxcfl.argtypes = [p, d]
        lib.fomf_threads.restype = i
        lib.fomf_set_threads.argtypes = [i]
        _lib = lib
    return _lib


class MultiphaseC:
    """Two-phase solver state in C; fields go in and out as Fortran-ordered (nx+2, ny+2) ghosted arrays."""

    def __init__(self, nx, ny, delta, rho_0, rho_1, mu_0, mu_1, sigma, beta=1.0, threads=0):
        self.lib = load()
        if threads:
            self.lib.fomf_set_threads(int(threads))
        self.shape = (nx + 2, ny + 2)
        self.h = C.c_void_p(self.lib.fomf_create(nx, ny, float(delta), float(rho_0), float(rho_1), float(mu_0),
                                                 float(mu_1), float(sigma), float(beta)))
        if not self.h:
            raise ValueError("fen_oracle_mf_c: nx must be a power of two")
        assert self.lib.fomf_field_size(self.h) == self.shape[0] * self.shape[1]
        self.dt_o = 0.0
        self.g = [0.0, 0.0]

    @classmethod
    def from_oracle(cls, ns, threads=0):
        """A C twin of a fen_oracle_mf.MultiphaseNavierStokes in its current state (x periodic, walls in y)."""
        G = ns.G
        assert G.ndim == 2 and G.boundary_conditions[:4] == ["Periodic", "Periodic", "Wall", "Wall"]
        c = cls(G.Nx, G.Ny, G.delta, ns.rho_0, ns.rho_1, ns.mu_0, ns.mu_1, ns.sigma, ns.vf.beta, threads)
        pairs = ((P, ns.p), (PHI, ns.phi), (RHO, ns.rho), (MU, ns.mu), (U, ns.v.x), (V, ns.v.y), (VOF, ns.vf.vof),
                 (H, ns.vf.h), (D, ns.vf.d), (CURV, ns.vf.curv), (NORMX, ns.vf.norm.x), (NORMY, ns.vf.norm.y),
                 (LX, ns.vf.l.x), (LY, ns.vf.l.y), (PHAT, ns.p_hat), (PO, ns.p_o))
        for fid, s in pairs:
            c.set(fid, s.f[:, :, s.gl] if s.gl else None)
        for fid, s in ((DVOX, ns.dv_o.x), (DVOY, ns.dv_o.y)):       # gl = 0 in the oracle: interior only
            a = np.zeros(c.shape, order="F")
            a[1:-1, 1:-1] = s.f[:, :, 0]
            c.set(fid, a)
        c.dt_o = ns.dt_o
        c.g = [float(ns.g[0]), float(ns.g[1])]
        c.lib.fomf_set_vof_state(c.h, 1 if ns.vf.x_first else 0, int(ns.vf.vof.bc_type["bottom"]))
        bot, top = ns.v.x.bc["bottom"], ns.v.x.bc["top"]          # uniform wall velocities only
        assert (bot == bot.flat[0]).all() and (top == top.flat[0]).all()
        assert not ns.v.y.bc["bottom"].any() and not ns.v.y.bc["top"].any()
        c.lib.fomf_set_wall_velocity(c.h, float(bot.flat[0]), float(top.flat[0]))
        return c

    @property
    def threads(self):
        return int(self.lib.fomf_threads())

    @property
    def x_first(self):
        return bool(self.lib.fomf_x_first(self.h))

    def set(self, fid, a):
        a = np.asfortranarray(a, dtype=np.float64)
        assert a.shape == self.shape
        self.lib.fomf_set_field(self.h, fid, a.ctypes.data_as(C.c_void_p))

    def get(self, fid):
        a = np.empty(self.shape, dtype=np.float64, order="F")
        self.lib.fomf_get_field(self.h, fid, a.ctypes.data_as(C.c_void_p))
        return a

    def navier_stokes_solver(self, step, dt):
        self.lib.fomf_set_params(self.h, float(self.dt_o), float(self.g[0]), float(self.g[1]))
        self.lib.fomf_step(self.h, float(dt))
        return dt

    def advect_vof(self, dt):
        self.lib.fomf_advect_vof(self.h, float(dt))

    @property
    def maxdiv(self):
        return float(self.lib.fomf_maxdiv(self.h))

    def maxCFL(self, dt):
        return float(self.lib.fomf_maxcfl(self.h, float(dt)))

    def destroy(self):
        if self.h:
            self.lib.fomf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
