"""ctypes front end of the plain-C restatement (oracle/fen_oracle_c.c): the periodic-box (ppp) fractional step and,
with ``zwalls=True``, the channel (x / y periodic, walls in z, ppn Poisson).

TEST INFRASTRUCTURE ONLY -- see the header of fen_oracle_c.c.  Mirrors the small part of fen_oracle.NavierStokes
the checks and the CPU baseline need."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import build_c

_lib = None
P, PHI, U, V, W, DVOX, DVOY, DVOZ = range(8)


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build_c.build())
        lib.foc_create.restype = C.c_void_p
        lib.foc_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        lib.foc_destroy.argtypes = [C.c_void_p]
        lib.foc_field_size.restype = C.c_long
        lib.foc_field_size.argtypes = [C.c_void_p]
        lib.foc_set_params.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double]
        lib.foc_set_field.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.foc_get_field.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.foc_step.argtypes = [C.c_void_p, C.c_double]
        lib.foc_poisson_solve.argtypes = [C.c_void_p, C.c_void_p]
        lib.foc_maxdiv.restype = C.c_double
        lib.foc_maxdiv.argtypes = [C.c_void_p]
        lib.foc_maxcfl.restype = C.c_double
        lib.foc_maxcfl.argtypes = [C.c_void_p, C.c_double]
        lib.foc_set_zwalls.argtypes = [C.c_void_p, C.c_int]
        lib.foc_threads.restype = C.c_int
        lib.foc_set_threads.argtypes = [C.c_int]
        _lib = lib
    return _lib


class NavierStokesC:
    """3-D single-phase solver state in C (periodic box or channel); fields go in and out as Fortran-ordered ghosted
    arrays."""

    def __init__(self, nx, ny, nz, delta, density=1.0, viscosity=1.0, threads=0, zwalls=False):
        self.lib = load()
        if threads:
            self.lib.foc_set_threads(int(threads))
        self.shape = (nx + 2, ny + 2, nz + 2)
        self.delta = float(delta)
        self.h = C.c_void_p(self.lib.foc_create(nx, ny, nz, float(delta), float(density), float(viscosity)))
        assert self.lib.foc_field_size(self.h) == self.shape[0] * self.shape[1] * self.shape[2]
        self.dt_o = 0.0
        self.g = [0.0, 0.0, 0.0]
        if zwalls:                       # channel: bc(5:6) = 'Wall' -> ppn Poisson, no-slip walls in z
            self.lib.foc_set_zwalls(self.h, 1)

    @property
    def threads(self):
        return int(self.lib.foc_threads())

    def set(self, fid, a):
        a = np.asfortranarray(a, dtype=np.float64)
        assert a.shape == self.shape
        self.lib.foc_set_field(self.h, fid, a.ctypes.data_as(C.c_void_p))

    def get(self, fid):
        a = np.empty(self.shape, dtype=np.float64, order="F")
        self.lib.foc_get_field(self.h, fid, a.ctypes.data_as(C.c_void_p))
        return a

    def navier_stokes_solver(self, step, dt):
        self.lib.foc_set_params(self.h, float(self.dt_o), *[float(x) for x in self.g])
        self.lib.foc_step(self.h, float(dt))
        return dt

    def solve_poisson(self, phi):
        """phi: Fortran-ordered ghosted array, solved in place (interior)."""
        assert phi.flags.f_contiguous and phi.shape == self.shape
        self.lib.foc_poisson_solve(self.h, phi.ctypes.data_as(C.c_void_p))

    @property
    def maxdiv(self):
        return float(self.lib.foc_maxdiv(self.h))

    def maxCFL(self, dt):
        return float(self.lib.foc_maxcfl(self.h, float(dt)))

    def destroy(self):
        if self.h:
            self.lib.foc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass
