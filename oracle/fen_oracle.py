"""CPU restatement (numpy/scipy, fp64) of FEN's fractional-step hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``fen_b200/`` may import this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and only as the checker or as the timed CPU
baseline -- never as the product path.

The reference (fradevita/FEN, Fortran + MPI + FFTW3 + 2decomp&FFT) cannot be
built in this image (no Fortran compiler, no MPI, no FFTW), so this file
restates its algorithm function by function, in the reference's own operation
order, citing the reference ``file:line`` each function follows.  Paths are
relative to ``/root/reference``.

Parity pinning: the reference ships no golden fields.  This oracle is pinned by
the reference's own pass criteria, reproduced in ``tests/test_oracle_*.py``:
  * Taylor-Green 2-D convergence slope > 1.8
    (test/small_test/navier_stokes/taylor_green_vortex/postpro.py:62)
  * Poisson convergence slope > 1.8 for every variant
    (test/small_test/poisson/convergence_rate/postpro.py:28)
  * projection max|div| <= 1e-11 (test/small_test/poisson/projection/projection.f90:125)
  * Poiseuille slope > 1.8 (test/small_test/navier_stokes/poiseuille/postpro.py:62)
  * periodic ghost fill within 1e-14 (test/small_test/fields/methods.f90:208-280)
  * advection slope > 1.8 (test/small_test/navier_stokes/advection/postpro.py:40)
  * ABC flow slope > 1.5 (test/large_test/ABC/postpro.py:42)

Third-party arithmetic restated from published definitions (not vendored in the
reference): FFTW 3.3.10 (INSTALL.sh:16-17) r2c/c2r/c2c unnormalised transforms
== scipy.fft.rfft / irfft(norm="forward") / fft / ifft(norm="forward");
REDFT10 == scipy.fft.dct(type=2), REDFT01 == scipy.fft.dct(type=3), both
unnormalised.  2decomp&FFT (unpinned HEAD, INSTALL.sh:28-31) only moves data.

Array convention: every field is a numpy array indexed ``f[i, j, k]`` in
Fortran (column-major) memory order with ``gl`` ghost layers on every side, so
reference index ``(i, j, k)`` (1-based interior) lives at ``f[i-1+gl, j-1+gl,
k-1+gl]`` and ``f.tobytes(order="F")`` matches the reference's raw dumps.
Single rank only (prow = pcol = 1): the reference's result does not depend on
the rank count except for the order of two scalar reductions.
"""
from __future__ import annotations

import math
import os

import numpy as np
import scipy.fft as sfft

PI = math.acos(-1.0)  # global.f90:14

_WORKERS = int(os.environ.get("FEN_ORACLE_WORKERS", "0")) or (os.cpu_count() or 1)


def set_workers(n: int) -> None:
    """Number of threads scipy.fft may use (CPU-baseline timing only)."""
    global _WORKERS
    _WORKERS = max(1, int(n))


def _f32(n) -> float:
    """Fortran ``float(n)``: default-real (single precision) conversion (hazard H1)."""
    return float(np.float32(n))


# --------------------------------------------------------------------------------------
# grid.f90
# --------------------------------------------------------------------------------------
class Grid:
    """``type grid`` and ``grid%setup`` -- src/grid.f90:22-62, 67-200."""

    def __init__(self, Nx, Ny, Nz, Lx, Ly, Lz, x0=(0.0, 0.0, 0.0), prow=1, pcol=1,
                 bc=None, ndim=None):
        self.ndim = ndim if ndim is not None else (2 if Nz == 1 else 3)  # cpp macro DIM
        nb = 6 if self.ndim == 3 else 4
        self.boundary_conditions = ["Periodic"] * nb                     # grid.f90:97-105
        if bc is not None:
            bc = list(bc)
            assert len(bc) == nb, "bc must have %d entries" % nb
            self.boundary_conditions = bc
        b = self.boundary_conditions
        pbc = [False, False, True]                                       # grid.f90:107
        pbc[0] = b[0] == "Periodic" and b[1] == "Periodic"
        pbc[1] = b[2] == "Periodic" and b[3] == "Periodic"
        if self.ndim == 3:
            pbc[2] = b[4] == "Periodic" and b[5] == "Periodic"
        self.periodic_bc = pbc
        self.Nx, self.Ny, self.Nz = int(Nx), int(Ny), int(Nz)
        self.Lx, self.Ly, self.Lz = float(Lx), float(Ly), float(Lz)
        self.origin = tuple(float(v) for v in x0)
        self.delta = self.Lx / _f32(Nx)                                  # grid.f90:140
        self.spacing_ok = (self.Lx / _f32(Nx)) == (self.Ly / _f32(Ny))   # grid.f90:143 (warn only)
        # cell-centre coordinates, index 0..N+1                          # grid.f90:153-164
        self.x = self.origin[0] + (np.arange(0, Nx + 2) - 0.5) * self.delta
        self.y = self.origin[1] + (np.arange(0, Ny + 2) - 0.5) * self.delta
        self.z = self.origin[2] + (np.arange(0, Nz + 2) - 0.5) * self.delta
        if prow != 1 or pcol != 1:
            raise ValueError("the oracle is single-rank (prow = pcol = 1)")
        self.prow, self.pcol, self.rank, self.nranks = 1, 1, 0, 1
        self.lo = (1, 1, 1)
        self.hi = (self.Nx, self.Ny, self.Nz)

    @property
    def shape(self):
        return (self.Nx, self.Ny, self.Nz)


# --------------------------------------------------------------------------------------
# scalar.f90 / vector.f90
# --------------------------------------------------------------------------------------
FACES = ("left", "right", "bottom", "top", "front", "back")


class Scalar:
    """``type scalar`` -- src/scalar.f90:40-58; ``allocate`` :63-133."""

    def __init__(self, G: Grid, gl: int = 0, c: str = "c", name: str = "unset"):
        self.G, self.gl, self.c, self.name = G, int(gl), c, name
        n = [G.Nx + 2 * gl, G.Ny + 2 * gl, G.Nz + 2 * gl]
        self.f = np.zeros(n, dtype=np.float64, order="F")                # scalar.f90:79-84
        self.bc_type = {}
        self.bc = {}
        if gl > 0:                                                       # scalar.f90:87-116
            self.bc["left"] = np.zeros((n[1], n[2]), order="F")
            self.bc["right"] = np.zeros((n[1], n[2]), order="F")
            self.bc["bottom"] = np.zeros((n[0], n[2]), order="F")
            self.bc["top"] = np.zeros((n[0], n[2]), order="F")
            if G.ndim == 3:
                self.bc["front"] = np.zeros((n[0], n[1]), order="F")
                self.bc["back"] = np.zeros((n[0], n[1]), order="F")
            for face in FACES[: 2 * G.ndim]:
                self.bc_type[face] = 0

    # interior view f(lo:hi, lo:hi, lo:hi)
    @property
    def I(self):
        g = self.gl
        G = self.G
        return self.f[g:g + G.Nx, g:g + G.Ny, g:g + G.Nz]

    def sh(self, di=0, dj=0, dk=0):
        """Interior-shaped view shifted by (di, dj, dk); needs gl >= |shift|."""
        g = self.gl
        G = self.G
        return self.f[g + di:g + di + G.Nx, g + dj:g + dj + G.Ny, g + dk:g + dk + G.Nz]

    def max_value(self) -> float:
        """scalar.f90:179-197 -- signed maximum over the interior (hazard H6)."""
        return float(self.I.max())

    def integral(self) -> float:
        """scalar.f90:201-219."""
        return float(self.I.sum()) * self.G.delta ** 3

    def copy(self):
        s = Scalar(self.G, self.gl, self.c, self.name)
        s.f[...] = self.f
        s.bc_type = dict(self.bc_type)
        s.bc = {k: v.copy() for k, v in self.bc.items()}
        return s

    def update_ghost_nodes(self) -> None:
        """scalar.f90:223-396, single rank (prow = pcol = 1; the halo call :251 is a no-op).

        Order x-left, x-right, y-bottom, y-top, z-front, z-back over the FULL transverse
        extent (ghosts included): it defines the edge/corner ghosts (hazard H3).
        """
        f, gl, c, G = self.f, self.gl, self.c, self.G
        if gl == 0:
            raise ValueError("scalar %s has no ghost nodes" % self.name)
        t, bc = self.bc_type, self.bc
        lo = gl            # array index of reference index lo(d) = 1
        hx, hy, hz = gl + G.Nx - 1, gl + G.Ny - 1, gl + G.Nz - 1

        # left :255-271
        if t["left"] == 0:
            for l in range(1, gl + 1):
                f[lo - l, :, :] = f[hx - l + 1, :, :]
        elif t["left"] == 1:
            if c in "cyz":
                f[lo - 1, :, :] = 2.0 * bc["left"] - f[lo, :, :]
            else:
                f[lo - 1, :, :] = bc["left"]
        elif t["left"] == 2:
            f[lo - 1, :, :] = f[lo, :, :]
        else:
            raise ValueError("wrong left bc type")
        # right :274-291
        if t["right"] == 0:
            for l in range(1, gl + 1):
                f[hx + l, :, :] = f[lo + l - 1, :, :]
        elif t["right"] == 1:
            if c in "cyz":
                f[hx + 1, :, :] = 2.0 * bc["right"] - f[hx, :, :]
            else:
                f[hx, :, :] = bc["right"]
                f[hx + 1, :, :] = bc["right"]
        elif t["right"] == 2:
            f[hx + 1, :, :] = f[hx, :, :]
        else:
            raise ValueError("wrong right bc type")
        # bottom :294-316 (only one layer is ever written, scalar.f90:298-300)
        if t["bottom"] == 0:
            f[:, lo - 1, :] = f[:, hy, :]
        elif t["bottom"] == 1:
            if c in "cxz":
                f[:, lo - 1, :] = 2.0 * bc["bottom"] - f[:, lo, :]
            else:
                f[:, lo - 1, :] = bc["bottom"]
        elif t["bottom"] == 2:
            f[:, lo - 1, :] = f[:, lo, :]
        elif t["bottom"] != -1:
            raise ValueError("wrong bottom bc type")
        # top :319-342
        if t["top"] == 0:
            f[:, hy + 1, :] = f[:, lo, :]
        elif t["top"] == 1:
            if c in "cxz":
                f[:, hy + 1, :] = 2.0 * bc["top"] - f[:, hy, :]
            else:
                f[:, hy, :] = bc["top"]
                f[:, hy + 1, :] = bc["top"]
        elif t["top"] == 2:
            f[:, hy + 1, :] = f[:, hy, :]
        elif t["top"] != -1:
            raise ValueError("wrong top bc type")
        if G.ndim == 3:
            # front :345-365
            if t["front"] == 0:
                f[:, :, lo - 1] = f[:, :, hz]
            elif t["front"] == 1:
                if c in "cxy":
                    f[:, :, lo - 1] = 2.0 * bc["front"] - f[:, :, lo]
                else:
                    f[:, :, lo - 1] = bc["front"]
            elif t["front"] == 2:
                f[:, :, lo - 1] = f[:, :, lo]
            elif t["front"] != -1:
                raise ValueError("wrong front bc type")
            # back :367-388
            if t["back"] == 0:
                f[:, :, hz + 1] = f[:, :, lo]
            elif t["back"] == 1:
                if c in "cxy":
                    f[:, :, hz + 1] = 2.0 * bc["back"] - f[:, :, hz]
                else:
                    f[:, :, hz] = bc["back"]
                    f[:, :, hz + 1] = bc["back"]
            elif t["back"] == 2:
                f[:, :, hz + 1] = f[:, :, hz]
            elif t["back"] != -1:
                raise ValueError("wrong back bc type")


class Vector:
    """``type vector`` -- src/vector.f90:15-63."""

    def __init__(self, G: Grid, gl: int = 0, name: str = "unset"):
        self.G, self.name = G, name
        self.x = Scalar(G, gl, "x", name + "_x")
        self.y = Scalar(G, gl, "y", name + "_y")
        self.z = Scalar(G, gl, "z", name + "_z") if G.ndim == 3 else None

    @property
    def comps(self):
        return [self.x, self.y] + ([self.z] if self.z is not None else [])

    def update_ghost_nodes(self) -> None:
        """vector.f90:82-109."""
        for s in self.comps:
            s.update_ghost_nodes()


# --------------------------------------------------------------------------------------
# fields.f90
# --------------------------------------------------------------------------------------
def gradient(s: Scalar, grad_s: Vector) -> None:
    """gradient_of_scalar -- src/fields.f90:31-64 (forward difference to the + face)."""
    idelta = 1.0 / s.G.delta
    c0 = s.sh()
    grad_s.x.I[...] = (s.sh(1, 0, 0) - c0) * idelta
    grad_s.y.I[...] = (s.sh(0, 1, 0) - c0) * idelta
    if s.G.ndim == 3:
        grad_s.z.I[...] = (s.sh(0, 0, 1) - c0) * idelta


def divergence(v: Vector, div_v: Scalar) -> None:
    """divergence_of_vector -- src/fields.f90:120-153."""
    idelta = 1.0 / v.G.delta
    d = (v.x.sh() - v.x.sh(-1, 0, 0)) * idelta + (v.y.sh() - v.y.sh(0, -1, 0)) * idelta
    if v.G.ndim == 3:
        d = d + (v.z.sh() - v.z.sh(0, 0, -1)) * idelta
    div_v.I[...] = d


def center_to_face(s: Scalar, v: Vector) -> None:
    """src/fields.f90:175-206."""
    c0 = s.sh()
    v.x.I[...] = 0.5 * (s.sh(1, 0, 0) + c0)
    v.y.I[...] = 0.5 * (s.sh(0, 1, 0) + c0)
    if s.G.ndim == 3:
        v.z.I[...] = 0.5 * (s.sh(0, 0, 1) + c0)


def face_to_center(sf: Scalar, sc: Scalar, face: str) -> None:
    """src/fields.f90:210-252: average of a face field with its low-side neighbour."""
    d = {"x": (-1, 0, 0), "y": (0, -1, 0), "z": (0, 0, -1)}[face]
    sc.I[...] = 0.5 * (sf.sh() + sf.sh(*d))


def curl(v: Vector, curl_v: Vector) -> None:
    """src/fields.f90:347-392, defined on the top right cell vertex; in 2-D the (z) result goes to curl_v%x."""
    idelta = 1.0 / v.G.delta
    if v.G.ndim == 3:
        curl_v.x.I[...] = (v.z.sh(0, 1, 0) - v.z.sh()) * idelta - (v.y.sh(0, 0, 1) - v.y.sh()) * idelta
        curl_v.y.I[...] = (v.x.sh(0, 0, 1) - v.x.sh()) * idelta - (v.z.sh(1, 0, 0) - v.z.sh()) * idelta
        curl_v.z.I[...] = (v.y.sh(1, 0, 0) - v.y.sh()) * idelta - (v.x.sh(0, 1, 0) - v.x.sh()) * idelta
    else:
        curl_v.x.I[...] = (v.y.sh(1, 0, 0) - v.y.sh()) * idelta - (v.x.sh(0, 1, 0) - v.x.sh()) * idelta


def _lap_terms(s: Scalar):
    c0 = s.sh()
    lx = s.sh(1, 0, 0) - 2.0 * c0 + s.sh(-1, 0, 0)
    ly = s.sh(0, 1, 0) - 2.0 * c0 + s.sh(0, -1, 0)
    lz = (s.sh(0, 0, 1) - 2.0 * c0 + s.sh(0, 0, -1)) if s.G.ndim == 3 else None
    return lx, ly, lz


def laplacian_scalar(s: Scalar, lap_s: Scalar) -> None:
    """laplacian_of_scalar -- src/fields.f90:256-294."""
    id2 = 1.0 / s.G.delta ** 2
    lx, ly, lz = _lap_terms(s)
    out = (lx + ly) * id2
    if lz is not None:
        out = out + lz * id2
    lap_s.I[...] = out


def laplacian_vector(v: Vector, lap_v: Vector) -> None:
    """laplacian_of_vector -- src/fields.f90:298-343 (note the different grouping for z)."""
    id2 = 1.0 / v.G.delta ** 2
    for comp, out in ((v.x, lap_v.x), (v.y, lap_v.y)):
        lx, ly, lz = _lap_terms(comp)
        r = (lx + ly) * id2
        if lz is not None:
            r = r + lz * id2
        out.I[...] = r
    if v.G.ndim == 3:
        lx, ly, lz = _lap_terms(v.z)
        lap_v.z.I[...] = (lx + ly + lz) * id2


# --------------------------------------------------------------------------------------
# poisson.f90
# --------------------------------------------------------------------------------------
def _mwn(n: int, delta: float, periodic: bool) -> np.ndarray:
    """Modified wavenumbers -- poisson.f90:627-629 (periodic), :794 (Neumann/DCT)."""
    m = np.arange(n, dtype=np.float64)      # (i - 1.0_dp), i = 1..n
    fac = 2.0 if periodic else 1.0
    return 2.0 * (np.cos(fac * PI * m / _f32(n)) - 1.0) / delta ** 2


def _tri_coeffs(n: int, delta: float, inflow_outflow: bool):
    """Tridiagonal coefficients -- poisson.f90:219-232 (2-D), :744-757 (3-D)."""
    a = np.full(n, 1.0 / delta ** 2)
    b = np.full(n, -2.0 / delta ** 2)
    c = np.full(n, 1.0 / delta ** 2)
    b[0] = b[0] + a[0]
    if inflow_outflow:
        b[n - 1] = b[n - 1] - c[n - 1]
    else:
        b[n - 1] = b[n - 1] + c[n - 1]
    a[0] = 0.0
    c[n - 1] = 0.0
    return a, b, c


def _seq_sum(x: np.ndarray) -> float:
    """Serial running sum in Fortran loop order (k outer, i inner) -- poisson.f90:398-405."""
    flat = np.ravel(x, order="F")
    if flat.size == 0:
        return 0.0
    return float(np.cumsum(flat)[-1])


class PoissonSolver:
    """``init_poisson_solver`` + the seven ``poisson_solver_*`` -- src/poisson.f90:57-113."""

    def __init__(self, phi: Scalar):
        G = phi.G
        self.G = G
        p = G.periodic_bc
        nx, ny, nz, d = G.Nx, G.Ny, G.Nz, G.delta
        b = G.boundary_conditions
        if G.ndim == 3:
            key = tuple(bool(q) for q in p)
            table = {(True, True, True): "ppp", (True, True, False): "ppn",
                     (False, True, False): "npn", (False, False, False): "nnn"}
            if key not in table:                                        # poisson.f90:91-95 (stop)
                raise RuntimeError("Unable to find the proper poisson solver with the selected "
                                   "boundary conditions")
            self.variant = table[key]
            io = b[4] == "Inflow" and b[5] == "Outflow"                 # :750-751
            self.mwn_x = _mwn(nx, d, self.variant[0] == "p")
            self.mwn_y = _mwn(ny, d, self.variant[1] == "p")
            if self.variant == "ppp":
                self.mwn_z = _mwn(nz, d, True)                           # :663-665
            else:
                self.a, self.b, self.c = _tri_coeffs(nz, d, io)
        else:
            key = (bool(p[0]), bool(p[1]))
            table = {(True, True): "pp", (True, False): "pn", (False, False): "nn"}
            if key not in table:                                        # poisson.f90:106-109
                raise RuntimeError("Unable to find the proper poisson solver with the selected "
                                   "boundary conditions")
            self.variant = table[key]
            io = b[2] == "Inflow" and b[3] == "Outflow"                 # :225-226
            self.mwn_x = _mwn(nx, d, self.variant[0] == "p")
            if self.variant == "pp":
                self.mwn_y = _mwn(ny, d, True)
            else:
                self.a, self.b, self.c = _tri_coeffs(ny, d, io)

    def solve(self, phi: Scalar) -> None:
        getattr(self, "_solve_" + self.variant)(phi)

    # ---- 2-D ------------------------------------------------------------------------
    def _solve_pp(self, phi: Scalar) -> None:
        """poisson_solver_pp -- poisson.f90:416-505."""
        G = self.G
        nx, ny, nz = G.shape
        w = _WORKERS
        cx = sfft.rfft(phi.I, axis=0, workers=w)                         # :441-445
        cy = sfft.fft(cx, axis=1, workers=w)                             # :451-455
        cy = cy / _f32(nx * ny)                                          # :458
        lam = self.mwn_x[: nx // 2 + 1, None, None] + self.mwn_y[None, :, None]
        zero = lam == 0.0                                                # :464-468
        cy = np.where(zero, 0.0, cy / np.where(zero, 1.0, lam))
        cxy = sfft.ifft(cy, axis=1, norm="forward", workers=w)           # :474-478
        phi.I[...] = sfft.irfft(cxy, n=nx, axis=0, norm="forward", workers=w)   # :484-488
        mean_phi = _seq_sum(phi.I)                                       # :491-498
        phi.f[...] = phi.f - mean_phi / _f32(nx * ny * nz)               # :503 (whole array)

    def _thomas_2d(self, rhs: np.ndarray, lam: np.ndarray) -> np.ndarray:
        """Thomas along axis 1 (y) -- poisson.f90:346-385 (pn) / :542-581 (nn)."""
        a, b, c = self.a, self.b, self.c
        ny = rhs.shape[1]
        lam = lam[:, None]
        c1 = np.empty(rhs.shape, dtype=np.float64)
        d1 = np.empty_like(rhs)
        c1[:, 0, :] = c[0] / (b[0] + lam)                                # :350
        d1[:, 0, :] = rhs[:, 0, :] / (b[0] + lam)                        # :351
        for j in range(1, ny - 1):                                       # :355-360
            c1[:, j, :] = c[j] / (b[j] - a[j] * c1[:, j - 1, :] + lam)
            d1[:, j, :] = (rhs[:, j, :] - a[j] * d1[:, j - 1, :]) / \
                          (b[j] + lam - a[j] * c1[:, j - 1, :])
        j = ny - 1                                                       # :362-371
        frac = b[j] + lam - a[j] * c1[:, j - 1, :]
        ok = frac != 0.0
        d1[:, j, :] = np.where(ok, (rhs[:, j, :] - a[j] * d1[:, j - 1, :]) / np.where(ok, frac, 1.0),
                               0.0)
        out = np.empty_like(rhs)
        out[:, ny - 1, :] = d1[:, ny - 1, :]                             # :374-378
        for j in range(ny - 2, -1, -1):                                  # :379-385
            out[:, j, :] = d1[:, j, :] - c1[:, j, :] * out[:, j + 1, :]
        return out

    def _solve_pn(self, phi: Scalar) -> None:
        """poisson_solver_pn -- poisson.f90:306-412."""
        G = self.G
        nx, ny, nz = G.shape
        w = _WORKERS
        cx = sfft.rfft(phi.I, axis=0, workers=w) / _f32(nx)              # :333-341
        sol = self._thomas_2d(cx, self.mwn_x[: nx // 2 + 1])
        phi.I[...] = sfft.irfft(sol, n=nx, axis=0, norm="forward", workers=w)   # :391-395
        mean_phi = _seq_sum(phi.I)                                       # :398-405
        phi.I[...] = phi.I - mean_phi / _f32(nx * ny * nz)               # :409-410 (interior only)

    def _solve_nn(self, phi: Scalar) -> None:
        """poisson_solver_nn -- poisson.f90:509-596."""
        G = self.G
        nx = G.Nx
        w = _WORKERS
        rx = sfft.dct(phi.I, type=2, axis=0, workers=w)                  # REDFT10 :533-537
        sol = self._thomas_2d(rx, self.mwn_x)
        phi.I[...] = sfft.dct(sol, type=3, axis=0, workers=w)            # REDFT01 :587-591
        phi.f[...] = phi.f / _f32(nx * 2)                                # :594 (whole array)

    # ---- 3-D ------------------------------------------------------------------------
    def _solve_ppp(self, phi: Scalar) -> None:
        """poisson_solver_ppp -- poisson.f90:941-1034 (no mean removal)."""
        G = self.G
        nx, ny, nz = G.shape
        w = _WORKERS
        cx = sfft.rfft(phi.I, axis=0, workers=w)                         # :965-969
        cy = sfft.fft(cx, axis=1, workers=w)                             # :975-979
        cz = sfft.fft(cy, axis=2, workers=w)                             # :985-989
        cz = cz / _f32(nx * ny * nz)                                     # :992
        lam = (self.mwn_x[: nx // 2 + 1, None, None] + self.mwn_y[None, :, None]) \
            + self.mwn_z[None, None, :]                                  # :998
        zero = lam == 0.0
        cz = np.where(zero, 0.0, cz / np.where(zero, 1.0, lam))          # :998-1002
        cy = sfft.ifft(cz, axis=2, norm="forward", workers=w)            # :1008-1012
        cx = sfft.ifft(cy, axis=1, norm="forward", workers=w)            # :1018-1022
        phi.I[...] = sfft.irfft(cx, n=nx, axis=0, norm="forward", workers=w)    # :1028-1032

    def _thomas_3d(self, rhs: np.ndarray, lx: np.ndarray, ly: np.ndarray) -> np.ndarray:
        """Thomas along axis 2 (z) -- poisson.f90:1092-1135 (ppn; same in npn/nnn).

        ``lx``/``ly`` are mwn_x(i) and mwn_y(j) shaped (n1, 1) and (1, n2); the pivot is
        evaluated as ``((b(k) + mwn_x(i)) + mwn_y(j)) - a(k)*c1`` like the reference."""
        a, b, c = self.a, self.b, self.c
        nz = rhs.shape[2]
        c1 = np.empty(rhs.shape, dtype=np.float64)
        d1 = np.empty_like(rhs)
        factor = 1.0 / ((b[0] + lx) + ly)                                # :1096
        c1[:, :, 0] = c[0] * factor
        d1[:, :, 0] = rhs[:, :, 0] * factor
        for k in range(1, nz - 1):                                       # :1102-1110
            factor = 1.0 / (((b[k] + lx) + ly) - a[k] * c1[:, :, k - 1])
            c1[:, :, k] = c[k] * factor
            d1[:, :, k] = (rhs[:, :, k] - a[k] * d1[:, :, k - 1]) * factor
        k = nz - 1                                                       # :1112-1121
        factor = ((b[k] + lx) + ly) - a[k] * c1[:, :, k - 1]
        ok = factor != 0.0                                               # exact-zero pivot (hazard H5)
        d1[:, :, k] = np.where(ok, (rhs[:, :, k] - a[k] * d1[:, :, k - 1]) / np.where(ok, factor, 1.0),
                               0.0)
        out = np.empty_like(rhs)
        out[:, :, nz - 1] = d1[:, :, nz - 1]                             # :1124-1128
        for k in range(nz - 2, -1, -1):                                  # :1129-1135
            out[:, :, k] = d1[:, :, k] - c1[:, :, k] * out[:, :, k + 1]
        return out

    def _solve_ppn(self, phi: Scalar) -> None:
        """poisson_solver_ppn -- poisson.f90:1038-1173."""
        G = self.G
        nx, ny, nz = G.shape
        w = _WORKERS
        cx = sfft.rfft(phi.I, axis=0, workers=w) / _f32(nx)              # :1066-1074
        cy = sfft.fft(cx, axis=1, workers=w) / _f32(ny)                  # :1080-1087
        sol = self._thomas_3d(cy, self.mwn_x[: nx // 2 + 1, None], self.mwn_y[None, :])
        cx = sfft.ifft(sol, axis=1, norm="forward", workers=w)           # :1141-1145
        phi.f[...] = 0.0                                                 # :1151
        phi.I[...] = sfft.irfft(cx, n=nx, axis=0, norm="forward", workers=w)    # :1152-1156
        mean_phi = _seq_sum(phi.I)                                       # :1159-1166
        phi.f[...] = phi.f - mean_phi / _f32(nx * ny * nz)               # :1171 (whole array)

    def _solve_npn(self, phi: Scalar) -> None:
        """poisson_solver_npn -- poisson.f90:1177-1312 (mean computed, NOT subtracted :1310)."""
        G = self.G
        nx, ny, nz = G.shape
        w = _WORKERS
        rx = sfft.dct(phi.I, type=2, axis=0, workers=w) / float(nx * 2)  # :1205-1213 real(nx*2,dp)
        cy = sfft.rfft(rx, axis=1, workers=w) / float(ny)                # :1219-1226
        sol = self._thomas_3d(cy, self.mwn_x[:, None], self.mwn_y[None, : ny // 2 + 1])
        rxy = sfft.irfft(sol, n=ny, axis=1, norm="forward", workers=w)   # :1280-1284
        phi.f[...] = 0.0                                                 # :1290
        phi.I[...] = sfft.dct(rxy, type=3, axis=0, workers=w)            # :1291-1295

    def _solve_nnn(self, phi: Scalar) -> None:
        """poisson_solver_nnn -- poisson.f90:1316-1451 (mean computed, NOT subtracted :1449)."""
        G = self.G
        nx, ny, nz = G.shape
        w = _WORKERS
        rx = sfft.dct(phi.I, type=2, axis=0, workers=w) / float(nx * 2)  # :1344-1352
        ry = sfft.dct(rx, type=2, axis=1, workers=w) / float(ny * 2)     # :1358-1365
        sol = self._thomas_3d(ry, self.mwn_x[:, None], self.mwn_y[None, :])
        rxy = sfft.dct(sol, type=3, axis=1, workers=w)                   # :1419-1423
        phi.f[...] = 0.0                                                 # :1429
        phi.I[...] = sfft.dct(rxy, type=3, axis=0, workers=w)            # :1430-1434


# --------------------------------------------------------------------------------------
# navier_stokes.f90
# --------------------------------------------------------------------------------------
_BC_TABLE = {
    # grid BC string -> (p/phi type, rho/mu type, v normal type, v tangential type)
    # navier_stokes.f90:780-1017 ; Inflow tangential is 2 on x faces, 1 on y/z faces.
    "Periodic": (0, 0, 0, 0),
    "Wall": (2, 2, 1, 1),
    "Inflow": (2, 2, 1, None),
    "Outflow": (1, 2, 2, 2),
}


class NavierStokes:
    """Module state + procedures of ``navier_stokes_mod`` (src/navier_stokes.f90) with the
    ``init_solver`` wiring of src/solver.f90:34-99 (single phase, no IBM/FSI/MF)."""

    def __init__(self, G: Grid, density: float = 1.0, viscosity: float = 1.0):
        self.G = G
        self.density, self.viscosity = float(density), float(viscosity)   # :18
        self.g = [0.0, 0.0, 0.0]                                          # :21
        self.maxdiv = 0.0
        self.maxCFL = 0.0
        self.CFL = 1.0                                                    # :24
        self.dt_visc = 0.0
        self.dt_conv = 0.0
        self.dt_o = 0.0
        self.constant_CFL = False                                         # :42
        self.constant_viscosity = True                                    # :45
        # allocate_navier_stokes_fields :752-775
        self.p = Scalar(G, 1, "c", "p")
        self.phi = Scalar(G, 1, "c", "phi")
        self.rho = Scalar(G, 1, "c", "rho")
        self.mu = Scalar(G, 1, "c", "mu")
        self.rhof = Vector(G, 0, "rhof")
        self.v = Vector(G, 1, "v")
        self.dv = Vector(G, 0, "dv")
        self.dv_o = Vector(G, 0, "dv_o")
        self.grad_p = Vector(G, 0, "grad_p")
        self.S = Vector(G, 0, "S")
        self.rho.f[...] = self.density                                    # :774
        self.mu.f[...] = self.viscosity                                   # :775
        self._wire_bc()
        self.poisson = PoissonSolver(self.phi)                            # solver.f90:61

    def _wire_bc(self) -> None:
        """navier_stokes.f90:780-1017 (Appendix B of SURVEY.md)."""
        G = self.G
        vel = self.v.comps
        normal_of_face = {"left": 0, "right": 0, "bottom": 1, "top": 1, "front": 2, "back": 2}
        for n, face in enumerate(FACES[: 2 * G.ndim]):
            s = G.boundary_conditions[n]
            if s not in _BC_TABLE:
                continue            # prints an error and leaves type 0 (:821)
            if face == "front" and s == "Outflow":
                continue            # not accepted (:961-987)
            if face == "back" and s == "Inflow":
                continue            # not accepted (:990-1016)
            tp, tr, tn, tt = _BC_TABLE[s]
            if s == "Inflow":
                tt = 2 if face in ("left", "right") else 1
            for sc in (self.p, self.phi):
                sc.bc_type[face] = tp
            for sc in (self.rho, self.mu):
                sc.bc_type[face] = tr
            for d, comp in enumerate(vel):
                comp.bc_type[face] = tn if d == normal_of_face[face] else tt

    # ---- timestep ---------------------------------------------------------------------
    def set_timestep(self, U: float) -> float:
        """navier_stokes.f90:623-666. Returns dt; sets dt_o = dt."""
        d = self.G.delta
        self.dt_conv = self.CFL * d / U
        self.dt_visc = 0.125 * d * d * self.density / self.viscosity
        if self.G.ndim == 3:
            self.dt_visc = (1.0 / 6.0) * d * d * self.density / self.viscosity
        dt = min(self.dt_conv, self.dt_visc)
        self.dt_o = dt
        return dt

    def _max_vel(self) -> float:
        vel = np.abs(self.v.x.I) + np.abs(self.v.y.I)
        if self.G.ndim == 3:
            vel = vel + np.abs(self.v.z.I)
        return float(max(0.0, vel.max()))

    def update_timestep(self, dt: float) -> float:
        """navier_stokes.f90:670-730."""
        self.dt_o = dt
        max_vel = self._max_vel()
        if max_vel > 0.0:
            self.dt_conv = self.CFL * self.G.delta / max_vel
        else:
            self.dt_conv = 1.0
        dt = min(self.dt_conv, self.dt_visc)
        if dt > 1.1 * self.dt_o:
            dt = 1.1 * self.dt_o
        return dt

    # ---- explicit terms ---------------------------------------------------------------
    def add_advection(self, RHS: Vector) -> None:
        """navier_stokes.f90:261-353."""
        G = self.G
        idelta = 1.0 / G.delta
        u, v = self.v.x, self.v.y
        U = u.sh
        V = v.sh
        three = G.ndim == 3
        # u terms :297-310
        uuip = 0.25 * (U(1, 0, 0) + U()) ** 2
        uuim = 0.25 * (U(-1, 0, 0) + U()) ** 2
        uvjp = (U(0, 1, 0) + U()) * (V(1, 0, 0) + V()) * 0.25
        uvjm = (U() + U(0, -1, 0)) * (V(1, -1, 0) + V(0, -1, 0)) * 0.25
        RHS.x.I[...] = RHS.x.I - (uuip - uuim) * idelta - (uvjp - uvjm) * idelta
        if three:
            W = self.v.z.sh
            uwkp = (U(0, 0, 1) + U()) * (W(1, 0, 0) + W()) * 0.25
            uwkm = (U() + U(0, 0, -1)) * (W(1, 0, -1) + W(0, 0, -1)) * 0.25
            RHS.x.I[...] = RHS.x.I - (uwkp - uwkm) * idelta
        # v terms :315-327
        vuip = (V(1, 0, 0) + V()) * (U(0, 1, 0) + U()) * 0.25
        vuim = (V() + V(-1, 0, 0)) * (U(-1, 1, 0) + U(-1, 0, 0)) * 0.25
        vvjp = 0.25 * (V(0, 1, 0) + V()) ** 2
        vvjm = 0.25 * (V(0, -1, 0) + V()) ** 2
        RHS.y.I[...] = RHS.y.I - (vuip - vuim) * idelta - (vvjp - vvjm) * idelta
        if three:
            vwkp = (V(0, 0, 1) + V()) * (W(0, 1, 0) + W()) * 0.25
            vwkm = (V() + V(0, 0, -1)) * (W(0, 1, -1) + W(0, 0, -1)) * 0.25
            RHS.y.I[...] = RHS.y.I - (vwkp - vwkm) * idelta
            # w terms :334-347
            wuip = (W() + W(1, 0, 0)) * (U() + U(0, 0, 1)) * 0.25
            wuim = (W() + W(-1, 0, 0)) * (U(-1, 0, 0) + U(-1, 0, 1)) * 0.25
            wvjp = (W() + W(0, 1, 0)) * (V() + V(0, 0, 1)) * 0.25
            wvjm = (W() + W(0, -1, 0)) * (V(0, -1, 0) + V(0, -1, 1)) * 0.25
            wwkp = (W() + W(0, 0, 1)) * (W() + W(0, 0, 1)) * 0.25
            wwkm = (W() + W(0, 0, -1)) * (W() + W(0, 0, -1)) * 0.25
            RHS.z.I[...] = RHS.z.I - (wuip - wuim) * idelta - (wvjp - wvjm) * idelta \
                - (wwkp - wwkm) * idelta

    def add_diffusion(self, RHS: Vector) -> None:
        """navier_stokes.f90:357-404, constant-viscosity branch."""
        if not self.constant_viscosity:
            raise NotImplementedError("variable viscosity is outside the hot-path scope")
        G = self.G
        lap_v = Vector(G, 1, "lap_v")                                    # :387
        laplacian_vector(self.v, lap_v)
        mu = self.mu.sh()
        for comp, lap, rf in zip(RHS.comps, lap_v.comps, self.rhof.comps):
            comp.I[...] = comp.I + mu * lap.I / rf.I                    # :394-397

    def compute_explicit_terms(self, RHS: Vector) -> None:
        """navier_stokes.f90:217-257."""
        for comp in RHS.comps:
            comp.f[...] = 0.0
        self.add_advection(RHS)
        self.add_diffusion(RHS)
        for comp, s, rf in zip(RHS.comps, self.S.comps, self.rhof.comps):
            comp.I[...] = comp.I + s.I / rf.I                            # :248-251

    def predicted_velocity_field(self, dt: float) -> None:
        """navier_stokes.f90:140-213."""
        A = 1.0 + 0.5 * dt / self.dt_o                                   # :157
        B = -0.5 * dt / self.dt_o                                        # :158
        center_to_face(self.rho, self.rhof)                              # :161
        self.compute_explicit_terms(self.dv)                             # :164
        gradient(self.p, self.grad_p)                                    # :165
        for d, (vc, gp, rf, dv, dvo) in enumerate(zip(self.v.comps, self.grad_p.comps,
                                                      self.rhof.comps, self.dv.comps,
                                                      self.dv_o.comps)):
            RHS = -gp.I / rf.I + A * dv.I + B * dvo.I + self.g[d]        # :169-172
            vc.I[...] = vc.I + dt * RHS                                  # :187-198
        for dv, dvo in zip(self.dv.comps, self.dv_o.comps):
            dvo.f[...] = dv.f                                            # :201-205
        self.v.update_ghost_nodes()                                      # :208

    def correct_velocity_field(self, dt: float) -> None:
        """navier_stokes.f90:505-546."""
        gradient(self.phi, self.grad_p)                                  # :521
        for vc, gp, rf in zip(self.v.comps, self.grad_p.comps, self.rhof.comps):
            vc.I[...] = vc.I - gp.I * dt / rf.I                          # :533-536
        self.v.update_ghost_nodes()                                      # :544

    def update_pressure(self) -> None:
        """navier_stokes.f90:550-566."""
        self.p.f[...] = self.p.f + self.phi.f                            # :561 (whole array)
        self.p.update_ghost_nodes()                                      # :564

    def checks(self, dt: float) -> None:
        """navier_stokes.f90:570-619."""
        div = Scalar(self.G, 0)
        divergence(self.v, div)
        self.maxdiv = div.max_value()                                    # signed max (H6)
        self.maxCFL = dt * self._max_vel() / self.G.delta                # :617

    def navier_stokes_solver(self, step: int, dt: float) -> float:
        """navier_stokes.f90:50-136. Returns the (possibly updated) dt."""
        if self.constant_CFL:
            dt = self.update_timestep(dt)                                # :78
        self.predicted_velocity_field(dt)                                # :105
        divergence(self.v, self.phi)                                     # :111
        self.phi.I[...] = self.phi.I * self.rho.sh() / dt                # :115-121
        self.poisson.solve(self.phi)                                     # :123
        self.phi.update_ghost_nodes()                                    # :124
        self.correct_velocity_field(dt)                                  # :127
        self.update_pressure()                                           # :130
        self.checks(dt)                                                  # :134
        return dt

    advance_solution = navier_stokes_solver                              # solver.f90:75

    def status_line(self, step: int, time: float, dt: float) -> str:
        """print_navier_stokes_solver_status -- navier_stokes.f90:734-748 (format :746)."""
        def e(x):
            return "%13s" % _fortran_e(x)
        return ("step: %7d time: %s dt: %s maxdiv: %s maxCFL:  %s"
                % (step, e(time), e(dt), e(self.maxdiv), e(self.maxCFL)))


def _fortran_e(x: float) -> str:
    """Fortran ``E13.6`` rendering (0.dddddd E+xx)."""
    if x == 0.0:
        return "0.000000E+00"
    exp = int(math.floor(math.log10(abs(x)))) + 1
    man = x / 10.0 ** exp
    if abs(round(man, 6)) >= 1.0:
        man /= 10.0
        exp += 1
    return "%s0.%06dE%+03d" % ("-" if man < 0 else "", int(round(abs(man) * 1e6)), exp)


# --------------------------------------------------------------------------------------
# raw I/O -- scalar%write / decomp_2d_write_one (scalar.f90:428-452): interior, x fastest
# --------------------------------------------------------------------------------------
def raw_bytes(s: Scalar) -> bytes:
    return np.ascontiguousarray(s.I.transpose(2, 1, 0)).tobytes()


# --------------------------------------------------------------------------------------
# initial conditions used by the reference drivers and the BASELINE configs
# --------------------------------------------------------------------------------------
def init_tgv2d(ns: NavierStokes) -> None:
    """test/small_test/navier_stokes/taylor_green_vortex/taylor_green_vortex.f90:86-113."""
    G = ns.G
    d = G.delta
    i = np.arange(1, G.Nx + 1, dtype=np.float64)[:, None, None]
    j = np.arange(1, G.Ny + 1, dtype=np.float64)[None, :, None]
    one = np.ones((1, 1, G.Nz))
    ns.v.x.I[...] = -np.cos(i * d) * np.sin((j - 0.5) * d) * one
    ns.v.y.I[...] = np.sin((i - 0.5) * d) * np.cos(j * d) * one
    ns.p.I[...] = -0.25 * (np.cos(2.0 * ((i - 0.5) * d)) + np.cos(2.0 * ((j - 0.5) * d))) * one
    ns.p.update_ghost_nodes()
    ns.v.update_ghost_nodes()


def init_tgv3d(ns: NavierStokes) -> None:
    """BASELINE config 2: 3-D periodic Taylor-Green vortex on [0, 2pi]^3."""
    G = ns.G
    d = G.delta
    i = np.arange(1, G.Nx + 1, dtype=np.float64)[:, None, None]
    j = np.arange(1, G.Ny + 1, dtype=np.float64)[None, :, None]
    k = np.arange(1, G.Nz + 1, dtype=np.float64)[None, None, :]
    ns.v.x.I[...] = np.sin(i * d) * np.cos((j - 0.5) * d) * np.cos((k - 0.5) * d)
    ns.v.y.I[...] = -np.cos((i - 0.5) * d) * np.sin(j * d) * np.cos((k - 0.5) * d)
    ns.v.z.I[...] = 0.0
    ns.p.I[...] = (1.0 / 16.0) * (np.cos(2.0 * (i - 0.5) * d) + np.cos(2.0 * (j - 0.5) * d)) \
        * (np.cos(2.0 * (k - 0.5) * d) + 2.0)
    ns.p.update_ghost_nodes()
    ns.v.update_ghost_nodes()


def init_abc(ns: NavierStokes, A=1.0, B=1.0, C=1.0) -> None:
    """test/large_test/ABC/ABC.f90:94-125."""
    G = ns.G
    x = G.x[1:G.Nx + 1][:, None, None]
    y = G.y[1:G.Ny + 1][None, :, None]
    z = G.z[1:G.Nz + 1][None, None, :]
    ns.v.x.I[...] = A * np.sin(z) + C * np.cos(y) + 0.0 * x
    ns.v.y.I[...] = B * np.sin(x) + A * np.cos(z) + 0.0 * y
    ns.v.z.I[...] = B * np.cos(x) + C * np.sin(y) + 0.0 * z
    ns.p.I[...] = -(B * C * np.cos(x) * np.sin(y) + A * B * np.sin(x) * np.cos(z)
                    + A * C * np.cos(y) * np.sin(z))
    ns.v.update_ghost_nodes()
    ns.p.update_ghost_nodes()


def init_channel(ns: NavierStokes, amp: float = 0.05) -> None:
    """BASELINE config 3: laminar Poiseuille profile u(z) between z walls plus a deterministic
    sinusoidal perturbation (no RNG), body force g(1) = 1 set by the caller."""
    G = ns.G
    d = G.delta
    Lz = G.Nz * d
    i = np.arange(1, G.Nx + 1, dtype=np.float64)[:, None, None]
    j = np.arange(1, G.Ny + 1, dtype=np.float64)[None, :, None]
    k = np.arange(1, G.Nz + 1, dtype=np.float64)[None, None, :]
    zc = (k - 0.5) * d
    kx = 2.0 * PI / (G.Nx * d)
    ky = 2.0 * PI / (G.Ny * d)
    kz = PI / Lz
    prof = 4.0 * zc * (Lz - zc) / (Lz * Lz)
    ns.v.x.I[...] = prof + amp * np.sin(kx * i * d) * np.cos(ky * (j - 0.5) * d) * np.sin(kz * zc)
    ns.v.y.I[...] = amp * np.cos(kx * (i - 0.5) * d) * np.sin(ky * j * d) * np.sin(kz * zc)
    ns.v.z.I[...] = amp * np.cos(kx * (i - 0.5) * d) * np.cos(ky * (j - 0.5) * d) \
        * np.sin(2.0 * kz * (k * d))
    ns.v.z.I[:, :, G.Nz - 1] = 0.0
    ns.p.I[...] = 0.0
    ns.p.update_ghost_nodes()
    ns.v.update_ghost_nodes()
