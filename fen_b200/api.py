"""Host-side mirror of FEN's solver API over the C ABI of libfen_gpu.so.

Names, argument meaning and error behaviour follow the reference's Fortran modules so that the
parity tests read like the reference's own test programs:

    grid            <- type grid / grid%setup                  (src/grid.f90:22-62, 67-200)
    scalar, vector  <- type scalar / type vector               (src/scalar.f90:40-58, src/vector.f90:15-26)
    gradient, divergence, laplacian, center_to_face            (src/fields.f90)
    Solver          <- solver_mod + navier_stokes_mod + poisson_mod
                       (src/solver.f90:34-99, src/navier_stokes.f90, src/poisson.f90:51-52)

The host arrays ``scalar.f`` have the reference's layout (Fortran order, ``gl`` ghost layers) and
stay owned by the host; ``push()`` / ``pull()`` are the explicit transfer points that replace the
reference's direct pokes into module-global arrays.  All arithmetic runs in
the CUDA library; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import sys

import numpy as np

from . import _lib
from ._lib import FenError, GridDesc, MfParams, NsParams, check

_BC_CODE = {"Periodic": 0, "Wall": 1, "Inflow": 2, "Outflow": 3}
_FACES = ("left", "right", "bottom", "top", "front", "back")
_LOC = {"c": 0, "x": 1, "y": 2, "z": 3}

# enum fen_field
P, PHI, RHO, MU, VX, VY, VZ, DVX, DVY, DVZ, DVOX, DVOY, DVOZ, GPX, GPY, GPZ, SX, SY, SZ = range(19)
VOF, H, D, CURV, NORMX, NORMY, NORMZ, LX, LY, LZ, PHAT, PO, VOF1 = range(19, 32)


def _f32(n):
    return float(np.float32(n))


def _fortran_e(x, w=16, d=8):
    """Fortran ``Ew.d`` edit descriptor: 0.dddddddd E+ee, right-justified in w columns (gfortran's output form)."""
    x = float(x)
    if x == 0.0:
        body = "0." + "0" * d + "E+00"
    else:
        ex = int(math.floor(math.log10(abs(x)))) + 1
        man = abs(x) / 10.0 ** ex
        digits = int(round(man * 10 ** d))
        if digits >= 10 ** d:                     # 0.99999999.. rounded up to 1.0
            digits //= 10
            ex += 1
        body = ("-" if x < 0 else "") + "0.%0*d" % (d, digits) + "E%+03d" % ex
    return body.rjust(w)


def grid_json(Nx, Ny, Nz, origin, Lx, Ly, Lz):
    """The text grid%print_json writes (grid.f90:246-257), line for line."""
    e = _fortran_e
    return "\n".join([
        "{",
        "    " + '"Grid": {',
        "        " + '"Nx": ' + " " + "%7d" % Nx + ",",
        "        " + '"Ny": ' + " " + "%7d" % Ny + ",",
        "        " + '"Nz": ' + " " + "%7d" % Nz + ",",
        "        " + '"origin": [' + " " + e(origin[0]) + "," + e(origin[1]) + "," + e(origin[2]) + "],",
        "        " + '"Lx": ' + " " + e(Lx) + ",",
        "        " + '"Ly": ' + " " + e(Ly) + ",",
        "        " + '"Lz": ' + " " + e(Lz),
        "    " + "  }",
        "}",
    ]) + "\n"


class grid:
    """``type grid``; ``setup`` mirrors grid%setup(Nx,Ny,Nz,Lx,Ly,Lz,x0,prow,pcol,bc)."""

    def __init__(self):
        self.ctx = None

    def setup(self, Nx, Ny, Nz, Lx, Ly, Lz, x0=(0.0, 0.0, 0.0), prow=1, pcol=1, bc=None,
              ndim=None, rank=0, device=-1):
        lib = _lib.load()
        self.ndim = ndim if ndim is not None else (2 if Nz == 1 else 3)
        nb = 2 * self.ndim
        self.boundary_conditions = ["Periodic"] * nb if bc is None else list(bc)
        if len(self.boundary_conditions) != nb:
            raise ValueError("bc must have %d entries" % nb)
        if prow != 1:
            raise FenError(3, "the GPU path uses slabs: prow must be 1 (pcol = number of GPUs)")
        b = self.boundary_conditions
        self.periodic_bc = [b[0] == "Periodic" and b[1] == "Periodic",
                            b[2] == "Periodic" and b[3] == "Periodic",
                            (b[4] == "Periodic" and b[5] == "Periodic") if self.ndim == 3 else True]
        self.Nx, self.Ny, self.Nz = int(Nx), int(Ny), int(Nz)
        self.Lx, self.Ly, self.Lz = float(Lx), float(Ly), float(Lz)
        self.origin = tuple(float(v) for v in x0)
        self.delta = self.Lx / _f32(Nx)                                   # grid.f90:140
        if self.Lx / _f32(Nx) != self.Ly / _f32(Ny) and rank == 0:        # grid.f90:143-147
            print("The grid spacing must be equal in all directions", file=sys.stderr)
        self.x = self.origin[0] + (np.arange(0, Nx + 2) - 0.5) * self.delta
        self.y = self.origin[1] + (np.arange(0, Ny + 2) - 0.5) * self.delta
        self.z = self.origin[2] + (np.arange(0, Nz + 2) - 0.5) * self.delta
        self.prow, self.pcol, self.rank, self.nranks = 1, int(pcol), int(rank), int(pcol)
        d = GridDesc()
        d.nx, d.ny, d.nz, d.ndim, d.delta = self.Nx, self.Ny, self.Nz, self.ndim, self.delta
        for i in range(6):
            d.bc[i] = _BC_CODE.get(b[i], -1) if i < nb else 0
        d.rank, d.nranks, d.device = self.rank, self.nranks, device
        ctx = C.c_void_p()
        check(lib.fen_gpu_create(C.byref(d), C.byref(ctx)))
        self.ctx = ctx
        self.lib = lib
        lo = (C.c_int * 3)()
        hi = (C.c_int * 3)()
        check(lib.fen_gpu_local_bounds(ctx, lo, hi))
        self.lo, self.hi = tuple(lo), tuple(hi)
        self.nloc = tuple(h - l + 1 for l, h in zip(self.lo, self.hi))
        return self

    name = "grid"                                                         # grid.f90:57
    # global.f90:23-28: offset of a location from the cell centre in units of delta; 0 centre, 1..3 x / y / z face, 4 corner
    _STAGGER = ((0.0, 0.0, 0.0), (0.5, 0.0, 0.0), (0.0, 0.5, 0.0), (0.0, 0.0, 0.5), (0.5, 0.5, 0.5))

    def closest_grid_node(self, xl, ind):
        """grid%closest_grid_node (grid.f90:204-229): 1-based indices of the grid point of location ``ind`` closest to the
        point ``xl`` (first minimum, as Fortran's minloc)."""
        st = self._STAGGER[ind]
        ie = [int(np.argmin(np.abs(self.x[1:self.Nx + 1] + st[0] * self.delta - xl[0]))) + 1,
              int(np.argmin(np.abs(self.y[1:self.Ny + 1] + st[1] * self.delta - xl[1]))) + 1, 1]
        if self.ndim == 3:
            ie[2] = int(np.argmin(np.abs(self.z[1:self.Nz + 1] + st[2] * self.delta - xl[2]))) + 1
        return ie

    def print_json(self, dirname="."):
        """grid%print_json (grid.f90:233-264): ``<name>.json`` with Nx, Ny, Nz, origin, Lx, Ly, Lz in the reference's
        own formats (I7, E16.8) -- the file every postpro.py of the reference opens first.  Rank 0 writes.  The
        reference calls it at the end of ``setup`` (:198); here it is an explicit call so that creating a grid has no
        side effect in the working directory."""
        if getattr(self, "rank", 0) != 0:
            return None
        path = os.path.join(dirname, self.name + ".json")
        with open(path, "w") as fh:
            fh.write(grid_json(self.Nx, self.Ny, self.Nz, self.origin, self.Lx, self.Ly, self.Lz))
        return path

    def connect(self, all_gather):
        """Wire the z-slab neighbours: ``all_gather(bytes) -> list[bytes]`` over all ranks."""
        n = self.lib.fen_gpu_comm_handle_bytes()
        buf = C.create_string_buffer(n)
        check(self.lib.fen_gpu_comm_export(self.ctx, buf))
        from .decomp import gather_handles
        allh = gather_handles(all_gather, buf.raw, self.nranks)
        check(self.lib.fen_gpu_comm_connect(self.ctx, C.c_char_p(allh)))

    def synchronize(self):
        check(self.lib.fen_gpu_synchronize(self.ctx))

    def pull_wait(self):
        check(self.lib.fen_gpu_pull_wait(self.ctx))

    def destroy(self):
        if self.ctx is not None:
            self.lib.fen_gpu_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class scalar:
    """``type scalar``: host array ``f`` (reference layout) + a device twin addressed by ``id``."""

    def __init__(self, G: grid = None, l: int = 0, c: str = "c", field_id=None):
        self.G = None
        self.id = None
        if G is not None:
            self.allocate(G, l, c, field_id)

    def allocate(self, G: grid, l: int = 0, c: str = "c", field_id=None):
        self.G, self.gl, self.c = G, int(l), c
        n = [G.nloc[0] + 2 * l, G.nloc[1] + 2 * l, G.nloc[2] + 2 * l]
        self.f = np.zeros(n, dtype=np.float64, order="F")
        self._owned = field_id is None
        if field_id is None:
            fid = C.c_int()
            check(G.lib.fen_gpu_scalar_allocate(G.ctx, self.gl, _LOC[c], C.byref(fid)))
            self.id = fid.value
        else:
            self.id = int(field_id)
        return self

    # interior view f(lo:hi, lo:hi, lo:hi)
    @property
    def I(self):
        g = self.gl
        n = self.G.nloc
        return self.f[g:g + n[0], g:g + n[1], g:g + n[2]]

    def push(self):
        check(self.G.lib.fen_gpu_push(self.G.ctx, self.id, self.f.ctypes.data_as(C.c_void_p), self.gl))
        return self

    def pull(self):
        check(self.G.lib.fen_gpu_pull(self.G.ctx, self.id, self.f.ctypes.data_as(C.c_void_p), self.gl))
        return self

    def pull_async(self):
        """pull() that does not wait: ``f`` is valid after ``G.pull_wait()`` / ``G.synchronize()``."""
        check(self.G.lib.fen_gpu_pull_async(self.G.ctx, self.id, self.f.ctypes.data_as(C.c_void_p), self.gl))
        return self

    def set_from_function(self, fp, args=None):
        """scalar%set_from_function (scalar.f90:137-164): ``f(i,j,k) = fp([x(i), y(j)(, z(k))], args)`` over the interior
        at the grid's CELL-CENTRE coordinates (whatever the location tag, as in the reference), then the ghost update
        when the scalar has ghost nodes.  The host array and its device twin both hold the result."""
        G, g = self.G, self.gl
        lo, n = G.lo, G.nloc
        for k in range(n[2]):
            for j in range(n[1]):
                for i in range(n[0]):
                    pt = [G.x[lo[0] + i], G.y[lo[1] + j]] + ([G.z[lo[2] + k]] if G.ndim == 3 else [])
                    self.f[g + i, g + j, g + k] = fp(pt, args)
        self.push()
        if g > 0:
            self.update_ghost_nodes()
            self.pull()
        return self

    def setToValue(self, val):
        self.f[...] = val
        check(self.G.lib.fen_gpu_set_to_value(self.G.ctx, self.id, float(val)))

    def set_bc_type(self, face, t):
        check(self.G.lib.fen_gpu_set_bc_type(self.G.ctx, self.id, _FACES.index(face), int(t)))

    def get_bc_type(self, face):
        t = C.c_int()
        check(self.G.lib.fen_gpu_get_bc_type(self.G.ctx, self.id, _FACES.index(face), C.byref(t)))
        return t.value

    def set_bc(self, face, value):
        """bc%<face> = value: a scalar (broadcast) or a plane including ghosts (Fortran order)."""
        if np.isscalar(value):
            v = C.c_double(float(value))
            check(self.G.lib.fen_gpu_set_bc_plane(self.G.ctx, self.id, _FACES.index(face), C.byref(v), 1))
        else:
            a = np.asfortranarray(value, dtype=np.float64)
            check(self.G.lib.fen_gpu_set_bc_plane(self.G.ctx, self.id, _FACES.index(face),
                                                  a.ctypes.data_as(C.c_void_p), 0))

    def update_ghost_nodes(self):
        check(self.G.lib.fen_gpu_update_ghost_nodes(self.G.ctx, self.id, 1))

    def max_value(self):
        out = C.c_double()
        check(self.G.lib.fen_gpu_max_value(self.G.ctx, self.id, C.byref(out)))
        return out.value

    def integral(self):
        out = C.c_double()
        check(self.G.lib.fen_gpu_integral(self.G.ctx, self.id, C.byref(out)))
        return out.value

    def write(self, filename):
        """scalar%write (scalar.f90:428): the global interior array, x fastest, real(dp), no header; every rank
        writes its own z-slab range of the file."""
        check(self.G.lib.fen_gpu_scalar_write(self.G.ctx, self.id, str(filename).encode()))

    def read(self, filename):
        """scalar%read (scalar.f90:400) into the device field (interior; ghosts are not touched)."""
        check(self.G.lib.fen_gpu_scalar_read(self.G.ctx, self.id, str(filename).encode()))

    def destroy(self):
        if self.id is not None and self._owned and self.G is not None and self.G.ctx is not None:
            self.G.lib.fen_gpu_scalar_destroy(self.G.ctx, self.id)
        self.id = None


class vector:
    """``type vector``: components x, y, (z) with consecutive device ids."""

    def __init__(self, G: grid = None, l: int = 0, first_id=None):
        if G is not None:
            self.allocate(G, l, first_id)

    def allocate(self, G: grid, l: int = 0, first_id=None):
        self.G = G
        ids = [None] * 3 if first_id is None else [first_id, first_id + 1, first_id + 2]
        self.x = scalar(G, l, "x", ids[0])
        self.y = scalar(G, l, "y", ids[1])
        self.z = scalar(G, l, "z", ids[2]) if G.ndim == 3 else None
        if first_id is None:
            want = list(range(self.x.id, self.x.id + len(self.comps)))
            if [s.id for s in self.comps] != want:
                raise FenError(1, "vector components must get consecutive ids")
        return self

    @property
    def comps(self):
        return [self.x, self.y] + ([self.z] if self.z is not None else [])

    def push(self):
        for s in self.comps:
            s.push()
        return self

    def pull(self):
        for s in self.comps:
            s.pull()
        return self

    def update_ghost_nodes(self):
        check(self.G.lib.fen_gpu_update_ghost_nodes(self.G.ctx, self.x.id, len(self.comps)))

    def destroy(self):
        for s in self.comps:
            s.destroy()


def gradient(s: scalar, grad_s: vector):
    check(s.G.lib.fen_gpu_gradient(s.G.ctx, s.id, grad_s.x.id))


def divergence(v: vector, div_v: scalar):
    check(v.G.lib.fen_gpu_divergence(v.G.ctx, v.x.id, div_v.id))


def laplacian(v, lap_v):
    """fields_mod's generic ``laplacian``: of a vector (fields.f90:298) or of a scalar (:256)."""
    if isinstance(v, scalar):
        check(v.G.lib.fen_gpu_laplacian_scalar(v.G.ctx, v.id, lap_v.id))
    else:
        check(v.G.lib.fen_gpu_laplacian(v.G.ctx, v.x.id, lap_v.x.id))


def face_to_center(sf: scalar, sc: scalar, face: str):
    """fields.f90:210-252; ``face`` is 'x', 'y' or 'z' as in the reference."""
    check(sf.G.lib.fen_gpu_face_to_center(sf.G.ctx, sf.id, sc.id, "xyz".index(face)))


def curl(v: vector, curl_v: vector):
    """fields.f90:347-392 (2-D: the result is in ``curl_v.x``)."""
    check(v.G.lib.fen_gpu_curl(v.G.ctx, v.x.id, curl_v.x.id))


def center_to_face(s: scalar, v: vector):
    check(s.G.lib.fen_gpu_center_to_face(s.G.ctx, s.id, v.x.id))


class PoissonSolver:
    """init_Poisson_Solver / solve_Poisson / destroy_Poisson_solver (poisson.f90:51-52)."""

    def __init__(self, phi: scalar):
        self.G = phi.G
        check(self.G.lib.fen_gpu_init_poisson_solver(self.G.ctx))

    @property
    def variant(self):
        return self.G.lib.fen_gpu_poisson_variant(self.G.ctx).decode()

    def solve(self, phi: scalar):
        check(self.G.lib.fen_gpu_solve_poisson(self.G.ctx, phi.id))

    def destroy(self):
        check(self.G.lib.fen_gpu_destroy_poisson_solver(self.G.ctx))


class Solver:
    """solver_mod + navier_stokes_mod: ``init_solver``, ``set_timestep``, ``advance_solution``.

    Module scalars (density, viscosity, g, CFL, constant_CFL, dt_o ...) are attributes; as in the
    reference, ``density``/``viscosity`` are copied into rho/mu at ``init_solver`` time only
    (navier_stokes.f90:774-775)."""

    _PARAMS = ("density", "viscosity", "CFL", "dt_o", "dt_visc", "dt_conv", "constant_CFL")

    def __init__(self, G: grid, density=1.0, viscosity=1.0):
        object.__setattr__(self, "G", G)
        object.__setattr__(self, "_ready", False)
        p = self._get()
        p.density, p.viscosity = density, viscosity
        self._set(p)

    def _get(self):
        p = NsParams()
        check(self.G.lib.fen_gpu_get_params(self.G.ctx, C.byref(p)))
        return p

    def _set(self, p):
        check(self.G.lib.fen_gpu_set_params(self.G.ctx, C.byref(p)))

    def __getattr__(self, name):
        if name in Solver._PARAMS:
            v = getattr(self._get(), name)
            return bool(v) if name == "constant_CFL" else v
        if name == "g":
            return list(self._get().g)
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in Solver._PARAMS:
            p = self._get()
            setattr(p, name, int(value) if name == "constant_CFL" else float(value))
            self._set(p)
        elif name == "g":
            p = self._get()
            for i in range(3):
                p.g[i] = float(value[i])
            self._set(p)
        else:
            object.__setattr__(self, name, value)

    def init_solver(self):
        G = self.G
        check(G.lib.fen_gpu_init_solver(G.ctx))
        self.p = scalar(G, 1, "c", P)
        self.phi = scalar(G, 1, "c", PHI)
        self.rho = scalar(G, 1, "c", RHO)
        self.mu = scalar(G, 1, "c", MU)
        self.v = vector(G, 1, VX)
        self.dv_o = vector(G, 0, DVOX)
        self.S = vector(G, 0, SX)
        self._ready = True
        return self

    @property
    def poisson_variant(self):
        return self.G.lib.fen_gpu_poisson_variant(self.G.ctx).decode()

    def set_timestep(self, U):
        dt = C.c_double()
        check(self.G.lib.fen_gpu_set_timestep(self.G.ctx, float(U), C.byref(dt)))
        return dt.value

    def navier_stokes_solver(self, step, dt):
        d = C.c_double(dt)
        check(self.G.lib.fen_gpu_navier_stokes_solver(self.G.ctx, int(step), C.byref(d)))
        return d.value

    advance_solution = navier_stokes_solver          # solver.f90:75

    def status(self):
        a, b = C.c_double(), C.c_double()
        check(self.G.lib.fen_gpu_get_status(self.G.ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def maxdiv(self):
        return self.status()[0]

    @property
    def maxCFL(self):
        return self.status()[1]

    def print_solver_status(self, step, time, dt):
        buf = C.create_string_buffer(256)
        check(self.G.lib.fen_gpu_status_line(self.G.ctx, int(step), float(time), float(dt), buf, 256))
        return buf.value.decode()

    def add_advection(self, rhs: vector):
        check(self.G.lib.fen_gpu_add_advection(self.G.ctx, rhs.x.id))

    def compute_explicit_terms(self, rhs: vector):
        check(self.G.lib.fen_gpu_compute_explicit_terms(self.G.ctx, rhs.x.id))

    def destroy_solver(self):
        check(self.G.lib.fen_gpu_destroy_solver(self.G.ctx))
        self._ready = False

    # ---- solver_mod output / restart (solver.f90:103-329) -----------------------------------
    def save_state(self, step, dirname="data"):
        """save_state(step): ``<dirname>/state_<step7>.raw`` = p, v_x, v_y, dv_o_x, dv_o_y, [v_z, dv_o_z]."""
        path = "%s/state_%07d.raw" % (dirname, int(step))
        check(self.G.lib.fen_gpu_save_state(self.G.ctx, path.encode()))
        return path

    def load_state(self, step, dirname="data"):
        path = "%s/state_%07d.raw" % (dirname, int(step))
        check(self.G.lib.fen_gpu_load_state(self.G.ctx, path.encode()))
        return path

    def save_fields(self, step, dirname="data"):
        """save_fields(step): cell-centred vx_, vy_, [vz_] and p_ raw files."""
        check(self.G.lib.fen_gpu_save_fields(self.G.ctx, int(step), str(dirname).encode()))

    def set_forcing_hook(self, fn):
        """Host callback ``fn(step, dt)`` between the predictor and the Poisson solve -- where the reference calls
        apply_ibm_forcing(v, dt) (navier_stokes.f90:106-108).  ``None`` removes it."""
        if fn is None:
            self._hook = None
            check(self.G.lib.fen_gpu_set_forcing_hook(self.G.ctx, _lib.FORCING_FN(0), None))
            return

        def tramp(_user, step, dt):
            try:
                fn(step, dt)
                return 0
            except Exception as e:          # noqa: BLE001 -- reported through the C return code
                print("forcing hook failed: %r" % (e,), file=sys.stderr)
                return 1
        self._hook = _lib.FORCING_FN(tramp)      # keep the trampoline alive
        check(self.G.lib.fen_gpu_set_forcing_hook(self.G.ctx, self._hook, None))

    # ---- measurement -----------------------------------------------------------------------
    def profile(self, on=True):
        check(self.G.lib.fen_gpu_profile_enable(self.G.ctx, 1 if on else 0))

    def profile_read(self):
        n = 64
        names = ((C.c_char * 32) * n)()
        ms = (C.c_double * n)()
        cnt = (C.c_int * n)()
        nout = C.c_int()
        check(self.G.lib.fen_gpu_profile_read(self.G.ctx, n, names, ms, cnt, C.byref(nout)))
        return {names[i].value.decode(): (ms[i], cnt[i]) for i in range(nout.value)}

    def launch_count(self):
        return int(self.G.lib.fen_gpu_launch_count(self.G.ctx))


# ---------------------------------------------------------------------------------------------------
# two-phase build (-DMF): volume_of_fluid_mod, multiphase_mod and the MF branches of navier_stokes_mod
# ---------------------------------------------------------------------------------------------------
class _MfState:
    """Module variables of multiphase_mod / volume_of_fluid_mod as attributes (rho_0, rho_1, mu_0, mu_1, sigma,
    beta, cut, quadratic, x_first, dt_surf, rhomin, irhomin)."""

    _MF = tuple(n for n, _ in MfParams._fields_)

    def _mf_get(self):
        p = MfParams()
        check(self.G.lib.fen_gpu_mf_get_params(self.G.ctx, C.byref(p)))
        return p

    def _mf_attr(self, name):
        v = getattr(self._mf_get(), name)
        return bool(v) if name in ("quadratic", "x_first") else v

    def _mf_setattr(self, name, value):
        p = self._mf_get()
        setattr(p, name, int(bool(value)) if name in ("quadratic", "x_first") else float(value))
        check(self.G.lib.fen_gpu_mf_set_params(self.G.ctx, C.byref(p)))

    def _mf_fields(self):
        G = self.G
        self.vof = scalar(G, 1, "c", VOF)
        self.h = scalar(G, 1, "c", H)
        self.d = scalar(G, 1, "c", D)
        self.curv = scalar(G, 1, "c", CURV)
        self.norm = vector(G, 1, NORMX)
        self.l = vector(G, 1, LX)

    def _tramp(self, distance):
        def tramp(_user, x, y):
            return float(distance(x, y))
        self._dist = _lib.DISTANCE_FN(tramp)             # keep the trampoline alive
        return self._dist

    # volume_of_fluid_mod procedures
    def get_vof_from_distance(self, distance):
        """``distance => f; call get_vof_from_distance`` (volume_of_fluid.f90:676); f(x, y) -> signed distance."""
        G = self.G
        check(G.lib.fen_gpu_get_vof_from_distance(G.ctx, self._tramp(distance), None, G.origin[0], G.origin[1]))

    def get_h_from_vof(self):
        check(self.G.lib.fen_gpu_get_h_from_vof(self.G.ctx))

    def advect_vof(self, v: vector, dt):
        check(self.G.lib.fen_gpu_advect_vof(self.G.ctx, v.x.id, float(dt)))

    def check_vof_integral(self):
        a, b = C.c_double(), C.c_double()
        check(self.G.lib.fen_gpu_check_vof_integral(self.G.ctx, C.byref(a), C.byref(b)))
        return a.value, b.value


class VoF(_MfState):
    """volume_of_fluid_mod on its own, as the reference's VoF-only tests use it (allocate_vof_fields +
    get_vof_from_distance + advect_vof with a prescribed velocity; test/small_test/volume_of_fluid/*)."""

    def __init__(self, G: grid):
        object.__setattr__(self, "G", G)
        check(G.lib.fen_gpu_allocate_vof_fields(G.ctx))
        self._mf_fields()

    def __getattr__(self, name):
        if name in _MfState._MF:
            return self._mf_attr(name)
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in _MfState._MF:
            self._mf_setattr(name, value)
        else:
            object.__setattr__(self, name, value)

    def destroy_vof(self):
        check(self.G.lib.fen_gpu_destroy_vof(self.G.ctx))


class MultiphaseSolver(Solver, _MfState):
    """solver_mod + navier_stokes_mod compiled with -DMF.  Set rho_0, rho_1, mu_0, mu_1, sigma, beta (module
    variables) before ``init_solver(distance)``, exactly as the reference's drivers do
    (test/small_test/multiphase/viscous_decay/viscous_decay.f90:40-53)."""

    def __init__(self, G: grid):
        Solver.__init__(self, G, 1.0, 1.0)

    def __getattr__(self, name):
        if name in _MfState._MF:
            return self._mf_attr(name)
        return Solver.__getattr__(self, name)

    def __setattr__(self, name, value):
        if name in _MfState._MF:
            self._mf_setattr(name, value)
        else:
            Solver.__setattr__(self, name, value)

    def init_solver(self, distance=None):
        """init_solver (solver.f90:34-99, MF).  ``distance`` is the reference's ``distance`` procedure pointer; with
        None the vof field is left zero (the reference prints an error): push ``vof`` and call
        ``update_material_properties`` instead."""
        G = self.G
        fn = self._tramp(distance) if distance is not None else _lib.DISTANCE_FN(0)
        check(G.lib.fen_gpu_init_solver_mf(G.ctx, fn, None, G.origin[0], G.origin[1]))
        self.p = scalar(G, 1, "c", P)
        self.phi = scalar(G, 1, "c", PHI)
        self.rho = scalar(G, 1, "c", RHO)
        self.mu = scalar(G, 1, "c", MU)
        self.v = vector(G, 1, VX)
        self.dv_o = vector(G, 0, DVOX)
        self.S = vector(G, 0, SX)
        self.p_hat = scalar(G, 1, "c", PHAT)
        self.p_o = scalar(G, 1, "c", PO)
        self._mf_fields()
        self._ready = True
        return self

    def update_material_properties(self):
        check(self.G.lib.fen_gpu_update_material_properties(self.G.ctx))
