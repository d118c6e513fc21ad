"""A CPU stand-in for the fen_b200 API backed by the numpy oracle (test infrastructure).

The GPU tests written after round 1's GPU budget was spent (tests/test_gpu_zy_*.py, tests/test_gpu_zz_*.py) have not met
a device yet.  tests/test_first_run_logic.py runs their bodies against this stand-in -- the same names and call
sequences, host arrays shared with "the device", push / pull as no-ops -- so that a typo, a wrong shape or a wrong loop
bound in a test shows up here and not as a wasted GPU run.  Both sides of every comparison are then the oracle, so the
comparisons themselves prove nothing: only the test logic is exercised."""
import numpy as np
from oracle import fen_oracle as fo, fen_oracle_mf as mf

class FenError(RuntimeError): pass

class _G(fo.Grid):
    def destroy(self): pass
    def pull_wait(self): pass
    def synchronize(self): pass

class grid:
    def setup(self, Nx, Ny, Nz, Lx, Ly, Lz, x0=(0.0,0.0,0.0), prow=1, pcol=1, bc=None, ndim=None, rank=0, device=-1):
        kw = {} if ndim is None else {"ndim": ndim}
        return _G(Nx, Ny, Nz, Lx, Ly, Lz, bc=bc, **kw)

class _S:
    """wraps an oracle Scalar"""
    def __init__(self, s): self.s = s; self.G = s.G; self.gl = s.gl
    @property
    def f(self): return self.s.f
    @f.setter
    def f(self, v): self.s.f = v
    @property
    def I(self): return self.s.I
    def push(self): return self
    def pull(self): return self
    def pull_async(self): return self
    def update_ghost_nodes(self): self.s.update_ghost_nodes()
    def set_bc(self, face, value): self.s.bc[face][...] = value
    def set_bc_type(self, face, t): self.s.bc_type[face] = t
    def get_bc_type(self, face): return self.s.bc_type[face]
    def destroy(self): pass
    def set_from_function(self, fp, args=None):
        G = self.G
        for k in range(G.Nz):
            for j in range(G.Ny):
                for i in range(G.Nx):
                    pt = [G.x[1+i], G.y[1+j]] + ([G.z[1+k]] if G.ndim == 3 else [])
                    self.s.I[i, j, k] = fp(pt, args)
        if self.gl: self.s.update_ghost_nodes()
        return self

def scalar(G, l=0, c="c", field_id=None): return _S(fo.Scalar(G, l, c))

class _V:
    def __init__(self, v): self.v = v; self.x = _S(v.x); self.y = _S(v.y); self.z = _S(v.z) if v.G.ndim == 3 else None
    @property
    def comps(self): return [self.x, self.y] + ([self.z] if self.z is not None else [])
    def push(self): return self
    def pull(self): return self
    def update_ghost_nodes(self): self.v.update_ghost_nodes()
def vector(G, l=0, first_id=None): return _V(fo.Vector(G, l))

def gradient(s, g): fo.gradient(s.s, g.v)
def laplacian(a, b):
    if isinstance(a, _S):
        if a.s is b.s: raise FenError("output is input")
        fo.laplacian_scalar(a.s, b.s)
    else: fo.laplacian_vector(a.v, b.v)
def face_to_center(a, b, face): fo.face_to_center(a.s, b.s, face)
def curl(a, b): fo.curl(a.v, b.v)

class PoissonSolver:
    def __init__(self, phi):
        n = phi.G
        for L in (n.Nx, n.Ny, n.Nz):
            for p in (67, 71, 73):
                if L % p == 0: raise FenError("transform sizes must be products of primes <= 61")
        self.ps = fo.PoissonSolver(phi.s)
    @property
    def variant(self): return self.ps.variant
    def solve(self, phi): self.ps.solve(phi.s)

class Solver:
    _cls = fo.NavierStokes
    def __init__(self, G, density=1.0, viscosity=1.0): self.G = G; self._a = (density, viscosity); self.ns = None
    def init_solver(self):
        self.ns = fo.NavierStokes(self.G, *self._a); self._wrap(); return self
    def _wrap(self):
        ns = self.ns
        self.v = _V(ns.v); self.p = _S(ns.p); self.phi = _S(ns.phi)
        if hasattr(ns, "S"): self.S = _V(ns.S)
    @property
    def poisson_variant(self): return self.ns.poisson.variant
    def set_timestep(self, U): return self.ns.set_timestep(U)
    def navier_stokes_solver(self, step, dt): return self.ns.navier_stokes_solver(step, dt)
    def status(self): return self.ns.maxdiv, self.ns.maxCFL
    @property
    def maxdiv(self): return self.ns.maxdiv
    def destroy_solver(self): pass
    def __setattr__(self, k, v):
        if k in ("CFL", "g", "constant_CFL") and self.__dict__.get("ns") is not None: setattr(self.ns, k, list(v) if k == "g" else v)
        else: object.__setattr__(self, k, v)

class MultiphaseSolver(Solver):
    def __init__(self, G): self.G = G; self.ns = None; self._p = {}
    def __setattr__(self, k, v):
        if k in ("rho_0","rho_1","mu_0","mu_1","sigma","beta","g") :
            if self.__dict__.get("ns") is None: self.__dict__.setdefault("_p", {})[k] = v
            elif k == "beta": self.ns.vf.beta = v
            else: setattr(self.ns, k, v)
        else: object.__setattr__(self, k, v)
    def init_solver(self, distance=None):
        p = self._p
        d = np.vectorize(distance)
        self.ns = mf.MultiphaseNavierStokes(self.G, p.get("rho_0",1.0), p.get("rho_1",1.0), p.get("mu_0",1.0), p.get("mu_1",1.0),
                                            p.get("sigma",0.0), distance=lambda x, y: d(x + 0*y, y + 0*x), beta=p.get("beta",1.0))
        if "g" in p: self.ns.g[:len(p["g"][:2])] = p["g"][:2]
        self._wrap(); self.vof = _S(self.ns.vof); self.p_hat = _S(self.ns.p_hat); self.rho = _S(self.ns.rho)
        return self

class VoF:
    def __init__(self, G): self.vf = mf.VoF(G); self.vof = _S(self.vf.vof)
    def get_vof_from_distance(self, fn):
        d = np.vectorize(fn); self.vf.distance = lambda x, y: d(x + 0*y, y + 0*x); self.vf.get_vof_from_distance()
    def advect_vof(self, v, dt): self.vf.advect_vof(v.v, dt)
    def check_vof_integral(self): return self.vf.check_vof_integral()
