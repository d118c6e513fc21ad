"""The plain-C restatement (oracle/fen_oracle_c.c) against the numpy oracle and the committed golden fixture.
Both restate the same reference loops independently (numpy whole-array expressions + scipy.fft vs explicit triple
loops + a textbook radix-2 FFT), so agreement to round-off pins each of them."""
import os

import numpy as np
import pytest

from oracle import fen_oracle as fo
from oracle import fen_oracle_c as foc

PI = fo.PI
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    n = np.linalg.norm(b.ravel())
    return np.linalg.norm((a - b).ravel()) / (n if n > 0 else 1.0)


@pytest.mark.parametrize("n", [(16, 16, 16), (32, 16, 64)])
def test_c_poisson_matches_numpy_oracle(n):
    G = fo.Grid(n[0], n[1], n[2], 1.0, n[1] / n[0], n[2] / n[0])
    rng = np.random.default_rng(4)
    rhs = rng.standard_normal(n)
    rhs -= rhs.mean()
    po = fo.Scalar(G, 1)
    po.I[...] = rhs
    fo.PoissonSolver(po).solve(po)
    c = foc.NavierStokesC(n[0], n[1], n[2], G.delta)
    a = np.zeros((n[0] + 2, n[1] + 2, n[2] + 2), order="F")
    a[1:-1, 1:-1, 1:-1] = rhs
    c.solve_poisson(a)
    assert rel(a[1:-1, 1:-1, 1:-1], po.I) < 1e-12
    c.destroy()


@pytest.mark.parametrize("threads", [1, 4])
def test_c_steps_match_numpy_oracle_and_golden(threads):
    g = np.load(os.path.join(GOLD, "ns_tgv3d_16_2steps.npz"))
    n = [int(x) for x in g["n"]]
    Go = fo.Grid(n[0], n[1], n[2], 2 * PI, 2 * PI, 2 * PI)
    nso = fo.NavierStokes(Go, 1.0, float(g["nu"]))
    nso.CFL = float(g["cfl"])
    fo.init_tgv3d(nso)
    dt = nso.set_timestep(float(g["U"]))
    c = foc.NavierStokesC(n[0], n[1], n[2], Go.delta, 1.0, float(g["nu"]), threads=threads)
    c.dt_o = dt
    for fid, f in ((foc.P, nso.p), (foc.U, nso.v.x), (foc.V, nso.v.y), (foc.W, nso.v.z)):
        c.set(fid, f.f)
    for s in range(1, int(g["steps"]) + 1):
        nso.navier_stokes_solver(s, dt)
        c.navier_stokes_solver(s, dt)
    for key, fid, f in (("u", foc.U, nso.v.x), ("v", foc.V, nso.v.y), ("w", foc.W, nso.v.z), ("p", foc.P, nso.p)):
        got = c.get(fid)
        assert rel(got, f.f) < 1e-12, key                  # ghosts included
        assert rel(got, g[key]) < 1e-12, key               # committed fixture
    assert abs(c.maxCFL(dt) - nso.maxCFL) < 1e-13 and abs(c.maxdiv) < 1e-13
    # a few more steps: divergence stays at round-off, results independent of the thread count
    for s in range(3, 8):
        nso.navier_stokes_solver(s, dt)
        c.navier_stokes_solver(s, dt)
    assert rel(c.get(foc.U), nso.v.x.f) < 1e-12 and abs(c.maxdiv) < 1e-13
    c.destroy()


ZWALLS = ["Periodic"] * 4 + ["Wall", "Wall"]


@pytest.mark.parametrize("n", [(16, 16, 8), (32, 16, 24)])
def test_c_poisson_ppn_matches_numpy_oracle(n):
    """poisson_solver_ppn restated twice: explicit per-system Thomas loops in C vs whole-array numpy, incl. the
    exactly-zero last pivot of the singular mode (hazard H5) and the mean removal."""
    G = fo.Grid(n[0], n[1], n[2], 1.0, n[1] / n[0], n[2] / n[0], bc=ZWALLS)
    rng = np.random.default_rng(6)
    rhs = rng.standard_normal(n)
    rhs -= rhs.mean()
    po = fo.Scalar(G, 1)
    po.I[...] = rhs
    ps = fo.PoissonSolver(po)
    assert ps.variant == "ppn"
    ps.solve(po)
    c = foc.NavierStokesC(n[0], n[1], n[2], G.delta, zwalls=True)
    a = np.zeros((n[0] + 2, n[1] + 2, n[2] + 2), order="F")
    a[1:-1, 1:-1, 1:-1] = rhs
    c.solve_poisson(a)
    assert rel(a[1:-1, 1:-1, 1:-1], po.I) < 1e-12
    assert abs(a[1:-1, 1:-1, 1:-1].mean()) < 1e-14
    c.destroy()


@pytest.mark.parametrize("threads", [1, 4])
def test_c_channel_steps_match_numpy_oracle_and_golden(threads):
    """BASELINE configs[2] in miniature (the committed channel fixture): walls in z, body force, ppn Poisson."""
    g = np.load(os.path.join(GOLD, "ns_channel_16x16x8_3steps.npz"))
    n, L = [int(x) for x in g["n"]], [float(x) for x in g["L"]]
    Go = fo.Grid(n[0], n[1], n[2], L[0], L[1], L[2], bc=[str(b) for b in g["bc"]])
    nso = fo.NavierStokes(Go, 1.0, float(g["nu"]))
    nso.g = [float(x) for x in g["g"]]
    fo.init_channel(nso)
    assert np.array_equal(nso.v.x.f, g["u0"])
    dt = nso.set_timestep(float(g["U"]))
    c = foc.NavierStokesC(n[0], n[1], n[2], Go.delta, 1.0, float(g["nu"]), threads=threads, zwalls=True)
    c.dt_o = dt
    c.g = list(nso.g)
    for fid, f in ((foc.P, nso.p), (foc.U, nso.v.x), (foc.V, nso.v.y), (foc.W, nso.v.z)):
        c.set(fid, f.f)
    for s in range(1, int(g["steps"]) + 1):
        nso.navier_stokes_solver(s, dt)
        c.navier_stokes_solver(s, dt)
    for key, fid, f in (("u", foc.U, nso.v.x), ("v", foc.V, nso.v.y), ("w", foc.W, nso.v.z), ("p", foc.P, nso.p)):
        got = c.get(fid)
        assert rel(got, f.f) < 1e-12, key                  # ghosts included: the wall wiring is the same
        assert rel(got, g[key]) < 1e-12, key               # committed fixture
    assert abs(c.maxCFL(dt) - nso.maxCFL) < 1e-13 and abs(c.maxdiv - nso.maxdiv) < 1e-13
    for s in range(4, 10):
        nso.navier_stokes_solver(s, dt)
        c.navier_stokes_solver(s, dt)
    assert rel(c.get(foc.U), nso.v.x.f) < 1e-12 and rel(c.get(foc.P), nso.p.f) < 1e-11
    c.destroy()


# ---- the two-phase restatement (oracle/fen_oracle_mf_c.c) ------------------------------------------------------------
def _wave(Nx, Ny, sigma):
    """The wave case of tests/golden/make_golden.py::mf_case (viscous_decay.f90 with a larger amplitude)."""
    import math
    from oracle import fen_oracle_mf as mf
    Lx, Ly = 1.0, float(Ny) / Nx
    G = fo.Grid(Nx, Ny, 1, Lx, Ly, Lx / Nx, bc=["Periodic", "Periodic", "Wall", "Wall"])
    rho_0 = 1000.0
    mu_0 = rho_0 * Lx * math.sqrt(mf.GRAVITY * Lx) / 1.0e4
    ns = mf.MultiphaseNavierStokes(G, rho_0, rho_0 / 850.0, mu_0, mu_0 * 1.9e-2, sigma,
                                   distance=lambda x, y: y - 0.05 * np.cos(2.0 * PI * x / Lx) - Ly / 2.0)
    ns.g[1] = -mf.GRAVITY
    i = np.arange(1, Nx + 1)[:, None]
    j = np.arange(1, Ny + 1)[None, :]
    d = G.delta
    wn = 2.0 * PI / Lx
    om = math.sqrt(mf.GRAVITY * wn)
    F = ns.vof.sh
    x, y = i * d, (j - 0.5) * d - Ly / 2.0
    f = ((F(1, 0) + F()) * 0.5)[..., 0]
    ns.v.x.I[..., 0] = (1.0 - f) * 0.05 * om * np.exp(wn * y) * np.cos(wn * x) - f * 0.05 * om * np.exp(-wn * y) * np.cos(wn * x)
    x, y = (i - 0.5) * d, j * d - Ly / 2.0
    f = ((F(0, 1) + F()) * 0.5)[..., 0]
    ns.v.y.I[..., 0] = (1.0 - f) * 0.05 * om * np.exp(wn * y) * np.sin(wn * x) + f * 0.05 * om * np.exp(-wn * y) * np.sin(wn * x)
    ns.v.update_ghost_nodes()
    return G, ns, 0.1 * ns.set_timestep(1.0)


@pytest.mark.parametrize("sigma,threads", [(0.0, 1), (0.07, 4)])
def test_c_two_phase_steps_match_numpy_oracle(sigma, threads):
    """MTHINC advection, material properties, variable-viscosity predictor with the pressure splitting, pn Poisson,
    correction: the C loops (written from the Fortran) and the numpy whole-array expressions agree to round-off over
    seven steps -- 1e-15 after the first, 1e-12 after the seventh (u, v are O(0.1); vof is O(1), held absolutely)."""
    from oracle import fen_oracle_mf_c as mfc
    G, ns, dt = _wave(32, 64, sigma)
    c = mfc.MultiphaseC.from_oracle(ns, threads=threads)
    for s in range(1, 8):
        ns.navier_stokes_solver(s, dt)
        c.navier_stokes_solver(s, dt)
        tol = 1e-14 if s == 1 else 2e-12
        for name, fid, o in (("u", mfc.U, ns.v.x), ("v", mfc.V, ns.v.y), ("p", mfc.P, ns.p), ("rho", mfc.RHO, ns.rho),
                             ("mu", mfc.MU, ns.mu), ("p_o", mfc.PO, ns.p_o)):
            assert rel(c.get(fid), o.f[:, :, 1]) < tol, (s, name)
        assert np.abs(c.get(mfc.VOF) - ns.vof.f[:, :, 1]).max() < 1e-14 * s
        assert c.x_first == ns.vf.x_first
        assert abs(c.maxdiv - ns.maxdiv) < 1e-12 and abs(c.maxCFL(dt) - ns.maxCFL) < 1e-12 * ns.maxCFL
    c.destroy()


def test_c_two_phase_reproduces_the_golden_fixture():
    """tests/golden/mf_wave_16x32_3steps.npz (frozen numpy-oracle output, also the GPU fixture) from its stored inputs."""
    from oracle import fen_oracle_mf_c as mfc
    g = np.load(os.path.join(GOLD, "mf_wave_16x32_3steps.npz"))
    Nx, Ny = (int(v) for v in g["n"])
    G, ns, dt = _wave(Nx, Ny, float(g["sigma"]))
    assert dt == float(g["dt"]) and np.array_equal(ns.vof.f, g["vof0"]) and np.array_equal(ns.v.x.f, g["u0"])
    c = mfc.MultiphaseC.from_oracle(ns)
    for s in range(1, int(g["steps"]) + 1):
        c.navier_stokes_solver(s, dt)
    for name, fid in (("u", mfc.U), ("v", mfc.V), ("p", mfc.P), ("vof", mfc.VOF), ("rho", mfc.RHO)):
        assert rel(c.get(fid), g[name][:, :, 1]) < 1e-12, name
    assert abs(c.maxdiv - float(g["maxdiv"])) < 1e-12 and abs(c.maxCFL(dt) - float(g["maxCFL"])) < 1e-13
    c.destroy()
