"""Committed golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the oracle).

CPU half: the oracle still reproduces them (guards the checker against drift; scipy's FFT may differ in the last
bits between builds, hence 1e-13 and not bit equality).  GPU half: the CUDA path through the C ABI reproduces
them within the north_star tolerance (rel-L2 <= 1e-12) without the oracle in the loop.
"""
import glob
import os

import numpy as np
import pytest

from oracle import fen_oracle as fo

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
NS_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "ns_*.npz")))
PS_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "poisson_*.npz")))
INIT = {"ns_tgv3d": "init_tgv3d", "ns_tgv2d": "init_tgv2d", "ns_channel": "init_channel"}


def rel_l2(a, b):
    n = np.linalg.norm(np.asarray(b).ravel())
    d = np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel())
    return d / n if n > 0 else d


def test_fixtures_exist():
    assert len(NS_CASES) >= 3 and len(PS_CASES) >= 4


def _fields(g):
    return [k for k in ("u", "v", "w", "p") if k in g.files]


@pytest.mark.parametrize("case", NS_CASES)
def test_oracle_reproduces_ns_golden(case):
    g = np.load(os.path.join(GOLD, case + ".npz"))
    n, L = [int(x) for x in g["n"]], [float(x) for x in g["L"]]
    G = fo.Grid(n[0], n[1], n[2], L[0], L[1], L[2], bc=[str(b) for b in g["bc"]])
    ns = fo.NavierStokes(G, 1.0, float(g["nu"]))
    ns.CFL = float(g["cfl"])
    ns.g = [float(x) for x in g["g"]]
    getattr(fo, INIT[case.rsplit("_", 2)[0]])(ns)
    # the stored initial condition is what the init function produces
    assert np.array_equal(ns.v.x.f, g["u0"]) and np.array_equal(ns.p.f, g["p0"])
    dt = ns.set_timestep(float(g["U"]))
    assert dt == float(g["dt"])
    for s in range(1, int(g["steps"]) + 1):
        ns.navier_stokes_solver(s, dt)
    got = {"u": ns.v.x.f, "v": ns.v.y.f, "p": ns.p.f}
    if G.ndim == 3:
        got["w"] = ns.v.z.f
    for k in _fields(g):
        assert rel_l2(got[k], g[k]) < 1e-13, k
    assert abs(ns.maxCFL - float(g["maxCFL"])) < 1e-13


@pytest.mark.parametrize("case", PS_CASES)
def test_oracle_reproduces_poisson_golden(case):
    g = np.load(os.path.join(GOLD, case + ".npz"))
    n = [int(x) for x in g["n"]]
    G = fo.Grid(n[0], n[1], n[2], 1.0, n[1] / n[0], n[2] / n[0], bc=[str(b) for b in g["bc"]])
    phi = fo.Scalar(G, 1)
    phi.I[...] = g["rhs"]
    ps = fo.PoissonSolver(phi)
    assert ps.variant == str(g["variant"])
    ps.solve(phi)
    assert rel_l2(phi.I, g["sol"]) < 1e-13


# ---- GPU half -----------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", NS_CASES)
def test_gpu_reproduces_ns_golden(case):
    import fen_b200 as fb
    g = np.load(os.path.join(GOLD, case + ".npz"))
    n, L = [int(x) for x in g["n"]], [float(x) for x in g["L"]]
    ndim = 2 if n[2] == 1 else 3
    Gg = fb.grid().setup(n[0], n[1], n[2], L[0], L[1], L[2], bc=[str(b) for b in g["bc"]], ndim=ndim)
    ns = fb.Solver(Gg, 1.0, float(g["nu"])).init_solver()
    ns.CFL = float(g["cfl"])
    ns.g = [float(x) for x in g["g"]][:3]
    comps = [("u0", ns.v.x), ("v0", ns.v.y), ("p0", ns.p)] + ([("w0", ns.v.z)] if ndim == 3 else [])
    for key, f in comps:
        f.f[...] = g[key]
        f.push()
    dt = ns.set_timestep(float(g["U"]))
    assert dt == float(g["dt"])
    for s in range(1, int(g["steps"]) + 1):
        ns.navier_stokes_solver(s, dt)
    ns.v.pull(); ns.p.pull()
    got = {"u": ns.v.x.f, "v": ns.v.y.f, "p": ns.p.f}
    if ndim == 3:
        got["w"] = ns.v.z.f
    for k in _fields(g):
        a, b = got[k], g[k]
        if k == "p" and "Wall" in [str(x) for x in g["bc"]]:
            # pn / ppn remove the mean of phi; compare pressures without their mean (hazard H4)
            a = a[1:-1, 1:-1, 1:-1] if ndim == 3 else a[1:-1, 1:-1, :]
            b = b[1:-1, 1:-1, 1:-1] if ndim == 3 else b[1:-1, 1:-1, :]
            a, b = a - a.mean(), b - b.mean()
        if np.linalg.norm(b) == 0.0:
            assert np.abs(a).max() < 1e-14, k
        else:
            assert rel_l2(a, b) < 1e-12, k
    md, mc = ns.status()
    assert abs(mc - float(g["maxCFL"])) < 1e-12 and abs(md) < 1e-12
    Gg.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("case", PS_CASES)
def test_gpu_reproduces_poisson_golden(case):
    import fen_b200 as fb
    g = np.load(os.path.join(GOLD, case + ".npz"))
    n = [int(x) for x in g["n"]]
    Gg = fb.grid().setup(n[0], n[1], n[2], 1.0, n[1] / n[0], n[2] / n[0], bc=[str(b) for b in g["bc"]],
                         ndim=int(g["ndim"]))
    phi = fb.scalar(Gg, 1)
    phi.I[...] = g["rhs"]
    ps = fb.PoissonSolver(phi)
    assert ps.variant == str(g["variant"])
    phi.push()
    ps.solve(phi)
    phi.pull()
    a, b = phi.I, g["sol"]
    if str(g["variant"]) in ("pn", "ppn"):
        a, b = a - a.mean(), b - b.mean()
    assert rel_l2(a, b) < 1e-12
    Gg.destroy()
