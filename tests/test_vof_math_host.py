"""The cell arithmetic the CUDA kernels of multiphase.cu compile (fen_b200/csrc/vof_math.cuh), run from host loops
(tests/cpu/vof_math_host.cpp, g++) and held against the numpy oracle of the two-phase path.  No GPU needed: this is
the transcription check; the kernels themselves are checked by tests/test_gpu_multiphase.py.

Built twice: plain x86-64 (every operation rounded once, like the oracle) and with FMA contraction (-mfma
-ffp-contract=fast), which is what nvcc does on the device -- the second build shows how far contraction alone moves
the results and sets the tolerance the GPU parity tests use."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from oracle import fen_oracle as fo
from oracle import fen_oracle_mf as mf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpu", "vof_math_host.cpp")
PD = C.POINTER(C.c_double)


def _build(tag, extra):
    out = os.path.join(ROOT, "build", "libvof_math_host_%s.so" % tag)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    deps = [SRC, os.path.join(ROOT, "fen_b200", "csrc", "vof_math.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC"] + extra + ["-x", "c++", SRC, "-o", out],
                       check=True)
    return C.CDLL(out)


def _has_fma():
    try:
        return "fma" in open("/proc/cpuinfo").read()
    except OSError:
        return False


VARIANTS = [("plain", [])] + ([("fma", ["-mfma", "-ffp-contract=fast"])] if _has_fma() else [])


def _p(a):
    return a.ctypes.data_as(PD)


def _g2(s):     # ghosted 2-D slice, Fortran order, contiguous
    return np.asfortranarray(s.f[:, :, s.gl])


def _i2(G):
    return np.zeros((G.Nx, G.Ny), order="F")


def _circle_case(N=48, walls=False):
    bc = ["Periodic", "Periodic", "Wall", "Wall"] if walls else None
    G = fo.Grid(N, N, 1, fo.PI, fo.PI, fo.PI / N, bc=bc)
    vf = mf.VoF(G)
    x0, y0, r = 0.5 * fo.PI, 0.2 * (fo.PI + 1), 0.2 * fo.PI
    vf.distance = lambda x, y: np.sqrt((x - x0) ** 2 + (y - y0) ** 2) - r
    vf.get_vof_from_distance()
    v = fo.Vector(G, 1)
    i = np.arange(1, N + 1)[:, None]
    j = np.arange(1, N + 1)[None, :]
    d = G.delta
    v.x.I[..., 0] = np.sin(i * d) * np.cos((j - 0.5) * d)
    v.y.I[..., 0] = -np.cos((i - 0.5) * d) * np.sin(j * d)
    v.update_ghost_nodes()
    return G, vf, v


def _host_recon(lib, G, vf, src):
    outs = [_i2(G) for _ in range(7)]
    lib.host_recon(G.Nx, G.Ny, _p(_g2(src)), C.c_double(G.delta), C.c_double(vf.beta), C.c_double(vf.cut),
                   1 if vf.quadratic else 0, *[_p(o) for o in outs])
    return outs


@pytest.mark.parametrize("tag,extra", VARIANTS)
def test_reconstruction_matches_oracle(tag, extra):
    lib = _build(tag, extra)
    G, vf, _ = _circle_case()
    vf.get_h_from_vof()
    nx, ny, lx, ly, curv, h, d = _host_recon(lib, G, vf, vf.vof)
    ncut = int(((vf.vof.I > vf.cut) & (vf.vof.I < 1 - vf.cut)).sum())
    assert ncut > 100                                   # the interface band is really exercised
    tol = 1e-13 if tag == "plain" else 1e-10
    for name, a, b in (("nx", nx, vf.norm.x), ("ny", ny, vf.norm.y), ("lx", lx, vf.l.x), ("ly", ly, vf.l.y),
                       ("h", h, vf.h)):
        err = float(np.abs(a - b.I[..., 0]).max())
        assert err < tol, (name, err)
    # d comes out of a quadratic that loses digits as vof -> cut or 1 - cut (|d| ~ 9, where the profile is saturated
    # and d no longer matters): weigh its error with the sensitivity dh/dd = 2 beta h (1 - h)
    f = vf.vof.I[..., 0]
    err = float((np.abs(d - vf.d.I[..., 0]) * f * (1.0 - f)).max())
    assert err < tol, ("d", err)
    assert float(np.abs(curv - vf.curv.I[..., 0]).max()) < tol * float(np.abs(vf.curv.I).max() + 1.0)


@pytest.mark.parametrize("tag,extra", VARIANTS)
def test_tiled_reconstruction_is_bit_identical(tag, extra):
    """k_vof_recon_tile's three phases (tile load, corners once per corner, cells) run from loops: the same bits as the
    per-cell path on a grid that is not a multiple of the 64 x 4 tile, with walls, a noisy '1 +- eps' phase and exactly
    uniform regions."""
    lib = _build(tag, extra)
    nx, ny = 150, 37
    G = fo.Grid(nx, ny, 1, 1.0, float(ny) / nx, 1.0 / nx, bc=["Wall", "Wall", "Periodic", "Periodic"])
    vf = mf.VoF(G)
    vf.distance = lambda x, y: 0.11 - np.sqrt((x - 0.43) ** 2 + (y - 0.12) ** 2)
    vf.get_vof_from_distance()
    rng = np.random.default_rng(1)
    noisy = vf.vof.I[..., 0] >= 1.0 - 1e-9
    vf.vof.I[..., 0][noisy] *= 1.0 + 1e-16 * rng.integers(-3, 4, size=int(noisy.sum()))
    vf.vof.update_ghost_nodes()
    a = _host_recon(lib, G, vf, vf.vof)
    b = [_i2(G) for _ in range(7)]
    lib.host_recon_tile(G.Nx, G.Ny, _p(_g2(vf.vof)), C.c_double(G.delta), C.c_double(vf.beta), C.c_double(vf.cut),
                        1 if vf.quadratic else 0, *[_p(o) for o in b])
    for name, x, y in zip(("nx", "ny", "lx", "ly", "curv", "h", "d"), a, b):
        assert np.array_equal(x, y), name
    assert (np.abs(a[0]) > 0).sum() > 200 and (a[0] == 0).sum() > 200        # both kinds of cells were present


@pytest.mark.parametrize("tag,extra", VARIANTS)
def test_advect_vof_matches_oracle(tag, extra):
    """Twenty advect_vof calls (both sweep orders, the H13 boundary-type switch, negative and positive face
    velocities) driven with the host cell functions in the kernels' own sequence (multiphase.cu: advect_vof)."""
    lib = _build(tag, extra)
    G, vf, v = _circle_case(walls=True)
    dt = 0.4 * G.delta
    u2, v2 = _g2(v.x), _g2(v.y)
    # shadow state driven by the host functions
    vof = vf.vof.copy()
    x_first = True
    worst = 0.0
    for step in range(20):
        vf.advect_vof(v, dt)
        # --- the device sequence -----------------------------------------------------------------------------
        rec = _host_recon(lib, G, vf, vof)

        def ghosted(arr, like):
            s = like.copy()
            s.I[..., 0] = arr
            s.update_ghost_nodes()
            return _g2(s)
        wired = vf.h                                     # boundary types of allocate_vof_fields (never reassigned)
        gn = [ghosted(rec[q], wired) for q in (0, 1, 2, 3, 6)]       # nx, ny, lx, ly, d
        vof1 = fo.Scalar(G, 1, "c")                      # fresh scalar: default boundary types
        out = _i2(G)
        d1, d2 = (1, 2) if x_first else (2, 1)
        lib.host_sweep(G.Nx, G.Ny, d1, 0, int(x_first), _p(_g2(vof)), *[_p(a) for a in gn], _p(u2), _p(v2),
                       C.c_double(dt), C.c_double(G.delta), C.c_double(vf.beta), C.c_double(vf.cut), _p(out))
        vof1.I[..., 0] = out
        vof.bc_type = dict(vof1.bc_type)                 # hazard H13
        vof1.update_ghost_nodes()
        rec = _host_recon(lib, G, vf, vof1)
        gn = [ghosted(rec[q], wired) for q in (0, 1, 2, 3, 6)]
        lib.host_sweep(G.Nx, G.Ny, d2, 1, int(x_first), _p(_g2(vof1)), *[_p(a) for a in gn], _p(u2), _p(v2),
                       C.c_double(dt), C.c_double(G.delta), C.c_double(vf.beta), C.c_double(vf.cut), _p(out))
        vof.I[..., 0] = out
        vof.update_ghost_nodes()
        x_first = not x_first
        worst = max(worst, float(np.abs(vof.f - vf.vof.f).max()))
    tol = 1e-13 if tag == "plain" else 1e-10
    assert worst < tol, worst
    assert vf.vof.bc_type["bottom"] == 0                 # the oracle switched to the default types as well


@pytest.mark.parametrize("tag,extra", VARIANTS)
def test_predictor_cell_matches_oracle(tag, extra):
    """mf_predict_cell against predicted_velocity_field of the two-phase oracle (density ratio 850, surface tension,
    gravity, a source term and a non-zero p_hat)."""
    lib = _build(tag, extra)
    Nx, Ny = 32, 64
    G = fo.Grid(Nx, Ny, 1, 1.0, 2.0, 1.0 / Nx, bc=["Periodic", "Periodic", "Wall", "Wall"])
    wave = lambda x, y: y - 0.05 * np.cos(2 * fo.PI * x) - 1.0          # noqa: E731
    ns = mf.MultiphaseNavierStokes(G, 1000.0, 1000.0 / 850.0, 1.0, 0.019, 0.07, distance=wave, beta=2.0)
    ns.g[1] = -9.80665
    rng = np.random.default_rng(7)
    for s in (ns.v.x, ns.v.y, ns.p, ns.p_o):
        s.I[...] = rng.standard_normal(s.I.shape) * 0.1
    ns.v.update_ghost_nodes(); ns.p.update_ghost_nodes()
    for s in ns.S.comps + ns.dv_o.comps:
        s.I[...] = rng.standard_normal(s.I.shape)
    dt = ns.set_timestep(1.0)
    ns.dt_o = 0.8 * dt
    ns.vf.advect_vof(ns.v, dt)
    ns.update_material_properties()
    ns.p_hat.f[...] = 2.0 * ns.p.f - ns.p_o.f
    ns.p_hat.update_ghost_nodes()
    ins = [_g2(s) for s in (ns.v.x, ns.v.y, ns.p, ns.p_hat, ns.rho, ns.mu, ns.vf.vof, ns.vf.curv)]
    ints = [np.asfortranarray(s.I[..., 0].copy()) for s in (ns.S.x, ns.S.y, ns.dv_o.x, ns.dv_o.y)]
    outs = [_i2(G) for _ in range(4)]
    A = 1.0 + 0.5 * dt / ns.dt_o
    B = -0.5 * dt / ns.dt_o
    cd = C.c_double
    lib.host_predict(Nx, Ny, *[_p(a) for a in ins], *[_p(a) for a in ints], cd(1.0 / G.delta), cd(dt), cd(A), cd(B),
                     cd(ns.g[0]), cd(ns.g[1]), cd(ns.sigma), cd(ns.irhomin), *[_p(o) for o in outs])
    ns.predicted_velocity_field(dt)
    # the oracle ends with v%update_ghost_nodes (navier_stokes.f90:208): the wall Dirichlet condition of the normal
    # component also overwrites the last interior face (hazard H3) -- do the same to the host result
    for a, b in ((outs[0], ns.v.x), (outs[1], ns.v.y)):
        t = b.copy()
        t.I[..., 0] = a
        t.update_ghost_nodes()
        a[...] = t.I[..., 0]
    for name, a, b in (("u", outs[0], ns.v.x), ("v", outs[1], ns.v.y), ("dvx", outs[2], ns.dv_o.x),
                       ("dvy", outs[3], ns.dv_o.y)):
        ref = b.I[..., 0]
        rel = float(np.sqrt(((a - ref) ** 2).sum()) / np.sqrt((ref ** 2).sum()))
        assert rel < (1e-14 if tag == "plain" else 1e-12), (name, rel)


def test_gauss_points_are_the_references():
    """rp, rm of volume_of_fluid.f90:33-34 as literals in vof_math.cuh."""
    src = open(os.path.join(ROOT, "fen_b200", "csrc", "vof_math.cuh")).read()
    rp = float(src.split("VOF_RP =")[1].split(";")[0])
    rm = float(src.split("VOF_RM =")[1].split(";")[0])
    assert rp == 0.5 * (1.0 + 1.0 / math.sqrt(3.0)) and rm == 0.5 * (1.0 - 1.0 / math.sqrt(3.0))
