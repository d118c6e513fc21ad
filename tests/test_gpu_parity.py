"""GPU parity tests: the CUDA path (through the C ABI of libfen_gpu.so) against the CPU oracle.

Tolerances are BASELINE.json's: relative L2 <= 1e-12 on u, v, w, p after one step, <= 1e-9 after
100 steps of a laminar Taylor-Green case, divergence at machine precision.
"""
import numpy as np
import pytest

import fen_b200 as fb
from oracle import fen_oracle as fo

pytestmark = pytest.mark.gpu
PI = fo.PI


def rel_l2(a, b):
    n = np.linalg.norm(b.ravel())
    return np.linalg.norm((a - b).ravel()) / (n if n > 0 else 1.0)


def make_pair(n, bc=None, ndim=3, L=1.0):
    nz = n[2]
    Go = fo.Grid(n[0], n[1], nz, L, L * n[1] / n[0], L * nz / n[0], bc=bc, ndim=ndim)
    Gg = fb.grid().setup(n[0], n[1], nz, L, L * n[1] / n[0], L * nz / n[0], bc=bc, ndim=ndim)
    assert Gg.delta == Go.delta
    return Go, Gg


BC_SETS_3D = {
    "periodic": ["Periodic"] * 6,
    "zwalls": ["Periodic"] * 4 + ["Wall", "Wall"],
}


# ---- scalar%update_ghost_nodes ---------------------------------------------------------------
@pytest.mark.parametrize("types", [(0, 0, 0, 0, 0, 0), (1, 1, 2, 2, 1, 2), (2, 1, 1, 1, 2, 2), (0, 0, 1, 1, 2, 1)])
@pytest.mark.parametrize("loc", ["c", "x", "y", "z"])
def test_ghost_nodes_all_bc_types(types, loc):
    Go, Gg = make_pair((16, 12, 8))
    rng = np.random.default_rng(7)
    so = fo.Scalar(Go, 1, loc)
    sg = fb.scalar(Gg, 1, loc)
    so.f[...] = rng.standard_normal(so.f.shape)
    sg.f[...] = so.f
    for face, t in zip(fo.FACES, types):
        so.bc_type[face] = t
        sg.set_bc_type(face, t)
        plane = rng.standard_normal(so.bc[face].shape)
        so.bc[face][...] = plane
        sg.set_bc(face, plane)
    so.update_ghost_nodes()
    sg.push()
    sg.update_ghost_nodes()
    sg.pull()
    assert np.array_equal(sg.f, so.f)          # pure copies / 2*bc - f: bit exact
    Gg.destroy()


def test_ghost_nodes_uniform_bc_value_and_2d():
    Go, Gg = make_pair((16, 16, 1), bc=["Periodic", "Periodic", "Wall", "Wall"], ndim=2)
    so = fo.Scalar(Go, 1, "x")
    sg = fb.scalar(Gg, 1, "x")
    so.f[...] = np.random.default_rng(3).standard_normal(so.f.shape)
    sg.f[...] = so.f
    for face, t in (("bottom", 1), ("top", 1)):
        so.bc_type[face] = t
        sg.set_bc_type(face, t)
    so.bc["top"][...] = 1.5                      # lid_driven.f90:59  v%x%bc%top = U
    sg.set_bc("top", 1.5)
    so.update_ghost_nodes()
    sg.push().update_ghost_nodes()
    sg.pull()
    assert np.array_equal(sg.f, so.f)
    Gg.destroy()


# ---- fields_mod operators --------------------------------------------------------------------
@pytest.mark.parametrize("n,ndim", [((16, 16, 16), 3), ((32, 16, 1), 2)])
def test_field_operators(n, ndim):
    Go, Gg = make_pair(n, ndim=ndim)
    rng = np.random.default_rng(11)
    vo, vg = fo.Vector(Go, 1), fb.vector(Gg, 1)
    so, sg = fo.Scalar(Go, 1), fb.scalar(Gg, 1)
    for co, cg in zip(vo.comps + [so], vg.comps + [sg]):
        co.I[...] = rng.standard_normal(co.I.shape)
        co.update_ghost_nodes()
        cg.f[...] = co.f
        cg.push()
    go, gg = fo.Vector(Go, 0), fb.vector(Gg, 0)
    fo.gradient(so, go); fb.gradient(sg, gg); gg.pull()
    do, dg = fo.Scalar(Go, 0), fb.scalar(Gg, 0)
    fo.divergence(vo, do); fb.divergence(vg, dg); dg.pull()
    lo, lg = fo.Vector(Go, 1), fb.vector(Gg, 1)
    fo.laplacian_vector(vo, lo); fb.laplacian(vg, lg); lg.pull()
    co_, cg_ = fo.Vector(Go, 0), fb.vector(Gg, 0)
    fo.center_to_face(so, co_); fb.center_to_face(sg, cg_); cg_.pull()
    for a, b in zip(gg.comps + [dg] + cg_.comps, go.comps + [do] + co_.comps):
        assert rel_l2(a.I, b.I) < 1e-14
    for a, b in zip(lg.comps, lo.comps):
        assert rel_l2(a.I, b.I) < 1e-13
    assert abs(dg.max_value() - do.max_value()) <= 1e-13 * abs(do.max_value())
    assert abs(dg.integral() - do.integral()) <= 1e-10 * max(1.0, abs(do.integral()))
    Gg.destroy()


# ---- poisson_mod -----------------------------------------------------------------------------
POISSON_CASES = [
    ("ppp", (32, 32, 32), ["Periodic"] * 6, 3),
    ("ppp", (64, 16, 32), ["Periodic"] * 6, 3),
    ("ppp", (8, 128, 16), ["Periodic"] * 6, 3),
    # register-path transforms (length >= 64): every radix plan 8.8, 8.8.2, 8.8.4, 8.8.8, 8.8.8.2 in x, y and z
    ("ppp", (128, 64, 64), ["Periodic"] * 6, 3),
    ("ppp", (256, 128, 256), ["Periodic"] * 6, 3),
    ("ppp", (512, 256, 8), ["Periodic"] * 6, 3),
    ("ppp", (1024, 512, 8), ["Periodic"] * 6, 3),
    ("ppp", (16, 8, 1024), ["Periodic"] * 6, 3),
    ("ppp", (2048, 8, 64), ["Periodic"] * 6, 3),
    ("ppn", (128, 1024, 16), ["Periodic"] * 4 + ["Wall", "Wall"], 3),
    ("pp", (256, 2048, 1), ["Periodic"] * 4, 2),
    ("ppn", (32, 32, 32), ["Periodic"] * 4 + ["Wall", "Wall"], 3),
    ("ppn", (16, 64, 128), ["Periodic"] * 4 + ["Wall", "Wall"], 3),
    ("ppn", (32, 16, 16), ["Periodic"] * 4 + ["Inflow", "Outflow"], 3),
    ("pp", (64, 64, 1), ["Periodic"] * 4, 2),
    ("pp", (256, 32, 1), ["Periodic"] * 4, 2),
    ("pn", (64, 64, 1), ["Periodic", "Periodic", "Wall", "Wall"], 2),
    ("pn", (4, 32, 1), ["Periodic", "Periodic", "Wall", "Wall"], 2),
    ("pn", (16, 16, 1), ["Periodic", "Periodic", "Inflow", "Outflow"], 2),
    # Neumann directions (DCT-II / DCT-III): nn, npn, nnn  
    ("nn", (32, 32, 1), ["Wall"] * 4, 2),
    ("nn", (64, 16, 1), ["Wall"] * 4, 2),
    ("nn", (16, 32, 1), ["Wall", "Wall", "Inflow", "Outflow"], 2),
    ("npn", (16, 32, 16), ["Wall", "Wall", "Periodic", "Periodic", "Wall", "Wall"], 3),
    ("npn", (32, 128, 8), ["Wall", "Wall", "Periodic", "Periodic", "Wall", "Wall"], 3),
    ("nnn", (16, 16, 16), ["Wall"] * 6, 3),
    ("nnn", (64, 32, 16), ["Wall"] * 6, 3),
    ("nnn", (8, 128, 32), ["Wall"] * 6, 3),
]


@pytest.mark.parametrize("variant,n,bc,ndim", POISSON_CASES)
def test_poisson_variants_match_oracle(variant, n, bc, ndim):
    Go, Gg = make_pair(n, bc=bc, ndim=ndim)
    rng = np.random.default_rng(5)
    rhs = rng.standard_normal(n)
    if variant in ("ppp", "pp"):
        rhs -= rhs.mean()
    po, pg = fo.Scalar(Go, 1), fb.scalar(Gg, 1)
    po.I[...] = rhs
    pg.I[...] = rhs
    pso, psg = fo.PoissonSolver(po), fb.PoissonSolver(pg)
    assert psg.variant == pso.variant == variant
    pso.solve(po)
    pg.push()
    psg.solve(pg)
    pg.pull()
    assert rel_l2(pg.I, po.I) < 1e-12
    # solving twice gives the same answer (no state carried between solves)
    pg.I[...] = rhs
    pg.push(); psg.solve(pg); pg.pull()
    assert rel_l2(pg.I, po.I) < 1e-12
    Gg.destroy()


def test_lid_driven_cavity_nn_steps_match_oracle():
    """test/small_test/navier_stokes/lid_driven: 2-D cavity, walls everywhere, moving lid v%x%bc%top = 1
    (lid_driven.f90:59), nn Poisson (DCT in x, tridiagonal in y)."""
    n = 32
    bc = ["Wall"] * 4
    Go = fo.Grid(n, n, 1, 1.0, 1.0, 1.0 / n, bc=bc)
    Gg = fb.grid().setup(n, n, 1, 1.0, 1.0, 1.0 / n, bc=bc)
    nso = fo.NavierStokes(Go, 1.0, 1.0e-2)
    nsg = fb.Solver(Gg, 1.0, 1.0e-2).init_solver()
    assert nsg.poisson_variant == "nn"
    nso.v.x.bc["top"][...] = 1.0
    nsg.v.x.set_bc("top", 1.0)
    nso.CFL = nsg.CFL = 0.25
    dt = nso.set_timestep(1.0)
    assert nsg.set_timestep(1.0) == dt
    for step in range(1, 31):
        nso.navier_stokes_solver(step, dt)
        nsg.navier_stokes_solver(step, dt)
    _compare(nso, nsg, 1e-11)
    assert abs(nsg.maxdiv) < 1e-11 and np.abs(nsg.v.x.I).max() > 1e-3
    Gg.destroy()


def test_cavity_3d_nnn_steps_match_oracle():
    """test/large_test/lid3D in miniature: 3-D cavity, nnn Poisson (DCT in x and y, tridiagonal in z)."""
    n = 16
    bc = ["Wall"] * 6
    Go = fo.Grid(n, n, n, 1.0, 1.0, 1.0, bc=bc)
    Gg = fb.grid().setup(n, n, n, 1.0, 1.0, 1.0, bc=bc)
    nso = fo.NavierStokes(Go, 1.0, 1.0e-2)
    nsg = fb.Solver(Gg, 1.0, 1.0e-2).init_solver()
    assert nsg.poisson_variant == "nnn"
    nso.v.x.bc["top"][...] = 1.0
    nsg.v.x.set_bc("top", 1.0)
    nso.CFL = nsg.CFL = 0.25
    dt = nso.set_timestep(1.0)
    assert nsg.set_timestep(1.0) == dt
    for step in range(1, 11):
        nso.navier_stokes_solver(step, dt)
        nsg.navier_stokes_solver(step, dt)
    _compare(nso, nsg, 1e-11)
    assert abs(nsg.maxdiv) < 1e-11 and np.abs(nsg.v.x.I).max() > 1e-3
    Gg.destroy()


def test_poisson_unsupported_bc_combo_is_an_error():
    Gg = fb.grid().setup(16, 16, 16, 1.0, 1.0, 1.0,
                         bc=["Periodic", "Periodic", "Wall", "Wall", "Periodic", "Periodic"])
    phi = fb.scalar(Gg, 1)
    with pytest.raises(fb.FenError):            # poisson.f90:91-95 `stop`
        fb.PoissonSolver(phi)
    Gg.destroy()


@pytest.mark.parametrize("case", ["ppp", "ppn"])
def test_projection_makes_velocity_divergence_free(case):
    """test/small_test/poisson/projection/projection.f90: random 32^3 field, max div <= 1e-11."""
    bc = ["Periodic"] * 6 if case == "ppp" else ["Periodic"] * 4 + ["Wall", "Wall"]
    Gg = fb.grid().setup(32, 32, 32, 1.0, 1.0, 1.0, bc=bc)
    v, div, phi, gphi = fb.vector(Gg, 1), fb.scalar(Gg, 0), fb.scalar(Gg, 1), fb.vector(Gg, 1)
    if case == "ppn":
        for f in ("front", "back"):
            phi.set_bc_type(f, 2)
            for c in v.comps:
                c.set_bc_type(f, 1)
    ps = fb.PoissonSolver(phi)
    rng = np.random.default_rng(99)
    div_max = 0.0
    for _ in range(5):
        for c in v.comps:
            c.I[...] = rng.random(c.I.shape)
            c.push()
        v.update_ghost_nodes()
        fb.divergence(v, div)
        div.pull()
        phi.I[...] = -div.I
        phi.push()
        ps.solve(phi)
        phi.update_ghost_nodes()
        fb.gradient(phi, gphi)
        v.pull(); gphi.pull()
        for c, g in zip(v.comps, gphi.comps):
            c.f[...] = c.f + g.f
            c.push()
        v.update_ghost_nodes()
        fb.divergence(v, div)
        div_max = max(div_max, abs(div.max_value()))
    assert div_max <= 1e-11
    Gg.destroy()


# ---- navier_stokes_solver --------------------------------------------------------------------
def _mirror_state(nso, nsg):
    for a, b in ((nso.p, nsg.p), (nso.v.x, nsg.v.x), (nso.v.y, nsg.v.y)):
        b.f[...] = a.f
        b.push()
    if nso.G.ndim == 3:
        nsg.v.z.f[...] = nso.v.z.f
        nsg.v.z.push()


def _compare(nso, nsg, tol):
    nsg.v.pull(); nsg.p.pull()
    errs = {}
    for name, a, b in [("u", nsg.v.x, nso.v.x), ("v", nsg.v.y, nso.v.y), ("p", nsg.p, nso.p)] + \
            ([("w", nsg.v.z, nso.v.z)] if nso.G.ndim == 3 else []):
        if np.linalg.norm(b.I) == 0.0:
            errs[name] = np.abs(a.I).max()
        else:
            errs[name] = rel_l2(a.f, b.f)          # ghosts included
    assert all(e <= tol for e in errs.values()), errs
    return errs


def _setup_ns(n, bc, ndim, L, nu, init, U, g=None, cfl=1.0):
    Go, Gg = make_pair(n, bc=bc, ndim=ndim, L=L)
    nso = fo.NavierStokes(Go, 1.0, nu)
    nsg = fb.Solver(Gg, 1.0, nu).init_solver()
    nso.CFL = cfl
    nsg.CFL = cfl
    if g is not None:
        nso.g = list(g)
        nsg.g = list(g)
    init(nso)
    _mirror_state(nso, nsg)
    dt = nso.set_timestep(U)
    dtg = nsg.set_timestep(U)
    assert dtg == dt
    return Go, Gg, nso, nsg, dt


def _unfused_checks(nsg, dt):
    """checks (navier_stokes.f90:570) recomputed by the stand-alone k_check kernel from the stored fields"""
    from fen_b200._lib import check
    check(nsg.G.lib.fen_gpu_checks(nsg.G.ctx, dt))
    return nsg.status()


@pytest.mark.parametrize("n", [(32, 32, 32), (64, 64, 64), (128, 32, 64), (256, 64, 128)])
def test_one_step_tgv3d_matches_oracle(n):
    Go, Gg, nso, nsg, dt = _setup_ns(n, ["Periodic"] * 6, 3, 2 * PI, 0.01, fo.init_tgv3d, 1.0)
    assert nsg.poisson_variant == "ppp"
    nso.navier_stokes_solver(1, dt)
    nsg.navier_stokes_solver(1, dt)
    _compare(nso, nsg, 1e-12)
    md, mc = nsg.status()
    # divergence at round-off level: the reference's own bound (projection.f90:125) and no worse than the
    # oracle's round-off on the same grid (1.2e-12 at 256x64x128, where delta is 40x smaller than the velocity)
    assert abs(md) < 1e-11 and abs(md) <= max(1e-12, 4.0 * abs(nso.maxdiv))
    assert abs(mc - nso.maxCFL) < 1e-12
    # the checks fused into the correction kernel are bit-identical to the stand-alone ones
    assert _unfused_checks(nsg, dt) == (md, mc)
    Gg.destroy()


def test_100_steps_tgv3d_matches_oracle():
    """100 steps of the laminar Taylor-Green case of BASELINE config 2.  At 64^3 the time step is set
    to the one the 512^3 grid gets from set_timestep(U=1) (CFL = 1/8 here), i.e. the same physical
    horizon t = 1.23: at CFL = 1 on a 64^3 grid AB2 itself amplifies round-off by ~2x per step (in the
    oracle as much as here), which says nothing about parity."""
    Go, Gg, nso, nsg, dt = _setup_ns((64, 64, 64), ["Periodic"] * 6, 3, 2 * PI, 0.01, fo.init_tgv3d, 1.0,
                                     cfl=0.125)
    for step in range(1, 101):
        nso.navier_stokes_solver(step, dt)
        nsg.navier_stokes_solver(step, dt)
    _compare(nso, nsg, 1e-9)
    md, _ = nsg.status()
    assert abs(md) < 1e-11            # reference bound, projection.f90:125
    assert abs(md) < 5e-13            # machine precision at this size
    Gg.destroy()


def test_steps_tgv2d_matches_oracle_config1():
    """BASELINE config 1 (2-D Taylor-Green, pp Poisson) at 64^2: 250 steps to t = 0.3."""
    n = 64
    Go, Gg, nso, nsg, dt = _setup_ns((n, n, 1), ["Periodic"] * 4, 2, 2 * PI, 1.0, fo.init_tgv2d, 2.0)
    assert nsg.poisson_variant == "pp"
    nso.navier_stokes_solver(1, dt)
    nsg.navier_stokes_solver(1, dt)
    _compare(nso, nsg, 1e-12)
    t, step = dt, 1
    while t < 0.3:
        step += 1
        t += dt
        nso.navier_stokes_solver(step, dt)
        nsg.navier_stokes_solver(step, dt)
    assert step == 250
    _compare(nso, nsg, 1e-9)
    # the reference's own check: compare u with the analytic solution (postpro.py:48)
    d = Go.delta
    i = np.arange(1, n + 1)[:, None]; j = np.arange(1, n + 1)[None, :]
    sol = -np.cos(i * d) * np.sin((j - 0.5) * d) * np.exp(-0.6)
    assert np.abs(nsg.v.x.I[:, :, 0] - sol).max() < 5e-3
    Gg.destroy()


def test_channel_ppn_steps_match_oracle():
    """BASELINE config 3 shape in miniature: walls in z, body force, ppn Poisson."""
    n = (32, 32, 16)
    bc = ["Periodic"] * 4 + ["Wall", "Wall"]
    Go, Gg, nso, nsg, dt = _setup_ns(n, bc, 3, 2.0, 0.05, fo.init_channel, 1.0, g=(1.0, 0.0, 0.0), cfl=0.05)
    assert nsg.poisson_variant == "ppn"
    nso.navier_stokes_solver(1, dt)
    nsg.navier_stokes_solver(1, dt)
    _compare(nso, nsg, 1e-12)
    for step in range(2, 41):
        nso.navier_stokes_solver(step, dt)
        nsg.navier_stokes_solver(step, dt)
    _compare(nso, nsg, 1e-12)
    assert abs(nsg.maxdiv) < 1e-12
    md, mc = nsg.status()
    assert abs(mc - nso.maxCFL) < 1e-12 and abs(md - nso.maxdiv) < 1e-12
    assert _unfused_checks(nsg, dt) == (md, mc)        # walls in z inside the fused correction + checks kernel
    Gg.destroy()


def test_poiseuille_pn_steps_match_oracle():
    """test/small_test/navier_stokes/poiseuille: nx = 4, walls in y, g(1) = 1, pn Poisson."""
    ny = 16
    Lx = 1.0 * fo._f32(4) / fo._f32(ny)
    Go = fo.Grid(4, ny, 1, Lx, 1.0, 1.0 / ny, bc=["Periodic", "Periodic", "Wall", "Wall"])
    Gg = fb.grid().setup(4, ny, 1, Lx, 1.0, 1.0 / ny, bc=["Periodic", "Periodic", "Wall", "Wall"])
    nso = fo.NavierStokes(Go)
    nsg = fb.Solver(Gg).init_solver()
    nso.g[0] = 1.0
    nsg.g = [1.0, 0.0, 0.0]
    dt = nso.set_timestep(1.0)
    assert nsg.set_timestep(1.0) == dt
    for step in range(1, 51):
        nso.navier_stokes_solver(step, dt)
        nsg.navier_stokes_solver(step, dt)
    _compare(nso, nsg, 1e-10)
    Gg.destroy()


def test_general_property_path_matches_uniform_and_oracle():
    """rho, mu pushed as (non-uniform) fields and a source S: the general kernels (hazard H11)."""
    n = (32, 32, 32)
    Go, Gg, nso, nsg, dt = _setup_ns(n, ["Periodic"] * 6, 3, 2 * PI, 0.05, fo.init_tgv3d, 1.0)
    x = Go.x[:, None, None]; y = Go.y[None, :, None]; z = Go.z[None, None, :]
    nso.rho.f[...] = 1.0 + 0.2 * np.sin(x) * np.cos(y) * np.cos(z)
    nso.mu.f[...] = 0.05 * (1.0 + 0.3 * np.cos(x) * np.sin(z))
    nso.S.x.f[...] = 0.1 * np.sin(y[:, 1:-1, :]) * np.ones((n[0], 1, n[2]))
    for a, b in ((nso.rho, nsg.rho), (nso.mu, nsg.mu), (nso.S.x, nsg.S.x), (nso.S.y, nsg.S.y), (nso.S.z, nsg.S.z)):
        b.f[...] = a.f
        b.push()
    for step in (1, 2, 3):
        nso.navier_stokes_solver(step, dt)
        nsg.navier_stokes_solver(step, dt)
    _compare(nso, nsg, 1e-11)
    Gg.destroy()


def test_constant_cfl_timestep_control():
    Go, Gg, nso, nsg, dt = _setup_ns((32, 32, 32), ["Periodic"] * 6, 3, 2 * PI, 0.01, fo.init_tgv3d, 1.0)
    nso.constant_CFL = True
    nsg.constant_CFL = True
    nso.CFL = 0.5
    nsg.CFL = 0.5
    dto, dtg = dt, dt
    for step in range(1, 6):
        dto = nso.navier_stokes_solver(step, dto)
        dtg = nsg.navier_stokes_solver(step, dtg)
        assert abs(dtg - dto) <= 1e-13 * dto
    _compare(nso, nsg, 1e-11)
    Gg.destroy()


def test_advection_operator_matches_oracle():
    """test/small_test/navier_stokes/advection/advection.f90 calls add_advection directly."""
    Go, Gg, nso, nsg, dt = _setup_ns((32, 32, 1), ["Periodic"] * 4, 2, 2 * PI, 1.0, fo.init_tgv2d, 2.0)
    ro, rg = fo.Vector(Go, 0), fb.vector(Gg, 0)
    nso.add_advection(ro)
    nsg.add_advection(rg)
    rg.pull()
    for a, b in zip(rg.comps, ro.comps):
        assert rel_l2(a.I, b.I) < 1e-13
    Gg.destroy()


def test_status_line_and_errors():
    Gg = fb.grid().setup(16, 16, 16, 1.0, 1.0, 1.0)
    ns = fb.Solver(Gg)
    with pytest.raises(fb.FenError):            # step before init_solver
        ns.navier_stokes_solver(1, 1e-3)
    ns.init_solver()
    with pytest.raises(fb.FenError):            # dt_o unset: set_timestep not called
        ns.navier_stokes_solver(1, 1e-3)
    dt = ns.set_timestep(1.0)
    ns.navier_stokes_solver(1, dt)
    line = ns.print_solver_status(1, dt, dt)
    assert line.startswith("step:       1 time: ") and "maxdiv:" in line and "maxCFL:" in line
    Gg.destroy()


# ---- BASELINE full size: size-independent properties -------------------------------------------
def test_full_size_512_properties():
    """BASELINE configs[1] at its full size (512^3, ppp): checked through properties that need no oracle run --
    the discrete Laplacian of the Poisson solution gives the right-hand side back, the solve is linear, a
    Navier-Stokes step leaves a divergence-free field and conserves momentum in the periodic box."""
    n = 512
    Gg = fb.grid().setup(n, n, n, 2 * PI, 2 * PI, 2 * PI)
    d = Gg.delta
    ns = fb.Solver(Gg, 1.0, 0.01).init_solver()
    phi = ns.phi
    rng = np.random.default_rng(2026)
    rhs = rng.standard_normal((n, n, n))
    rhs -= rhs.mean()
    phi.I[...] = rhs
    phi.push()
    ps = fb.PoissonSolver(phi)
    assert ps.variant == "ppp"
    ps.solve(phi)
    phi.pull()
    sol = phi.I.copy()
    lap = -6.0 * sol
    for ax in range(3):
        lap += np.roll(sol, 1, axis=ax) + np.roll(sol, -1, axis=ax)
    lap /= d * d
    assert np.linalg.norm(lap - rhs) <= 1e-10 * np.linalg.norm(rhs)
    # linearity: solve(-2.5 rhs) == -2.5 solve(rhs)
    phi.I[...] = -2.5 * rhs
    phi.push(); ps.solve(phi); phi.pull()
    assert np.linalg.norm(phi.I + 2.5 * sol) <= 1e-13 * np.linalg.norm(2.5 * sol)
    del lap, sol, rhs
    # three steps of the Taylor-Green case
    i = np.arange(1, n + 1, dtype=np.float64)
    sx, cxh = np.sin(i * d), np.cos((i - 0.5) * d)
    for kk in range(n):
        cz = np.cos((kk + 0.5) * d)
        ns.v.x.I[:, :, kk] = (sx[:, None] * cxh[None, :]) * cz
        ns.v.y.I[:, :, kk] = (-cxh[:, None] * sx[None, :]) * cz
    ns.v.z.I[...] = 0.0
    ns.p.I[...] = 0.0
    ns.v.push(); ns.p.push()
    ns.v.update_ghost_nodes(); ns.p.update_ghost_nodes()
    ns.CFL = 0.25
    dt = ns.set_timestep(1.0)
    mom0 = [float(c.I.sum()) for c in ns.v.comps]
    ke0 = sum(float((c.I ** 2).sum()) for c in ns.v.comps)
    for s in range(1, 4):
        ns.navier_stokes_solver(s, dt)
    md, mc = ns.status()
    assert abs(md) < 1e-12 and 0.0 < mc < 0.3
    assert _unfused_checks(ns, dt) == (md, mc)
    ns.v.pull()
    scale = float(np.abs(ns.v.x.I).sum())
    for c, m0 in zip(ns.v.comps, mom0):
        assert abs(float(c.I.sum()) - m0) <= 1e-9 * scale           # momentum is conserved in the periodic box
    ke1 = sum(float((c.I ** 2).sum()) for c in ns.v.comps)
    assert 0.99 * ke0 < ke1 < ke0                                   # viscous decay, no blow-up
    Gg.destroy()


@pytest.mark.parametrize("n", [64, 512])
def test_one_step_512_matches_c_oracle(n):
    """BASELINE configs[1] at its FULL size (and at 64^3, the size the CPU logic check runs) against the oracle: one navier_stokes_solver step of a 512^3 periodic box
    (Taylor-Green u, v, p plus an O(1) w so that every component carries signal; the projection removes its divergence)
    on the GPU and by the plain-C restatement (oracle/fen_oracle_c.c, ~1 s per step on the box's host cores), north_star's
    bound: relative L2 <= 1e-12 on u, v, w and p, divergence at machine precision."""
    from oracle import fen_oracle_c as foc
    Gg = fb.grid().setup(n, n, n, 2 * PI, 2 * PI, 2 * PI)
    delta = Gg.delta
    assert delta == 2 * PI / float(np.float32(n))
    i = np.arange(0, n + 2, dtype=np.float64)
    s_f, c_c = np.sin(i * delta), np.cos((i - 0.5) * delta)
    c2 = np.cos(2.0 * (i - 0.5) * delta)
    ns = fb.Solver(Gg, 1.0, 0.01).init_solver()
    ns.CFL = 0.25
    dt = ns.set_timestep(1.0)
    co = foc.NavierStokesC(n, n, n, delta, 1.0, 0.01)
    co.dt_o = dt
    # fields are built once, plane by plane, straight into the GPU side's host arrays (ghosts analytic: periodic)
    for kk in range(n + 2):
        ns.v.x.f[:, :, kk] = (s_f[:, None] * c_c[None, :]) * c_c[kk]
        ns.v.y.f[:, :, kk] = -(c_c[:, None] * s_f[None, :]) * c_c[kk]
        ns.v.z.f[:, :, kk] = 0.5 * (c_c[:, None] * c_c[None, :]) * s_f[kk]
        ns.p.f[:, :, kk] = (1.0 / 16.0) * (c2[:, None] + c2[None, :]) * (c2[kk] + 2.0)
    for fid, a in ((foc.U, ns.v.x), (foc.V, ns.v.y), (foc.W, ns.v.z), (foc.P, ns.p)):
        co.set(fid, a.f)
        a.push()
    ns.navier_stokes_solver(1, dt)
    co.navier_stokes_solver(1, dt)
    md, mc = ns.status()
    assert abs(md) < 1e-11 and abs(co.maxdiv) < 1e-11
    errs = {}
    for name, fid, a in (("u", foc.U, ns.v.x), ("v", foc.V, ns.v.y), ("w", foc.W, ns.v.z), ("p", foc.P, ns.p)):
        a.pull()
        ref = co.get(fid)[1:-1, 1:-1, 1:-1]
        errs[name] = rel_l2(a.I, ref)
        del ref
    assert max(errs.values()) <= 1e-12, errs
    assert abs(mc - co.maxCFL(dt)) <= 1e-12 * mc
    co.destroy()
    Gg.destroy()


def test_staged_x_passes_match_oracle():
    """FEN_X_R2C=1 FEN_X_C2R=1: the block-wide staged x kernels (what every x length below 128 uses) at the sizes where
    the row-private kernels are the default -- the cross-check switch of poisson.cu: launch_x -- in a child process,
    because the switches are read once per process."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FEN_X_R2C="1", FEN_X_C2R="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "tests/test_gpu_parity.py", "-k",
                        "test_one_step_tgv3d_matches_oracle or test_poisson_variants_match_oracle or "
                        "test_steps_tgv2d_matches_oracle_config1"], cwd=root, env=env, capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("n", [(64, 64, 64), (128, 64, 32)])
def test_poisson_ppp_matches_cufft(n):
    """north_star: cuFFT only as a correctness (and, in bench.py's `extra.cufft_crosscheck`, performance) cross-check of the
    hand-written transforms.  The ppp solve of a random right-hand side against the same spectral solve done with
    torch.fft (cuFFT): forward 3-D transform, division by the modified wavenumbers of the discrete Laplacian
    (poisson.f90:627-629, :998-1001, singular mode zeroed), inverse transform.  Library code is on this side of the
    comparison only; the product path never calls it (tests/test_host_logic.py checks the .so does not link cuFFT)."""
    import torch
    L = (2 * PI, 2 * PI * n[1] / n[0], 2 * PI * n[2] / n[0])
    Gg = fb.grid().setup(n[0], n[1], n[2], *L)
    d = Gg.delta
    phi = fb.scalar(Gg, 1)
    rng = np.random.default_rng(23)
    rhs = rng.standard_normal(n)
    rhs -= rhs.mean()
    phi.I[...] = rhs
    phi.push()
    ps = fb.PoissonSolver(phi)
    assert ps.variant == "ppp"
    ps.solve(phi)
    phi.pull()
    t = torch.tensor(rhs, device="cuda", dtype=torch.float64)
    R = torch.fft.fftn(t)
    lam = [torch.tensor(2.0 * (np.cos(2.0 * PI * np.arange(m) / m) - 1.0) / (d * d), device="cuda") for m in n]
    lam3 = lam[0][:, None, None] + lam[1][None, :, None] + lam[2][None, None, :]
    lam3[0, 0, 0] = 1.0
    R = R / lam3
    R[0, 0, 0] = 0.0
    ref = torch.fft.ifftn(R).real.cpu().numpy()
    assert rel_l2(phi.I, ref) <= 1e-12
    Gg.destroy()


def test_c_driver_runs():
    """examples/tgv_driver.c: the 2-D Taylor-Green case driven from plain C through the C ABI; the error against the
    analytic solution is the second-order one the reference's test plots (postpro.py:48-62)."""
    import re
    import subprocess
    from tests.test_host_logic import _build_c_driver
    exe = _build_c_driver()
    errs = []
    for n in (32, 64):
        r = subprocess.run([exe, str(n), "40"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        assert "maxdiv" in r.stdout
        errs.append(float(re.search(r"u_exact\| after 40 steps: ([0-9.eE+-]+)", r.stdout).group(1)))
    # the oracle gives 2.4544e-3 and 1.0255e-3 for these two runs (dt = 0.125 delta^2, so t differs with N)
    assert abs(errs[0] - 2.4544e-3) < 2e-6 and abs(errs[1] - 1.0255e-3) < 2e-6, errs


def test_cpp_driver_runs():
    """examples/tgv_driver.cpp: the same Taylor-Green case through include/fen_gpu.hpp, the C++ mirror of the reference's
    API; same error figures as the C driver (the oracle gives 2.4544e-3 and 1.0255e-3)."""
    import re
    import subprocess
    from tests.test_host_logic import _build_cpp_driver
    exe = _build_cpp_driver()
    errs = []
    for n in (32, 64):
        r = subprocess.run([exe, str(n), "40"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        assert "maxdiv" in r.stdout and "poisson variant pp" in r.stdout
        errs.append(float(re.search(r"u_exact\| after 40 steps: ([0-9.eE+-]+)", r.stdout).group(1)))
    assert abs(errs[0] - 2.4544e-3) < 2e-6 and abs(errs[1] - 1.0255e-3) < 2e-6, errs
