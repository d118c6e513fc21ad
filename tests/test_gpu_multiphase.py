"""GPU parity tests of the two-phase path: MTHINC volume of fluid + one-fluid Navier-Stokes
through the C ABI of libfen_gpu.so against oracle/fen_oracle_mf.py.  The tests read like the reference's own drivers
(test/small_test/volume_of_fluid/reversed, .../multiphase/viscous_decay, .../capillary_wave).

Tolerances.  The VoF cell arithmetic goes through exp / log / cosh / tanh, whose device and host implementations
differ in the last bit; tests/test_vof_math_host.py shows the reference's formulas keep that at ~3e-15 on vof after
20 advections, so vof is held to 1e-12 (absolute; vof is O(1)) and u, v, p to BASELINE's 1e-12 relative L2 after one
step.  Index work (boundary types, x_first, ghost cells of copies) is bit-exact."""
import math
import os

import numpy as np
import pytest

import fen_b200 as fb
from oracle import fen_oracle as fo
from oracle import fen_oracle_mf as mf

pytestmark = pytest.mark.gpu
PI = fo.PI


def rel_l2(a, b):
    n = np.linalg.norm(b.ravel())
    return np.linalg.norm((a - b).ravel()) / (n if n > 0 else 1.0)


def circle(x, y):            # test/small_test/volume_of_fluid/reversed/reversed.f90:100-113
    x0, y0, r = 0.5 * PI, 0.2 * (PI + 1.0), 0.2 * PI
    return np.sqrt((x - x0) ** 2 + (y - y0) ** 2) - r


def vortex(G, v):            # reversed.f90:117-140
    N = G.Nx
    i = np.arange(1, N + 1)[:, None]
    j = np.arange(1, G.Ny + 1)[None, :]
    d = G.delta
    v.x.I[..., 0] = np.sin(i * d) * np.cos((j - 0.5) * d)
    v.y.I[..., 0] = -np.cos((i - 0.5) * d) * np.sin(j * d)


def vof_pair(N, bc=None):
    Go = fo.Grid(N, N, 1, PI, PI, PI / N, bc=bc)
    Gg = fb.grid().setup(N, N, 1, PI, PI, PI / N, bc=bc)
    vo = mf.VoF(Go)
    vg = fb.VoF(Gg)
    vo.distance = circle
    vo.get_vof_from_distance()
    vg.get_vof_from_distance(lambda x, y: float(circle(x, y)))
    uo = fo.Vector(Go, 1)
    ug = fb.vector(Gg, 1)
    vortex(Go, uo)
    uo.update_ghost_nodes()
    for a, b in zip(ug.comps, uo.comps):
        a.f[...] = b.f
        a.push()
    return Go, Gg, vo, vg, uo, ug


def test_get_vof_from_distance_and_reconstruction():
    """reconstruction.f90: get_vof_from_distance + get_h_from_vof, every field incl. ghosts."""
    Go, Gg, vo, vg, _, _ = vof_pair(32)
    vg.vof.pull(); vg.h.pull()
    assert np.abs(vg.vof.f - vo.vof.f).max() < 1e-14
    assert np.abs(vg.h.f - vo.h.f).max() < 1e-14
    # the normals divide by sqrt(mx^2 + my^2 + 1e-14): where the profile is saturated (gradients ~ 1e-8) a last-bit
    # difference of vof (device tanh vs libm tanh) moves them by 1e-8, so the reconstruction is compared on
    # bit-identical input
    vg.vof.f[...] = vo.vof.f
    vg.vof.push()
    vo.get_h_from_vof()
    vg.get_h_from_vof()
    f = vo.vof.f
    for name, a, b in (("nx", vg.norm.x, vo.norm.x), ("ny", vg.norm.y, vo.norm.y), ("lx", vg.l.x, vo.l.x),
                       ("ly", vg.l.y, vo.l.y), ("h", vg.h, vo.h)):
        a.pull()
        assert np.abs(a.f - b.f).max() < 1e-12, name
    vg.curv.pull(); vg.d.pull()
    assert np.abs(vg.curv.f - vo.curv.f).max() < 1e-12 * (1.0 + np.abs(vo.curv.f).max())
    # d loses digits where the profile is saturated (see tests/test_vof_math_host.py): sensitivity-weighted
    assert (np.abs(vg.d.f - vo.d.f) * f * (1.0 - f)).max() < 1e-12
    Gg.destroy()


@pytest.mark.parametrize("bc", [None, ["Periodic", "Periodic", "Wall", "Wall"], ["Wall"] * 4])
def test_advect_vof_matches_oracle(bc):
    """advect_vof with a prescribed vortex (reversed.f90), both sweep orders, walls and the boundary-type switch of
    `vof = vof1` (hazard H13)."""
    Go, Gg, vo, vg, uo, ug = vof_pair(48, bc)
    dt = 0.4 * Go.delta
    m0 = vg.check_vof_integral()
    for step in range(1, 13):
        vo.advect_vof(uo, dt)
        vg.advect_vof(ug, dt)
        if step in (1, 2, 12):
            vg.vof.pull()
            assert np.abs(vg.vof.f - vo.vof.f).max() < 1e-12, step
            assert vg.x_first == vo.x_first
            assert [vg.vof.get_bc_type(f) for f in fo.FACES[:4]] == [vo.vof.bc_type[f] for f in fo.FACES[:4]]
    m1 = vg.check_vof_integral()
    o1 = vo.check_vof_integral()
    assert abs(m1[0] - o1[0]) < 1e-12 * abs(o1[0]) and abs(m1[1] - o1[1]) < 1e-12 * abs(o1[1])
    if bc is None:
        assert abs(m1[0] - m0[0]) < 1e-12 * abs(m0[0])          # the divergence-free vortex conserves the phase volume
    Gg.destroy()


def test_reversed_vortex_returns():
    """The property the reference's reversed test plots (postpro.py: contours of the first, middle and last file):
    after the flow reversal the drop comes back -- checked on the GPU path alone at the reference's time step."""
    N = 64
    Gg = fb.grid().setup(N, N, 1, PI, PI, PI / N)
    vg = fb.VoF(Gg)
    vg.get_vof_from_distance(lambda x, y: float(circle(x, y)))
    ug = fb.vector(Gg, 1)
    vortex(Gg, ug)
    ug.push(); ug.update_ghost_nodes()
    vg.vof.pull()
    f0 = vg.vof.I.copy()
    m0 = vg.check_vof_integral()[0]
    dt = 0.00125 * PI * 200 / N
    nstep = int(2 * PI / dt)
    for step in range(1, nstep + 1):
        vg.advect_vof(ug, dt)
        if step == nstep // 2:
            ug.pull()
            for s in ug.comps:
                s.f *= -1.0
                s.push()
    vg.vof.pull()
    m1 = vg.check_vof_integral()[0]
    assert abs(m1 - m0) < 1e-11 * m0
    err = np.abs(vg.vof.I - f0).sum() * Gg.delta ** 2
    assert err < 0.03 * (PI * (0.2 * PI) ** 2)              # L1 shape error below 3 % of the drop area (oracle: 1.9 %)
    assert vg.vof.I.min() > -1e-6 and vg.vof.I.max() < 1.0 + 1e-6
    Gg.destroy()


# ---- the full two-phase step ------------------------------------------------------------------------------------
def wave_case(Nx, Ny, sigma=0.0, walls=True, beta=1.0, constant_CFL=False):
    """A gravity / capillary wave between two fluids of density ratio 850 (viscous_decay.f90 with a larger amplitude
    so that every term is exercised after a few steps).  walls=False: fully periodic box with an off-centre light
    drop instead (a wave would wrap into a one-cell density jump at the y boundary, and a drop centred on the grid has
    |n_x| == |n_y| exactly on its diagonals, where the reference's x/y-dominant branch flips on a last-bit difference:
    both amplify 1e-16 perturbations of the ORACLE's own input to 1e-5 within five steps)."""
    Lx, Ly = 1.0, float(Ny) / Nx
    bc = ["Periodic", "Periodic", "Wall", "Wall"] if walls else None
    Go = fo.Grid(Nx, Ny, 1, Lx, Ly, Lx / Nx, bc=bc)
    Gg = fb.grid().setup(Nx, Ny, 1, Lx, Ly, Lx / Nx, bc=bc)

    def wave(x, y):
        if not walls:
            return 0.2317 - np.sqrt((x - 0.4631) ** 2 + (y - 0.9173) ** 2)
        return y - 0.05 * np.cos(2.0 * PI * x / Lx) - Ly / 2.0
    rho_0 = 1000.0
    rho_1 = rho_0 / 850.0
    mu_0 = rho_0 * Lx * math.sqrt(mf.GRAVITY * Lx) / 1.0e4
    mu_1 = mu_0 * 1.9e-2
    ons = mf.MultiphaseNavierStokes(Go, rho_0, rho_1, mu_0, mu_1, sigma, distance=wave, beta=beta)
    ons.g[1] = -mf.GRAVITY
    gns = fb.MultiphaseSolver(Gg)
    gns.rho_0, gns.rho_1, gns.mu_0, gns.mu_1, gns.sigma, gns.beta = rho_0, rho_1, mu_0, mu_1, sigma, beta
    gns.g = [0.0, -mf.GRAVITY, 0.0]
    gns.init_solver(lambda x, y: float(wave(x, y)))
    odt = ons.set_timestep(1.0)
    gdt = gns.set_timestep(1.0)
    assert gdt == odt
    odt = 0.1 * odt          # viscous_decay.f90:56-57: dt is scaled after set_timestep, dt_o keeps the unscaled value
    # init_velocity of viscous_decay.f90:104-131
    wn = 2.0 * PI / Lx
    om = math.sqrt(mf.GRAVITY * wn)
    i = np.arange(1, Nx + 1)[:, None]
    j = np.arange(1, Ny + 1)[None, :]
    d = Go.delta
    F = ons.vof.sh
    x = i * d
    y = (j - 0.5) * d - Ly / 2.0
    f = ((F(1, 0) + F()) * 0.5)[..., 0]
    ons.v.x.I[..., 0] = (1.0 - f) * 0.05 * om * np.exp(wn * y) * np.cos(wn * x) - f * 0.05 * om * np.exp(-wn * y) * np.cos(wn * x)
    x = (i - 0.5) * d
    y = j * d - Ly / 2.0
    f = ((F(0, 1) + F()) * 0.5)[..., 0]
    ons.v.y.I[..., 0] = (1.0 - f) * 0.05 * om * np.exp(wn * y) * np.sin(wn * x) + f * 0.05 * om * np.exp(-wn * y) * np.sin(wn * x)
    if not walls:       # a solenoidal vortex array for the drop
        kx, ky = 2.0 * PI / Lx, 2.0 * PI / Ly
        ons.v.x.I[..., 0] = 0.2 * np.sin(kx * i * d) * np.cos(ky * (j - 0.5) * d)
        ons.v.y.I[..., 0] = -0.2 * (kx / ky) * np.cos(kx * (i - 0.5) * d) * np.sin(ky * j * d)
    ons.v.update_ghost_nodes()
    for a, b in zip(gns.v.comps, ons.v.comps):
        a.f[...] = b.f
        a.push()
    ons.constant_CFL = constant_CFL
    gns.constant_CFL = constant_CFL
    return Go, Gg, ons, gns, odt


def compare_state(gns, ons, tol, tag=""):
    gns.v.pull(); gns.p.pull(); gns.vof.pull(); gns.rho.pull(); gns.mu.pull(); gns.p_o.pull()
    errs = {"u": rel_l2(gns.v.x.I, ons.v.x.I), "v": rel_l2(gns.v.y.I, ons.v.y.I), "p": rel_l2(gns.p.I, ons.p.I),
            "rho": rel_l2(gns.rho.I, ons.rho.I), "mu": rel_l2(gns.mu.I, ons.mu.I),
            "vof": float(np.abs(gns.vof.I - ons.vof.I).max()), "p_o": rel_l2(gns.p_o.I, ons.p_o.I)}
    assert max(errs.values()) < tol, (tag, errs)
    return errs


def test_init_solver_mf_matches_oracle():
    Go, Gg, ons, gns, dt = wave_case(32, 64)
    assert gns.poisson_variant == "pn"
    assert gns.rhomin == ons.rhomin and gns.irhomin == ons.irhomin
    for a, b in ((gns.vof, ons.vof), (gns.rho, ons.rho), (gns.mu, ons.mu)):
        a.pull()        # plane 1 = the 2-D field with its x / y ghosts (the two z ghost planes of a 2-D array are never used)
        assert np.abs(a.f[:, :, 1] - b.f[:, :, 1]).max() < 1e-13 * np.abs(b.f).max()
    for face in fo.FACES[:4]:
        for a, b in ((gns.vof, ons.vof), (gns.p_hat, ons.p_hat), (gns.p_o, ons.p_o), (gns.curv, ons.vf.curv),
                     (gns.norm.x, ons.vf.norm.x), (gns.l.y, ons.vf.l.y), (gns.rho, ons.rho)):
            assert a.get_bc_type(face) == b.bc_type[face]
    Gg.destroy()


@pytest.mark.parametrize("sigma,walls", [(0.0, True), (0.07, True), (0.07, False)])
def test_mf_step_matches_oracle(sigma, walls):
    """One and five two-phase steps: rel-L2 <= 1e-12 on u, v, p after one step (BASELINE's tolerance), 1e-10 after
    five; divergence at machine precision."""
    Go, Gg, ons, gns, dt = wave_case(32, 64, sigma=sigma, walls=walls)
    ons.navier_stokes_solver(1, dt)
    gns.navier_stokes_solver(1, dt)
    compare_state(gns, ons, 1e-12, "step 1")
    md, mc = gns.status()
    assert abs(md - ons.maxdiv) < 1e-11 and abs(mc - ons.maxCFL) < 1e-12 * ons.maxCFL
    for s in range(2, 6):
        ons.navier_stokes_solver(s, dt)
        gns.navier_stokes_solver(s, dt)
    compare_state(gns, ons, 1e-10, "step 5")
    md, _ = gns.status()
    assert abs(md) < 1e-9        # |div| ~ eps * |u| / delta * rho ratio; the oracle's own value is of this size
    assert abs(md - ons.maxdiv) < 1e-9
    Gg.destroy()


def test_mf_constant_cfl_and_source():
    """constant_CFL (update_timestep with dt_surf, the extrapolated p_hat of navier_stokes.f90:89) and a body-force
    field S."""
    Go, Gg, ons, gns, dt = wave_case(32, 64, sigma=0.07, constant_CFL=True)
    ons.CFL = 0.3
    gns.CFL = 0.3
    rng = np.random.default_rng(3)
    for a, b in zip(gns.S.comps, ons.S.comps):
        b.I[...] = rng.standard_normal(b.I.shape) * 10.0
        a.f[...] = b.f
        a.push()
    odt = gdt = dt
    for s in range(1, 5):
        odt = ons.navier_stokes_solver(s, odt)
        gdt = gns.navier_stokes_solver(s, gdt)
        assert abs(gdt - odt) <= 1e-12 * odt, (s, gdt, odt)
    compare_state(gns, ons, 1e-10, "constant CFL")
    Gg.destroy()


def test_mf_graph_replay_matches_eager():
    """The captured CUDA graph of the two-phase step (x_first alternates, so two graphs) against eager launches:
    bit-identical fields after 9 steps."""
    out = []
    for no_graph in (False, True):
        Go, Gg, ons, gns, dt = wave_case(32, 64, sigma=0.07)
        if no_graph:
            gns.profile(True)            # per-kernel profiling forces eager launches (context.cu: graph_ok)
        for s in range(1, 10):
            gns.navier_stokes_solver(s, dt)
        gns.v.pull(); gns.p.pull(); gns.vof.pull()
        out.append([gns.v.x.f.copy(), gns.v.y.f.copy(), gns.p.f.copy(), gns.vof.f.copy(), gns.x_first,
                    gns.launch_count()])
        Gg.destroy()
    for a, b in zip(out[0][:4], out[1][:4]):
        assert np.array_equal(a, b)
    assert out[0][4] == out[1][4]
    assert out[0][5] == out[1][5]            # same number of kernels either way


def test_capillary_wave_tracks_prosperetti():
    """test/small_test/multiphase/capillary_wave at Nx = 16 on the GPU path: the maximum interface amplitude against
    the reference's Prosperetti data (tests/golden/prosperetti_capillary.npz, sub-sampled from the reference's
    prosperetti.csv), error below the N^-1 guide line of the reference's postpro.py (0.4 / N)."""
    import os
    pros = np.load(os.path.join(os.path.dirname(__file__), "golden", "prosperetti_capillary.npz"))["curve"]
    a, lam = 0.01, 1.0
    wn = 2.0 * PI / lam
    Nx = 16
    Ny = 3 * Nx
    Gg = fb.grid().setup(Nx, Ny, 1, lam, 3.0 * lam, lam / Nx, x0=(0.0, -1.5 * lam, 0.0),
                         bc=["Periodic", "Periodic", "Wall", "Wall"])
    dl = Gg.delta

    def wave(x, y):          # capillary.f90:117-135
        x1 = x - dl / 2.0; y1 = a * math.cos(wn * x1)
        x2 = x + dl / 2.0; y2 = a * math.cos(wn * x2)
        return -((x2 - x1) * (y1 - y) - (x1 - x) * (y2 - y1)) / math.sqrt((x2 - x1) ** 2 + (y2 - y1) ** 2)
    ns = fb.MultiphaseSolver(Gg)
    ns.rho_0 = ns.rho_1 = 1.0
    ns.mu_0 = ns.mu_1 = 0.0182571749236
    ns.sigma = 1.0
    ns.init_solver(wave)
    dt = ns.set_timestep(1.0)
    Tmax = 25.0 / 11.1366559937
    nprint = int(Tmax / 80 / dt)
    omega0 = math.sqrt(1.0 * wn ** 3 / 2.0)
    Y = Gg.y[1:Ny + 1]
    t, step, L1 = 0.0, 0, 0.0
    while t < Tmax:
        step += 1
        t += dt
        ns.navier_stokes_solver(step, dt)
        if step % nprint == 0:
            ns.vof.pull()
            vof = ns.vof.I[:, :, 0]
            amp = np.array([np.interp(0.5, vof[i, :], Y) for i in range(Nx)])
            L1 = max(L1, abs(np.abs(amp).max() - np.interp(t * omega0, pros[:, 0], pros[:, 1])))
    assert L1 < 0.4 / Nx, L1
    assert abs(ns.maxdiv) < 1e-12
    Gg.destroy()


def test_gpu_reproduces_mf_golden():
    """The committed two-phase fixture (tests/golden/mf_wave_16x32_3steps.npz) without the oracle in the loop: the
    initial vof, u, v are pushed (the route a driver takes when it restarts from files) and three steps compared."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mf_wave_16x32_3steps.npz"))
    Nx, Ny = [int(x) for x in g["n"]]
    Gg = fb.grid().setup(Nx, Ny, 1, 1.0, float(Ny) / Nx, 1.0 / Nx, bc=["Periodic", "Periodic", "Wall", "Wall"])
    ns = fb.MultiphaseSolver(Gg)
    ns.rho_0, ns.rho_1, ns.mu_0, ns.mu_1 = [float(x) for x in g["props"]]
    ns.sigma = float(g["sigma"])
    ns.g = [0.0, -mf.GRAVITY, 0.0]
    ns.init_solver(None)                       # no distance function: vof comes from the fixture
    ns.vof.f[...] = g["vof0"]
    ns.vof.push()
    ns.update_material_properties()
    ns.v.x.f[...] = g["u0"]; ns.v.y.f[...] = g["v0"]
    ns.v.push()
    dt = 0.1 * ns.set_timestep(1.0)
    assert dt == float(g["dt"])
    for s in range(1, int(g["steps"]) + 1):
        ns.navier_stokes_solver(s, dt)
    ns.v.pull(); ns.p.pull(); ns.vof.pull(); ns.rho.pull()
    for k, a in (("u", ns.v.x), ("v", ns.v.y), ("p", ns.p), ("vof", ns.vof), ("rho", ns.rho)):
        assert rel_l2(a.I, g[k][1:-1, 1:-1, 1:2]) < 1e-12, k
    md, mc = ns.status()
    assert abs(md) < 1e-11 and abs(mc - float(g["maxCFL"])) < 1e-12
    Gg.destroy()


@pytest.mark.parametrize("n", [(32, 48), (64, 3000)])
def test_poisson_pn_any_length_in_the_thomas_direction(n):
    """The wall-normal direction of the *n variants is solved by the Thomas algorithm (poisson.f90:306-412): no
    power-of-two or 2048-point limit there, as in the reference (the two-phase wave cases use ny = 2 nx, 3 nx)."""
    bc = ["Periodic", "Periodic", "Wall", "Wall"]
    Go = fo.Grid(n[0], n[1], 1, 1.0, float(n[1]) / n[0], 1.0 / n[0], bc=bc)
    Gg = fb.grid().setup(n[0], n[1], 1, 1.0, float(n[1]) / n[0], 1.0 / n[0], bc=bc)
    rng = np.random.default_rng(5)
    rhs = rng.standard_normal((n[0], n[1], 1))
    rhs -= rhs.mean()
    po = fo.Scalar(Go, 1)
    for face, t in zip(fo.FACES[:4], (0, 0, 2, 2)):
        po.bc_type[face] = t
    po.I[...] = rhs
    pso = fo.PoissonSolver(po)
    pso.solve(po)
    gns = fb.Solver(Gg).init_solver()
    assert gns.poisson_variant == "pn" == pso.variant
    gns.phi.f[...] = 0.0
    gns.phi.I[...] = rhs
    gns.phi.push()
    fb.api.check(Gg.lib.fen_gpu_solve_poisson(Gg.ctx, gns.phi.id))
    gns.phi.pull()
    assert rel_l2(gns.phi.I, po.I) < 1e-12
    Gg.destroy()


def test_mf_state_files(tmp_path):
    """save_state / load_state / save_fields of a -DMF build (solver.f90:201-209, 283-297, 145-148): the state file
    carries p, v_x, v_y, dv_o_x, dv_o_y, vof, rho, mu, p_o in that order, and loading it restores them bit for bit
    (test/small_test/io/test_MF.f90)."""
    Go, Gg, ons, gns, dt = wave_case(16, 32, sigma=0.07)
    for s in range(1, 4):
        gns.navier_stokes_solver(s, dt)
    d = str(tmp_path)
    path = gns.save_state(3, d)
    gns.save_fields(3, d)
    raw = np.fromfile(path).reshape((9, 32, 16))
    fields = (gns.p, gns.v.x, gns.v.y, gns.dv_o.x, gns.dv_o.y, gns.vof, gns.rho, gns.mu, gns.p_o)
    for f in fields:
        f.pull()
    for blk, f in zip(raw, fields):
        assert np.array_equal(blk.T, f.I[:, :, 0])
    assert np.array_equal(np.fromfile(os.path.join(d, "vof_0000003.raw")).reshape((16, 32), order="F"), gns.vof.I[:, :, 0])
    want = [f.I.copy() for f in fields]
    Gg.destroy()
    Go, Gg, ons, gns, dt = wave_case(16, 32, sigma=0.07)
    gns.load_state(3, d)
    fields = (gns.p, gns.v.x, gns.v.y, gns.dv_o.x, gns.dv_o.y, gns.vof, gns.rho, gns.mu, gns.p_o)
    for f, w in zip(fields, want):
        f.pull()
        assert np.array_equal(f.I, w)
    # vof's ghost cells come from its WIRED boundary types again: a restarted run starts with a fresh module state
    assert [gns.vof.get_bc_type(f) for f in fo.FACES[:4]] == [0, 0, 2, 2] and gns.x_first
    Gg.destroy()
