"""GPU parity of the any-length Poisson path (fen_b200/csrc/fft_any.cuh): grid sizes that are not powers of two.

The reference has no size restriction (FFTW plans of any n) and its own drivers use such grids: 96 x 96
(test/small_test/fsi/Pan_Eulerian/Pan.f90:33-34), 16 x 48 (test/small_test/io/test_MF.f90), 3072 x 4608
(test/large_test/startup_flow_cylinder/main.f90:37-38), 10^3 (test/small_test/fields/memory.f90).  Powers of two run
the tuned register-path kernels; every other length whose prime factors are <= 61 runs the mixed-radix kernels, whose
phases are executed on the CPU by tests/cpu/test_fft_any.cu (global indexing included).  Oracle: the numpy restatement,
which takes any n like FFTW does.

First run on a B200: the driver's round-1 GPU tier (19/19) and round 2's first session (profiles/r02a_pytest_firstrun.log)."""
import numpy as np
import pytest

import fen_b200 as fb
from oracle import fen_oracle as fo
from tests.test_gpu_parity import PI, _compare, _setup_ns, make_pair, rel_l2

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

P4, P6 = ["Periodic"] * 4, ["Periodic"] * 6
ANY_CASES = [
    ("pp", (96, 96, 1), P4, 2),                       # Pan.f90
    ("pp", (48, 20, 1), P4, 2),
    ("pp", (15, 9, 1), P4, 2),                        # odd lengths: no Nyquist mode
    ("pp", (64, 100, 1), P4, 2),                      # tuned x pass + any-length fused y solve
    ("pn", (12, 10, 1), P4[:2] + ["Wall", "Wall"], 2),
    ("pn", (96, 48, 1), P4[:2] + ["Wall", "Wall"], 2),
    ("nn", (48, 40, 1), ["Wall"] * 4, 2),
    ("nn", (10, 16, 1), ["Wall", "Wall", "Inflow", "Outflow"], 2),
    ("ppp", (24, 12, 20), P6, 3),
    ("ppp", (10, 10, 10), P6, 3),                     # memory.f90's grid
    ("ppp", (32, 6, 16), P6, 3),                      # only y goes through the any-length kernels
    ("ppp", (16, 8, 22), P6, 3),                      # only the fused z solve does (prime factor 11)
    ("pp", (74, 6, 1), P4, 2),                        # prime factor 37: the generic radix butterfly
    ("ppn", (12, 24, 10), P4 + ["Wall", "Wall"], 3),
    ("npn", (10, 6, 8), ["Wall", "Wall", "Periodic", "Periodic", "Wall", "Wall"], 3),
    ("nnn", (12, 20, 6), ["Wall"] * 6, 3),
    ("nnn", (16, 18, 8), ["Wall"] * 6, 3),            # tuned DCT in x, any-length DCT in y
    ("pp", (3072, 8, 1), P4, 2),                      # startup_flow_cylinder's x length: one row per block
    ("pn", (4608, 6, 1), P4[:2] + ["Wall", "Wall"], 2),
]


@pytest.mark.parametrize("variant,n,bc,ndim", ANY_CASES)
def test_poisson_any_length_matches_oracle(variant, n, bc, ndim):
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
    for _ in range(2):                                # twice: no state carried between solves
        pg.I[...] = rhs
        pg.push(); psg.solve(pg); pg.pull()
        assert rel_l2(pg.I, po.I) < 1e-12
    Gg.destroy()


def test_unsupported_lengths_are_rejected_loudly():
    """A prime factor above 61, or more than 6144 points, has no kernel: init_poisson_solver must say so (no silent
    fallback), as must an any-length grid on several ranks (tests/test_gpu_multirank.py keeps to powers of two)."""
    for n in ((134, 16, 1), (16, 2 * 67, 1)):
        Gg = fb.grid().setup(n[0], n[1], 1, 1.0, 1.0 * n[1] / n[0], 1.0, bc=P4, ndim=2)
        with pytest.raises(fb.FenError) as e:
            fb.PoissonSolver(fb.scalar(Gg, 1))
        assert "primes" in str(e.value)
        Gg.destroy()


def test_steps_tgv2d_96_match_oracle():
    """The 2-D Taylor-Green driver of BASELINE config 1 on Pan.f90's 96 x 96 grid: 1e-12 after one step, 1e-10 after
    forty, divergence at machine precision."""
    n = 96
    Go, Gg, nso, nsg, dt = _setup_ns((n, n, 1), P4, 2, 2 * PI, 1.0, fo.init_tgv2d, 2.0)
    assert nsg.poisson_variant == "pp"
    nso.navier_stokes_solver(1, dt)
    nsg.navier_stokes_solver(1, dt)
    _compare(nso, nsg, 1e-12)
    for step in range(2, 41):
        nso.navier_stokes_solver(step, dt)
        nsg.navier_stokes_solver(step, dt)
    _compare(nso, nsg, 1e-10)
    assert abs(nsg.maxdiv) < 1e-12
    Gg.destroy()


def test_steps_tgv3d_24x48x24_match_oracle():
    """The 3-D periodic step (BASELINE config 2's path) on a grid with a factor 3 in every direction: the fused right-hand side is
    off (power-of-two x only), so this also covers the separate divergence kernel in front of the any-length x pass."""
    n = (24, 48, 24)                  # box sides 2 pi x 4 pi x 2 pi: whole Taylor-Green wavelengths
    Go, Gg, nso, nsg, dt = _setup_ns(n, P6, 3, 2 * PI, 0.01, fo.init_tgv3d, 1.0)
    assert nsg.poisson_variant == "ppp"
    nso.navier_stokes_solver(1, dt)
    nsg.navier_stokes_solver(1, dt)
    _compare(nso, nsg, 1e-12)
    for step in range(2, 6):
        nso.navier_stokes_solver(step, dt)
        nsg.navier_stokes_solver(step, dt)
    _compare(nso, nsg, 1e-11)
    assert abs(nsg.maxdiv) < 1e-12
    Gg.destroy()


def test_steps_cavity_48x40_match_oracle():
    """Lid-driven cavity (nn: any-length DCT in x, Thomas in y) on 48 x 40."""
    bc = ["Wall"] * 4
    Go = fo.Grid(48, 40, 1, 1.2, 1.0, 1.2 / 48, bc=bc)
    Gg = fb.grid().setup(48, 40, 1, 1.2, 1.0, 1.2 / 48, bc=bc)
    nso = fo.NavierStokes(Go, 1.0, 1.0e-2)
    nsg = fb.Solver(Gg, 1.0, 1.0e-2).init_solver()
    assert nsg.poisson_variant == "nn"
    nso.v.x.bc["top"][...] = 1.0
    nsg.v.x.set_bc("top", 1.0)
    nso.CFL = nsg.CFL = 0.25
    dt = nso.set_timestep(1.0)
    assert nsg.set_timestep(1.0) == dt
    for step in range(1, 21):
        nso.navier_stokes_solver(step, dt)
        nsg.navier_stokes_solver(step, dt)
    _compare(nso, nsg, 1e-11)
    assert abs(nsg.maxdiv) < 1e-11 and np.abs(nsg.v.x.I).max() > 1e-3
    Gg.destroy()


@pytest.mark.parametrize("n,ndim", [((16, 12, 8), 3), ((32, 16, 1), 2)])
def test_scalar_laplacian_face_to_center_curl(n, ndim):
    """The operators of fields_mod's generic interfaces that the step itself does not use: laplacian(s, lap_s)
    (the reference's fields test, methods.f90:326), face_to_center (save_fields) and curl (lid_driven.f90:86).
    Same expressions in the same order as the oracle; the compiler contracts a*b + c into one fused operation where
    numpy rounds twice, hence 1e-13 on the Laplacian and the curl (as tests/test_gpu_parity.py::test_field_operators
    holds the vector Laplacian) and bit-exactness on the average."""
    Go, Gg = make_pair(n, ndim=ndim)
    rng = np.random.default_rng(9)
    so, sg = fo.Scalar(Go, 1), fb.scalar(Gg, 1)
    so.I[...] = rng.standard_normal(so.I.shape)
    so.update_ghost_nodes()
    sg.f[...] = so.f
    sg.push()
    lo, lg = fo.Scalar(Go, 0), fb.scalar(Gg, 0)
    fo.laplacian_scalar(so, lo)
    fb.laplacian(sg, lg)
    lg.pull()
    assert rel_l2(lg.I, lo.I) < 1e-13
    for face in "xyz"[:ndim]:
        co, cg = fo.Scalar(Go, 0), fb.scalar(Gg, 0)
        fo.face_to_center(so, co, face)
        fb.face_to_center(sg, cg, face)
        cg.pull()
        assert np.array_equal(cg.I, co.I), face
    vo, vg = fo.Vector(Go, 1), fb.vector(Gg, 1)
    for a, b in zip(vg.comps, vo.comps):
        b.I[...] = rng.standard_normal(b.I.shape)
    vo.update_ghost_nodes()
    for a, b in zip(vg.comps, vo.comps):
        a.f[...] = b.f
        a.push()
    wo, wg = fo.Vector(Go, 0), fb.vector(Gg, 0)
    fo.curl(vo, wo)
    fb.curl(vg, wg)
    wg.pull()
    for a, b in zip(wg.comps[:3 if ndim == 3 else 1], wo.comps):
        assert rel_l2(a.I, b.I) < 1e-13
    with pytest.raises(fb.FenError):
        fb.laplacian(sg, sg)                      # an output must not be an input
    Gg.destroy()


def test_set_from_function_fills_interior_and_ghosts():
    """scalar%set_from_function (scalar.f90:137-164): interior from the cell-centre coordinates, then the ghost update."""
    Go, Gg = make_pair((16, 8, 4))
    sg = fb.scalar(Gg, 1)
    sg.set_from_function(lambda x, a: a[0] * np.sin(2 * PI * x[0]) + x[1] - 2.0 * x[2], [3.0])
    so = fo.Scalar(Go, 1)
    X = Go.x[1:-1, None, None]; Y = Go.y[None, 1:-1, None]; Z = Go.z[None, None, 1:-1]
    so.I[...] = 3.0 * np.sin(2 * PI * X) + Y - 2.0 * Z
    so.update_ghost_nodes()
    assert np.abs(sg.f - so.f).max() < 1e-14
    sg.f[...] = 0.0
    sg.pull()
    assert np.abs(sg.f - so.f).max() < 1e-14
    Gg.destroy()


def test_poiseuille_inflow_outflow_steps_match_oracle():
    """test/small_test/navier_stokes/poiseuille_io: walls, Inflow with a parabolic boundary-value PLANE on the normal
    velocity, Outflow -- the nn variant with the Inflow/Outflow tridiagonal (poisson.f90:226-231).  Forty steps against
    the oracle, then on to the driver's steady state on the GPU alone and the reference's own comparison with the
    parabola at its coarsest resolution."""
    from tests.test_oracle import poiseuille_io_case
    nx = 8
    Go, nso, plane = poiseuille_io_case(nx, fo.Grid, fo.NavierStokes)
    Gg, nsg, _ = poiseuille_io_case(nx, lambda *a, **k: fb.grid().setup(*a, **k), lambda G: fb.Solver(G).init_solver())
    assert nsg.poisson_variant == "nn"
    for face in fo.FACES[:4]:
        for a, b in ((nsg.v.x, nso.v.x), (nsg.v.y, nso.v.y), (nsg.p, nso.p), (nsg.phi, nso.phi)):
            assert a.get_bc_type(face) == b.bc_type[face], face
    nso.v.y.bc["bottom"][...] = plane
    nsg.v.y.set_bc("bottom", plane)
    dt = nso.set_timestep(1.0)
    assert nsg.set_timestep(1.0) == dt
    for step in range(1, 41):
        nso.navier_stokes_solver(step, dt)
        nsg.navier_stokes_solver(step, dt)
        if step == 1:
            _compare(nso, nsg, 1e-12)
    _compare(nso, nsg, 1e-11)
    uo = nsg.v.y.f.copy()
    step = 40
    while True:
        step += 1
        nsg.navier_stokes_solver(step, dt)
        nsg.v.y.pull()
        if (nsg.v.y.f - uo).max() < 1e-8:
            break
        uo = nsg.v.y.f.copy()
        assert step < 2000
    Xf = (np.arange(nx) + 0.5) * (1.0 / nx)
    assert np.abs(nsg.v.y.I[:, 2, 0] + (Xf ** 2 - Xf) / 2.0).max() < 2e-3          # 1.1e-3 for the oracle at nx = 8
    Gg.destroy()


def test_no_device_memory_is_left_behind():
    """The reference's memory tests (test/small_test/grid, test/small_test/fields/memory.f90 on a 10^3 grid: valgrind,
    0 bytes in use at exit) in device terms: twenty rounds of grid set-up, field allocation, init_solver, a few steps,
    a Poisson solve, asynchronous pulls and destruction must give every byte back to the driver (the first round is the
    warm-up: module loading and the runtime's own pools are one-off)."""
    import torch

    def one_round(n):
        G = fb.grid().setup(n[0], n[1], n[2], 1.0, 1.0 * n[1] / n[0], 1.0 * n[2] / n[0])
        s = [fb.scalar(G, 1) for _ in range(3)]
        v = fb.vector(G, 1)
        ns = fb.Solver(G, 1.0, 0.1).init_solver()
        ns.v.x.f[...] = 0.1
        ns.v.push()
        ns.v.update_ghost_nodes()
        dt = ns.set_timestep(1.0)
        for step in range(1, 4):
            ns.navier_stokes_solver(step, dt)
        ns.status()
        ns.p.pull_async()
        G.pull_wait()
        fb.gradient(s[0], v)
        s[1].destroy()
        ns.destroy_solver()
        G.destroy()

    free = []
    for r in range(21):
        one_round((10, 10, 10) if r % 2 == 0 else (32, 16, 8))      # memory.f90's grid (any-length path) and a tuned one
        torch.cuda.synchronize()
        free.append(torch.cuda.mem_get_info()[0])
    assert free[-1] >= free[2] - (1 << 20), [f - free[2] for f in free]      # within 1 MiB of the level after warm-up
