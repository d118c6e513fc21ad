"""Multi-rank (z-slab) parity: P ranks must give the SAME BITS as one rank.

The reference never asserts rank-count invariance; here it is exact because the
decomposition only moves data: the halos are copies, the fused transposes deliver every spectral
coefficient to the rank that needs it unchanged, the max-reductions are exact and the mean of the ppn
solver is taken on the one rank that owns the (kx, ky) = (0, 0) line.  Ranks run as threads of this
process on one device (tests/ranks.py); tests/test_gpu_multiprocess.py repeats the step with one process
per GPU over CUDA IPC when the box has at least two GPUs.
"""
import numpy as np
import pytest

import fen_b200 as fb
from oracle import fen_oracle as fo
from tests.ranks import gather_interior, run_ranks, slab_grid, slab_of

pytestmark = pytest.mark.gpu
PI = fo.PI
ZWALLS = ["Periodic"] * 4 + ["Wall", "Wall"]


@pytest.mark.parametrize("P", [2, 4])
@pytest.mark.parametrize("bc", [None, ZWALLS])
def test_ghost_nodes_match_single_rank(P, bc):
    n = (16, 8, 16)
    rng = np.random.default_rng(5)
    glob = [np.asfortranarray(rng.random((n[0] + 2, n[1] + 2, n[2] + 2))) for _ in range(3)]
    L = (1.0, 0.5, 1.0)

    def program(rank, P, comm):
        G = slab_grid(rank, P, comm, n, L, bc)
        v = fb.vector(G, 1)
        if bc is not None:
            # what allocate_navier_stokes_fields wires for Wall faces (Dirichlet), on the ranks that own them
            for c in v.comps:
                if rank == 0:
                    c.set_bc_type("front", 1)
                if rank == P - 1:
                    c.set_bc_type("back", 1)
        for c, g in zip(v.comps, glob):
            c.f[...] = slab_of(g, rank, P, 1)
            c.push()
        G.synchronize()
        comm.sync()
        v.update_ghost_nodes()
        mx = v.x.max_value()
        v.pull()
        out = [c.f.copy() for c in v.comps]
        comm.sync()
        G.destroy()
        return out, mx

    one = run_ranks(1, program)[0]
    many = run_ranks(P, program)
    nzl = n[2] // P
    for r in range(P):
        for m in range(3):
            want = one[0][m][:, :, r * nzl: r * nzl + nzl + 2]
            assert np.array_equal(many[r][0][m], want), (r, m)
        assert many[r][1] == one[1]


def _poisson_program(n, L, bc, rhs):
    def program(rank, P, comm):
        G = slab_grid(rank, P, comm, n, L, bc)
        phi = fb.scalar(G, 1)
        if bc is not None:
            for face, s in zip(fo.FACES[:4], bc[:4]):
                if s == "Wall":
                    phi.set_bc_type(face, 2)
            if rank == 0:
                phi.set_bc_type("front", 2)
            if rank == P - 1:
                phi.set_bc_type("back", 2)
        ps = fb.PoissonSolver(phi)
        phi.f[...] = slab_of(rhs, rank, P, 1)
        phi.push()
        G.synchronize()
        comm.sync()
        ps.solve(phi)
        phi.update_ghost_nodes()
        phi.pull()
        out = phi.f.copy()
        var = ps.variant
        comm.sync()
        G.destroy()
        return out, var
    return program


@pytest.mark.parametrize("P", [2, 4])
@pytest.mark.parametrize("n,L", [((32, 16, 16), (2.0, 1.0, 1.0)),
                                 ((128, 64, 64), (2.0, 1.0, 1.0))])     # second: register-path transforms + fused transposes
@pytest.mark.parametrize("bc,variant", [(None, "ppp"), (ZWALLS, "ppn")])
def test_poisson_matches_single_rank_and_oracle(P, bc, variant, n, L):
    rng = np.random.default_rng(11)
    rhs = np.zeros((n[0] + 2, n[1] + 2, n[2] + 2), order="F")
    rhs[1:-1, 1:-1, 1:-1] = rng.standard_normal(n)
    rhs[1:-1, 1:-1, 1:-1] -= rhs[1:-1, 1:-1, 1:-1].mean()
    prog = _poisson_program(n, L, bc, rhs)
    one = run_ranks(1, prog)[0]
    many = run_ranks(P, prog)
    assert one[1] == variant and all(m[1] == variant for m in many)
    got = gather_interior([m[0] for m in many], 1)
    assert np.array_equal(got, one[0][1:-1, 1:-1, 1:-1])
    # and the single-rank answer is the oracle's
    Go = fo.Grid(n[0], n[1], n[2], L[0], L[1], L[2], bc=bc)
    po = fo.Scalar(Go, 1)
    if bc is not None:
        po.bc_type["front"] = po.bc_type["back"] = 2
    po.f[...] = rhs
    fo.PoissonSolver(po).solve(po)
    ref = po.I
    assert np.linalg.norm(got - ref) <= 1e-12 * np.linalg.norm(ref)


@pytest.mark.parametrize("bc,variant", [(None, "ppp"), (ZWALLS, "ppn")])
def test_poisson_1024_point_lines_on_8_ranks(bc, variant):
    """The line lengths and rank count of the 8-GPU bench (1024-point y and z lines, 128 lines / planes per rank: the
    blocked slab path of slab_bulk.cuh with 16 KB bulk stores to 8 destinations), x shrunk to 16 cells so that the case
    fits one device and the oracle: same bits as one rank, and the oracle's answer."""
    n, L = (16, 1024, 1024), (0.125, 8.0, 8.0)
    rng = np.random.default_rng(17)
    rhs = np.zeros((n[0] + 2, n[1] + 2, n[2] + 2), order="F")
    rhs[1:-1, 1:-1, 1:-1] = rng.standard_normal(n)
    rhs[1:-1, 1:-1, 1:-1] -= rhs[1:-1, 1:-1, 1:-1].mean()
    prog = _poisson_program(n, L, bc, rhs)
    one = run_ranks(1, prog)[0]
    many = run_ranks(8, prog)
    assert one[1] == variant and all(m[1] == variant for m in many)
    got = gather_interior([m[0] for m in many], 1)
    assert np.array_equal(got, one[0][1:-1, 1:-1, 1:-1])
    Go = fo.Grid(n[0], n[1], n[2], L[0], L[1], L[2], bc=bc)
    po = fo.Scalar(Go, 1)
    if bc is not None:
        po.bc_type["front"] = po.bc_type["back"] = 2
    po.f[...] = rhs
    fo.PoissonSolver(po).solve(po)
    ref = po.I
    assert np.linalg.norm(got - ref) <= 1e-12 * np.linalg.norm(ref)


@pytest.mark.parametrize("env", [{"FEN_SLAB_BULK": "0"}, {"FEN_SLAB_CHUNKS": "3", "FEN_SLAB_SMS": "8"},
                                 {"FEN_SLAB_CHUNKS": "1"}])
def test_slab_transpose_forms_give_the_same_bits(env):
    """The forms of the slab transposes deliver the same coefficients as one rank: FEN_SLAB_BULK=0 (round 1's register
    stores into the peers' row-layout arrays), bulk stores in three overlapped pieces with the persistent transposing
    kernels capped at 8 SMs' worth of blocks (also on 2 ranks and for ppn, where one piece is the default), and
    unchunked.  The switches are read per process, so each runs in a child."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "tests/test_gpu_multirank.py", "-k",
                        "test_ns_steps_tgv3d or test_ns_steps_channel_ppn or test_poisson_1024"], cwd=root,
                       env=dict(os.environ, **env),
                       capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("P,n", [(2, (24, 12, 12)), (3, (20, 12, 9)), (3, (32, 24, 48))])
@pytest.mark.parametrize("bc,variant", [(None, "ppp"), (ZWALLS, "ppn")])
def test_poisson_any_length_on_slabs(P, n, bc, variant):
    """Grid sizes that are not powers of two (the any-length kernels of fft_any.cuh) and a rank count that is not one
    either, on z slabs: the passes run in place and the y <-> z transposes are staggered row copies (k_a2a_scatter with
    the owner of a line found by division) -- same bits as one rank, and the oracle's answer."""
    L = (1.0, 1.0 * n[1] / n[0], 1.0 * n[2] / n[0])
    rng = np.random.default_rng(19)
    rhs = np.zeros((n[0] + 2, n[1] + 2, n[2] + 2), order="F")
    rhs[1:-1, 1:-1, 1:-1] = rng.standard_normal(n)
    rhs[1:-1, 1:-1, 1:-1] -= rhs[1:-1, 1:-1, 1:-1].mean()
    prog = _poisson_program(n, L, bc, rhs)
    one = run_ranks(1, prog)[0]
    many = run_ranks(P, prog)
    assert one[1] == variant and all(m[1] == variant for m in many)
    got = gather_interior([m[0] for m in many], 1)
    assert np.array_equal(got, one[0][1:-1, 1:-1, 1:-1])
    Go = fo.Grid(n[0], n[1], n[2], L[0], L[1], L[2], bc=bc)
    po = fo.Scalar(Go, 1)
    if bc is not None:
        po.bc_type["front"] = po.bc_type["back"] = 2
    po.f[...] = rhs
    fo.PoissonSolver(po).solve(po)
    ref = po.I
    assert np.linalg.norm(got - ref) <= 1e-12 * np.linalg.norm(ref)


def test_ns_steps_any_length_on_three_slabs():
    """Five Taylor-Green steps on a 24 x 12 x 18 grid split over three ranks: the whole step (halos, right-hand side,
    any-length Poisson passes, row-copy transposes, correction, checks) against one rank, bit for bit."""
    n = (24, 12, 18)
    L = (2 * PI, 2 * PI * n[1] / n[0], 2 * PI * n[2] / n[0])
    prog, _ = _ns_program(n, L, None, 0.01, fo.init_tgv3d, 1.0, None, 0.25, 5)
    one = run_ranks(1, prog)[0]
    many = run_ranks(3, prog)
    nzl = n[2] // 3
    for r in range(3):
        for m in range(4):
            assert np.array_equal(many[r][0][m], one[0][m][:, :, r * nzl: r * nzl + nzl + 2]), (r, m)
        assert many[r][1] == one[1]


XZWALLS = ["Wall", "Wall", "Periodic", "Periodic", "Wall", "Wall"]
ALLWALLS = ["Wall"] * 6


@pytest.mark.parametrize("P", [2, 4])
@pytest.mark.parametrize("bc,variant", [(XZWALLS, "npn"), (ALLWALLS, "nnn")])
def test_poisson_dct_variants_on_slabs(P, bc, variant):
    """The Neumann-in-x variants (DCT-II / III in x, and in y for nnn; poisson.f90:1177-1451) on z slabs: same bits
    as one rank, and the oracle's answer."""
    n, L = (32, 16, 16), (2.0, 1.0, 1.0)
    rng = np.random.default_rng(13)
    rhs = np.zeros((n[0] + 2, n[1] + 2, n[2] + 2), order="F")
    rhs[1:-1, 1:-1, 1:-1] = rng.standard_normal(n)
    rhs[1:-1, 1:-1, 1:-1] -= rhs[1:-1, 1:-1, 1:-1].mean()
    prog = _poisson_program(n, L, bc, rhs)
    one = run_ranks(1, prog)[0]
    many = run_ranks(P, prog)
    assert one[1] == variant and all(m[1] == variant for m in many)
    got = gather_interior([m[0] for m in many], 1)
    assert np.array_equal(got, one[0][1:-1, 1:-1, 1:-1])
    Go = fo.Grid(n[0], n[1], n[2], L[0], L[1], L[2], bc=bc)
    po = fo.Scalar(Go, 1)
    po.f[...] = rhs
    fo.PoissonSolver(po).solve(po)
    ref = po.I
    # npn / nnn do not remove the mean (poisson.f90:1310, :1449 are commented out): compare as computed
    assert np.linalg.norm(got - ref) <= 1e-11 * np.linalg.norm(ref)


def _ns_program(n, L, bc, nu, init, U, g, cfl, steps, constant_cfl=False):
    Go = fo.Grid(n[0], n[1], n[2], L[0], L[1], L[2], bc=bc)
    nso = fo.NavierStokes(Go, 1.0, nu)
    if g is not None:
        nso.g = list(g)
    init(nso)
    state = [a.f.copy() for a in (nso.v.x, nso.v.y, nso.v.z, nso.p)]

    def program(rank, P, comm):
        G = slab_grid(rank, P, comm, n, L, bc)
        ns = fb.Solver(G, 1.0, nu)
        if g is not None:
            ns.g = list(g)
        ns.init_solver()
        ns.CFL = cfl
        ns.constant_CFL = constant_cfl
        dt = ns.set_timestep(U)
        for a, s in zip((ns.v.x, ns.v.y, ns.v.z, ns.p), state):
            a.f[...] = slab_of(s, rank, P, 1)
            a.push()
        G.synchronize()
        comm.sync()
        ns.v.update_ghost_nodes()
        ns.p.update_ghost_nodes()
        hist = []
        for step in range(1, steps + 1):
            dt = ns.navier_stokes_solver(step, dt)
            hist.append((dt,) + ns.status())
        ns.v.pull(); ns.p.pull()
        out = [a.f.copy() for a in (ns.v.x, ns.v.y, ns.v.z, ns.p)]
        comm.sync()
        G.destroy()
        return out, hist
    return program, nso


@pytest.mark.parametrize("P,n", [(2, (32, 32, 32)), (4, (32, 32, 32)), (2, (128, 64, 64))])
def test_ns_steps_tgv3d_rank_count_invariant(P, n):
    # (128, 64, 64): register-path FFTs, rhs fused into the x pass, fused correction + checks across a rank boundary
    prog, nso = _ns_program(n, (2 * PI,) * 3, None, 0.01, fo.init_tgv3d, 1.0, None, 0.25, 5)
    one = run_ranks(1, prog)[0]
    many = run_ranks(P, prog)
    nzl = n[2] // P
    for r in range(P):
        for m in range(4):
            assert np.array_equal(many[r][0][m], one[0][m][:, :, r * nzl: r * nzl + nzl + 2]), (r, m)
        assert many[r][1] == one[1]          # dt, maxdiv, maxCFL of every step: identical on all ranks
    # the oracle agrees to the one-step tolerance
    nso.CFL = 0.25
    dt = nso.set_timestep(1.0)
    for step in range(1, 6):
        nso.navier_stokes_solver(step, dt)
    for m, a in enumerate((nso.v.x, nso.v.y, nso.v.z, nso.p)):
        got = gather_interior([x[0][m] for x in many], 1)
        nrm = np.linalg.norm(a.I)
        assert np.linalg.norm(got - a.I) <= 1e-12 * max(nrm, 1.0)


@pytest.mark.parametrize("P,n", [(2, (32, 16, 16)), (4, (32, 16, 16)), (4, (64, 64, 32))])
def test_ns_steps_channel_ppn_rank_count_invariant(P, n):
    # (64, 64, 32): 64-point y lines -> blocked slab path (bulk-store transposes, Thomas on the blocked layout)
    L = (2.0, 2.0 * n[1] / n[0], 2.0 * n[2] / n[0])
    prog, _ = _ns_program(n, L, ZWALLS, 0.05, fo.init_channel, 1.0, (1.0, 0.0, 0.0), 0.05, 6)
    one = run_ranks(1, prog)[0]
    many = run_ranks(P, prog)
    nzl = n[2] // P
    for r in range(P):
        for m in range(4):
            assert np.array_equal(many[r][0][m], one[0][m][:, :, r * nzl: r * nzl + nzl + 2]), (r, m)
        assert many[r][1] == one[1]


@pytest.mark.parametrize("bc,variant", [(XZWALLS, "npn"), (ALLWALLS, "nnn")])
def test_ns_steps_closed_box_rank_count_invariant(bc, variant):
    """A closed box driven by a body force (walls in x and z, or everywhere: the lid3D / cavity wiring without the
    lid) on 2 slabs: Dirichlet / Neumann ghosts on the x and y walls of every rank, the DCT Poisson variants and
    their slab transposes inside the full step."""
    n = (32, 16, 16)

    def init(ns):
        i = np.arange(1, n[0] + 1)[:, None, None]
        k = np.arange(1, n[2] + 1)[None, None, :]
        ns.v.y.I[...] = 0.05 * np.sin(PI * (i - 0.5) / n[0]) * np.sin(PI * (k - 0.5) / n[2]) * np.ones((1, n[1], 1))
        if variant == "nnn":
            ns.v.y.I[:, -1, :] = 0.0          # the wall-normal component on the top wall face
        ns.v.update_ghost_nodes()
    prog, nso = _ns_program(n, (2.0, 1.0, 1.0), bc, 0.05, init, 1.0, (1.0, 0.5, 0.25), 0.2, 4)
    one = run_ranks(1, prog)[0]
    many = run_ranks(2, prog)
    nzl = n[2] // 2
    for r in range(2):
        for m in range(4):
            assert np.array_equal(many[r][0][m], one[0][m][:, :, r * nzl: r * nzl + nzl + 2]), (r, m)
        assert many[r][1] == one[1]
    nso.CFL = 0.2
    dt = nso.set_timestep(1.0)
    for step in range(1, 5):
        nso.navier_stokes_solver(step, dt)
    for m, a in enumerate((nso.v.x, nso.v.y, nso.v.z, nso.p)):
        got = gather_interior([x[0][m] for x in many], 1)
        ref = a.I
        if m == 3:                                # npn / nnn leave the mean of the pressure undetermined
            got, ref = got - got.mean(), ref - ref.mean()
        assert np.linalg.norm(got - ref) <= 1e-11 * max(np.linalg.norm(ref), 1.0), m


def test_constant_cfl_allreduce_multirank():
    """update_timestep's max-velocity all-reduce (navier_stokes.f90:712) with 2 ranks."""
    n = (32, 32, 32)
    prog, _ = _ns_program(n, (2 * PI,) * 3, None, 0.01, fo.init_tgv3d, 1.0, None, 0.25, 4, constant_cfl=True)
    one = run_ranks(1, prog)[0]
    many = run_ranks(2, prog)
    assert many[0][1] == one[1] and many[1][1] == one[1]


def test_missing_connect_is_an_error():
    G = fb.grid().setup(16, 16, 16, 1.0, 1.0, 1.0, pcol=2, rank=0)
    with pytest.raises(fb.FenError):
        fb.Solver(G).init_solver()
    G.destroy()
