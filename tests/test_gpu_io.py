"""Raw field files in the reference's on-disk format and the post-predictor host hook.

Mirrors test/small_test/io (restart must be bit-identical: run.sh:6-16 diffs the raw state files) and
test/small_test/fields/methods.f90:101-116 (write -> read round trip is exact)."""
import os

import numpy as np
import pytest

import fen_b200 as fb
from oracle import fen_oracle as fo
from tests.ranks import run_ranks, slab_grid, slab_of

pytestmark = pytest.mark.gpu
PI = fo.PI


def _tgv(n, device=0):
    G = fb.grid().setup(n, n, n, 2 * PI, 2 * PI, 2 * PI, device=device)
    ns = fb.Solver(G, 1.0, 0.01).init_solver()
    ns.CFL = 0.25
    Go = fo.Grid(n, n, n, 2 * PI, 2 * PI, 2 * PI)
    nso = fo.NavierStokes(Go, 1.0, 0.01)
    fo.init_tgv3d(nso)
    for a, b in ((ns.v.x, nso.v.x), (ns.v.y, nso.v.y), (ns.v.z, nso.v.z), (ns.p, nso.p)):
        a.f[...] = b.f
        a.push()
    return G, ns


def test_restart_is_bit_identical(tmp_path):
    """test/small_test/io/test_NS.f90: 10 steps  ==  5 steps + save_state + load_state + 5 steps."""
    n = 16
    d = str(tmp_path)
    G, ns = _tgv(n)
    dt = ns.set_timestep(1.0)
    for s in range(1, 11):
        ns.navier_stokes_solver(s, dt)
    a = ns.save_state(10, d)
    ref = open(a, "rb").read()
    os.rename(a, a + ".full")
    G.destroy()

    G, ns = _tgv(n)
    dt = ns.set_timestep(1.0)
    for s in range(1, 6):
        ns.navier_stokes_solver(s, dt)
    ns.save_state(5, d)
    G.destroy()

    G = fb.grid().setup(n, n, n, 2 * PI, 2 * PI, 2 * PI)
    ns = fb.Solver(G, 1.0, 0.01).init_solver()
    ns.CFL = 0.25
    dt = ns.set_timestep(1.0)          # dt and dt_o are not in the file: the driver re-derives them (test_NS.f90:54-63)
    ns.load_state(5, d)
    for s in range(6, 11):
        ns.navier_stokes_solver(s, dt)
    b = ns.save_state(10, d)
    assert open(b, "rb").read() == ref
    assert len(ref) == 7 * n ** 3 * 8
    # layout: field blocks in the order p, v_x, v_y, dv_o_x, dv_o_y, v_z, dv_o_z, each x fastest (solver.f90:201-210)
    ns.p.pull(); ns.v.pull(); ns.dv_o.pull()
    raw = np.frombuffer(ref, dtype=np.float64).reshape((7, n, n, n))
    for blk, f in zip(raw, (ns.p, ns.v.x, ns.v.y, ns.dv_o.x, ns.dv_o.y, ns.v.z, ns.dv_o.z)):
        assert np.array_equal(blk.transpose(2, 1, 0), f.I)      # (k, j, i) on disk == Fortran order
    G.destroy()


def test_scalar_write_read_roundtrip_and_save_fields(tmp_path):
    n = (32, 16, 8)
    G = fb.grid().setup(n[0], n[1], n[2], 2.0, 1.0, 0.5)
    s = fb.scalar(G, 1)
    rng = np.random.default_rng(7)
    s.I[...] = rng.standard_normal(n)
    want = s.I.copy()
    s.push()
    path = str(tmp_path / "s.raw")
    s.write(path)
    # postpro.py convention: np.fromfile(...).reshape((Nx, Ny, Nz), order='F')
    assert np.array_equal(np.fromfile(path).reshape(n, order="F"), want)
    s.I[...] = 0.0
    s.push()
    s.read(path)
    s.pull()
    assert np.array_equal(s.I, want)                              # methods.f90:101-116: exact
    # save_fields: cell-centred velocities 0.5 (f(i) + f(i-1)) (fields.f90:210) and p
    ns = fb.Solver(G, 1.0, 0.1).init_solver()
    for f in (ns.v.x, ns.v.y, ns.v.z, ns.p):
        f.f[...] = rng.standard_normal(f.f.shape)
        f.push()
    ns.v.update_ghost_nodes(); ns.v.pull()
    ns.save_fields(3, str(tmp_path))
    u = ns.v.x.f
    vx = np.fromfile(str(tmp_path / "vx_0000003.raw")).reshape(n, order="F")
    assert np.array_equal(vx, 0.5 * (u[1:-1, 1:-1, 1:-1] + u[0:-2, 1:-1, 1:-1]))
    w = ns.v.z.f
    vz = np.fromfile(str(tmp_path / "vz_0000003.raw")).reshape(n, order="F")
    assert np.array_equal(vz, 0.5 * (w[1:-1, 1:-1, 1:-1] + w[1:-1, 1:-1, 0:-2]))
    assert np.array_equal(np.fromfile(str(tmp_path / "p_0000003.raw")).reshape(n, order="F"), ns.p.I)
    with pytest.raises(fb.FenError):
        ns.load_state(99, str(tmp_path))                          # missing file is an error, not a silent no-op
    G.destroy()


def test_slab_ranks_write_one_global_file(tmp_path):
    """Two z-slab ranks write their byte ranges of ONE file; it equals the single-rank file (decomp_2d_write_one)."""
    n, L = (16, 16, 16), (1.0, 1.0, 1.0)
    rng = np.random.default_rng(3)
    glob = np.asfortranarray(rng.standard_normal((n[0] + 2, n[1] + 2, n[2] + 2)))
    path = str(tmp_path / "g.raw")

    def prog(rank, P, comm):
        G = slab_grid(rank, P, comm, n, L)
        s = fb.scalar(G, 1)
        s.f[...] = slab_of(glob, rank, P, 1)
        s.push()
        G.synchronize()
        comm.sync()
        s.write(path)
        comm.sync()
        s.I[...] = 0.0
        s.push()
        s.read(path)
        s.pull()
        out = s.I.copy()
        comm.sync()
        G.destroy()
        return out

    parts = run_ranks(2, prog)
    assert np.array_equal(np.fromfile(path).reshape(n, order="F"), glob[1:-1, 1:-1, 1:-1])
    assert np.array_equal(np.concatenate(parts, axis=2), glob[1:-1, 1:-1, 1:-1])


def test_forcing_hook_runs_between_predictor_and_poisson():
    """The place of apply_ibm_forcing(v, dt) (navier_stokes.f90:106-108): a host callback forces v after the
    predictor; the oracle does the same by hand."""
    n = 16
    G, ns = _tgv(n)
    Go = fo.Grid(n, n, n, 2 * PI, 2 * PI, 2 * PI)
    nso = fo.NavierStokes(Go, 1.0, 0.01)
    fo.init_tgv3d(nso)
    nso.CFL = 0.25
    dt = nso.set_timestep(1.0)
    assert ns.set_timestep(1.0) == dt
    mask = np.zeros((n, n, n))
    mask[4:8, 4:8, 4:8] = 1.0          # a "solid" block where the velocity is forced to zero
    calls = []

    def hook(step, hdt):
        calls.append((step, hdt))
        ns.v.pull()
        for c in ns.v.comps:
            c.I[...] = c.I * (1.0 - mask)
        ns.v.push()

    ns.set_forcing_hook(hook)
    for s in (1, 2, 3):
        ns.navier_stokes_solver(s, dt)
        # oracle: the same sequence with the forcing between predictor and divergence
        nso.predicted_velocity_field(dt)
        for c in nso.v.comps:
            c.I[...] = c.I * (1.0 - mask)
        nso.v.update_ghost_nodes()
        fo.divergence(nso.v, nso.phi)
        nso.phi.I[...] = nso.phi.I * nso.rho.sh() / dt
        nso.poisson.solve(nso.phi)
        nso.phi.update_ghost_nodes()
        nso.correct_velocity_field(dt)
        nso.update_pressure()
        nso.checks(dt)
    assert calls == [(1, dt), (2, dt), (3, dt)]
    ns.v.pull(); ns.p.pull()
    for a, b in ((ns.v.x, nso.v.x), (ns.v.y, nso.v.y), (ns.v.z, nso.v.z), (ns.p, nso.p)):
        assert np.linalg.norm(a.f - b.f) <= 1e-12 * np.linalg.norm(b.f)
    ns.set_forcing_hook(None)
    ns.navier_stokes_solver(4, dt)
    assert len(calls) == 3
    G.destroy()


def test_pull_async_and_push_ordering():
    """fen_gpu_pull_async: the copy runs on its own stream; pull_wait makes the host array valid, and a push of the same
    array right behind it waits for the download on the device side (the e2e loop of bench.py relies on both)."""
    import torch
    n = (64, 32, 16)
    G = fb.grid().setup(n[0], n[1], n[2], 2.0, 1.0, 0.5)
    rng = np.random.default_rng(3)
    fields = [fb.scalar(G, 1) for _ in range(6)]
    want = []
    for s in fields:
        # pinned host arrays, so that the copies really are asynchronous
        t = torch.empty(s.f.size, dtype=torch.float64, pin_memory=True)
        s.f = t.numpy().reshape(s.f.shape, order="F")
        s._keep = t
        s.f[...] = rng.standard_normal(s.f.shape)
        want.append(s.f.copy())
        s.push()
    for s in fields:
        s.f[...] = 0.0
    for s in fields:                       # more pulls in flight than staging buffers (4)
        s.pull_async()
    G.pull_wait()
    for s, w in zip(fields, want):
        assert np.array_equal(s.f, w)
    # download, then upload the same array without waiting on the host, twice round
    for rep in range(2):
        for s in fields:
            s.pull_async()
        for s in fields:
            s.push()
    G.synchronize()
    for s in fields:
        s.f[...] = -1.0
        s.pull()
    for s, w in zip(fields, want):
        assert np.array_equal(s.f, w)
    G.destroy()
