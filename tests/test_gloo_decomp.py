"""World-size-2 (and 4) `gloo` tests of the N > 1 HOST logic on CPU: slab / z-pencil bounds, neighbour wiring,
transpose block map, rank-ordered handle gather, slab-wise initial condition, max-over-ranks timing rule, and the
granule-blocked layouts + run descriptors of the bulk-store transposes (carried out between the ranks for real).
The device side of the multi-GPU path is covered by tests/test_gpu_multirank.py and test_gpu_multiprocess.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fen_b200 import decomp
        import bench
        nx, ny, nz = 16, 8, 8 * world
        lo, hi = decomp.slab_bounds(nx, ny, nz, world, rank)
        # 1. slabs tile the z range without gap or overlap
        allb = [None] * world
        dist.all_gather_object(allb, (lo, hi))
        ks = []
        for (l, h) in allb:
            assert l[:2] == (1, 1) and h[:2] == (nx, ny)
            ks += list(range(l[2], h[2] + 1))
        assert ks == list(range(1, nz + 1))
        # 2. neighbour wiring is symmetric (my back neighbour's front neighbour is me), periodic and not
        for periodic in (True, False):
            nb = decomp.z_neighbours(world, rank, periodic)
            alln = [None] * world
            dist.all_gather_object(alln, nb)
            if nb[1] >= 0:
                assert alln[nb[1]][0] == rank
            if nb[0] >= 0:
                assert alln[nb[0]][1] == rank
            if not periodic:
                assert (nb[0] == -1) == (rank == 0) and (nb[1] == -1) == (rank == world - 1)
        # 3. handle gather is in rank order whatever the collective returns first
        blob = bytes([rank]) * 24

        def ag(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        allh = decomp.gather_handles(ag, blob, world)
        assert allh == b"".join(bytes([r]) * 24 for r in range(world))
        # 4. transpose: what I send to `dest` is exactly what `dest` expects from me, and the blocks cover the pencil
        mine = decomp.transpose_blocks(ny, nz, world, rank)
        allt = [None] * world
        dist.all_gather_object(allt, mine)
        zlo, zhi = decomp.zpencil_bounds(nx, ny, nz, world, rank)
        cover = np.zeros((ny, nz), dtype=int)
        for src in range(world):
            dest, (j0, j1), (k0, k1) = allt[src][rank]
            assert dest == rank and (j0, j1) == (zlo[1], zhi[1])
            cover[j0 - 1:j1, k0 - 1:k1] += 1
        assert (cover[zlo[1] - 1:zhi[1], :] == 1).all() and cover.sum() == (zhi[1] - zlo[1] + 1) * nz
        # 5. the slab-wise initial condition of bench.py assembles to the single-rank one
        delta = 2 * np.pi / nx
        _, (u, v, w, p) = bench.init_tgv3d_slab((nx, ny, nz // world), delta, lo[2] - 1, pinned=False)
        parts = [None] * world
        dist.all_gather_object(parts, (u[1:-1, 1:-1, 1:-1].copy(), p[1:-1, 1:-1, 1:-1].copy()))
        _, (ug, vg, wg, pg) = bench.init_tgv3d_slab((nx, ny, nz), delta, 0, pinned=False)
        assert np.array_equal(np.concatenate([a for a, _ in parts], axis=2), ug[1:-1, 1:-1, 1:-1])
        assert np.array_equal(np.concatenate([b for _, b in parts], axis=2), pg[1:-1, 1:-1, 1:-1])
        # 6. timing rule: the reported time is the max over ranks
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t) == 10.0 + world - 1
        # 7. bytes of one all-to-all
        assert decomp.alltoall_bytes_per_gpu(1024, 1024, 1024, 8) == 513 * 1024 * 128 * 16 * 7 // 8
        assert decomp.alltoall_bytes_per_gpu(64, 64, 64, 8, x_periodic=False) == 64 * 64 * 8 * 16 * 7 // 8
        # 8. the row-copy transpose (k_a2a_scatter): every rank covers every (dest, index) pair exactly once, and at
        #    every block position the ranks address pairwise different destinations (no receiver is shared)
        blk = nz // world
        sched = decomp.scatter_schedule(world, rank, blk)
        assert sorted(i for _, i in sched) == list(range(nz))
        assert all(dst == i // blk for dst, i in sched)
        allsched = [None] * world
        dist.all_gather_object(allsched, [dst for dst, _ in sched])
        for pos in range(len(sched)):
            assert len({allsched[r][pos] for r in range(world)}) == world
        # 9. the granule-blocked transposes of slab_bulk.cuh, carried out for real between the ranks: every rank fills the
        #    tiles its kernels would hold with a global index code, ships the runs decomp.forward_runs / backward_runs
        #    describe (gloo all_to_all of the concatenated runs), places them at the stated offsets, and finds at every
        #    blocked offset of its z-pencil / y-slab array exactly the element the layout formulas name -- in 1, 2 and 3
        #    pieces of the chunked solve
        NG, nyb, nzb = 2, 4 * world, 2 * world                    # granules, ny, nz
        nyl, nzl = nyb // world, nzb // world
        code = lambda g_, kxi, j, k: ((g_ * 8 + kxi) * nyb + j) * nzb + k      # noqa: E731
        for nq in (1, 2, 3):
            Cz = np.full(NG * nzb * nyl * 8, -1, dtype=np.int64)
            send = [[] for _ in range(world)]
            for (z0, z1) in decomp.pieces(nzl, nq):               # forward pieces = z planes
                for zl in range(z0, z1):
                    for g_ in range(NG):
                        tile = np.array([code(g_, kxi, j, rank * nzl + zl) for j in range(nyb) for kxi in range(8)])
                        for dest, j0, n, off in decomp.forward_runs(nyb, nzb, world, rank, g_, zl):
                            send[dest].append((off, tile[j0 * 8: j0 * 8 + n]))
            allsend = [None] * world
            dist.all_gather_object(allsend, send)
            for src in range(world):
                for off, run in allsend[src][rank]:
                    assert (Cz[off: off + len(run)] == -1).all()  # nobody else writes there
                    Cz[off: off + len(run)] = run
            for g_ in range(NG):
                for k in range(nzb):
                    for jl in range(nyl):
                        for kxi in range(8):
                            assert Cz[decomp.zpencil_blocked_offset(g_, k, jl, kxi, nzb, nyl)] == \
                                code(g_, kxi, rank * nyl + jl, k)
            Cy = np.full(NG * nyb * nzl * 8, -1, dtype=np.int64)
            send = [[] for _ in range(world)]
            for (g0, g1) in decomp.pieces(NG, nq):                # backward pieces = granules
                for g_ in range(g0, g1):
                    for jl in range(nyl):
                        tile = np.array([Cz[decomp.zpencil_blocked_offset(g_, k, jl, kxi, nzb, nyl)]
                                         for k in range(nzb) for kxi in range(8)])
                        for dest, k0, n, off in decomp.backward_runs(nyb, nzb, world, rank, g_, jl):
                            send[dest].append((off, tile[k0 * 8: k0 * 8 + n]))
            dist.all_gather_object(allsend, send)
            for src in range(world):
                for off, run in allsend[src][rank]:
                    assert (Cy[off: off + len(run)] == -1).all()
                    Cy[off: off + len(run)] = run
            for g_ in range(NG):
                for j in range(nyb):
                    for zl in range(nzl):
                        for kxi in range(8):
                            assert Cy[decomp.yslab_blocked_offset(g_, j, zl, kxi, nyb, nzl)] == \
                                code(g_, kxi, j, rank * nzl + zl)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: %r\n%s" % (e, traceback.format_exc())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_host_decomposition_logic_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_decomposition_argument_errors():
    sys.path.insert(0, ROOT)
    from fen_b200 import decomp
    with pytest.raises(ValueError):
        decomp.slab_bounds(8, 8, 9, 2, 0)
    with pytest.raises(ValueError):
        decomp.slab_bounds(8, 8, 8, 2, 2)
    with pytest.raises(ValueError):
        decomp.gather_handles(lambda b: [b], b"x", 2)
