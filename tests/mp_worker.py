"""SPMD worker for tests/test_gpu_multiprocess.py: one process per GPU (torch.distributed.run), peers mapped
through CUDA IPC.  Every rank advances the 3-D Taylor-Green case on its z slab; rank 0 also advances the whole
domain on a single-rank context and checks that the slabs carry the same bits.  Prints MP_WORKER_OK on success."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import fen_b200 as fb
    from oracle import fen_oracle as fo

    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    case = sys.argv[1] if len(sys.argv) > 1 else "ppp"
    backend = "nccl" if torch.cuda.device_count() >= world else "gloo"
    dev = local % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, device_id=torch.device("cuda", dev) if backend == "nccl" else None)

    def all_gather(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out

    two_pi = 2.0 * fo.PI
    if case == "ppp":
        n, L, bc, nu, g, U, cfl, init = (128, 64, 64), (two_pi, two_pi / 2, two_pi / 2), None, 0.01, None, 1.0, 0.25, fo.init_tgv3d
    elif case == "ppp1024":
        # the y / z line lengths of the 8-GPU bench (1024 points, 128 per rank at world = 8), x shrunk to 16 cells
        n, L, bc, nu, g, U, cfl, init = (16, 1024, 1024), (two_pi / 64, two_pi, two_pi), None, 0.01, None, 1.0, 0.25, fo.init_tgv3d
    else:
        n, L, bc, nu, g, U, cfl, init = (32, 32, 16), (2.0, 2.0, 1.0), ["Periodic"] * 4 + ["Wall", "Wall"], 0.05, (1.0, 0.0, 0.0), 1.0, 0.05, fo.init_channel
    Go = fo.Grid(n[0], n[1], n[2], L[0], L[1], L[2], bc=bc)
    nso = fo.NavierStokes(Go, 1.0, nu)
    if g is not None:
        nso.g = list(g)
    init(nso)
    state = [a.f.copy() for a in (nso.v.x, nso.v.y, nso.v.z, nso.p)]
    steps = 2 if case == "ppp1024" else 5

    def advance(P, r, connect):
        G = fb.grid().setup(n[0], n[1], n[2], L[0], L[1], L[2], pcol=P, rank=r, bc=bc, device=dev)
        if connect:
            G.connect(all_gather)
        ns = fb.Solver(G, 1.0, nu)
        if g is not None:
            ns.g = list(g)
        ns.init_solver()
        ns.CFL = cfl
        dt = ns.set_timestep(U)
        nzl = n[2] // P
        for a, s in zip((ns.v.x, ns.v.y, ns.v.z, ns.p), state):
            a.f[...] = s[:, :, r * nzl: r * nzl + nzl + 2]
            a.push()
        ns.v.update_ghost_nodes(); ns.p.update_ghost_nodes()
        hist = []
        for step in range(1, steps + 1):
            ns.navier_stokes_solver(step, dt)
            hist.append(ns.status())
        ns.v.pull(); ns.p.pull()
        out = [a.f.copy() for a in (ns.v.x, ns.v.y, ns.v.z, ns.p)]
        return G, out, hist

    G, mine, hist = advance(world, rank, True)
    dist.barrier()
    G.destroy()
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, hist))
    ok = True
    if rank == 0:
        G1, one, hist1 = advance(1, 0, False)
        G1.destroy()
        nzl = n[2] // world
        for r in range(world):
            for m in range(4):
                if not np.array_equal(gathered[r][0][m], one[m][:, :, r * nzl: r * nzl + nzl + 2]):
                    ok = False
                    print("rank %d field %d differs: max %.3e" % (
                        r, m, np.abs(gathered[r][0][m] - one[m][:, :, r * nzl: r * nzl + nzl + 2]).max()))
            if gathered[r][1] != hist1:
                ok = False
                print("rank %d status history differs" % r)
        print("MP_WORKER_OK" if ok else "MP_WORKER_FAIL", case, "world", world, "backend", backend)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
