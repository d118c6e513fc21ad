"""Test helper: run an SPMD "rank program" on P z-slab ranks inside ONE process, one thread per rank.

Each thread owns one fen_ctx (rank r of P) on the CUDA device ``devices[r]`` (all on device 0 by default,
so the whole multi-rank path -- halo mailboxes, fused transposes, flag waits, all-reduce -- runs on a
single-GPU box).  ctypes releases the GIL inside every library call, so the ranks really run
concurrently and collective calls may block on each other exactly as MPI ranks would.
"""
from __future__ import annotations

import os
import threading
from concurrent.futures import ThreadPoolExecutor

os.environ.setdefault("FEN_GPU_SPIN_LIMIT_MS", "3000")    # a missed flag fails the test in seconds
# Every rank drives two streams (the blocked slab path overlaps its transposes on an auxiliary one).  With all ranks on
# ONE device the default 8 hardware work queues would alias 16 streams: a rank's solve kernel could sit in a queue behind
# another rank's spinning flag wait and never start (seen: 8 ranks, "peer wait timed out").  One process per GPU -- the
# production layout -- has three streams per device and cannot alias.  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np

import fen_b200 as fb


class ThreadComm:
    def __init__(self, P):
        self.P = P
        self.barrier = threading.Barrier(P, timeout=120)
        self.box = [None] * P

    def sync(self):
        """Thread barrier.  Rank programs call it between their set-up phase (cudaMalloc / cudaMemcpy are
        device-wide synchronisations) and their collective phase: with all ranks on ONE device a rank that
        is already waiting in a flag kernel would otherwise block a late rank's cudaMalloc, which in turn
        could never raise the flag.  With one process per GPU (the real layout) the hazard does not exist."""
        self.barrier.wait()

    def all_gather(self, rank, obj):
        self.box[rank] = obj
        self.barrier.wait()
        out = list(self.box)
        self.barrier.wait()
        return out


def run_ranks(P, program, devices=None, timeout=300):
    """program(rank, P, comm) -> result; returns the list of results in rank order."""
    comm = ThreadComm(P)

    def guarded(r):
        try:
            return program(r, P, comm)
        except BaseException:
            comm.barrier.abort()          # release the other ranks instead of leaving them in wait()
            raise

    with ThreadPoolExecutor(max_workers=P) as ex:
        futs = [ex.submit(guarded, r) for r in range(P)]
        results, first = [], None
        for f in futs:
            try:
                results.append(f.result(timeout=timeout))
            except threading.BrokenBarrierError as e:
                first = first or e
            except BaseException as e:
                if first is None or isinstance(first, threading.BrokenBarrierError):
                    first = e
        if first is not None:
            raise first
        return results


def slab_grid(rank, P, comm, n, L, bc=None, device=0):
    """grid%setup with (prow, pcol) = (1, P) + the peer-memory wiring."""
    G = fb.grid().setup(n[0], n[1], n[2], L[0], L[1], L[2], pcol=P, rank=rank, bc=bc, device=device)
    if P > 1:
        G.connect(lambda b: comm.all_gather(rank, b))
    return G


def slab_of(a_global, rank, P, gl):
    """Slab of a global Fortran-ordered array with gl ghost layers (ghost planes included)."""
    nz = a_global.shape[2] - 2 * gl
    nzl = nz // P
    return np.asfortranarray(a_global[:, :, rank * nzl: rank * nzl + nzl + 2 * gl])


def gather_interior(parts, gl):
    """Concatenate the interiors of per-rank slabs along z."""
    ins = [p[gl:p.shape[0] - gl, gl:p.shape[1] - gl, gl:p.shape[2] - gl] if gl else p for p in parts]
    return np.concatenate(ins, axis=2)
