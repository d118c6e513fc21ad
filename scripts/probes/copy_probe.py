"""Probe: host<->device copy rates on this box -- flat pinned copies (torch) against fen_gpu_push / fen_gpu_pull
(cudaMemcpy3DAsync between the Fortran-ordered host array and the padded device layout)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import fen_b200 as fb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
G = fb.grid().setup(n, n, n, 1.0, 1.0, 1.0, device=0)
s = fb.scalar(G, 1)
nel = (n + 2) ** 3
t = torch.empty(nel, dtype=torch.float64, pin_memory=True)
s.f = t.numpy().reshape((n + 2, n + 2, n + 2), order="F")
s.f[...] = 1.0
d = torch.empty(nel, dtype=torch.float64, device="cuda")
gb = nel * 8 / 1e9


def rate(fn, reps=5):
    fn()
    torch.cuda.synchronize(); G.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize(); G.synchronize()
    return gb * reps / (time.perf_counter() - t0)


print("flat H2D  %.1f GB/s" % rate(lambda: d.copy_(t, non_blocking=True)))
print("flat D2H  %.1f GB/s" % rate(lambda: t.copy_(d, non_blocking=True)))
print("fen push  %.1f GB/s" % rate(lambda: s.push()))
print("fen pull  %.1f GB/s" % rate(lambda: s.pull()))
# both directions at once on two streams (what a double-buffered driver could reach)
t2 = torch.empty(nel, dtype=torch.float64, pin_memory=True)
d2 = torch.empty(nel, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def both():
    with torch.cuda.stream(s1):
        d.copy_(t, non_blocking=True)
    with torch.cuda.stream(s2):
        t2.copy_(d2, non_blocking=True)


print("flat H2D + D2H concurrently  %.1f GB/s each" % rate(both))
G.destroy()
