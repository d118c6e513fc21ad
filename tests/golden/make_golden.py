#!/usr/bin/env python
"""Generates tests/golden/*.npz from the CPU oracle (oracle/fen_oracle.py).

The reference (Fortran + MPI + FFTW + 2decomp) cannot be built or imported in this image and ships no golden
fields of its own, so these fixtures are NOT reference outputs: they freeze the oracle's results for fixed,
seeded inputs.  They (1) catch any drift of the oracle itself (tests/test_golden.py, CPU) and (2) give the GPU
parity tests a committed target that does not depend on the oracle at run time.  The oracle's correctness is
pinned separately by the reference's own pass criteria (tests/test_oracle.py).

    python tests/golden/make_golden.py        # rewrites the fixtures in place
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import fen_oracle as fo  # noqa: E402

PI = fo.PI


def ns_case(name, n, bc, L, nu, init, U, steps, g=None, cfl=1.0):
    G = fo.Grid(n[0], n[1], n[2], L[0], L[1], L[2], bc=bc)
    ns = fo.NavierStokes(G, 1.0, nu)
    ns.CFL = cfl
    if g is not None:
        ns.g = list(g)
    init(ns)
    out = {"n": np.array(n), "L": np.array(L), "nu": nu, "U": U, "cfl": cfl, "steps": steps,
           "g": np.array(g if g is not None else [0.0, 0.0, 0.0]), "bc": np.array(bc),
           "u0": ns.v.x.f.copy(), "v0": ns.v.y.f.copy(), "p0": ns.p.f.copy()}
    if G.ndim == 3:
        out["w0"] = ns.v.z.f.copy()
    dt = ns.set_timestep(U)
    out["dt"] = dt
    for s in range(1, steps + 1):
        ns.navier_stokes_solver(s, dt)
    out.update({"u": ns.v.x.f.copy(), "v": ns.v.y.f.copy(), "p": ns.p.f.copy(),
                "maxdiv": ns.maxdiv, "maxCFL": ns.maxCFL})
    if G.ndim == 3:
        out["w"] = ns.v.z.f.copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "maxdiv %.3e maxCFL %.6f" % (ns.maxdiv, ns.maxCFL))


def poisson_case(name, n, bc, seed):
    ndim = 2 if n[2] == 1 else 3
    G = fo.Grid(n[0], n[1], n[2], 1.0, n[1] / n[0], n[2] / n[0], bc=bc)
    rng = np.random.default_rng(seed)
    rhs = rng.standard_normal(n)
    rhs -= rhs.mean()
    phi = fo.Scalar(G, 1)
    phi.I[...] = rhs
    ps = fo.PoissonSolver(phi)
    ps.solve(phi)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), n=np.array(n), bc=np.array(bc), rhs=rhs,
                        sol=phi.I.copy(), variant=ps.variant, ndim=ndim)
    print(name, ps.variant)


if __name__ == "__main__":
    P6, P4 = ["Periodic"] * 6, ["Periodic"] * 4
    ns_case("ns_tgv3d_16_2steps", (16, 16, 16), P6, (2 * PI,) * 3, 0.01, fo.init_tgv3d, 1.0, 2, cfl=0.25)
    ns_case("ns_tgv2d_32_5steps", (32, 32, 1), P4, (2 * PI, 2 * PI, 2 * PI / 32), 1.0, fo.init_tgv2d, 2.0, 5)
    ns_case("ns_channel_16x16x8_3steps", (16, 16, 8), P4 + ["Wall", "Wall"], (2.0, 2.0, 1.0), 0.05,
            fo.init_channel, 1.5, 3, g=(1.0, 0.0, 0.0))
    poisson_case("poisson_ppp_16x8x32", (16, 8, 32), P6, 1)
    poisson_case("poisson_ppn_8x16x16", (8, 16, 16), P4 + ["Wall", "Wall"], 2)
    poisson_case("poisson_pp_32x16", (32, 16, 1), P4, 3)
    poisson_case("poisson_pn_16x32", (16, 32, 1), ["Periodic", "Periodic", "Wall", "Wall"], 4)
