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


def mf_case(name, Nx, Ny, sigma, steps):
    """Two-phase wave (density ratio 850, gravity, optional surface tension): freezes the two-phase oracle
    (oracle/fen_oracle_mf.py).  The initial vof, u, v are stored so that the GPU test needs no oracle."""
    from oracle import fen_oracle_mf as mf
    import math
    Lx, Ly = 1.0, float(Ny) / Nx
    G = fo.Grid(Nx, Ny, 1, Lx, Ly, Lx / Nx, bc=["Periodic", "Periodic", "Wall", "Wall"])
    rho_0 = 1000.0
    mu_0 = rho_0 * Lx * math.sqrt(mf.GRAVITY * Lx) / 1.0e4
    ns = mf.MultiphaseNavierStokes(G, rho_0, rho_0 / 850.0, mu_0, mu_0 * 1.9e-2, sigma,
                                   distance=lambda x, y: y - 0.05 * np.cos(2.0 * PI * x / Lx) - Ly / 2.0)
    ns.g[1] = -mf.GRAVITY
    i = np.arange(1, Nx + 1)[:, None]
    j = np.arange(1, Ny + 1)[None, :]
    d = G.delta
    # init_velocity of viscous_decay.f90:104-131 (potential-flow wave, amplitude 0.05)
    wn = 2.0 * PI / Lx
    om = math.sqrt(mf.GRAVITY * wn)
    F = ns.vof.sh
    x, y = i * d, (j - 0.5) * d - Ly / 2.0
    f = ((F(1, 0) + F()) * 0.5)[..., 0]
    ns.v.x.I[..., 0] = (1.0 - f) * 0.05 * om * np.exp(wn * y) * np.cos(wn * x) - f * 0.05 * om * np.exp(-wn * y) * np.cos(wn * x)
    x, y = (i - 0.5) * d, j * d - Ly / 2.0
    f = ((F(0, 1) + F()) * 0.5)[..., 0]
    ns.v.y.I[..., 0] = (1.0 - f) * 0.05 * om * np.exp(wn * y) * np.sin(wn * x) + f * 0.05 * om * np.exp(-wn * y) * np.sin(wn * x)
    ns.v.update_ghost_nodes()
    dt = 0.1 * ns.set_timestep(1.0)          # as the reference's driver does (viscous_decay.f90:56-57); dt_o stays
    ins = {"vof0": ns.vof.f.copy(), "u0": ns.v.x.f.copy(), "v0": ns.v.y.f.copy(),
           "props": np.array([ns.rho_0, ns.rho_1, ns.mu_0, ns.mu_1])}
    for s in range(1, steps + 1):
        ns.navier_stokes_solver(s, dt)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), n=np.array([Nx, Ny]), sigma=sigma, steps=steps, dt=dt,
                        u=ns.v.x.f.copy(), v=ns.v.y.f.copy(), p=ns.p.f.copy(), vof=ns.vof.f.copy(),
                        rho=ns.rho.f.copy(), curv=ns.vf.curv.f.copy(), maxdiv=ns.maxdiv, maxCFL=ns.maxCFL, **ins)
    print(name, "maxdiv %.3e maxCFL %.6f" % (ns.maxdiv, ns.maxCFL))


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
    mf_case("mf_wave_16x32_3steps", 16, 32, 0.07, 3)
