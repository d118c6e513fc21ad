"""Generates tests/golden/lid3d_re1000_64.npz: the reference's 3-D lid-driven cavity (test/large_test/lid3D/main.f90)
replayed by the CPU oracle to the driver's own steady-state criterion, next to the data set the reference ships for it.

  * Uref.csv / Vref.csv (Ku et al. 1987: centreline u(y) and v(x) of the cubic cavity at Re = 1000) are copied as DATA
    (/root/reference does not exist on the GPU box);
  * the oracle run: 64^3, walls on the six faces (nnn Poisson: DCT in x and y, Thomas in z), viscosity = 1 / Re,
    v%x%bc%top = U, dt = set_timestep(U) / 2 (:53-57), time loop until max |v - v_old| < 1e-8 (:64-93).  Stored: the
    two centreline profiles exactly as postpro.py:48-52 forms them, the step count, the last change and the state of
    the run after 50 steps (u on the mid-plane) so that a short replay can be checked bit-tightly.

About 25 minutes of numpy on one core (64^3 x 7 680 steps to t = 60); run it in the build container:
    python tests/golden/make_lid3d.py [N] [t_end]
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import fen_oracle as fo  # noqa: E402

REF = "/root/reference/test/large_test/lid3D"


def centrelines(ns, N):
    u, v = ns.v.x.I, ns.v.y.I
    uc = 0.5 * (u[N // 2, :, N // 2] + u[N // 2 - 1, :, N // 2])          # postpro.py:48-49
    vc = 0.5 * (v[:, N // 2, N // 2] + v[:, N // 2 - 1, N // 2])          # postpro.py:51-52
    return uc.copy(), vc.copy()


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    # The driver runs until max |v - v_old| < 1e-8 or t = 2000 (:64-93).  At 64^3 the change per step is 3e-6 at
    # t = 40 and 1.2e-7 at t = 60 (the weak corner vortices of the cubic cavity settle slowly): the centreline profiles
    # the reference plots are converged to plotting accuracy long before, so the fixture stops at t_end (default 60).
    t_end = float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
    tol = 1.0e-8
    uref = np.genfromtxt(os.path.join(REF, "Uref.csv"), delimiter=",", skip_header=1)     # columns u, y
    vref = np.genfromtxt(os.path.join(REF, "Vref.csv"), delimiter=",", skip_header=1)     # columns x, v
    G = fo.Grid(N, N, N, 1.0, 1.0, 1.0 / N, bc=["Wall"] * 6)
    ns = fo.NavierStokes(G, 1.0, 1.0e-3)
    assert ns.poisson.variant == "nnn"
    ns.v.x.bc["top"][...] = 1.0
    dt = ns.set_timestep(1.0) / 2.0
    old = [c.f.copy() for c in ns.v.comps]
    step, t0, early = 0, time.time(), {}
    while True:
        step += 1
        ns.navier_stokes_solver(step, dt)
        diff = max(np.abs(c.f - o).max() for c, o in zip(ns.v.comps, old))
        for c, o in zip(ns.v.comps, old):
            o[...] = c.f
        if step == 50:
            early = {"u50_mid": ns.v.x.I[:, :, N // 2].copy(), "p50_mid": ns.p.I[:, :, N // 2].copy()}
        if step % 500 == 0:
            print(step, step * dt, diff, "%.0f s" % (time.time() - t0), flush=True)
        done = diff < tol or step * dt >= t_end
        if done or step % 500 == 0:
            uc, vc = centrelines(ns, N)
            np.savez_compressed(os.path.join(HERE, "lid3d_re1000_%d.npz" % N), N=N, dt=dt, steps=step, time=step * dt,
                                last_change=diff, maxdiv=ns.maxdiv, uc=uc, vc=vc, uref=uref, vref=vref, **early)
        if done:
            break
    print("steps", step, "time", step * dt, "last change", diff, "maxdiv", ns.maxdiv, "wrote lid3d_re1000_%d.npz" % N)
