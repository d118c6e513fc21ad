import sys; sys.path.insert(0, '.')
import numpy as np
import fen_b200 as fb
from oracle import fen_oracle as fo
from tests.test_gpu_parity import _setup_ns, rel_l2
PI = fo.PI

def run(name, n, bc, L, nu, init, U, g, nsteps, dtdiv=1.0):
    Go, Gg, nso, nsg, dt = _setup_ns(n, bc, 3, L, nu, init, U, g=g)
    dt = dt / dtdiv
    print(name, "dt", dt)
    for step in range(1, nsteps + 1):
        nso.navier_stokes_solver(step, dt)
        nsg.navier_stokes_solver(step, dt)
        if step in (1, 2, 3, 5, 10, 20, 50, 100):
            nsg.v.pull(); nsg.p.pull(); nsg.phi.pull()
            e = [rel_l2(a.f, b.f) for a, b in ((nsg.v.x, nso.v.x), (nsg.v.y, nso.v.y), (nsg.v.z, nso.v.z), (nsg.p, nso.p), (nsg.phi, nso.phi))]
            ea = [np.abs(a.f - b.f).max() for a, b in ((nsg.v.x, nso.v.x), (nsg.v.y, nso.v.y), (nsg.v.z, nso.v.z), (nsg.p, nso.p))]
            print(step, ["%.2e" % x for x in e], "abs", ["%.2e" % x for x in ea], "maxdiv %.2e %.2e" % (nsg.maxdiv, nso.maxdiv), flush=True)
    Gg.destroy()

run("tgv64 full dt", (64, 64, 64), ["Periodic"] * 6, 2 * PI, 0.01, fo.init_tgv3d, 1.0, None, 100)
run("tgv64 dt/8", (64, 64, 64), ["Periodic"] * 6, 2 * PI, 0.01, fo.init_tgv3d, 1.0, None, 100, 8.0)
run("channel", (32, 32, 16), ["Periodic"] * 4 + ["Wall", "Wall"], 2.0, 0.05, fo.init_channel, 1.0, (1.0, 0, 0), 20)
