"""test/large_test/isotropic_turbulence/isotropic.f90 on the GPU against the data the reference ships for it
(isotropic.basilisk, isotropic.hit3d -> tests/golden/isotropic_turbulence.npz; the reference only plots its output.txt
over them).  The driver loop is the reference's: before every step it reads v, rewrites the source field
S = 0.1 (v - <v>) on the HOST and hands it back (fen_gpu_pull of v, fen_gpu_push of S: the transfer points INTEGRATION.md
describes for this driver), and the time step follows constant_CFL = .true., CFL = 0.9.

(1) Laminar phase at the reference's own 128^3, t <= 3 (t <= 5 with FEN_TEST_LONG=1, as in the recorded run): the forced ABC flow grows as exp(2 (0.1 - nu) t); Basilisk, hit3d
    and the analytic rate agree to 0.1 % there.  The GPU run must reproduce the growth between t = 1 and the end to 0.5 %
    and the level to 2 %.
(2) The whole run through the transition (the energy peaks near t = 15 - 20 -- the two codes differ by a few time units
    there -- and collapses by t = 35) to t = 60: the mean energy over t in [45, 60] must lie in the band of the
    stationary state both codes settle in (Basilisk 0.178 +- 0.032 over t > 100, hit3d 0.194 +- 0.037): chaotic, so only
    statistics are comparable -- and the velocity stays divergence free.  In the suite this part runs at 64^3 (the host
    side of the driver loop, numpy on 6 x 17 MB arrays per step, makes the 8 733-step 128^3 run a five-minute test);
    FEN_TEST_LONG=1 runs it at 128^3, as recorded in profiles/r02p_pytest_isotropic_128.log: ke(1) 1.7862 (reference
    1.7954), ke(5) 3.6680 (3.6848), peak 18.97, mean over [45, 60] 0.216 inside (0.097, 0.286).
The oracle cannot run this; it is pinned on the same data at 32^3 by
tests/test_oracle.py::test_isotropic_forcing_growth_matches_the_reference_data."""
import os

import numpy as np
import pytest

import fen_b200 as fb
from oracle import fen_oracle as fo
from tests.test_oracle import isotropic_run

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1800)]
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run_case(n, tend, every=1):
    """the driver on an n^3 grid up to tend; returns (time, energy) rows, steps, last (maxdiv, maxCFL), the two curves"""
    ref = np.load(os.path.join(GOLD, "isotropic_turbulence.npz"))
    bas, hit = ref["basilisk"], ref["hit3d"]
    state = {}

    def make(n):
        G = fb.grid().setup(n, n, n, 2 * fo.PI, 2 * fo.PI, 2 * fo.PI)
        ns = fb.Solver(G, 1.0, 1.0e-2).init_solver()
        state["G"], state["ns"] = G, ns

        def pull_v(init=False):
            if init:
                ns.v.push()
                ns.v.update_ghost_nodes()
            else:
                ns.v.pull()
        return ns, ns.S.push, pull_v
    out, steps = isotropic_run(make, n, tend, every=every)
    ns, G = state["ns"], state["G"]
    st = ns.status()
    G.destroy()
    return out, steps, st, bas, hit


def test_isotropic_laminar_growth_at_128():
    tend = 5.0 if os.environ.get("FEN_TEST_LONG") else 3.0
    out, steps, (md, mc), bas, hit = run_case(128, tend + 0.2, every=10)      # samples every 10 steps: run past tend
    ke = lambda t: float(np.interp(t, out[:, 0], out[:, 1]))
    rb = lambda t: float(np.interp(t, bas[:, 0], bas[:, 1]))
    assert abs((ke(tend) / ke(1.0)) / (rb(tend) / rb(1.0)) - 1.0) < 5e-3, (ke(1.0), ke(tend))
    assert abs(ke(1.0) / rb(1.0) - 1.0) < 0.02 and abs(ke(tend) / rb(tend) - 1.0) < 0.02, (ke(1.0), ke(tend))
    assert abs(md) < 1e-10 and mc < 1.0


def test_isotropic_turbulence_reaches_the_stationary_band():
    n = 128 if os.environ.get("FEN_TEST_LONG") else 64
    out, steps, (md, mc), bas, hit = run_case(n, 60.0, every=5)
    # the transition happened (the energy rose well above its start and collapsed), then the stationary band
    assert out[:, 1].max() > 5.0
    m = out[:, 0] > 45.0
    mean = float(out[m, 1].mean())
    sb = bas[bas[:, 0] > 100.0, 1]
    sh = hit[hit[:, 0] > 100.0, 1]
    lo = min(sb.mean() - 2.5 * sb.std(), sh.mean() - 2.5 * sh.std())
    hi = max(sb.mean() + 2.5 * sb.std(), sh.mean() + 2.5 * sh.std())
    assert lo < mean < hi, (mean, lo, hi, steps)
    assert abs(md) < 1e-10 and mc < 1.0
    print("isotropic %d^3: %d steps, peak %.2f mean[45,60] %.4f band (%.3f, %.3f)" % (n, steps, out[:, 1].max(), mean, lo, hi))
