"""Runs the bodies of the first-run GPU tests against tests/mock_api.py (see there).  Executed in a child process by
tests/test_first_run_logic.py because it swaps the attributes of the fen_b200 package for the stand-ins."""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import fen_b200  # noqa: E402
from tests import mock_api  # noqa: E402

for name in ("grid", "scalar", "vector", "Solver", "MultiphaseSolver", "VoF", "PoissonSolver", "gradient", "laplacian",
             "face_to_center", "curl", "FenError"):
    setattr(fen_b200, name, getattr(mock_api, name))

import tests.test_gpu_zy_any_length as ty  # noqa: E402
import tests.test_gpu_zz_rising_bubble as tz  # noqa: E402
import tests.test_gpu_parity as tp  # noqa: E402
import tests.test_gpu_zx_isotropic as tx  # noqa: E402

failed = 0


def run(fn, *args):
    global failed
    t0 = time.time()
    try:
        fn(*args)
        print("PASS %s %s %.1fs" % (fn.__name__, args[:2], time.time() - t0), flush=True)
    except Exception as exc:                                     # noqa: BLE001
        failed += 1
        tb = traceback.extract_tb(exc.__traceback__)[-1]
        print("FAIL %s %s %s: %s (line %d: %s)" % (fn.__name__, args[:2], type(exc).__name__, str(exc)[:120], tb.lineno,
                                                    tb.line), flush=True)


for case in ty.ANY_CASES:
    if max(case[1]) <= 128:
        run(ty.test_poisson_any_length_matches_oracle, *case)
run(ty.test_unsupported_lengths_are_rejected_loudly)
run(ty.test_steps_tgv2d_96_match_oracle)
run(ty.test_steps_tgv3d_24x48x24_match_oracle)
run(ty.test_steps_cavity_48x40_match_oracle)
run(ty.test_scalar_laplacian_face_to_center_curl, (16, 12, 8), 3)
run(ty.test_scalar_laplacian_face_to_center_curl, (32, 16, 1), 2)
run(ty.test_poiseuille_inflow_outflow_steps_match_oracle)
run(tz.test_rising_bubble_steps_match_oracle)
run(tp.test_one_step_512_matches_c_oracle, 64)
run(tx.run_case, 16, 0.5, 2)
run(tz.test_rising_bubble_rises_and_keeps_its_volume)
print("FAILED %d" % failed if failed else "ALL PASS")
sys.exit(1 if failed else 0)
