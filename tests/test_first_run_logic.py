"""The first-run GPU tests (written after the round's GPU budget was spent) have their Python logic exercised on the CPU:
tests/first_run_logic_runner.py runs their bodies against tests/mock_api.py, a stand-in of the fen_b200 API backed by
the oracle.  Both sides of every comparison are then the oracle -- the point is only that no test dies of a typo, a wrong
shape or a wrong loop bound on its first GPU run.  (The long ones -- lid3D, shear drop, Zalesak, viscous decay, the
bubble at the reference's resolutions -- were run the same way once, by hand; they take minutes of numpy.)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_first_run_gpu_tests_are_logically_sound():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "first_run_logic_runner.py")], cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "ALL PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
    assert r.stdout.count("PASS ") >= 25
