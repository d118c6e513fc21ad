"""One process per GPU over CUDA IPC (the production layout): needs at least two GPUs on the box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("case", ["ppp", "ppn"])
def test_slabs_over_ipc_match_single_rank(case):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (the thread-per-rank tests cover the same code on one)")
    world = 4 if n >= 4 else 2
    port = 29500 + (os.getpid() % 500)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mp_worker.py"), case]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "MP_WORKER_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
