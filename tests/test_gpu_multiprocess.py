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


def test_1024_point_lines_on_8_gpus_match_single_rank():
    """world = 8 with the 1024-point y and z lines of the 8-GPU bench (16 KB bulk stores to 8 real peers over NVLink):
    same bits as one rank.  Needs an 8-GPU box; the thread-per-rank form of the same case runs on one device
    (tests/test_gpu_multirank.py::test_poisson_1024_point_lines_on_8_ranks)."""
    if _ngpu() < 8:
        pytest.skip("needs 8 GPUs")
    port = 29500 + (os.getpid() % 500)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "8",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mp_worker.py"), "ppp1024"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "MP_WORKER_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
