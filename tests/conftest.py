import os
import sys

import pytest

# kernels of different ranks must be able to start while a flag-wait kernel spins (tests/ranks.py runs several
# ranks on one device): no lazy module loading.  Must be set before the CUDA runtime initialises.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
# the C/OpenMP oracles: threads sleep at barriers instead of spinning, so a busy neighbour on the machine (or pytest-xdist)
# costs a little instead of a factor of fifty.  Must be set before libgomp is loaded (numpy / torch pull it in).
os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU oracle check")


def pytest_collection_modifyitems(config, items):
    """GPU tests must never silently pass on a box without a device."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
