import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# bounded device-side waits (csrc/tc_ptx.cuh): a protocol bug must fail a test within seconds, not after the 30 min
# production budget
os.environ.setdefault("XTAG_SPIN_TIMEOUT_MS", "20000")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) where no CUDA device exists, so a plain
    `pytest tests/` stays green on the CPU build box."""
    try:
        from xtag_clip_b200._cuda_probe import has_device_nodes, wait_for_cuda
        if has_device_nodes():
            wait_for_cuda()          # a fresh GPU box can fail its first cuInit; the failure is sticky per process
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
