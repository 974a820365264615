import os
import sys

# In-process multi-rank topologies run W spinning kernels concurrently on W streams; the default of 8
# hardware work queues would alias two streams onto one queue at W = 8 and serialise them.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# Exercise the TMA-staged local pass at test sizes too (the library only defaults to it from 256 MiB up).
os.environ.setdefault("B2_LOCAL_TMA_MIN_MB", "0")

# The local_cuda app registry defaults to ~/.torchx_b200: keep test runs out of the home directory.
import tempfile  # noqa: E402

os.environ.setdefault("TORCHX_HOME", tempfile.mkdtemp(prefix="torchx_b200_test_home_"))

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import pytest  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_count() -> int:
    try:
        import torch

        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.fixture(scope="session")
def cuda_count() -> int:
    return _cuda_count()


def pytest_collection_modifyitems(config, items):
    if _cuda_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
