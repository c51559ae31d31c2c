import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def gpu_host():
    """ntrace_b200 host layer initialised on cuda:0; fails loudly if the extension or the GPU is missing."""
    import torch
    assert torch.cuda.is_available(), "gpu-marked test started without a CUDA device"
    from ntrace_b200 import host
    host.init(0)
    return host
