import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library. GPU tests must exercise the real CUDA path: a missing library is an error, not a skip."""
    import torch
    from ffr_net_b200 import _lib
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return _lib.load()
