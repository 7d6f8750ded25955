import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with -m gpu")


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library; built on demand so the CPU suite also covers 'does it build'."""
    import __graft_entry__ as g
    g.build()
    import ss4k_b200
    return ss4k_b200._lib.load()


@pytest.fixture(scope="session")
def engine(lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ss4k_b200
    return ss4k_b200.Engine.get(0)
