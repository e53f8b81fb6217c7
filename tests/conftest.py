import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def dccm():
    """The product package (directory name is not a Python identifier)."""
    pkg = importlib.import_module("dennou-ccm_b200")
    pkg.build()
    return pkg


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle -- checker only."""
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def gpu(dccm):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    dccm._lib.check(dccm.lib().dccm_init(0))
    return torch.device("cuda:0")
