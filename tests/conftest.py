"""Shared fixtures. GPU tests are marked @pytest.mark.gpu and call the product through its C ABI
(ctypes over libdfb_b200.so); CPU tests cover the oracle, the host logic and the ABI surface."""
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "pointcloud-depthfusion_b200"))
sys.path.insert(0, os.path.join(REPO, "oracle"))
sys.path.insert(0, os.path.join(REPO, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def repo_root():
    return REPO


@pytest.fixture(scope="session")
def oracle():
    import pyoracle

    pyoracle.load()
    pyoracle.load(fast=True)
    return pyoracle


@pytest.fixture(scope="session")
def dfb():
    import dfb_b200

    dfb_b200.load()
    return dfb_b200


@pytest.fixture(scope="session")
def gpu(dfb):
    """Initialises device 0 through the C ABI; fails (does not skip) when the CUDA path is unusable."""
    import torch

    assert torch.cuda.is_available(), "GPU test selected but no CUDA device is visible"
    torch.cuda.set_device(0)
    dfb.init(0)
    return torch
