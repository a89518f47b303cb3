import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def build_product():
    from radiativetransfer_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def oracle():
    from oracle import ftte_oracle
    ftte_oracle.lib()
    return ftte_oracle


@pytest.fixture(scope="session")
def uvbg():
    from radiativetransfer_b200 import workloads
    return workloads.uvb_background(3.0)


def rel_err(a, b, floor=1e-290):
    """per-cell relative error max|a-b| / max(|b|, floor); `floor` keeps fp64 subnormals out of the ratio"""
    import numpy as np
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))
