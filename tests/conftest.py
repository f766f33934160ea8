"""pytest configuration: `gpu` marker + shared fixtures.

`-m "not gpu"` covers the oracle against the golden vectors, host logic and the C-ABI symbol check (no compute
calls without a GPU); `-m gpu` are the parity tests proper, calling the CUDA library through the C-ABI.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def small_golden():
    return dict(np.load(os.path.join(GOLDEN, "small_dtt.npz")))


@pytest.fixture(scope="session")
def r8_golden():
    return dict(np.load(os.path.join(GOLDEN, "r8_topic.npz")))
