import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run on the B200 box)")


@pytest.fixture(autouse=True)
def _seed():
    np.random.seed(42)
    try:
        import torch
        torch.manual_seed(42)
    except Exception:
        pass


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a CUDA device: the gpu-marked tests are skipped, not failed."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA GPU (B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
