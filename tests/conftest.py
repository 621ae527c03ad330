import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) on a box without CUDA, whatever -m says"""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_names():
    d = os.path.join(ROOT, "tests", "golden")
    return sorted(f[:-4] for f in os.listdir(d) if f.endswith(".npz"))


def load_golden(name):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


@pytest.fixture(scope="session")
def has_cuda():
    import torch
    return torch.cuda.is_available()
