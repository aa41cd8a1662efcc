import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, f"ref_{name}.npz")).items()}


@pytest.fixture(scope="session")
def ckpt():
    from puzzlefusion_plusplus_b200 import synthetic
    return synthetic.make_checkpoints(0)


@pytest.fixture(scope="session")
def lib():
    from puzzlefusion_plusplus_b200 import _lib
    return _lib
