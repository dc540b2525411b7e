import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
for _p in (str(ROOT), str(ROOT / "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_train():
    return np.load(GOLDEN / "muscle_train.npz")


@pytest.fixture(scope="session")
def golden_test():
    return np.load(GOLDEN / "muscle_test.npz")


@pytest.fixture(scope="session")
def golden_syn():
    return np.load(GOLDEN / "synthetic.npz")


@pytest.fixture(scope="session")
def golden_adipose():
    return np.load(GOLDEN / "adipose.npz")
