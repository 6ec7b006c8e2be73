import os
import sys

import numpy as np
import pytest

os.environ.setdefault("FSFB_CHECK_STATUS", "1")   # ops.frustum_rows reads the kernel's status word back (one sync per call)

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name: str):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture
def golden():
    return load_golden


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
