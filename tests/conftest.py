import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.ref import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref_l1():
    from oracle.ref import RefL1
    if not RefL1.available() and not os.path.isdir("/root/reference"):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return RefL1()


@pytest.fixture(scope="session")
def gold_l1():
    return np.load(os.path.join(GOLD, "l1_small.npz"))


@pytest.fixture(scope="session")
def gold_kpconv():
    return np.load(os.path.join(GOLD, "kpconv_small.npz"))


@pytest.fixture(scope="session")
def gold_encoder():
    return np.load(os.path.join(GOLD, "encoder_small.npz"))


@pytest.fixture(scope="session")
def gold_kpfcnn():
    return np.load(os.path.join(GOLD, "kpfcnn_small.npz"))


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible")
    from apr_b200 import _native
    _native.require_cuda()
    return torch.device("cuda", 0)
