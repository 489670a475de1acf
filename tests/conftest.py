import os
import sys

import pytest

# one hardware work queue per stream: tests that drive several group members from one process keep a flag-waiting kernel
# in one stream while another stream has to make progress (must be set before CUDA initialises)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def small_world():
    """Small synthetic corpus + HNSW files + mlp weights + queries shared by CPU and GPU tests."""
    from tests import util
    return util.small_world()
