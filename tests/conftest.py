import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gpu_lib():
    """The CUDA library, built in-tree.  GPU tests fail loudly if it is missing."""
    import restartsqp_b200 as r
    L = r.capi.lib()
    n = L.sqpb200_device_count()
    assert n > 0, "no CUDA device visible (sqpb200_device_count=%d)" % n
    return L
