import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def gpu_lib():
    import hash10x_b200
    L = hash10x_b200.load_library()
    if L.h10x_gpu_device_count() <= 0:
        pytest.fail("no CUDA device visible: the gpu tests have no CPU fallback")
    return L
