import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The built libraries are git-ignored: on a fresh checkout build them once (nvcc cross-compiles without
    a GPU); on the GPU box the snapshot already carries them."""
    need = [os.path.join(ROOT, "hash10x_b200", "libh10xgpu.so"), os.path.join(ROOT, "hash10x_b200", "libh10xsynth.so"),
            os.path.join(ROOT, "hash10x_b200", "bin", "hash10x-b200"), os.path.join(ROOT, "oracle", "liborc.so")]
    if all(os.path.exists(p) for p in need):
        return
    import subprocess
    try:
        subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.build()"], cwd=ROOT, check=True,
                       stdout=subprocess.DEVNULL)
    except Exception as e:     # the individual tests will say what is missing
        print("conftest: build failed: %r" % (e,), file=sys.stderr)


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
