"""Multi-GPU parity (gpu): runs tests/dist_worker.py on every visible GPU (up to 8) through torchrun;
on a one-GPU box it still runs the distributed code path with a single rank (NCCL self-exchange)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world):
    env = dict(os.environ)
    env.setdefault("NCCL_DEBUG", "WARN")
    if world == 1:
        cmd = [sys.executable, os.path.join(ROOT, "tests", "dist_worker.py")]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
               "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "dist_worker.py")]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)


def test_distributed_build_single_rank(gpu_lib):
    r = _run(1)
    assert r.returncode == 0 and "DIST_PARITY_OK world=1" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_distributed_build_all_gpus(gpu_lib):
    n = gpu_lib.h10x_gpu_device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(n, 8)
    r = _run(world)
    assert r.returncode == 0 and ("DIST_PARITY_OK world=%d" % world) in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
