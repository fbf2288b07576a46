"""Multi-GPU parity (gpu): runs tests/dist_worker.py on every visible GPU (up to 8) through torchrun;
on a one-GPU box it still runs the distributed code path with a single rank (NCCL self-exchange)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, extra_env=None):
    env = dict(os.environ)
    env.setdefault("NCCL_DEBUG", "WARN")
    env.update(extra_env or {})
    if world == 1:
        cmd = [sys.executable, os.path.join(ROOT, "tests", "dist_worker.py")]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
               "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "dist_worker.py")]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)


def test_distributed_build_single_rank(gpu_lib):
    r = _run(1)
    assert r.returncode == 0 and "DIST_PARITY_OK world=1" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


# the paths of the multi-GPU build: default = hand-written tail on every rank, owners merge their sorted runs, ranges cut at
# the density quantiles, triples pushed into peer memory; then each alternative that stays in the library as a fallback
VARIANTS = {
    "default": {},
    "library-tail+owner-sort+flat-ranges": {"H10X_LEGACY_TAIL": "1", "H10X_OWNER_SORT": "1", "H10X_FLAT_OWNERS": "1"},
    "nccl-send-recv": {"H10X_NO_PEER_PUSH": "1"},
    "copy-engines": {"H10X_PEER_COPY": "1"},
    "library-tail+nccl": {"H10X_LEGACY_TAIL": "1", "H10X_NO_PEER_PUSH": "1"},
    # a rank-local failure INSIDE the bin exchange (in the stretch before the 2nd / 3rd / 4th agreement of dist_bins: receive
    # arrays, owner merge, ids) must surface on all ranks instead of leaving the peers in a collective
    "fail-before-agreement-2": {"H10X_DIST_FAIL": "0:2"},
    "fail-before-agreement-3": {"H10X_DIST_FAIL": "0:3"},
    "fail-before-agreement-4": {"H10X_DIST_FAIL": "0:4"},
}


@pytest.mark.parametrize("variant", ["library-tail+owner-sort+flat-ranges", "fail-before-agreement-3"])
def test_distributed_build_single_rank_variants(gpu_lib, variant):
    r = _run(1, VARIANTS[variant])
    assert r.returncode == 0 and "DIST_PARITY_OK world=1" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_distributed_build_all_gpus(gpu_lib, variant):
    n = gpu_lib.h10x_gpu_device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(n, 8)
    env = dict(VARIANTS[variant])
    if "H10X_DIST_FAIL" in env:
        env["H10X_DIST_FAIL"] = "%d:%s" % (world - 1, env["H10X_DIST_FAIL"].split(":")[1])
    r = _run(world, env)
    assert r.returncode == 0 and ("DIST_PARITY_OK world=%d" % world) in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
