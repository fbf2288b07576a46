"""CPU test: the division-free arithmetic of hash10x_b200/csrc/h10x_common.cuh against plain % and /,
compiled as host code with nvcc (tests/hostcheck/arith_check.cu; no kernel is launched)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_divisibility_quotient_and_canonical_min(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "arith_check")
    subprocess.run([nvcc, "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "hostcheck", "arith_check.cu")],
                   check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok "), r.stdout + r.stderr
