"""hash10x_b200 - hash10x's `--readFQB` minhash index build on NVIDIA B200 (sm_100a).

The product is native: `libh10xgpu.so` (hand-written CUDA kernels behind the C ABI declared in
include/h10x_gpu.h) and `bin/hash10x-b200` (a C host program that keeps hash10x's command
chaining).  This Python package is only the harness-side binding used by tests/ and bench.py; it
calls the same C ABI through ctypes and never computes anything itself.  There is no CPU
fallback: importing works anywhere, but every build call raises without the library and a GPU.
"""
from .binding import (Hash10xGPU, H10xError, Index, Params, lib_path, load_library,  # noqa: F401
                      factor1_from_seed, DEFAULT_FACTOR1, FLAG_WIDE_B, FLAG_NO_TABLE, FLAG_NO_CODES,
                      FLAG_GENERIC_ONLY, FLAG_LEGACY_TAIL, FLAG_LAZY_CODES)
