"""Synthetic FQB generator bindings (bench / test infrastructure, not the product path).

`libh10xsynth.so` (csrc/synth_gpu.cu + csrc/synth_fqb.h) fills device memory with the same records the host
generator in oracle/synth_cpu.c produces; this module holds the parameter struct both sides share and the
ctypes wrapper of the device generator, so that bench.py's GPU arm does not touch the oracle directory.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("genomeLen", C.c_uint64), ("nBarcodes", C.c_uint32),
                ("pairsMin", C.c_uint32), ("pairsMax", C.c_uint32), ("molPerBarcode", C.c_uint32),
                ("molLen", C.c_uint32), ("snpPeriod", C.c_uint32), ("errThresh", C.c_uint32),
                ("readLen", C.c_uint32)]


def make_params(seed=1, genome_len=200_000, n_barcodes=40, pairs_min=20, pairs_max=120, mol_per_barcode=4,
                mol_len=20_000, snp_period=500, err_rate=0.002, read_len=151):
    return SynthParams(seed, genome_len, n_barcodes, pairs_min, pairs_max, mol_per_barcode, mol_len, snp_period,
                       int(err_rate * 2 ** 32), read_len)


_lib = None


def _load():
    global _lib
    if _lib is None:
        L = C.CDLL(os.path.join(_HERE, "libh10xsynth.so"))
        L.synth_layout_host.restype = C.c_uint64
        L.synth_layout_host.argtypes = [C.POINTER(SynthParams), C.c_void_p]
        L.synth_fqb_device.restype = C.c_int
        L.synth_fqb_device.argtypes = [C.POINTER(SynthParams), C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def layout(p):
    """(total records, record offset of every barcode run + total)"""
    off = np.zeros(p.nBarcodes + 1, np.uint64)
    n = _load().synth_layout_host(C.byref(p), off.ctypes.data)
    return int(n), off


def fill_device(p, off, r0, r1, dev_ptr, stream=None):
    """records r0..r1-1 of the data set into device memory at dev_ptr (30 * (r1 - r0) uint32)"""
    st = _load().synth_fqb_device(C.byref(p), off.ctypes.data, r0, r1, dev_ptr, stream)
    if st:
        raise RuntimeError("synthetic generator failed: cuda error %d" % st)
