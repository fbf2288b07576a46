"""Parity at the sizes bench.py times: the device build's index digests
(h10x_gpu_index_digest, hash10x_b200/csrc/h10x_digest.h) against tests/golden/golden_scale.json, which
tests/golden/make_golden_scale.py produced from the UNMODIFIED reference binary's `.hash` of the same data set;
and on small inputs against the oracle's arrays.  The synthetic FQB is generated in HBM by the same closed-form
record function the golden generator used on the host."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "golden_scale.json")


def _golden(key):
    with open(GOLDEN) as f:
        g = json.load(f)
    if key not in g:
        pytest.skip("no golden digests for %s" % key)
    return g[key]


def check_against_golden(dg, stats, gold, table=True):
    """dg: dict from Hash10xGPU.digest(); gold: one entry of golden_scale.json"""
    assert stats["nHashes"] == gold["nHashes"] and stats["nBins"] + 1 == gold["hashNumber"]
    assert stats["nBlocks"] + 1 == gold["nBlocksMax"] and stats["nRecords"] == gold["nReads"]
    names = ["hashValue", "hashDepth", "blkNRead", "blkNHash", "clusHash"] + (["hashIndex"] if table else [])
    for n in names:
        assert "%016x" % dg[n] == gold["dg_" + n], n
    if dg["haveCodes"]:
        assert dg["codesMissing"] == 0 and dg["codesUnordered"] == 0


def _build_workload(name, flags=0):
    import torch
    import bench
    import hash10x_b200
    from hash10x_b200 import synth as gsynth
    wl = bench.WORKLOADS[name]
    p = bench.synth_params(gsynth, wl, seed=3)
    n, off = gsynth.layout(p)
    fqb = torch.empty(n * 30, dtype=torch.int32, device="cuda:0")
    gsynth.fill_device(p, off, 0, n, fqb.data_ptr())
    torch.cuda.synchronize()
    with hash10x_b200.Hash10xGPU(B=wl["B"], device=0, flags=flags) as g:
        g.build_device(fqb.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
        st = g.stats()
        dg = g.digest()
    del fqb
    torch.cuda.empty_cache()
    return dg, st


def test_digest_equals_oracle_arrays(orc, gpu_lib):
    import hash10x_b200
    p = orc.synth_params(seed=21, n_barcodes=60, pairs_min=5, pairs_max=250)
    recs = orc.synth_fqb(p)
    want = orc.build(recs, B=21)
    wd = orc.index_digests(want)
    for flags in (0, hash10x_b200.FLAG_LEGACY_TAIL):
        with hash10x_b200.Hash10xGPU(B=21, device=0, flags=flags) as g:
            g.build_host(recs)
            dg = g.digest()
            assert g.stats()["tailPath"] == (1 if flags else 2)
        for k, v in wd.items():
            assert dg[k] == v, (k, flags)
        assert dg["codesMissing"] == 0 and dg["codesUnordered"] == 0 and dg["haveCodes"] == 1


@pytest.mark.parametrize("name", ["yeast", "gb10th"])
def test_workload_digests_match_reference(gpu_lib, name):
    import hash10x_b200
    gold = _golden(name)
    for flags in (0, hash10x_b200.FLAG_LEGACY_TAIL):
        dg, st = _build_workload(name, flags)
        check_against_golden(dg, st, gold)


def test_1gb_digests_match_reference(gpu_lib):
    """BASELINE configs[2], the workload bench.py times at N=1: 200M read pairs, -B 28"""
    gold = _golden("1gb")
    dg, st = _build_workload("1gb")
    assert st["tailPath"] == 2
    check_against_golden(dg, st, gold)


def test_human8_digests_match_reference(gpu_lib):
    """one GPU's share of BASELINE configs[3] (75M read pairs, -B 30): bench.py's `weak_base`"""
    gold = _golden("human8")
    dg, st = _build_workload("human8")
    check_against_golden(dg, st, gold)
