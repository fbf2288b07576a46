"""GPU parity against the committed golden digests (generated from the reference binary by
tests/golden/make_golden.py) and at a larger size through size-independent properties."""
import ctypes as C
import json
import os
import zlib

import numpy as np
import pytest

import hashfile

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
ROOT = os.path.dirname(HERE)


def _cases():
    with open(os.path.join(GOLD, "golden.json")) as f:
        return sorted(json.load(f).items())


def _crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


@pytest.mark.parametrize("name,d", _cases())
def test_gpu_reproduces_reference_golden(gpu_lib, tmp_path, name, d):
    import hash10x_b200
    kw = d["params"]
    with hash10x_b200.Hash10xGPU(B=kw.get("B", 20), k=kw.get("k", 21), w=kw.get("w", 31), r=kw.get("r", 17),
                                 N=kw.get("N", 0), chunkSize=kw.get("chunk", 100000)) as g:
        hp = str(tmp_path / "g.hash")
        g.build_file_to_hash(os.path.join(GOLD, name + ".fqb"), hp)
    hf = hashfile.parse(hp)
    assert hf.size == d["fileSize"] and (hf.depthDim, hf.blkDim) == (d["depthDim"], d["blkDim"])
    assert (hf.hashNumber, hf.nBlocksMax, hf.nHashes) == (d["hashNumber"], d["nBlocksMax"], d["nHashes"])
    assert _crc(hf.hashValue) == d["crc_hashValue"] and _crc(hf.hashDepth[:hf.depthMax]) == d["crc_hashDepth"]
    assert _crc(hf.hashIndex) == d["crc_hashIndex"]
    assert _crc(hf.blkNRead) == d["crc_blkNRead"] and _crc(hf.blkNHash) == d["crc_blkNHash"]
    assert _crc(hf.clusIdx) == d["crc_clusIdx"] and _crc(hf.clusRead) == d["crc_clusRead"]


def test_gpu_generator_matches_cpu_generator(orc, gpu_lib):
    import torch
    synth = C.CDLL(os.path.join(ROOT, "hash10x_b200", "libh10xsynth.so"))
    synth.synth_fqb_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    for read_len in (151, 160):
        p = orc.synth_params(seed=21, n_barcodes=300, pairs_min=1, pairs_max=40, read_len=read_len)
        n, off = orc.synth_layout(p)
        want = orc.synth_fqb(p)
        buf = torch.empty(n * 30, dtype=torch.int32, device="cuda")
        assert synth.synth_fqb_device(C.byref(p), off.ctypes.data, 0, n, buf.data_ptr(), None) == 0
        got = buf.cpu().numpy().view(np.uint32).reshape(n, 30)
        assert np.array_equal(got, want)


def test_large_build_properties(orc, gpu_lib):
    """~1.2M pairs (too slow for the Python-side strict compare to be the only check): the invariants
    the reference's -DCHECK tests (hash10x.c:341-345) and more, plus a strict compare with the oracle."""
    import hash10x_b200
    p = orc.synth_params(seed=33, n_barcodes=4000, pairs_min=100, pairs_max=500, genome_len=30_000_000,
                         mol_len=50_000, mol_per_barcode=8, err_rate=0.001)
    recs = orc.synth_fqb(p)
    with hash10x_b200.Hash10xGPU(B=24) as g:
        got = g.build_host(recs)
        st = g.stats()
    hn = got.hashNumber
    assert st["fusedBlocks"] == st["nBlocks"] - 1
    # depth = length of each hash->code list, lists ascending, block lists sorted by id and duplicate-free
    assert np.array_equal(np.diff(got.codeOff.astype(np.int64)), got.hashDepth)
    assert int(got.hashDepth.sum()) == got.nHashes == int(got.blkNHash.sum())
    seg = np.repeat(np.arange(hn), got.hashDepth)
    same = seg[1:] == seg[:-1]
    assert (np.diff(got.codes.astype(np.int64))[same] > 0).all()
    ids = (got.clus & np.uint64(0xFFFFFFFF)).astype(np.int64)
    blk = np.repeat(np.arange(got.nBlocksMax), got.blkNHash)
    sameb = blk[1:] == blk[:-1]
    assert (np.diff(ids)[sameb] > 0).all() and ids.min() >= 1 and ids.max() == hn - 1
    # ids are handed out in (first block, hash) order: first-occurrence ids are increasing
    first = np.full(hn, np.iinfo(np.int64).max)
    np.minimum.at(first, ids, blk)
    assert (np.diff(first[1:]) >= 0).all()
    hv = got.hashValue[1:].astype(np.uint64)
    grp = first[1:]
    sameg = grp[1:] == grp[:-1]
    assert (hv[1:][sameg] > hv[:-1][sameg]).all()
    assert (got.hashValue[1:] % np.uint64(31) == 0).all()
    hashfile.check_table(hashfile.from_index(got))
    want = orc.build(recs, B=24)
    hashfile.assert_strict_equal(hashfile.from_index(want), hashfile.from_index(got), table=True)
    assert np.array_equal(got.codes, want.codes)


def test_scale_invariants_device_resident(orc, gpu_lib):
    """5M read pairs generated on the GPU and built from device memory (the bench path): no oracle at this
    size, so the checks are the size-independent ones - the reference's -DCHECK invariant (hash10x.c:341-345),
    sortedness, id order, table reachability - plus agreement of the host-buffer entry point."""
    import ctypes as C
    import os
    import torch
    import hash10x_b200
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    synth = C.CDLL(os.path.join(root, "hash10x_b200", "libh10xsynth.so"))
    synth.synth_fqb_device.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    p = orc.synth_params(seed=91, n_barcodes=12500, pairs_min=300, pairs_max=500, genome_len=100_000_000,
                         mol_per_barcode=10, mol_len=50_000, snp_period=1000, err_rate=0.0005, read_len=160)
    n, off = orc.synth_layout(p)
    fqb = torch.empty(n * 30, dtype=torch.int32, device="cuda")
    assert synth.synth_fqb_device(C.byref(p), off.ctypes.data, 0, n, fqb.data_ptr(), None) == 0
    with hash10x_b200.Hash10xGPU(B=26) as g:
        g.build_device(fqb.data_ptr(), n)
        ix = g.download()
        st = g.stats()
        host = fqb.cpu().numpy().view(np.uint32)
        hn2, nh2, nb2 = g.build_host(host, want_index=False)
    assert (hn2, nh2, nb2) == (ix.hashNumber, ix.nHashes, ix.nBlocksMax)
    hn = ix.hashNumber
    assert st["fusedBlocks"] == ix.nBlocksMax - 2 and ix.blkNHash[-1] == 0
    assert np.array_equal(np.diff(ix.codeOff.astype(np.int64)), ix.hashDepth)
    assert int(ix.hashDepth.sum()) == ix.nHashes == int(ix.blkNHash.sum())
    ids = (ix.clus & np.uint64(0xFFFFFFFF)).astype(np.int64)
    assert (ix.clus >> np.uint64(48)).max() == 0
    blk = np.repeat(np.arange(ix.nBlocksMax), ix.blkNHash)
    sameb = blk[1:] == blk[:-1]
    assert (np.diff(ids)[sameb] > 0).all() and ids.min() >= 1 and ids.max() == hn - 1
    seg = np.repeat(np.arange(hn), ix.hashDepth)
    assert (np.diff(ix.codes.astype(np.int64))[seg[1:] == seg[:-1]] > 0).all()
    first = ix.codes[ix.codeOff[1:-1].astype(np.int64)].astype(np.int64)       # first block of every bin
    assert (np.diff(first) >= 0).all()
    hv = ix.hashValue[1:]
    sameg = first[1:] == first[:-1]
    assert (hv[1:][sameg] > hv[:-1][sameg]).all() and (hv % np.uint64(31) == 0).all()
    assert np.unique(hv).size == hv.size
    hashfile.check_table(hashfile.from_index(ix))
