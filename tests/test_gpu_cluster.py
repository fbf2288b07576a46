"""GPU parity for the "next" row f2: `--hashDepthRange` + `--cluster` through the C ABI
(h10x_gpu_depth_range, h10x_gpu_cluster) against the oracle's restatement of codeClusterFind + codeClusterReadMerge
(hash10x.c:770-868), which tests/test_oracle.py pins against the reference binary.  Bit-exact: every subCluster byte,
nSubCluster, and pointToMin as IEEE doubles."""
import os
import subprocess

import numpy as np
import pytest

import hashfile
from test_oracle import CLUSTER_CASES, _cluster_case

pytestmark = pytest.mark.gpu


def _gpu(**kw):
    import hash10x_b200
    return hash10x_b200.Hash10xGPU(**kw)


def _same(got, want):
    clus, nsub, ptm = got[:3]
    wclus, wnsub, wptm = want
    assert np.array_equal(nsub, wnsub)
    assert np.array_equal(ptm.view(np.uint64), wptm.view(np.uint64))
    assert np.array_equal(clus, wclus)


@pytest.mark.parametrize("seed,nb,pmin,pmax,genome,mol,mpb,dmin,dmax,thr,cmin,cmax", CLUSTER_CASES)
def test_cluster_matches_oracle(orc, gpu_lib, seed, nb, pmin, pmax, genome, mol, mpb, dmin, dmax, thr, cmin, cmax):
    recs = _cluster_case(orc, seed, nb, pmin, pmax, genome, mol, mpb)
    ix = orc.build(recs, B=20)
    _w, goff, good = orc.good_hashes(ix, dmin, dmax)
    want = orc.cluster(ix, goff, good, cmin, cmax, thr)
    assert int(want[1].sum()) > 0
    with _gpu(B=20) as g:
        g.build_host(recs)
        _gw, ggoff, ggood = g.depth_range(dmin, dmax)
        assert np.array_equal(ggoff, goff) and np.array_equal(ggood, good)
        got = g.cluster(cmin, cmax, thr)
        _same(got, want)
        assert got[3] > 0.0
        # the resident ClusterHash array was updated in place: a fresh download carries the labels
        assert np.array_equal(g.download().clus, want[0])


def test_cluster_twice_carries_state(orc, gpu_lib):
    recs = _cluster_case(orc, 36, 250, 150, 300, 100_000, 10_000, 6)
    ix = orc.build(recs, B=20)
    within, goff, good = orc.good_hashes(ix, 6, 60)
    c1 = orc.cluster(ix, goff, good, 0, 0, 4)
    within, goff2, good2 = orc.good_hashes(ix, 2, 5, within)
    c2 = orc.cluster(ix, goff2, good2, 10, 200, 2, clus=c1[0], n_sub=c1[1], point_to_min=c1[2])
    with _gpu(B=20) as g:
        g.build_host(recs)
        g.depth_range(6, 60)
        _same(g.cluster(0, 0, 4), c1)
        g.depth_range(2, 5)
        _same(g.cluster(10, 200, 2), c2)


def test_cluster_large_blocks_use_the_global_memory_label_arrays(orc, gpu_lib):
    # > 16384 good hashes and > 4096 read pairs in a block: both shared-memory label arrays fall back to global memory
    p = orc.synth_params(seed=44, n_barcodes=24, pairs_min=4300, pairs_max=4800, genome_len=1_200_000, mol_len=80_000,
                         mol_per_barcode=8)
    recs = orc.synth_fqb(p)
    ix = orc.build(recs, B=22)
    _w, goff, good = orc.good_hashes(ix, 1, 25)
    assert int(np.diff(goff.astype(np.int64)).max()) > 16384 and int(ix.blkNRead.max()) > 4096
    want = orc.cluster(ix, goff, good, 0, 0, 1)
    assert int(want[1].sum()) > 0
    with _gpu(B=22) as g:
        g.build_host(recs)
        g.depth_range(1, 25)
        _same(g.cluster(0, 0, 1), want)


def test_cluster_many_sharing_barcodes_use_the_big_table_and_deep_bins(orc, gpu_lib):
    # 10500 barcodes on a 3 kb genome: every block shares hashes with > 9800 others (the L2-resident table overflows and
    # the block is redone with the full-size one) and bins are thousands deep (per-warp counters instead of the
    # 128-entry shared-memory buffer)
    p = orc.synth_params(seed=61, n_barcodes=10500, pairs_min=6, pairs_max=8, genome_len=3000, mol_len=1500,
                         mol_per_barcode=2)
    recs = orc.synth_fqb(p)
    ix = orc.build(recs, B=20)
    assert int(ix.hashDepth.max()) > 128
    _w, goff, good = orc.good_hashes(ix, 2, 20000)
    want = orc.cluster(ix, goff, good, 0, 0, 5)
    assert int(want[1].sum()) > 0
    with _gpu(B=20) as g:
        g.build_host(recs)
        g.depth_range(2, 20000)
        _same(g.cluster(0, 0, 5), want)


def test_cluster_abandons_a_block_with_more_than_255_clusters(orc, gpu_lib):
    # hash10x.c:810-817 (the case is pinned against the reference binary in tests/test_oracle.py): the 256th founding step
    # of block 1 wipes its labels; 200 groups stay below the limit and keep theirs
    import fqbtools
    seen = set()
    for groups in (300, 200, 255, 256, 257, 258):          # around the limit: exactly 255 clusters stay, 256 abandon
        recs = fqbtools.abandonment_case(groups)
        ix = orc.build(recs, B=20)
        _w, goff, good = orc.good_hashes(ix, 2, 3)
        want = orc.cluster(ix, goff, good, 0, 0, 1)
        seen.add(int(want[1][1]))
        if groups == 300:
            assert int(want[1][1]) == 0
        if groups == 200:
            assert int(want[1][1]) >= 199
        with _gpu(B=20) as g:
            g.build_host(recs)
            g.depth_range(2, 3)
            _same(g.cluster(0, 0, 1), want)
    assert 255 in seen and 0 in seen


def _split_same(sp_gpu, n_new, want, nsub):
    assert n_new == int(nsub.sum()) and sp_gpu.nBlocksMax == want.nBlocksMax
    assert np.array_equal(sp_gpu.blkNRead, want.blkNRead) and np.array_equal(sp_gpu.blkNHash, want.blkNHash)
    assert np.array_equal(sp_gpu.blkOff, want.blkOff)
    assert not sp_gpu.blkNSub.any() and np.array_equal(sp_gpu.blkParent, want.blkParent)
    assert np.array_equal(sp_gpu.blkPointToMin.view(np.uint64), want.blkPointToMin.view(np.uint64))
    assert np.array_equal(sp_gpu.clus, want.clus)


@pytest.mark.parametrize("case", [(31, 200, 40, 160, 60_000, 20_000, 2, 3, 200, 3), (35, 500, 60, 160, 300_000, 15_000, 5, 3, 100, 2),
                                  (37, 120, 20, 80, 40_000, 8_000, 3, 2, 13, 1)])
def test_cluster_split_matches_oracle(orc, gpu_lib, case):
    """--clusterSplit (clusterSplitCodes, hash10x.c:956-1013) on the GPU against the oracle's restatement, which
    tests/test_oracle.py pins to the reference binary: the new block table, every ClusterHash word with its renumbered
    read, and the hash->code lists rebuilt over the new blocks; then --hashDepthRange / --cluster go on from the split
    index like a second clustering round of the reference."""
    seed, nb, pmin, pmax, genome, mol, mpb, dmin, dmax, thr = case
    recs = _cluster_case(orc, seed, nb, pmin, pmax, genome, mol, mpb)
    ix = orc.build(recs, B=20)
    _w, goff, good = orc.good_hashes(ix, dmin, dmax)
    clus, nsub, ptm = orc.cluster(ix, goff, good, 0, 0, thr)
    want = orc.cluster_split(ix, clus, nsub, ptm)
    assert int(nsub.sum()) > 0
    with _gpu(B=20) as g:
        g.build_host(recs)
        g.depth_range(dmin, dmax)
        g.cluster(0, 0, thr)
        got, n_new = g.cluster_split()
        _split_same(got, n_new, want, nsub)
        coff, codes = g.download_codes()
        assert np.array_equal(coff, ix.codeOff) and np.array_equal(codes, want.codes)
        dg = g.digest()
        assert dg["haveCodes"] and dg["codesMissing"] == 0 and dg["codesUnordered"] == 0
        import hash10x_b200
        with pytest.raises(hash10x_b200.H10xError, match="you must set hashDepthRange before cluster"):
            g.cluster(0, 0, thr)                      # the good lists were indexed by the old blocks: dropped
        g.depth_range(dmin, dmax)                     # a second round on the split index runs
        g.cluster(0, 0, thr)
        again, n2 = g.cluster_split()
        assert again.nBlocksMax == got.nBlocksMax + n2


def test_cli_cluster_split_writes_the_reference_file(orc, gpu_lib, tmp_path):
    """hash10x-b200 ... --cluster --clusterSplit --writeHash: the block table and ClusterHash stream of the file are the
    oracle's (= the reference's, tests/test_oracle.py), and the file reads back (--readHash --codeStats)."""
    import subprocess
    import hashfile
    recs = _cluster_case(orc, 34, 300, 100, 250, 200_000, 20_000, 4)
    fq = tmp_path / "x.fqb"
    recs.astype(np.uint32).tofile(fq)
    out = tmp_path / "s.hash"
    import hash10x_b200
    exe = os.path.join(os.path.dirname(hash10x_b200.__file__), "bin", "hash10x-b200")
    r = subprocess.run([exe, "-B", "20", "-ct", "3", "--readFQB", str(fq), "--hashDepthRange", "4", "400", "--cluster", "0", "0",
                        "--clusterSplit", "--writeHash", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    ix = orc.build(recs, B=20)
    _w, goff, good = orc.good_hashes(ix, 4, 400)
    clus, nsub, ptm = orc.cluster(ix, goff, good, 0, 0, 3)
    want = orc.cluster_split(ix, clus, nsub, ptm)
    assert "made %d additional new barcodes from clusters in %d original barcodes" % (int(nsub.sum()), ix.nBlocksMax) in r.stdout
    hf = hashfile.parse(str(out))
    assert hf.nBlocksMax == want.nBlocksMax
    assert np.array_equal(hf.blkNRead, want.blkNRead) and np.array_equal(hf.blkNHash, want.blkNHash)
    assert np.array_equal(hf.blkParent, want.blkParent) and not hf.blkNSub.any()
    assert np.array_equal(hf.blkPointToMin.view(np.uint64), want.blkPointToMin.view(np.uint64))
    assert np.array_equal(hf.clusRaw, want.clus)
    r2 = subprocess.run([exe, "-B", "20", "--readHash", str(out), "--codeStats"], capture_output=True, text=True, timeout=300)
    assert r2.returncode == 0, r2.stdout + r2.stderr
    if orc.ref_binary() is not None:        # the reference's own command chain on the same FQB: the same file
        ref = tmp_path / "r.hash"
        r3 = subprocess.run([orc.ref_binary(), "-B", "20", "-ct", "3", "--readFQB", str(fq), "--hashDepthRange", "4", "400",
                             "--cluster", "0", "0", "--clusterSplit", "--writeHash", str(ref)], capture_output=True, text=True, timeout=600)
        assert r3.returncode == 0, r3.stderr
        rf = hashfile.parse(str(ref))
        assert rf.size == hf.size and rf.nBlocksMax == hf.nBlocksMax
        assert np.array_equal(rf.blkNRead, hf.blkNRead) and np.array_equal(rf.blkNHash, hf.blkNHash)
        assert np.array_equal(rf.blkParent, hf.blkParent)
        # the reference's --readFQB leaves the two spare ClusterHash bytes uninitialised (hash10x.c:175): compare ids and reads
        assert np.array_equal(rf.clusIdx, hf.clusIdx) and np.array_equal(rf.clusRead, hf.clusRead)


def test_cluster_argument_checks(orc, gpu_lib):
    import hash10x_b200
    recs = _cluster_case(orc, 37, 120, 20, 80, 40_000, 8_000, 3)
    with _gpu(B=20) as g:
        g.build_host(recs)
        with pytest.raises(hash10x_b200.H10xError, match="you must set hashDepthRange before cluster"):
            g.cluster(0, 0, 5)                       # hash10x.c:1258
        g.depth_range(2, 13)
        with pytest.raises(hash10x_b200.H10xError, match="clusterThreshold"):
            g.cluster(0, 0, 0)
        with pytest.raises(hash10x_b200.H10xError, match="code range"):
            g.cluster(0, 10_000, 1)
        clus, nsub, ptm, _ms = g.cluster(5, 5, 1)    # empty range: nothing changes
        assert not nsub.any() and not ptm.any()
        g.build_host(recs)                            # a new build drops the good lists
        with pytest.raises(hash10x_b200.H10xError, match="you must set hashDepthRange before cluster"):
            g.cluster(0, 0, 5)


def test_cli_cluster_writes_the_reference_fields(orc, gpu_lib, tmp_path):
    """hash10x-b200 --readFQB --hashDepthRange -ct --cluster --writeHash: nSubCluster, pointToMin and the subCluster
    bytes land in the .hash exactly where the reference's --writeHash puts them (hash10x.c:62-70,256-261)."""
    import hash10x_b200
    exe = os.path.join(os.path.dirname(hash10x_b200.__file__), "bin", "hash10x-b200")
    recs = _cluster_case(orc, 34, 300, 100, 250, 200_000, 20_000, 4)
    fqb, out = str(tmp_path / "a.fqb"), str(tmp_path / "g.hash")
    recs.tofile(fqb)
    r = subprocess.run([exe, "-B", "20", "--readFQB", fqb, "--hashDepthRange", "4", "400", "-ct", "3", "--cluster", "0", "0",
                        "--writeHash", out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "  clustered codes 1 to %d" % (301 + 0) in r.stdout
    hf = hashfile.parse(out)
    ix = orc.build(recs, B=20)
    _w, goff, good = orc.good_hashes(ix, 4, 400)
    clus, nsub, ptm = orc.cluster(ix, goff, good, 0, 0, 3)
    assert np.array_equal(hf.blkNSub, nsub)
    assert np.array_equal(hf.blkPointToMin.view(np.uint64), ptm.view(np.uint64))
    assert np.array_equal(hf.clusRaw, clus)
    hashfile.assert_strict_equal(hashfile.from_index(ix), hf, table=True)
    # --codeStats after --cluster adds the CODE_CLUSTER histogram of nSubCluster (hash10x.c:388-402)
    r = subprocess.run([exe, "-B", "20", "--readFQB", fqb, "--hashDepthRange", "4", "400", "-ct", "3", "--cluster", "0", "0",
                        "--codeStats"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    hist = [ln.split() for ln in r.stdout.splitlines() if ln.startswith("CODE_CLUSTER_HIST")]
    want = np.bincount(nsub)
    assert [int(h[2]) for h in hist] == want.tolist() and [int(h[1]) for h in hist] == list(range(want.size))
    assert any(ln.startswith("CODE_CLUSTER_STATS MEAN") for ln in r.stdout.splitlines())
    # --cluster before --hashDepthRange: the reference's warning, not an error (hash10x.c:1258)
    r = subprocess.run([exe, "-B", "20", "--readFQB", fqb, "--cluster", "0", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "!! you must set hashDepthRange before cluster" in r.stdout


def test_loaded_index_clusters_like_a_built_one(orc, gpu_lib, tmp_path):
    """--readHash's counterpart on the device (h10x_gpu_load_index): a host index, here the oracle's with the labels of an
    earlier --cluster, becomes the resident one; --hashDepthRange / --cluster then continue from that state.  The CLI
    does the same after --readHash when a GPU is present (the README's read-then-cluster sessions)."""
    import hash10x_b200
    from hash10x_b200 import binding
    recs = _cluster_case(orc, 36, 250, 150, 300, 100_000, 10_000, 6)
    ix = orc.build(recs, B=20)
    within, goff, good = orc.good_hashes(ix, 6, 60)
    c1 = orc.cluster(ix, goff, good, 0, 0, 4)
    within2, goff2, good2 = orc.good_hashes(ix, 2, 5)         # a fresh session: the within[] flags start empty again
    c2 = orc.cluster(ix, goff2, good2, 10, 200, 2, clus=c1[0], n_sub=c1[1], point_to_min=c1[2])
    ix.clus, ix.blkNSub, ix.blkPointToMin = c1
    with _gpu(B=20) as g:
        g.load_index(ix)
        _gw, ggoff, ggood = g.depth_range(2, 5)
        assert np.array_equal(ggoff, goff2) and np.array_equal(ggood, good2)
        _same(g.cluster(10, 200, 2), c2)
    # the same through the host program: session 1 writes the clustered index, session 2 reads it and goes on
    exe = os.path.join(os.path.dirname(hash10x_b200.__file__), "bin", "hash10x-b200")
    s1, s2 = str(tmp_path / "s1.hash"), str(tmp_path / "s2.hash")
    binding.write_hash(ix, s1)
    r = subprocess.run([exe, "-B", "20", "--readHash", s1, "--hashDepthRange", "2", "5", "-ct", "2", "--cluster", "10", "200",
                        "--codeStats", "--writeHash", s2], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "  clustered codes 10 to 200" in r.stdout
    hf = hashfile.parse(s2)
    assert np.array_equal(hf.blkNSub, c2[1]) and np.array_equal(hf.clusRaw, c2[0])
    assert np.array_equal(hf.blkPointToMin.view(np.uint64), c2[2].view(np.uint64))


def _cluster_golden():
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_cluster.json")) as f:
        return sorted(json.load(f).items())


@pytest.mark.parametrize("name,d", _cluster_golden())
def test_gpu_reproduces_reference_cluster_golden(gpu_lib, name, d):
    """the committed digests of the REFERENCE's --hashDepthRange + --cluster output (tests/golden/make_golden_cluster.py)
    against the CUDA path alone - no oracle involved, no /root/reference needed on the GPU box"""
    import zlib
    crc = lambda a: zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF      # noqa: E731
    recs = np.fromfile(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".fqb"), np.uint32)
    with _gpu(B=d["params"]["B"]) as g:
        got = g.build_host(recs)
        assert (int(got.nBlocksMax), int(got.nHashes)) == (d["nBlocksMax"], d["nHashes"])
        g.depth_range(*d["depth_range"])
        clus, nsub, ptm, _ms = g.cluster(d["codes"][0], d["codes"][1], d["clusterThreshold"])
    sub = ((clus >> np.uint64(48)) & np.uint64(0xFF)).astype(np.uint8)
    assert crc(nsub) == d["crc_blkNSub"] and crc(ptm) == d["crc_pointToMin"] and crc(sub) == d["crc_clusSub"]
    assert crc(clus & np.uint64(0x00FFFFFFFFFFFFFF)) == d["crc_clusRaw"]
    assert int(nsub.sum()) == d["sub_clusters"] and int((sub > 0).sum()) == d["clustered_entries"]


def test_cli_multi_gpu_depth_range_and_cluster(orc, gpu_lib, tmp_path):
    """--gpus N --readFQB --hashDepthRange -ct --cluster --writeHash: the commands that follow the build run on every GPU
    for its own barcode blocks (h10x_multi_*); the .hash carries the oracle's nSubCluster / pointToMin / subCluster bytes,
    also for a cluster range that starts and ends inside different GPUs' block ranges"""
    import hash10x_b200
    n = gpu_lib.h10x_gpu_device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    exe = os.path.join(os.path.dirname(hash10x_b200.__file__), "bin", "hash10x-b200")
    recs = _cluster_case(orc, 36, 300, 100, 250, 200_000, 20_000, 4)
    fqb = str(tmp_path / "a.fqb")
    recs.tofile(fqb)
    ix = orc.build(recs, B=20)
    _w, goff, good = orc.good_hashes(ix, 4, 400)
    for g in sorted({2, min(n, 8)}):
        for cmin, cmax in ((0, 0), (40, 260)):
            out = str(tmp_path / ("g%d_%d.hash" % (g, cmin)))
            r = subprocess.run([exe, "--gpus", str(g), "-B", "20", "--readFQB", fqb, "--hashDepthRange", "4", "400", "-ct", "3",
                                "--cluster", str(cmin), str(cmax), "--writeHash", out], capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stderr
            clus, nsub, ptm = orc.cluster(ix, goff, good, cmin, cmax, 3)
            hf = hashfile.parse(out)
            assert np.array_equal(hf.blkNSub, nsub), (g, cmin)
            assert np.array_equal(hf.blkPointToMin.view(np.uint64), ptm.view(np.uint64))
            assert np.array_equal(hf.clusRaw, clus)
            hashfile.assert_strict_equal(hashfile.from_index(ix), hf, table=True)
