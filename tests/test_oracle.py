"""CPU tests (no GPU): the oracle against (i) the known-answer vectors produced by the reference's
seqhash.c, (ii) the golden digests generated from the reference binary
(tests/golden/make_golden.py), (iii) the reference binary itself when oracle/_ref is present."""
import json
import os
import zlib

import numpy as np
import pytest

import fqbtools
import hashfile

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")

KAT_SEQ = ("CCCCACCACCAGGACACTTTCAGAGTTCTCCGTCATCGTTAGCAGCACGGTTAGCTTGTCTCTGTCTATTTCACGACTGGTGGAGTCGTATAGTC"
           "ACGAGCTGGGTATGTCCTGAAGCTAGACATTGGAACTATGTAAGTCCTCTTCCGG")


def test_factor1_from_glibc_random(orc):
    # seqhash.c:29 after srandom(17): the constant every other vector depends on
    assert orc.factor1(17) == 0x49308BB9003CB3AD
    assert orc.DEFAULT_FACTOR1 == 0x49308BB9003CB3AD


def test_kat_sequence_generator_matches_appendix_e():
    x, out = 42, []
    for _ in range(150):
        x = (x * 1103515245 + 12345) & 0x7FFFFFFF
        out.append("ACGT"[(x >> 16) & 3])
    assert "".join(out) == KAT_SEQ


def test_kat_kmer0_hashes(orc):
    codes = [fqbtools.CODE[c] for c in KAT_SEQ]
    h = 0
    hrc = 0
    for i, c in enumerate(codes[:21]):
        h = (h << 2) | c
        hrc |= (3 - c) << (2 * i)
    assert h == 0x154514A11FD and hrc == 0x202ED7AEBAA
    hf, hr = orc.kmer_hashes(h, hrc)
    assert (hf, hr) == (0x32DB51C0BC3, 0x1FF0AB2C2AA)


def test_kat_moshes(orc):
    codes = np.array([fqbtools.CODE[c] for c in KAT_SEQ], np.uint8)
    h, pos, fwd = orc.seq_moshes(codes)
    want = [(4, 0x21C5F360BD8, 1), (25, 0x29099B43123, 0), (75, 0x10091E2CC3B, 1), (102, 0xEAAE4BA6A5, 1),
            (106, 0xD50A11DD64, 1)]
    assert [(int(p), int(x), int(f)) for x, p, f in zip(h, pos, fwd)] == want
    # 30 x A -> 10 moshes, all hash 0; 30 x C -> none; shorter than k -> none
    h, pos, _ = orc.seq_moshes(np.zeros(30, np.uint8))
    assert list(pos) == list(range(10)) and not h.any()
    assert orc.seq_moshes(np.ones(30, np.uint8))[0].size == 0
    assert orc.seq_moshes(np.zeros(20, np.uint8))[0].size == 0


def _digest_matches(hf, d):
    def crc(a):
        return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF
    assert hf.hashNumber == d["hashNumber"] and hf.nBlocksMax == d["nBlocksMax"] and hf.nHashes == d["nHashes"]
    assert crc(hf.hashValue) == d["crc_hashValue"]
    assert crc(hf.hashDepth[:hf.depthMax]) == d["crc_hashDepth"]
    assert crc(hf.hashIndex) == d["crc_hashIndex"]
    assert crc(hf.blkNRead) == d["crc_blkNRead"] and crc(hf.blkNHash) == d["crc_blkNHash"]
    assert crc(hf.clusIdx) == d["crc_clusIdx"] and crc(hf.clusRead) == d["crc_clusRead"]


def _golden_cases():
    with open(os.path.join(GOLD, "golden.json")) as f:
        return sorted(json.load(f).items())


@pytest.mark.parametrize("name,d", _golden_cases())
def test_oracle_reproduces_reference_golden(orc, tmp_path, name, d):
    recs = np.fromfile(os.path.join(GOLD, name + ".fqb"), np.uint32)
    kw = d["params"]
    f1 = orc.factor1(kw.get("r", 17))
    path = str(tmp_path / "o.hash")
    st = orc.build_and_write(recs, path, B=kw.get("B", 20), k=kw.get("k", 21), w=kw.get("w", 31), factor1_=f1,
                             N=kw.get("N", 0), chunk=kw.get("chunk", 100000))
    assert st == 0
    hf = hashfile.parse(path)
    _digest_matches(hf, d)
    assert hf.size == d["fileSize"] and (hf.depthDim, hf.blkDim) == (d["depthDim"], d["blkDim"])
    hashfile.check_table(hf)


def test_golden_quirks_are_what_the_survey_says():
    with open(os.path.join(GOLD, "golden.json")) as f:
        g = json.load(f)
    q = g["quirks"]
    assert q["blkNHash_head"][2] == 1 and q["blkNHash_head"][4] == 1     # phantom entries
    assert q["blkNHash_head"][6] == 0                                    # last run never hashed
    assert g["single_run"]["hashNumber"] == 1 and g["single_run"]["nHashes"] == 0
    # the all-A barcode run glued to its successor only when it ends on the chunk boundary
    assert g["allA_barcode_c10"]["nBlocksMax"] == g["allA_barcode_c11"]["nBlocksMax"] - 1


# ------------------------------------------------------------------ live reference (this container only)

def _need_ref(orc):
    if orc.ref_binary() is None:
        pytest.skip("oracle/_ref/hash10x not built (no /root/reference on this machine)")


@pytest.mark.parametrize("seed,nb,pmax,kw", [
    (1, 30, 200, {}), (2, 80, 60, dict(N=1500)), (3, 50, 90, dict(chunk=97)), (4, 20, 300, dict(k=16, w=32, r=3)),
    (5, 20, 300, dict(k=31, w=5, r=99)), (6, 15, 100, dict(k=11, w=2, r=1))])
def test_oracle_equals_reference_binary(orc, tmp_path, seed, nb, pmax, kw):
    _need_ref(orc)
    p = orc.synth_params(seed=seed, n_barcodes=nb, pairs_min=2, pairs_max=pmax, genome_len=60_000, mol_len=8_000)
    recs = orc.synth_fqb(p)
    fqb = str(tmp_path / "a.fqb")
    recs.tofile(fqb)
    B = 21
    r = orc.run_reference(fqb, str(tmp_path / "r.hash"), B=B, k=kw.get("k"), w=kw.get("w"), r=kw.get("r"),
                          N=kw.get("N"), chunk=kw.get("chunk"))
    assert r.returncode == 0, r.stderr
    st = orc.build_and_write(recs, str(tmp_path / "o.hash"), B=B, k=kw.get("k", 21), w=kw.get("w", 31),
                             factor1_=orc.factor1(kw.get("r", 17)), N=kw.get("N", 0), chunk=kw.get("chunk", 100000))
    assert st == 0
    a, b = hashfile.parse(str(tmp_path / "r.hash")), hashfile.parse(str(tmp_path / "o.hash"))
    assert a.size == b.size
    hashfile.assert_strict_equal(a, b, table=True)
    hashfile.assert_canonical_equal(a, b)
    ix = orc.build(recs, B=B, k=kw.get("k", 21), w=kw.get("w", 31), factor1_=orc.factor1(kw.get("r", 17)),
                   N=kw.get("N", 0), chunk=kw.get("chunk", 100000))
    co, cd = hashfile.hash_to_code_lists(a)
    assert np.array_equal(co, ix.codeOff) and np.array_equal(cd, ix.codes)     # -DCHECK's invariant and more
    assert np.array_equal(np.diff(ix.codeOff.astype(np.int64)), ix.hashDepth)


def test_oracle_errors_equal_reference(orc, tmp_path):
    _need_ref(orc)
    p = orc.synth_params(seed=9, n_barcodes=30, pairs_min=20, pairs_max=90)
    recs = orc.synth_fqb(p)
    fqb = str(tmp_path / "a.fqb")
    recs.tofile(fqb)
    r = orc.run_reference(fqb, None, B=20, chunk=50)
    assert r.returncode != 0 and "chunkSize too small" in r.stderr
    assert orc.build(recs, B=20, chunk=50).status == 2
    r = orc.run_reference(fqb, None, B=19)
    assert r.returncode != 0 and "out of range 20-30" in r.stderr
    assert orc.build(recs, B=19).status == 3
    big = orc.synth_params(seed=8, n_barcodes=60, pairs_min=800, pairs_max=1000, genome_len=20_000_000,
                           mol_len=100_000, mol_per_barcode=20)
    recs = orc.synth_fqb(big)
    recs.tofile(fqb)
    r = orc.run_reference(fqb, None, B=20)
    assert r.returncode != 0 and "hashTableSize is too small" in r.stderr
    assert orc.build(recs, B=20).status == 1


def test_good_hashes_restatement_properties(orc):
    """hashWithinRangeBuild + goodHashesBuild (hash10x.c:528-539,738-766) as restated by the oracle: flags only
    ever set, lists hold exactly the in-range entries, ordered by depth with ties in list order."""
    p = orc.synth_params(seed=81, n_barcodes=50, pairs_min=5, pairs_max=200, genome_len=80_000, mol_len=10_000)
    ix = orc.build(orc.synth_fqb(p), B=20)
    within, off, good = orc.good_hashes(ix, 3, 9)
    depth = ix.hashDepth.astype(np.int64)
    assert np.array_equal(within.astype(bool), (depth >= 3) & (depth < 9))
    ids = (ix.clus & np.uint64(0xFFFFFFFF)).astype(np.int64)
    for c in range(1, ix.nBlocksMax):
        lo, n = int(ix.blkOff[c]), int(ix.blkNHash[c])
        lst = good[int(off[c]):int(off[c + 1])].astype(np.int64)
        want = np.flatnonzero(within[ids[lo:lo + n]])
        assert sorted(lst.tolist()) == want.tolist()
        d = depth[ids[lo + lst]]
        assert (np.diff(d) >= 0).all()
        same = np.diff(d) == 0
        assert (np.diff(lst)[same] > 0).all()
    w2, _, good2 = orc.good_hashes(ix, 20, 30, within.copy())
    assert (w2 >= within).all() and good2.size >= good.size       # ranges accumulate


CLUSTER_CASES = [
    # seed, barcodes, pairs min/max, genome, molecule length, molecules per barcode, depth range, threshold, code range
    (31, 200, 40, 160, 60_000, 20_000, 2, 3, 200, 3, 0, 0),
    (34, 300, 100, 250, 200_000, 20_000, 4, 4, 400, 3, 0, 0),
    (35, 500, 60, 160, 300_000, 15_000, 5, 3, 100, 2, 7, 450),
    (36, 250, 150, 300, 100_000, 10_000, 6, 6, 60, 4, 0, 0),
    (37, 120, 20, 80, 40_000, 8_000, 3, 2, 13, 1, 0, 0),
]


def _cluster_case(orc, seed, nb, pmin, pmax, genome, mol, mpb):
    p = orc.synth_params(seed=seed, n_barcodes=nb, pairs_min=pmin, pairs_max=pmax, genome_len=genome, mol_len=mol,
                         mol_per_barcode=mpb)
    return orc.synth_fqb(p)


@pytest.mark.parametrize("seed,nb,pmin,pmax,genome,mol,mpb,dmin,dmax,thr,cmin,cmax", CLUSTER_CASES)
def test_cluster_equals_reference_binary(orc, tmp_path, seed, nb, pmin, pmax, genome, mol, mpb, dmin, dmax, thr, cmin, cmax):
    """--hashDepthRange + --cluster (hash10x.c:528-539,738-868): the oracle's restatements against the reference
    binary.  The reference reads a .hash whose subCluster bytes are zero (written by the oracle), so its own
    uninitialised ClusterHash bytes (hash10x.c:175) never enter; its --writeHash then exposes nSubCluster,
    pointToMin and every subCluster byte, which also pins goodHashesBuild (the clustering walks the good lists
    in order)."""
    _need_ref(orc)
    import subprocess
    recs = _cluster_case(orc, seed, nb, pmin, pmax, genome, mol, mpb)
    B = 20
    src, dst = str(tmp_path / "o.hash"), str(tmp_path / "r.hash")
    assert orc.build_and_write(recs, src, B=B) == 0
    cmd = [orc.ref_binary(), "-B", str(B), "-ct", str(thr), "--readHash", src, "--hashDepthRange", str(dmin), str(dmax),
           "--cluster", str(cmin), str(cmax), "--writeHash", dst]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    ref = hashfile.parse(dst)
    ix = orc.build(recs, B=B)
    _within, goff, good = orc.good_hashes(ix, dmin, dmax)
    clus, nsub, ptm = orc.cluster(ix, goff, good, cmin, cmax, thr)
    assert np.array_equal(ref.blkNSub, nsub)
    assert np.array_equal(ref.blkPointToMin.view(np.uint64), ptm.view(np.uint64))      # bit-exact doubles
    assert np.array_equal(ref.clusRaw, clus)
    assert int(nsub.sum()) > 0 and int(ref.clusSub.max()) > 0          # the case does cluster something


def test_cluster_abandons_a_block_with_more_than_255_clusters_like_the_reference(orc, tmp_path):
    """hash10x.c:810-817: the 256th founding step of a block wipes its labels, sets nSubCluster = 0 and stops; the other
    blocks (one cluster each here) are not affected.  Reference binary against the oracle, bit for bit."""
    _need_ref(orc)
    import subprocess
    import fqbtools
    recs = fqbtools.abandonment_case(300)
    src, dst = str(tmp_path / "o.hash"), str(tmp_path / "r.hash")
    assert orc.build_and_write(recs, src, B=20) == 0
    cmd = [orc.ref_binary(), "-B", "20", "-ct", "1", "--readHash", src, "--hashDepthRange", "2", "3", "--cluster", "0", "0",
           "--writeHash", dst]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "has too many clusters" in r.stderr                     # the reference did abandon block 1
    ref = hashfile.parse(dst)
    ix = orc.build(recs, B=20)
    _within, goff, good = orc.good_hashes(ix, 2, 3)
    clus, nsub, ptm = orc.cluster(ix, goff, good, 0, 0, 1)
    assert int(nsub[1]) == 0 and int(nsub[2:].sum()) >= 250          # block 1 abandoned, the partners clustered
    assert np.array_equal(ref.blkNSub, nsub)
    assert np.array_equal(ref.blkPointToMin.view(np.uint64), ptm.view(np.uint64))
    assert np.array_equal(ref.clusRaw, clus)


@pytest.mark.parametrize("seed,nb,pmin,pmax,genome,mol,mpb,dmin,dmax,thr,cmin,cmax", CLUSTER_CASES[:3] + CLUSTER_CASES[4:])
def test_cluster_split_equals_reference_binary(orc, tmp_path, seed, nb, pmin, pmax, genome, mol, mpb, dmin, dmax, thr, cmin, cmax):
    """--clusterSplit (clusterSplitCodes, hash10x.c:956-1013): the reference's --readHash --hashDepthRange --cluster
    --clusterSplit --writeHash against the oracle's restatement: the whole new block table (nRead, nHash, nSubCluster,
    clusterParent, pointToMin) and every ClusterHash word, renumbered reads included."""
    _need_ref(orc)
    import subprocess
    recs = _cluster_case(orc, seed, nb, pmin, pmax, genome, mol, mpb)
    B = 20
    src, dst = str(tmp_path / "o.hash"), str(tmp_path / "r.hash")
    assert orc.build_and_write(recs, src, B=B) == 0
    cmd = [orc.ref_binary(), "-B", str(B), "-ct", str(thr), "--readHash", src, "--hashDepthRange", str(dmin), str(dmax),
           "--cluster", str(cmin), str(cmax), "--clusterSplit", "--writeHash", dst]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    ref = hashfile.parse(dst)
    ix = orc.build(recs, B=B)
    _within, goff, good = orc.good_hashes(ix, dmin, dmax)
    clus, nsub, ptm = orc.cluster(ix, goff, good, cmin, cmax, thr)
    sp = orc.cluster_split(ix, clus, nsub, ptm)
    assert "made %d additional new barcodes from clusters in %d original barcodes" % (int(nsub.sum()), ix.nBlocksMax) in r.stdout
    assert ref.nBlocksMax == sp.nBlocksMax == ix.nBlocksMax + int(nsub.sum())
    assert np.array_equal(ref.blkNRead, sp.blkNRead) and np.array_equal(ref.blkNHash, sp.blkNHash)
    assert np.array_equal(ref.blkNSub, sp.blkNSub) and np.array_equal(ref.blkParent, sp.blkParent)
    assert np.array_equal(ref.blkPointToMin.view(np.uint64), sp.blkPointToMin.view(np.uint64))
    assert np.array_equal(ref.clusRaw, sp.clus)
    assert int(sp.blkParent.max()) > 0


def test_cluster_twice_carries_state_like_the_reference(orc, tmp_path):
    """A second --hashDepthRange / --cluster pair works on what the first left behind: within[] flags accumulate
    (hash10x.c:535), only the good entries are wiped (:783) and a block without good hashes is merged again with
    its old labels (:780,840)."""
    _need_ref(orc)
    import subprocess
    recs = _cluster_case(orc, 36, 250, 150, 300, 100_000, 10_000, 6)
    src, dst = str(tmp_path / "o.hash"), str(tmp_path / "r.hash")
    assert orc.build_and_write(recs, src, B=20) == 0
    cmd = [orc.ref_binary(), "-B", "20", "-ct", "4", "--readHash", src, "--hashDepthRange", "6", "60", "--cluster", "0", "0",
           "--hashDepthRange", "2", "5", "-ct", "2", "--cluster", "10", "200", "--writeHash", dst]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    ref = hashfile.parse(dst)
    ix = orc.build(recs, B=20)
    within, goff, good = orc.good_hashes(ix, 6, 60)
    clus, nsub, ptm = orc.cluster(ix, goff, good, 0, 0, 4)
    within, goff, good = orc.good_hashes(ix, 2, 5, within)
    stale = orc.cluster_stale_labels()
    clus, nsub, ptm = orc.cluster(ix, goff, good, 10, 200, 2, clus=clus, n_sub=nsub, point_to_min=ptm)
    assert orc.cluster_stale_labels() == stale      # the case stays clear of the reference's out-of-bounds read
    assert np.array_equal(ref.blkNSub, nsub)
    assert np.array_equal(ref.blkPointToMin.view(np.uint64), ptm.view(np.uint64))
    assert np.array_equal(ref.clusRaw, clus)


def _cluster_golden():
    with open(os.path.join(GOLD, "golden_cluster.json")) as f:
        return sorted(json.load(f).items())


@pytest.mark.parametrize("name,d", _cluster_golden())
def test_oracle_reproduces_reference_cluster_golden(orc, name, d):
    """golden_cluster.json (tests/golden/make_golden_cluster.py: the reference's own --readFQB / --readHash ...
    --hashDepthRange --cluster --writeHash) against the oracle's build + good lists + clustering"""
    crc = lambda a: zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF      # noqa: E731
    recs = np.fromfile(os.path.join(GOLD, name + ".fqb"), np.uint32)
    ix = orc.build(recs, B=d["params"]["B"])
    assert (int(ix.nBlocksMax), int(ix.nHashes)) == (d["nBlocksMax"], d["nHashes"])
    _w, goff, good = orc.good_hashes(ix, *d["depth_range"])
    clus, nsub, ptm = orc.cluster(ix, goff, good, d["codes"][0], d["codes"][1], d["clusterThreshold"])
    sub = ((clus >> np.uint64(48)) & np.uint64(0xFF)).astype(np.uint8)
    assert crc(nsub) == d["crc_blkNSub"] and crc(ptm) == d["crc_pointToMin"] and crc(sub) == d["crc_clusSub"]
    assert crc(clus & np.uint64(0x00FFFFFFFFFFFFFF)) == d["crc_clusRaw"]
    assert int(nsub.sum()) == d["sub_clusters"] and int((sub > 0).sum()) == d["clustered_entries"]
