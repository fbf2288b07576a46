"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/h10x_gpu.h declares,
fails loudly without a device, and its host-only parts (.hash writer/reader, factor1) are right."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import hashfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "h10x_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(h10x_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import hash10x_b200
    L = hash10x_b200.load_library()
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), n
    assert L.h10x_abi_version() == 3      # 3: h10x_gpu_cluster_split, clusterParent in h10x_index


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "h10x_gpu.h"\nint main(void){ h10x_params p; h10x_index i; (void)p; (void)i; '
                   'return sizeof(h10x_cluster_hash) == 8 ? 0 : 1; }\n')
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_factor1_and_error_texts():
    import hash10x_b200
    L = hash10x_b200.load_library()
    assert hash10x_b200.factor1_from_seed(17) == 0x49308BB9003CB3AD
    assert L.h10x_strerror(1) == b"hashTableSize is too small"      # hash10x.c:149
    assert L.h10x_strerror(2) == b"chunkSize too small"             # hash10x.c:206
    assert [L.h10x_stage_name(i) for i in range(12)].count(b"") == 0


def test_no_device_means_loud_failure_not_fallback():
    import hash10x_b200
    L = hash10x_b200.load_library()
    if L.h10x_gpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(hash10x_b200.H10xError) as e:
        hash10x_b200.Hash10xGPU(B=20)
    assert e.value.code == 7 and "no CPU fallback" in e.value.msg
    cli = os.path.join(ROOT, "hash10x_b200", "bin", "hash10x-b200")
    if os.path.exists(cli):
        fq = os.path.join(ROOT, "tests", "golden", "synth_small.fqb")
        r = subprocess.run([cli, "-B", "20", "--readFQB", fq], capture_output=True, text=True)
        assert r.returncode != 0 and "FATAL ERROR: no CUDA device" in r.stderr
        # the commands after the build have no CPU version either: a host index (--readHash) without a device dies loudly
        from oracle import orc
        import tempfile
        recs = np.fromfile(fq, dtype=np.uint32).reshape(-1, 30)
        with tempfile.TemporaryDirectory() as d:
            h = os.path.join(d, "x.hash")
            assert orc.build_and_write(recs, h, B=20) == 0
            for cmd, text in ((["--hashDepthRange", "2", "40"], "--hashDepthRange runs on the GPU-resident index"),
                              (["--clusterSplit"], "--clusterSplit runs on the index resident on one GPU")):
                r = subprocess.run([cli, "-B", "20", "--readHash", h] + cmd, capture_output=True, text=True)
                assert r.returncode != 0 and text in r.stderr, r.stderr
    # entry points that need a context refuse a NULL one instead of doing anything on the host
    err = C.create_string_buffer(256)
    n = C.c_uint64(0)
    L.h10x_gpu_depth_range_device.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_char_p, C.c_size_t]
    assert L.h10x_gpu_depth_range_device(None, 2, 40, C.byref(n), err, 256) != 0
    L.h10x_gpu_cluster_split.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_size_t]
    assert L.h10x_gpu_cluster_split(None, None, None, err, 256) != 0


def test_product_never_imports_the_oracle():
    for dirpath, _dirs, files in os.walk(os.path.join(ROOT, "hash10x_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")) and f != "synth_fqb.h":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "from oracle" not in text and "import oracle" not in text and "liborc" not in text, f


def test_hash_writer_reader_roundtrip_and_reference_layout(orc, tmp_path):
    """h10x_write_hash on an oracle-built index gives the same bytes as the oracle's own writer
    (which is checked against the reference binary), and h10x_read_hash reads it back."""
    import hash10x_b200
    from hash10x_b200 import binding
    p = orc.synth_params(seed=77, n_barcodes=40, pairs_min=2, pairs_max=70)
    recs = orc.synth_fqb(p)
    ix = orc.build(recs, B=20)
    a = str(tmp_path / "a.hash")
    b = str(tmp_path / "b.hash")
    binding.write_hash(ix, a)
    assert orc.build_and_write(recs, b, B=20) == 0
    assert open(a, "rb").read() == open(b, "rb").read()
    L = hash10x_b200.load_library()
    ci = binding.CIndex()
    err = C.create_string_buffer(256)
    assert L.h10x_read_hash(a.encode(), 21, C.byref(ci), err, 256) == 3 and b"rerun with -B 20" in err.value
    assert L.h10x_read_hash(a.encode(), 20, C.byref(ci), err, 256) == 0
    got = binding.Index(ci, L)
    L.h10x_index_free(C.byref(ci))
    assert got.hashNumber == ix.hashNumber and got.nHashes == ix.nHashes and got.nReads == ix.nReads
    assert np.array_equal(got.hashValue, ix.hashValue) and np.array_equal(got.clus, ix.clus)
    assert np.array_equal(got.hashDepth, ix.hashDepth) and np.array_equal(got.blkOff, ix.blkOff)
    bad = str(tmp_path / "bad.hash")
    raw = open(a, "rb").read()
    open(bad, "wb").write(b"nope" + raw[4:])
    assert L.h10x_read_hash(bad.encode(), 20, C.byref(ci), err, 256) != 0 and b"not a 10X hash file" in err.value
    # version 1 files carry hashValue as an Array (hash10x.c:285-291): read like the reference reads them; newer ones die
    # with its message
    import struct
    t = 16 + 4 * (1 << 20)
    hn = struct.unpack_from("<I", raw, t)[0]
    arr = struct.pack("<QQiiii", 0, 0, hn, 8, hn, 0)          # array header: magic, base, dim, size, max, pad
    magic = raw[t + 4 + 8 * hn:t + 4 + 8 * hn + 8]            # the magic of the hashDepth Array that follows
    v1 = raw[:4] + struct.pack("<I", 1) + raw[8:t] + magic + arr[8:] + raw[t + 4:]
    open(bad, "wb").write(v1)
    assert L.h10x_read_hash(bad.encode(), 20, C.byref(ci), err, 256) == 0, err.value
    got1 = binding.Index(ci, L)
    L.h10x_index_free(C.byref(ci))
    assert got1.hashNumber == ix.hashNumber and np.array_equal(got1.hashValue, ix.hashValue) and np.array_equal(got1.clus, ix.clus)
    open(bad, "wb").write(raw[:4] + struct.pack("<I", 3) + raw[8:])
    assert L.h10x_read_hash(bad.encode(), 20, C.byref(ci), err, 256) != 0 and b"hash file version mismatch: file 3 > code 2" in err.value
    # a file whose entries point outside the bins is refused instead of indexing device memory with them
    e0 = len(raw) - 8 * int(ix.nHashes)
    open(bad, "wb").write(raw[:e0] + struct.pack("<I", ix.hashNumber + 5) + raw[e0 + 4:])
    assert L.h10x_read_hash(bad.encode(), 20, C.byref(ci), err, 256) != 0 and b"inconsistent hash file" in err.value


def test_synth_generator_is_deterministic_and_well_formed(orc):
    p = orc.synth_params(seed=5, n_barcodes=10, pairs_min=3, pairs_max=9)
    a, b = orc.synth_fqb(p), orc.synth_fqb(p)
    assert np.array_equal(a, b)
    n, off = orc.synth_layout(p)
    assert a.shape == (n, 30)
    runs = np.flatnonzero(np.r_[True, a[1:, 0] != a[:-1, 0]])
    assert np.array_equal(runs, off[:-1].astype(np.int64))           # grouped by barcode, distinct words
    assert (a[:, 9] < (1 << 14)).all() and (a[:, 24] < (1 << 14)).all()   # 151 bp: last word right-aligned
    assert (a[:, 14] == 0x7FFFFF).all() and (a[:, 10:14] == 0xFFFFFFFF).all()
    part = orc.synth_fqb(p, 5, 17)
    assert np.array_equal(part, a[5:17])


def test_hash_writer_reader_carry_the_cluster_fields(orc, tmp_path):
    """ClusterBlock.nSubCluster / .pointToMin and ClusterHash.subCluster (what --cluster leaves behind) go through
    h10x_write_hash / h10x_read_hash at the reference's offsets (hash10x.c:62-70,256-261): the file parses to the
    oracle's values and, where the reference binary is available, has the reference's bytes outside the raw
    pointers its reader overwrites."""
    import subprocess
    import hash10x_b200
    from hash10x_b200 import binding
    import hashfile
    p = orc.synth_params(seed=37, n_barcodes=120, pairs_min=20, pairs_max=80, genome_len=40_000, mol_len=8_000,
                         mol_per_barcode=3)
    recs = orc.synth_fqb(p)
    ix = orc.build(recs, B=20)
    _w, goff, good = orc.good_hashes(ix, 2, 13)
    clus, nsub, ptm = orc.cluster(ix, goff, good, 0, 0, 1)
    assert int(nsub.sum()) > 0
    ix.clus, ix.blkNSub, ix.blkPointToMin = clus, nsub, ptm
    ours = str(tmp_path / "ours.hash")
    binding.write_hash(ix, ours)
    hf = hashfile.parse(ours)
    assert np.array_equal(hf.blkNSub, nsub) and np.array_equal(hf.clusRaw, clus)
    assert np.array_equal(hf.blkPointToMin.view(np.uint64), ptm.view(np.uint64))
    L = hash10x_b200.load_library()
    ci = binding.CIndex()
    err = C.create_string_buffer(256)
    assert L.h10x_read_hash(ours.encode(), 20, C.byref(ci), err, 256) == 0
    got = binding.Index(ci, L)
    L.h10x_index_free(C.byref(ci))
    assert np.array_equal(got.blkNSub, nsub) and np.array_equal(got.clus, clus)
    assert np.array_equal(got.blkPointToMin.view(np.uint64), ptm.view(np.uint64))
    if orc.ref_binary() is None:
        return
    src, ref = str(tmp_path / "src.hash"), str(tmp_path / "ref.hash")
    assert orc.build_and_write(recs, src, B=20) == 0
    r = subprocess.run([orc.ref_binary(), "-B", "20", "-ct", "1", "--readHash", src, "--hashDepthRange", "2", "13",
                        "--cluster", "0", "0", "--writeHash", ref], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    a, b = bytearray(open(ours, "rb").read()), bytearray(open(ref, "rb").read())
    assert len(a) == len(b)
    # blank the raw pointers: ArrayStruct.base of both Arrays and ClusterBlock.clusHash of every block
    hn, nb = int(ix.hashNumber), int(ix.nBlocksMax)
    off = 16 + (4 << 20) + 4 + 8 * hn
    for buf in (a, b):
        buf[off + 8:off + 16] = bytes(8)
    depth_dim = hf.depthDim
    off2 = off + 32 + 4 * depth_dim
    for buf in (a, b):
        buf[off2 + 8:off2 + 16] = bytes(8)
        for blk in range(hf.blkDim):
            o = off2 + 32 + 32 * blk + 16
            buf[o:o + 8] = bytes(8)
    assert a == b


def test_owner_thresholds_are_monotone_and_balance_the_mosh_density():
    """h10x_dist_owner_thresholds (the hash ranges of a multi-GPU build): monotone with the right ends for every rank
    count; cut at the quantiles of min (hash, hashRC) so that owners hold equal shares (flat = 1: equal widths, where
    owner 0 of 2 holds three quarters)."""
    import hash10x_b200
    L = hash10x_b200.load_library()
    L.h10x_dist_owner_thresholds.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
    rng = np.random.default_rng(11)
    k = 21
    top = 1 << (2 * k)
    moshes = np.minimum(rng.integers(0, top, 2_000_000, dtype=np.int64), rng.integers(0, top, 2_000_000, dtype=np.int64))
    for nranks in (1, 2, 3, 4, 8, 16):
        for flat in (0, 1):
            thr = np.zeros(nranks + 1, np.uint64)
            assert L.h10x_dist_owner_thresholds(k, nranks, flat, thr.ctypes.data) == 0
            t = thr.astype(np.int64)
            assert t[0] == 0 and t[-1] == top and (np.diff(t) >= 0).all()
            share = np.diff(np.searchsorted(np.sort(moshes), t)) / moshes.size
            if flat:
                assert abs(share[0] - (1 - (1 - 1 / nranks) ** 2)) < 0.01
            else:
                assert np.abs(share - 1 / nranks).max() < 0.01, (nranks, share)
    assert L.h10x_dist_owner_thresholds(0, 2, 0, thr.ctypes.data) != 0


def test_hash_writer_carries_the_split_index_like_the_reference(orc, tmp_path):
    """--clusterSplit's result through h10x_write_hash: clusterParent (hash10x.c:66) sits where the reference puts it; the
    file equals the reference's --clusterSplit --writeHash outside the raw pointers, and reads back."""
    import hash10x_b200
    from hash10x_b200 import binding
    p = orc.synth_params(seed=37, n_barcodes=120, pairs_min=20, pairs_max=80, genome_len=40_000, mol_len=8_000,
                         mol_per_barcode=3)
    recs = orc.synth_fqb(p)
    ix = orc.build(recs, B=20)
    _w, goff, good = orc.good_hashes(ix, 2, 13)
    clus, nsub, ptm = orc.cluster(ix, goff, good, 0, 0, 1)
    sp = orc.cluster_split(ix, clus, nsub, ptm)
    sp.B, sp.hashNumber, sp.nReads, sp.reserved = ix.B, ix.hashNumber, ix.nReads, 1      # H10X_INDEX_EXACT_BLOCKS
    sp.hashIndex, sp.hashValue, sp.hashDepth = ix.hashIndex, ix.hashValue, ix.hashDepth
    ours = str(tmp_path / "ours.hash")
    binding.write_hash(sp, ours)
    hf = hashfile.parse(ours)
    assert np.array_equal(hf.blkParent, sp.blkParent) and int(hf.blkParent.max()) > 0
    assert np.array_equal(hf.blkNRead, sp.blkNRead) and np.array_equal(hf.clusRaw, sp.clus)
    L = hash10x_b200.load_library()
    ci = binding.CIndex()
    err = C.create_string_buffer(256)
    assert L.h10x_read_hash(ours.encode(), 20, C.byref(ci), err, 256) == 0, err.value
    got = binding.Index(ci, L)
    L.h10x_index_free(C.byref(ci))
    assert np.array_equal(got.blkParent, sp.blkParent) and np.array_equal(got.clus, sp.clus)
    if orc.ref_binary() is None:
        return
    src, ref = str(tmp_path / "src.hash"), str(tmp_path / "ref.hash")
    assert orc.build_and_write(recs, src, B=20) == 0
    r = subprocess.run([orc.ref_binary(), "-B", "20", "-ct", "1", "--readHash", src, "--hashDepthRange", "2", "13",
                        "--cluster", "0", "0", "--clusterSplit", "--writeHash", ref], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    rf = hashfile.parse(ref)
    assert rf.size == hf.size and np.array_equal(rf.blkParent, hf.blkParent) and np.array_equal(rf.clusRaw, hf.clusRaw)
    assert np.array_equal(rf.blkNRead, hf.blkNRead) and np.array_equal(rf.blkNHash, hf.blkNHash)
    assert np.array_equal(rf.blkPointToMin.view(np.uint64), hf.blkPointToMin.view(np.uint64))
