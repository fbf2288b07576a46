"""GPU parity: the CUDA build (through the C ABI) against the CPU oracle, bit-exact.

Strict mode: identical bin ids, hashValue order, depths, per-block (id, read)
lists, hash->code CSR and a byte-identical hashIndex table.
"""
import numpy as np
import pytest

import fqbtools
import hashfile

pytestmark = pytest.mark.gpu


def _gpu(**kw):
    import hash10x_b200
    return hash10x_b200.Hash10xGPU(**kw)


def _compare(orc, recs, B=20, k=21, w=31, r=17, N=0, chunk=100000, flags=0):
    f1 = orc.factor1(r)
    want = orc.build(recs, k=k, w=w, factor1_=f1, B=B, N=N, chunk=chunk)
    assert want.status == 0, want.status_text
    if not (flags & 16):       # the round-1 library-sort tail must give the same index as the hand-written one
        _compare(orc, recs, B=B, k=k, w=w, r=r, N=N, chunk=chunk, flags=flags | 16)
    with _gpu(k=k, w=w, r=r, B=B, N=N, chunkSize=chunk, flags=flags, factor1=f1) as g:
        got = g.build_host(recs)
        st = g.stats()
    a, b = hashfile.from_index(want), hashfile.from_index(got)
    assert got.nReads == want.nReads and got.nHashes == want.nHashes
    hashfile.assert_strict_equal(a, b, table=True)
    assert np.array_equal(got.clus, want.clus)          # bytes 6-7 are zero on both sides
    assert np.array_equal(got.blkOff, want.blkOff)
    assert np.array_equal(got.codeOff, want.codeOff)
    assert np.array_equal(got.codes, want.codes)
    hashfile.check_table(b)
    return want, got, st


def test_record_moshes_match_oracle(orc, gpu_lib):
    rng = np.random.default_rng(5)
    recs = fqbtools.random_records(rng, [0x1234567, 0x89ABCDE], [40, 24])
    recs = np.concatenate([recs, fqbtools.const_records(77, 3, 0, 0), fqbtools.const_records(78, 2, 1, 1)])
    with _gpu(B=20) as g:
        off, hashes = g.record_moshes(recs)
    for i, rec in enumerate(recs):
        h, _pos, _which = orc.record_moshes(rec)
        assert np.array_equal(hashes[int(off[i]):int(off[i + 1])], h), i
    # poly-A: every k-mer of both reads is a mosh with hash 0 (107 + 130); poly-C: none
    assert int(off[65] - off[64]) == 237 and int(off[68] - off[67]) == 0


@pytest.mark.parametrize("read_len", [151, 160])
def test_fused_kernel_moshes_match_oracle(orc, gpu_lib, read_len):
    """The hash loop of k_fused_block ITSELF (not the generic k_moshes of the test above): in lean mode the keys it
    stores for a block are the block's moshes as seqAddHashes / moshRCnext generate them (hash10x.c:123-132,
    seqhash.c:154-195) - same multiset of (hash, read index), record by record, duplicates of a k-mer included."""
    p = orc.synth_params(seed=17, n_barcodes=40, pairs_min=1, pairs_max=300, read_len=read_len)
    recs = orc.synth_fqb(p)
    rng = np.random.default_rng(8)
    recs = np.concatenate([recs, fqbtools.random_records(rng, [0x3456789, 0x1111111], [70, 1])])    # + a run, + the unhashed last run
    with _gpu(B=21) as g:
        off, hashes, reads, lean = g.block_keys(recs)
    assert lean, "the hand-written tail (and with it the lean fused kernel) should have been chosen"
    word0 = recs[:, 0]
    starts = np.flatnonzero(np.r_[True, word0[1:] != word0[:-1]])
    assert off.size - 1 == starts.size - 1                      # every run but the last
    checked = 0
    for b in range(off.size - 1):
        r0, r1 = int(starts[b]), int(starts[b + 1])
        got = sorted(zip(hashes[int(off[b]):int(off[b + 1])].tolist(), reads[int(off[b]):int(off[b + 1])].tolist()))
        if not got and r1 - r0 > 0 and int(off[b + 1]) == int(off[b]):
            pass                                                # generic path took the block (none here) or no mosh at all
        want = []
        for i in range(r0, r1):
            h, _pos, _which = orc.record_moshes(recs[i])
            want += [(int(x), (i - r0) & 0xFFFF) for x in h]
        if not want:
            want = [(0, 0)]                                     # the phantom entry of a block without moshes (hash10x.c:167-168)
        assert got == sorted(want), b
        checked += len(want)
    assert checked > 10_000


@pytest.mark.parametrize("seed,nb,pmin,pmax", [(1, 30, 1, 40), (2, 200, 20, 300), (3, 12, 500, 1500)])
def test_synthetic_strict(orc, gpu_lib, seed, nb, pmin, pmax):
    p = orc.synth_params(seed=seed, n_barcodes=nb, pairs_min=pmin, pairs_max=pmax)
    recs = orc.synth_fqb(p)
    want, got, st = _compare(orc, recs, B=21)
    assert st["nBins"] == want.hashNumber - 1


@pytest.mark.parametrize("k,w,r", [(21, 31, 17), (16, 32, 17), (31, 7, 3), (11, 1, 9), (25, 12, 17), (5, 3, 1),
                                   (21, 1, 17), (23, 31, 17), (13, 3, 2), (22, 64, 5), (15, 30, 17), (21, 62, 17)])
def test_parameters(orc, gpu_lib, k, w, r):
    p = orc.synth_params(seed=11, n_barcodes=25, pairs_min=3, pairs_max=60, genome_len=50_000, mol_len=5_000)
    recs = orc.synth_fqb(p)
    _compare(orc, recs, B=22, k=k, w=w, r=r)


def test_quirks_last_block_phantom_polyA(orc, gpu_lib):
    rng = np.random.default_rng(9)
    recs = np.concatenate([
        fqbtools.random_records(rng, [11], [5]),
        fqbtools.const_records(12, 4, 1, 1),      # poly-C block: no moshes -> phantom {hash 0, read 0}
        fqbtools.random_records(rng, [13], [7]),
        fqbtools.const_records(14, 3, 0, 0),      # poly-A block: hash 0 moshes, shares the phantom's bin
        fqbtools.const_records(15, 2, 1, 1),      # second phantom block
        fqbtools.random_records(rng, [16], [6]),  # last block: never hashed
    ])
    want, got, _ = _compare(orc, recs)
    assert got.blkNHash[2] == 1 and got.blkNHash[4] == 1 and got.blkNHash[6] == 0
    z = int(np.where(got.hashValue[1:] == 0)[0][0]) + 1
    assert got.hashDepth[z] == 3


def test_generic_only_flag_and_fallbacks(orc, gpu_lib):
    # the fused shared-memory path and the generic global-memory path must agree: force the generic
    # one (flag 8), and mix in blocks the fused path must hand back (poly-A flood, oversized block)
    rng = np.random.default_rng(17)
    p = orc.synth_params(seed=13, n_barcodes=30, pairs_min=5, pairs_max=400)
    recs = orc.synth_fqb(p)
    _, _, st = _compare(orc, recs, B=21, flags=8)
    assert st["fusedBlocks"] == 0
    _, _, st = _compare(orc, recs, B=21)
    assert st["fusedBlocks"] == 29 and st["genericBlocks"] == 0
    mixed = np.concatenate([
        recs[:3000],
        fqbtools.const_records(0x0AAAAAA1, 300, 0, 0),          # 300 poly-A pairs: 71100 identical moshes
        fqbtools.random_records(rng, [0x0AAAAAA2], [3]),
        np.concatenate([fqbtools.const_records(0x0AAAAAA3, 40, 0, 0), fqbtools.random_records(rng, [0x0AAAAAA3], [40])]),
        recs[3000:],
    ])
    _, _, st = _compare(orc, mixed, B=21)
    assert st["genericBlocks"] >= 1 and st["fusedBlocks"] >= 25
    # blocks of the largest shared-memory class (many duplicate hashes: a 2 Mb genome) stay on the fused path ...
    big = orc.synth_params(seed=14, n_barcodes=4, pairs_min=1500, pairs_max=2600, genome_len=2_000_000)
    _, _, st = _compare(orc, orc.synth_fqb(big), B=21)
    assert st["fusedBlocks"] + st["genericBlocks"] == 3
    # ... and blocks beyond it (> ~2800 read pairs) are handed to the generic path by the host
    huge = orc.synth_params(seed=15, n_barcodes=3, pairs_min=3000, pairs_max=3400, genome_len=2_000_000)
    _, _, st = _compare(orc, orc.synth_fqb(huge), B=21)
    assert st["genericBlocks"] == 2 and st["fusedBlocks"] == 0


def test_edge_sizes(orc, gpu_lib):
    rng = np.random.default_rng(2)
    one = fqbtools.random_records(rng, [5], [9])
    for recs in (np.zeros((0, 30), np.uint32), one, one[:1],
                 fqbtools.random_records(rng, [5, 6], [1, 1]),
                 fqbtools.random_records(rng, [9, 8, 9, 8], [2, 3, 2, 1])):   # barcode recurs: separate runs
        _compare(orc, recs)


def test_read_index_wraps_at_16_bits(orc, gpu_lib):
    # 66000 pairs in one barcode: read index > 65535 is stored truncated (hash10x.c:37,180)
    p = orc.synth_params(seed=4, n_barcodes=3, pairs_min=66000, pairs_max=66000, genome_len=3_000_000,
                         mol_per_barcode=50, mol_len=50_000)
    recs = orc.synth_fqb(p)
    _compare(orc, recs, B=22)


def test_N_limit_and_chunks(orc, gpu_lib):
    p = orc.synth_params(seed=6, n_barcodes=40, pairs_min=10, pairs_max=90)
    recs = orc.synth_fqb(p)
    for N in (1, 57, 1000, recs.shape[0], recs.shape[0] + 5):
        _compare(orc, recs, N=N)
    for chunk in (91, 100, 1000):
        _compare(orc, recs, chunk=chunk)
        _compare(orc, recs, chunk=chunk, N=777)


def test_chunk_too_small(orc, gpu_lib):
    import hash10x_b200
    p = orc.synth_params(seed=6, n_barcodes=40, pairs_min=10, pairs_max=90)
    recs = orc.synth_fqb(p)
    n, off = orc.synth_layout(p)
    big = int(np.diff(off.astype(np.int64)).max())
    for chunk in (big, big - 1, 10):
        want = orc.build(recs, B=20, chunk=chunk)
        assert want.status == 2
        with _gpu(B=20, chunkSize=chunk) as g:
            with pytest.raises(hash10x_b200.H10xError) as e:
                g.build_host(recs)
            assert e.value.code == 2 and "chunkSize too small" in e.value.msg
    # exactly at the -N limit the reference leaves the loop before the check (hash10x.c:202)
    first = int(off[1])
    for N in (first, first + 1):
        want = orc.build(recs, B=20, chunk=first, N=N)
        with _gpu(B=20, chunkSize=first, N=N) as g:
            if want.status == 2:
                with pytest.raises(hash10x_b200.H10xError):
                    g.build_host(recs)
            else:
                got = g.build_host(recs)
                hashfile.assert_strict_equal(hashfile.from_index(want), hashfile.from_index(got))


def test_table_too_small(orc, gpu_lib):
    import hash10x_b200
    # B=20 holds at most 2^18-2 bins; ~330k distinct hashes overflow it
    p = orc.synth_params(seed=8, n_barcodes=60, pairs_min=800, pairs_max=1000, genome_len=20_000_000,
                         mol_len=100_000, mol_per_barcode=20)
    recs = orc.synth_fqb(p)
    want = orc.build(recs, B=20)
    assert want.status == 1
    with _gpu(B=20) as g:
        with pytest.raises(hash10x_b200.H10xError) as e:
            g.build_host(recs)
        assert e.value.code == 1 and "hashTableSize is too small" in e.value.msg
    _compare(orc, recs, B=21)


def test_all_A_barcode_chunk_boundary(orc, gpu_lib):
    # barcode word 0 ending exactly on a chunk boundary is glued to the next run by the reference's
    # `if (!barcode) barcode = u[0]` (hash10x.c:212); elsewhere it is an ordinary run
    rng = np.random.default_rng(21)
    recs = fqbtools.random_records(rng, [7, 0, 9, 0, 5, 3], [6, 4, 5, 3, 4, 2])
    for chunk in (10, 11, 12, 7, 8, 9, 15, 18, 100):
        want = orc.build(recs, B=20, chunk=chunk)
        if want.status == 0:
            _compare(orc, recs, chunk=chunk)
    starts0 = fqbtools.random_records(rng, [0, 4, 0, 2], [5, 5, 5, 5])
    for chunk in (5, 6, 10, 11, 100):
        want = orc.build(starts0, B=20, chunk=chunk)
        if want.status == 0:
            _compare(orc, starts0, chunk=chunk)


def test_build_file_and_device_paths(orc, gpu_lib, tmp_path):
    p = orc.synth_params(seed=12, n_barcodes=50, pairs_min=5, pairs_max=80)
    recs = orc.synth_fqb(p)
    path = str(tmp_path / "x.fqb")
    with open(path, "wb") as f:
        f.write(recs.tobytes())
        f.write(b"\x01\x02\x03")           # trailing partial record is ignored (fread of whole records)
    want = orc.build(recs, B=20)
    with _gpu(B=20) as g:
        got = g.build_file(path)
        hashfile.assert_strict_equal(hashfile.from_index(want), hashfile.from_index(got))
        hp = str(tmp_path / "x.hash")
        g.build_file_to_hash(path, hp)
    op = str(tmp_path / "o.hash")
    assert orc.build_and_write(recs, op, B=20) == 0
    a, b = open(hp, "rb").read(), open(op, "rb").read()
    assert a == b                           # the two writers agree byte for byte
    hf = hashfile.parse(hp)
    hashfile.assert_strict_equal(hashfile.from_index(want), hf)


def test_bad_parameters(gpu_lib):
    import hash10x_b200
    for kw in (dict(B=19), dict(B=31), dict(k=0), dict(k=32), dict(w=0)):
        with pytest.raises(hash10x_b200.H10xError):
            _gpu(**kw)
    with _gpu(B=31, flags=1) as g:        # H10X_FLAG_WIDE_B models the "B bound relaxed" oracle
        assert g.ctx


def test_random_small_inputs_match_oracle(orc, gpu_lib):
    """seeded twin of tests/test_oracle_property.py on the GPU: barcode word 0 runs, low-complexity reads,
    random -c / -N; same index or the same error as the oracle (which that test pins to the reference)"""
    import hash10x_b200
    rng = np.random.default_rng(12345)
    pool = [0, 0, 5, 9, 77, 0x3FFFFFFF, 0xFFFFFFFF]
    kinds_pool = ["rand", "rand", "polyA", "polyC", "repeat"]
    n_ok = n_err = 0
    for _case in range(60):
        n_runs = int(rng.integers(1, 8))
        words = [pool[int(rng.integers(0, len(pool)))] for _ in range(n_runs)]
        counts = [int(rng.integers(1, 10)) for _ in range(n_runs)]
        kinds = [kinds_pool[int(rng.integers(0, len(kinds_pool)))] for _ in range(n_runs)]
        chunk = int(rng.integers(1, 15))
        N = [0, 0, 1, 3, 7, 11, 19, 40][int(rng.integers(0, 8))]
        recs = fqbtools.mixed_records(words, counts, kinds, int(rng.integers(0, 2 ** 31)))
        want = orc.build(recs, B=20, chunk=chunk, N=N)
        with _gpu(B=20, chunkSize=chunk, N=N) as g:
            if want.status != 0:
                with pytest.raises(hash10x_b200.H10xError) as e:
                    g.build_host(recs)
                assert e.value.code == want.status
                n_err += 1
                continue
            got = g.build_host(recs)
        hashfile.assert_strict_equal(hashfile.from_index(want), hashfile.from_index(got), table=True)
        assert np.array_equal(got.codes, want.codes) and np.array_equal(got.clus, want.clus)
        n_ok += 1
    assert n_ok >= 20 and n_err >= 5


def test_mosh_count_is_the_raw_selection_count(orc, gpu_lib):
    # stats.nMoshes = selected k-mers of the processed blocks (seqhash.c:171,189), before the per-block dedup
    p = orc.synth_params(seed=31, n_barcodes=12, pairs_min=20, pairs_max=300)
    recs = orc.synth_fqb(p)
    n, off = orc.synth_layout(p)
    want = sum(len(orc.record_moshes(rec)[0]) for rec in recs[:int(off[-2])])     # the last run is never hashed
    for flags in (0, 8):
        with _gpu(B=20, flags=flags) as g:
            g.build_host(recs)
            st = g.stats()
        assert st["nMoshes"] == want, flags
        assert st["nHashes"] < st["nMoshes"]


def test_tail_big_sub_ranges_and_deep_bins(orc, gpu_lib, monkeypatch):
    # the hand-written tail: sub-ranges beyond the shared-memory sort's capacity go through the library sort +
    # k_sr_heads_big (forced here by a tiny capacity), and a hash held by every block (poly-A read pairs) makes one
    # very deep bin; both must give the oracle's index
    rng = np.random.default_rng(23)
    p = orc.synth_params(seed=29, n_barcodes=150, pairs_min=10, pairs_max=120, genome_len=60_000, mol_len=5_000)
    recs = orc.synth_fqb(p)
    n, off = orc.synth_layout(p)
    parts = []
    for b in range(150):                      # one poly-A pair at the end of every barcode run: hash 0 in every block
        run = recs[int(off[b]):int(off[b + 1])]
        polyA = fqbtools.const_records(int(run[0, 0]), 1, 0, 0)
        polyA[0, 0] = run[0, 0]
        parts += [run, polyA]
    deep = np.concatenate(parts)
    want, got, st = _compare(orc, deep, B=21)
    assert st["tailPath"] == 2
    z = int(np.where(got.hashValue[1:] == 0)[0][0]) + 1
    assert got.hashDepth[z] == 149
    monkeypatch.setenv("H10X_SR_CAP", "48")
    want, got, st = _compare(orc, deep, B=21)
    assert st["tailPath"] == 2
    want, got, st = _compare(orc, recs, B=21)


@pytest.mark.parametrize("slab", [1, 97, 5000])
def test_streamed_host_build_matches_oracle(orc, gpu_lib, monkeypatch, slab):
    # h10x_gpu_build_host copies the file in slabs and hashes the runs of every slab that has landed (prefuse_streamed):
    # tiny slabs put run boundaries, slab boundaries and open runs in every relation; the index must be the oracle's
    p = orc.synth_params(seed=41, n_barcodes=60, pairs_min=1, pairs_max=260)
    recs = orc.synth_fqb(p)
    monkeypatch.setenv("H10X_STREAM_SLAB", str(slab))
    want, got, st = _compare(orc, recs, B=20)
    assert st["nBins"] == want.hashNumber - 1
    # a barcode word of 0 glues runs in the reference's chunk loop: the streamed lists are dropped, the classic stages run
    recs2 = np.concatenate([fqbtools.const_records(0, 3, 0, 1), recs])
    _compare(orc, recs2, B=20)
    monkeypatch.setenv("H10X_NO_STREAM", "1")
    _compare(orc, recs, B=20)


def test_lazy_codes(orc, gpu_lib):
    # H10X_FLAG_LAZY_CODES: the hash->code lists stay on the device (what hash10x-b200 asks for); fetched on demand
    import hash10x_b200
    p = orc.synth_params(seed=43, n_barcodes=40, pairs_min=5, pairs_max=200)
    recs = orc.synth_fqb(p)
    want = orc.build(recs, B=20)
    with _gpu(B=20, flags=hash10x_b200.FLAG_LAZY_CODES) as g:
        got = g.build_host(recs)
        assert got.codes is None and got.codeOff is None
        assert np.array_equal(got.clus, want.clus) and np.array_equal(got.hashDepth, want.hashDepth)
        code_off, codes = g.download_codes()
        _gw, goff, good = g.depth_range(2, 40)          # --hashDepthRange / --cluster read the resident lists
        clus, nsub, _ptm, _ms = g.cluster(0, 0, 2)
    assert np.array_equal(code_off, want.codeOff) and np.array_equal(codes, want.codes)
    _w, wgoff, wgood = orc.good_hashes(want, 2, 40)
    wclus, wnsub, _wptm = orc.cluster(want, wgoff, wgood, 0, 0, 2)
    assert np.array_equal(goff, wgoff) and np.array_equal(good, wgood)
    assert np.array_equal(clus, wclus) and np.array_equal(nsub, wnsub)


def test_wide_table_B31(orc, gpu_lib):
    # H10X_FLAG_WIDE_B (hash10x-b200 --wideB): -B 31..34 as README.md:55 / moshset.c:17 describe.  Bin ids, values, depths
    # and lists do not depend on B (they equal the B = 30 oracle's); the 8 GiB hashIndex is the table the oracle's
    # restatement of hashIndexFind (hash10x.c:139-152) fills when it is allowed 31 bits
    import psutil
    import hash10x_b200
    if psutil.virtual_memory().available < (56 << 30):
        pytest.skip("needs about 40 GB of host memory for the two 8 GiB tables and their copies")
    p = orc.synth_params(seed=47, n_barcodes=30, pairs_min=10, pairs_max=150)
    recs = orc.synth_fqb(p)
    want30 = orc.build(recs, B=30)
    with pytest.raises(hash10x_b200.H10xError):
        _gpu(B=31)
    with _gpu(B=31, flags=hash10x_b200.FLAG_WIDE_B) as g:
        got = g.build_host(recs)
    assert np.array_equal(got.hashValue, want30.hashValue) and np.array_equal(got.hashDepth, want30.hashDepth)
    assert np.array_equal(got.clus, want30.clus) and np.array_equal(got.codes, want30.codes)
    assert got.hashIndex.size == 1 << 31
    del want30
    want31 = orc.build(recs, B=31, maxB=34)
    assert want31.status == 0 and np.array_equal(got.hashIndex, want31.hashIndex)
