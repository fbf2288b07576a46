"""Property test (CPU, this container only): random small FQBs - runs of random length, barcode words that
include 0 (the all-A barcode whose chunk-boundary behaviour is the oddest thing readFQB does), low-complexity
reads, random -c and -N - built by the reference binary and by the oracle must give the same file or the same
die()."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import fqbtools
import hashfile


@st.composite
def fqb_case(draw):
    n_runs = draw(st.integers(1, 7))
    words = [draw(st.sampled_from([0, 0, 5, 9, 77, 0x3FFFFFFF, 0xFFFFFFFF])) for _ in range(n_runs)]
    counts = [draw(st.integers(1, 9)) for _ in range(n_runs)]
    kinds = [draw(st.sampled_from(["rand", "rand", "polyA", "polyC", "repeat"])) for _ in range(n_runs)]
    seed = draw(st.integers(0, 2 ** 31))
    chunk = draw(st.integers(1, 14))
    N = draw(st.sampled_from([0, 0, 1, 3, 7, 11, 19, 40]))
    return words, counts, kinds, seed, chunk, N


_records = fqbtools.mixed_records


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(case=fqb_case())
def test_oracle_equals_reference_on_random_small_inputs(orc, tmp_path_factory, case):
    if orc.ref_binary() is None:
        pytest.skip("oracle/_ref/hash10x not built (no /root/reference on this machine)")
    words, counts, kinds, seed, chunk, N = case
    recs = _records(words, counts, kinds, seed)
    d = tmp_path_factory.mktemp("prop")
    fqb = str(d / "a.fqb")
    recs.tofile(fqb)
    r = orc.run_reference(fqb, str(d / "r.hash"), B=20, chunk=chunk, N=N if N else None)
    st_ = orc.build_and_write(recs, str(d / "o.hash"), B=20, chunk=chunk, N=N)
    if r.returncode != 0:
        assert "chunkSize too small" in r.stderr and st_ == 2, (r.stderr, st_)
        return
    assert st_ == 0
    a, b = hashfile.parse(str(d / "r.hash")), hashfile.parse(str(d / "o.hash"))
    assert a.size == b.size
    hashfile.assert_strict_equal(a, b, table=True)


@st.composite
def cluster_case(draw):
    seed = draw(st.integers(1, 10 ** 6))
    nb = draw(st.integers(20, 160))
    pmin = draw(st.integers(5, 60))
    pmax = pmin + draw(st.integers(0, 80))
    genome = draw(st.sampled_from([20_000, 40_000, 90_000]))
    mol = draw(st.sampled_from([4_000, 9_000, 15_000]))
    mpb = draw(st.integers(1, 6))
    dmin = draw(st.integers(1, 6))
    dmax = dmin + draw(st.integers(2, 120))
    thr = draw(st.integers(1, 6))
    cmin = draw(st.sampled_from([0, 0, 1, 3, 10]))
    cmax = draw(st.sampled_from([0, 0, nb // 2 + 5, nb + 1]))
    return seed, nb, pmin, pmax, genome, mol, mpb, dmin, dmax, thr, cmin, cmax


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(case=cluster_case())
def test_cluster_oracle_equals_reference_on_random_cases(orc, tmp_path_factory, case):
    """--hashDepthRange + --cluster (hash10x.c:528-539,738-868): random synthetic data sets, depth ranges, thresholds
    and code ranges through the reference binary (from a .hash with zeroed subCluster bytes) and the oracle."""
    if orc.ref_binary() is None:
        pytest.skip("oracle/_ref/hash10x not built (no /root/reference on this machine)")
    import subprocess
    seed, nb, pmin, pmax, genome, mol, mpb, dmin, dmax, thr, cmin, cmax = case
    p = orc.synth_params(seed=seed, n_barcodes=nb, pairs_min=pmin, pairs_max=pmax, genome_len=genome, mol_len=mol,
                         mol_per_barcode=mpb)
    recs = orc.synth_fqb(p)
    d = tmp_path_factory.mktemp("cprop")
    src, dst = str(d / "o.hash"), str(d / "r.hash")
    assert orc.build_and_write(recs, src, B=20) == 0
    ix = orc.build(recs, B=20)
    if cmax > int(ix.nBlocksMax) or (cmax and cmin >= cmax):
        cmax = 0
    r = subprocess.run([orc.ref_binary(), "-B", "20", "-ct", str(thr), "--readHash", src, "--hashDepthRange", str(dmin),
                        str(dmax), "--cluster", str(cmin), str(cmax), "--writeHash", dst],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    ref = hashfile.parse(dst)
    _w, goff, good = orc.good_hashes(ix, dmin, dmax)
    clus, nsub, ptm = orc.cluster(ix, goff, good, cmin, cmax, thr)
    assert np.array_equal(ref.blkNSub, nsub)
    assert np.array_equal(ref.blkPointToMin.view(np.uint64), ptm.view(np.uint64))
    assert np.array_equal(ref.clusRaw, clus)
