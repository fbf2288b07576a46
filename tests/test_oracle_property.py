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
