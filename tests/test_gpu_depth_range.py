"""GPU --hashDepthRange ("next" row f1: hashWithinRangeBuild + goodHashesBuild, hash10x.c:528-539,738-766)
against the oracle's restatement (parity unpinned: the reference never prints goodHashes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(orc, recs, B, ranges):
    import hash10x_b200
    want_ix = orc.build(recs, B=B)
    within = None
    with hash10x_b200.Hash10xGPU(B=B) as g:
        got_ix = g.build_host(recs)
        assert np.array_equal(got_ix.clus, want_ix.clus)
        for dmin, dmax in ranges:
            within, off, good = orc.good_hashes(want_ix, dmin, dmax, within)
            w2, off2, good2 = g.depth_range(dmin, dmax)
            assert np.array_equal(w2, within)
            assert np.array_equal(off2, off)
            assert np.array_equal(good2, good)
    return off, good


def test_depth_range_matches_oracle(orc, gpu_lib):
    p = orc.synth_params(seed=71, n_barcodes=120, pairs_min=5, pairs_max=400, genome_len=150_000, mol_len=20_000)
    recs = orc.synth_fqb(p)
    off, good = _check(orc, recs, 21, [(2, 12), (2, 12), (30, 100), (1, 2)])     # flags accumulate over calls
    assert good.size > 0 and off[-1] == good.size


def test_depth_range_skips_blocks_over_65535_hashes(orc, gpu_lib):
    p = orc.synth_params(seed=4, n_barcodes=3, pairs_min=66000, pairs_max=66000, genome_len=3_000_000,
                         mol_per_barcode=50, mol_len=50_000)
    recs = orc.synth_fqb(p)
    off, good = _check(orc, recs, 22, [(1, 3)])
    assert off[1] == off[2] == off[3]        # both processed blocks have > 65535 hashes: empty lists
