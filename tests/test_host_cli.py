"""CPU tests of the C host program's host-only commands (no GPU): `hash10x-b200 --readHash ... --hashStats --codeStats
--writeHash` against the reference binary on the same .hash file (this container only: oracle/_ref)."""
import os
import subprocess

import numpy as np
import pytest

import hashfile


def _exe():
    import hash10x_b200
    return os.path.join(os.path.dirname(hash10x_b200.__file__), "bin", "hash10x-b200")


def _report_lines(text):
    return [ln for ln in text.splitlines() if "_HIST " in ln or "_STATS " in ln]


def test_read_hash_stats_and_rewrite_match_the_reference(orc, tmp_path):
    if orc.ref_binary() is None:
        pytest.skip("oracle/_ref/hash10x not built (no /root/reference on this machine)")
    from hash10x_b200 import binding
    p = orc.synth_params(seed=37, n_barcodes=120, pairs_min=20, pairs_max=80, genome_len=40_000, mol_len=8_000,
                         mol_per_barcode=3)
    recs = orc.synth_fqb(p)
    ix = orc.build(recs, B=20)
    _w, goff, good = orc.good_hashes(ix, 2, 13)
    ix.clus, ix.blkNSub, ix.blkPointToMin = orc.cluster(ix, goff, good, 0, 0, 1)
    src = str(tmp_path / "clustered.hash")
    binding.write_hash(ix, src)
    outs = {}
    for name, exe in (("ours", _exe()), ("ref", orc.ref_binary())):
        dst = str(tmp_path / (name + ".hash"))
        r = subprocess.run([exe, "-B", "20", "--readHash", src, "--hashStats", "--codeStats", "--writeHash", dst],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        outs[name] = (r.stdout, hashfile.parse(dst))
    a, b = _report_lines(outs["ours"][0]), _report_lines(outs["ref"][0])
    assert a == b and any(ln.startswith("CODE_CLUSTER_STATS") for ln in a) and any(ln.startswith("HASH_COUNT_STATS") for ln in a)
    ha, hb = outs["ours"][1], outs["ref"][1]
    assert ha.size == hb.size
    hashfile.assert_strict_equal(ha, hb, table=True)
    assert np.array_equal(ha.clusRaw, hb.clusRaw) and np.array_equal(ha.blkNSub, hb.blkNSub)
    assert np.array_equal(ha.blkPointToMin.view(np.uint64), hb.blkPointToMin.view(np.uint64))


def test_cluster_needs_the_gpu_index(orc, tmp_path):
    """--hashDepthRange / --cluster after --readHash without a CUDA device: there is no resident GPU index and no CPU
    version of either in this program, so it dies saying so rather than falling back."""
    assert orc.build_and_write(orc.synth_fqb(orc.synth_params(seed=3, n_barcodes=12, pairs_min=3, pairs_max=20)),
                               str(tmp_path / "a.hash"), B=20) == 0
    r = subprocess.run([_exe(), "-B", "20", "--readHash", str(tmp_path / "a.hash"), "--cluster", "0", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "!! you must set hashDepthRange before cluster" in r.stdout
    import hash10x_b200
    if hash10x_b200.load_library().h10x_gpu_device_count() > 0:
        return          # with a GPU the read index is loaded onto it and --cluster runs (tests/test_gpu_cluster.py)
    r = subprocess.run([_exe(), "-B", "20", "--readHash", str(tmp_path / "a.hash"), "--hashDepthRange", "1", "50",
                        "--cluster", "0", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "FATAL ERROR: --hashDepthRange runs on the GPU-resident index" in r.stderr
