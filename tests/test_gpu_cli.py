"""Drop-in check of the host program: `hash10x-b200` against the reference binary (oracle/_ref/hash10x)
on the same FQB with the same command chain - same stdout apart from resource lines, same .hash."""
import os
import re
import subprocess

import numpy as np
import pytest

import hashfile

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "hash10x_b200", "bin", "hash10x-b200")


def _clean(text):
    keep = []
    for ln in text.splitlines():
        if re.match(r"^\s*user\t", ln) or ln.startswith("total resources used"):
            continue
        keep.append(ln)
    return keep


def _run(exe, args, cwd):
    return subprocess.run([exe] + args, capture_output=True, text=True, cwd=cwd, timeout=600)


@pytest.mark.parametrize("extra", [[], ["-N", "1500"], ["-c", "400"], ["-k", "17", "-w", "13", "-r", "5"]])
def test_cli_matches_reference(orc, gpu_lib, tmp_path, extra):
    ref = orc.ref_binary("hash10x")
    if ref is None:
        pytest.skip("oracle/_ref/hash10x not built")
    assert os.path.exists(CLI), "hash10x-b200 not built"
    p = orc.synth_params(seed=31, n_barcodes=40, pairs_min=10, pairs_max=200)
    recs = orc.synth_fqb(p)
    recs.tofile(str(tmp_path / "in.fqb"))
    chain = extra + ["-B", "20", "--readFQB", "in.fqb", "--writeHash", "OUT.hash", "--hashStats", "--codeStats",
                     "--hashDepthRange", "2", "12"]
    a = _run(ref, [x.replace("OUT", "ref") for x in chain], str(tmp_path))
    b = _run(CLI, [x.replace("OUT", "gpu") for x in chain], str(tmp_path))
    assert a.returncode == 0 and b.returncode == 0, (a.stderr, b.stderr)
    la = [x.replace("ref.hash", "X.hash") for x in _clean(a.stdout)]
    lb = [x.replace("gpu.hash", "X.hash") for x in _clean(b.stdout)]
    assert la == lb
    ha, hb = hashfile.parse(str(tmp_path / "ref.hash")), hashfile.parse(str(tmp_path / "gpu.hash"))
    assert ha.size == hb.size and (ha.depthDim, ha.blkDim) == (hb.depthDim, hb.blkDim)
    hashfile.assert_strict_equal(ha, hb, table=True)
    # the GPU-written file goes back through the reference: --readHash then the same reports
    c = _run(ref, ["-B", "20", "--readHash", "gpu.hash", "--hashStats", "--codeStats"], str(tmp_path))
    d = _run(ref, ["-B", "20", "--readHash", "ref.hash", "--hashStats", "--codeStats"], str(tmp_path))
    assert c.returncode == 0
    assert [x.replace("gpu.hash", "X") for x in _clean(c.stdout)] == [x.replace("ref.hash", "X") for x in _clean(d.stdout)]
    # and our own --readHash reads both
    e = _run(CLI, ["-B", "20", "--readHash", "ref.hash", "--hashStats", "--codeStats"], str(tmp_path))
    assert e.returncode == 0
    assert [x.replace("ref.hash", "X") for x in _clean(e.stdout)] == [x.replace("ref.hash", "X") for x in _clean(d.stdout)]


def test_cli_errors_match_reference(orc, gpu_lib, tmp_path):
    ref = orc.ref_binary("hash10x")
    if ref is None:
        pytest.skip("oracle/_ref/hash10x not built")
    p = orc.synth_params(seed=31, n_barcodes=40, pairs_min=10, pairs_max=200)
    orc.synth_fqb(p).tofile(str(tmp_path / "in.fqb"))
    for chain in (["-B", "19", "--readFQB", "in.fqb"], ["-B", "31", "--readFQB", "in.fqb"],
                  ["-B", "20", "-c", "50", "--readFQB", "in.fqb"], ["-B", "20", "--readFQB", "missing.fqb"],
                  ["-B", "20", "--bogus"], ["-B", "21", "--readHash", "in.fqb"]):
        a, b = _run(ref, chain, str(tmp_path)), _run(CLI, chain, str(tmp_path))
        assert a.returncode != 0 and b.returncode == a.returncode, chain
        assert a.stderr.strip().splitlines()[-1] == b.stderr.strip().splitlines()[-1], chain


def test_cli_multi_gpu_matches_reference(orc, gpu_lib, tmp_path):
    """--gpus N (one process, a thread per GPU, NCCL inside the library) writes the reference's .hash"""
    ref = orc.ref_binary("hash10x")
    n = gpu_lib.h10x_gpu_device_count()
    if ref is None or n < 2:
        pytest.skip("needs oracle/_ref/hash10x and at least 2 GPUs")
    p = orc.synth_params(seed=61, n_barcodes=90, pairs_min=5, pairs_max=250)
    orc.synth_fqb(p).tofile(str(tmp_path / "in.fqb"))
    chain = ["-B", "21", "--readFQB", "in.fqb", "--writeHash", "OUT.hash", "--hashStats", "--codeStats"]
    a = _run(ref, [x.replace("OUT", "ref") for x in chain], str(tmp_path))
    assert a.returncode == 0, a.stderr
    for g in sorted({2, min(n, 8)}):
        b = _run(CLI, ["--gpus", str(g)] + [x.replace("OUT", "gpu%d" % g) for x in chain], str(tmp_path))
        assert b.returncode == 0, b.stderr
        la = [x.replace("ref.hash", "X.hash") for x in _clean(a.stdout)]
        lb = [x.replace("gpu%d.hash" % g, "X.hash") for x in _clean(b.stdout)
              if not x.startswith("COMMAND --gpus") and not x.startswith("NCCL version")]   # NCCL_DEBUG banner
        assert la == lb
        ha, hb = hashfile.parse(str(tmp_path / "ref.hash")), hashfile.parse(str(tmp_path / ("gpu%d.hash" % g)))
        assert ha.size == hb.size
        hashfile.assert_strict_equal(ha, hb, table=True)


def test_cli_crib_build_and_stats_match_reference(orc, gpu_lib, tmp_path):
    """--cribBuild genome1.fa genome2.fa (hash10x.c:426-510) and --hashStats / --codeStats counted on the device
    (h10x_gpu_crib_build, h10x_gpu_histogram): the reference's report lines, for genomes made of the reads themselves
    (hom), of reads only one genome has (het), of repeated pieces (mul), with N, lower case, wrapped lines and a
    sequence shorter than k"""
    import fqbtools
    ref = orc.ref_binary("hash10x")
    if ref is None:
        pytest.skip("oracle/_ref/hash10x not built")
    p = orc.synth_params(seed=71, n_barcodes=30, pairs_min=10, pairs_max=120, read_len=151)
    recs = orc.synth_fqb(p)
    recs.tofile(str(tmp_path / "in.fqb"))
    _f1, f2 = fqbtools.fastq_from_fqb(recs, 151)
    reads = f2.split(b"\n")[1::4]
    rng = np.random.default_rng(5)

    def fasta(chunks, wrap):
        out = []
        for i, c in enumerate(chunks):
            out.append(b">chr%d some description" % (i + 1))
            out += [c[j:j + wrap] for j in range(0, len(c), wrap)]
        return b"\n".join(out) + b"\n"
    a = b"".join(reads[0:300])
    b = b"".join(reads[300:500])
    c = b"".join(reads[500:600])
    g1 = fasta([a, b + b"NNNN" + b[:3000].lower(), b"ACGTACGT", c], 60)          # b's head twice: mul; a tiny sequence
    g2 = fasta([a[:20000] + b"N" * 7 + bytes(rng.choice(list(b"ACGT"), 5000).tolist()), b[4000:], b"".join(reads[600:700])], 71)
    (tmp_path / "g1.fa").write_bytes(g1)
    (tmp_path / "g2.fa").write_bytes(g2)
    chain = ["-B", "20", "--readFQB", "in.fqb", "--cribBuild", "g1.fa", "g2.fa", "--hashStats", "--codeStats"]
    ra, rb = _run(ref, chain, str(tmp_path)), _run(CLI, chain, str(tmp_path))
    assert ra.returncode == 0 and rb.returncode == 0, (ra.stderr, rb.stderr)
    la, lb = _clean(ra.stdout), _clean(rb.stdout)
    assert any("known and" in x for x in la) and any(x.startswith("    hom") for x in la)
    assert la == lb
    # a bad character ends that genome's walk with the reference's message
    (tmp_path / "g3.fa").write_bytes(g1.replace(b"NNNN", b"NN*N"))
    chain = ["-B", "20", "--readFQB", "in.fqb", "--cribBuild", "g3.fa", "g2.fa"]
    ra, rb = _run(ref, chain, str(tmp_path)), _run(CLI, chain, str(tmp_path))
    assert ra.returncode == rb.returncode == 0 and _clean(ra.stdout) == _clean(rb.stdout)
    assert [x for x in ra.stderr.splitlines() if x.startswith("Bad char")] == [x for x in rb.stderr.splitlines() if x.startswith("Bad char")]
