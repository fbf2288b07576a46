"""The fq2b stage (FASTQ pair -> FQB records, whitelist barcode correction, bsort grouping): fq2b.c:33-208, README.md:25-26.

CPU: the oracle restatement (oracle/fq2b_oracle.py) against the golden vectors the unmodified reference binary produced
(tests/golden/golden_fq2b.json, made by tests/golden/make_golden_fq2b.py).  GPU: h10x_gpu_fq2b through the C ABI against
the oracle, bit-exact: records, counters, the reference's die() texts; the sorted variant by its properties."""
import json
import os
import zlib

import numpy as np
import pytest

from oracle import fq2b_oracle as fo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_fq2b.json")))


def _case_inputs(c):
    wl = fo.synth_whitelist(*c["wl"]) if c["wl"] else None
    f1, f2 = fo.synth_fastq(c["seed"], c["n"], c["l1"], c["l2"] or 8, wl, lower=c.get("lower", False))
    return f1, (f2 if c["l2"] else None), ([fo.pack_barcode(s) for s in wl] if wl else None)


@pytest.mark.parametrize("name", sorted(GOLD["cases"]))
def test_oracle_matches_reference_binary_golden(name):
    g = GOLD["cases"][name]
    f1, f2, wl = _case_inputs(g["params"])
    recs, st = fo.fq2b(f1, f2, wl)
    data = recs.astype("<u4").tobytes()
    assert len(data) == g["nbytes"] and "%08x" % zlib.crc32(data) == g["crc32"]
    if "nBad" in g:
        assert (st["nBad"], st["nFixed"], st["nFixBase"]) == (g["nBad"], g["nFixed"], g["nFixBase"])


def test_oracle_pack_rules():
    # fq2b.c:36-41: `while (len > 16)` - a 32-base line is 16 + 16, a 17-base line 16 + 1 right-aligned; N and others -> A
    assert fo.seq_pack(b"ACGT" * 8) == [0x1B1B1B1B, 0x1B1B1B1B]
    assert fo.seq_pack(b"T" * 17) == [0xFFFFFFFF, 3]
    assert fo.seq_pack(b"NnXacgt") == [0b00000000011011]
    assert fo.qual_pack(bytes([55, 56, 33, 74])) == [0b0101]
    assert fo.switch_base(0, 1 + 4 * 15 + 3) == 0xC0000000


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GOLD["cases"]))
def test_gpu_fq2b_matches_oracle_and_golden(gpu_lib, name):
    import hash10x_b200
    g = GOLD["cases"][name]
    f1, f2, wl = _case_inputs(g["params"])
    want, wst = fo.fq2b(f1, f2, wl)
    with hash10x_b200.Hash10xGPU(B=20) as gp:
        got, st, dptr = gp.fq2b(f1, f2, wl)
        assert dptr
        assert got.shape == want.shape and np.array_equal(got, want)
        assert "%08x" % zlib.crc32(got.astype("<u4").tobytes()) == g["crc32"]
        for k in ("nRead", "nRecords", "nBad", "nFixed", "nFixBase", "s1Len", "recWords"):
            assert st[k] == wst[k], k
        # + bsort: the same records, grouped by their first four bytes, input order kept inside a group
        srt, sst, _ = gp.fq2b(f1, f2, wl, sort=True)
        assert np.array_equal(srt, fo.bsort(want)) and sst["nRecords"] == wst["nRecords"]
        key = srt[:, 0].astype(np.uint32).byteswap()
        assert np.all(key[1:] >= key[:-1])
        # a second call with the same whitelist reuses the table; without one nothing is dropped
        again, _, _ = gp.fq2b(f1, f2, wl)
        assert np.array_equal(again, want)
        plain, pst, _ = gp.fq2b(f1, f2, None)
        assert plain.shape[0] == wst["nRead"] and pst["nBad"] == 0


@pytest.mark.gpu
def test_gpu_fq2b_then_build_on_device(orc, gpu_lib):
    # FASTQ -> FQB -> index without the records leaving the device: the synthetic FQB of the build tests, written out
    # as FASTQ text, must give the index the oracle builds from the FQB itself
    import hash10x_b200
    import fqbtools
    import hashfile
    p = orc.synth_params(seed=51, n_barcodes=25, pairs_min=5, pairs_max=120, read_len=151)
    recs = orc.synth_fqb(p)
    f1, f2 = fqbtools.fastq_from_fqb(recs, 151)
    want = orc.build(recs, B=20)
    with hash10x_b200.Hash10xGPU(B=20) as gp:
        got, st, dptr = gp.fq2b(f1, f2, None, sort=False)
        assert np.array_equal(got, recs)
        gp.build_device(dptr, st["nRecords"])
        ix = gp.download()
    hashfile.assert_strict_equal(hashfile.from_index(want), hashfile.from_index(ix), table=True)


@pytest.mark.gpu
def test_gpu_fq2b_errors_are_the_references(gpu_lib):
    import hash10x_b200
    f1, f2 = fo.synth_fastq(7, 6, 40, 40)
    lines = f1.split(b"\n")
    cases = []
    bad = list(lines); bad[8] = b"read2"; cases.append((b"\n".join(bad), f2))                 # entry 3 of file 1 -> 5th call
    bad = list(lines); bad[5] = bad[5][:-1]; cases.append((b"\n".join(bad), f2))              # short sequence line
    bad = list(lines); bad[6] = b"+x"; cases.append((b"\n".join(bad), f2))
    bad = list(lines); bad[7] = bad[7] + b"I"; cases.append((b"\n".join(bad), f2))            # long quality line
    cases.append((f1, b"\n".join(f2.split(b"\n")[:8]) + b"\n"))                               # file 2 ends early
    with hash10x_b200.Hash10xGPU(B=20) as gp:
        for a, b in cases:
            with pytest.raises(fo.FastqError) as want:
                fo.fq2b(a, b, None)
            with pytest.raises(hash10x_b200.H10xError) as got:
                gp.fq2b(a, b, None)
            assert got.value.code == 5 and got.value.msg == want.value.text, (got.value.msg, want.value.text)


@pytest.mark.gpu
def test_fq2b_cli_matches_reference_binary(gpu_lib, tmp_path):
    # hash10x_b200/bin/fq2b-b200 against oracle/_ref/fq2b (the unmodified reference, when it travelled with the snapshot;
    # the golden CRC otherwise): the same .fqb bytes and the same report on stderr; -sort adds the bsort step
    import subprocess
    exe = os.path.join(ROOT, "hash10x_b200", "bin", "fq2b-b200")
    ref = os.path.join(ROOT, "oracle", "_ref", "fq2b")
    g = GOLD["cases"]["pairs151_wl"]
    c = g["params"]
    wl = fo.synth_whitelist(*c["wl"])
    f1, f2 = fo.synth_fastq(c["seed"], c["n"], c["l1"], c["l2"], wl)
    p1, p2, pw = tmp_path / "r1.fq", tmp_path / "r2.fq", tmp_path / "wl.txt"
    p1.write_bytes(f1)
    p2.write_bytes(f2)
    pw.write_text("".join(s + "\n" for s in wl))
    r = subprocess.run([exe, "-10x", str(pw), "-o", str(tmp_path / "a.fqb"), str(p1), str(p2)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    data = (tmp_path / "a.fqb").read_bytes()
    assert "%08x" % zlib.crc32(data) == g["crc32"]
    assert r.stderr.splitlines()[-4:] == g["stderr"]
    if os.path.exists(ref):
        rr = subprocess.run([ref, "-10x", str(pw), "-o", str(tmp_path / "b.fqb"), str(p1), str(p2)], capture_output=True, text=True)
        assert rr.returncode == 0 and (tmp_path / "b.fqb").read_bytes() == data and rr.stderr == r.stderr
    rs = subprocess.run([exe, "-10x", str(pw), "-sort", "-o", str(tmp_path / "s.fqb"), str(p1), str(p2)], capture_output=True, text=True)
    assert rs.returncode == 0, rs.stderr
    srt = np.frombuffer((tmp_path / "s.fqb").read_bytes(), dtype="<u4").reshape(-1, 30)
    assert np.array_equal(srt, fo.bsort(np.frombuffer(data, dtype="<u4").reshape(-1, 30)))
    # a malformed entry dies with the reference's text
    p1.write_bytes(f1.replace(b"\n+\n", b"\n-\n", 1))
    rb = subprocess.run([exe, "-o", str(tmp_path / "c.fqb"), str(p1), str(p2)], capture_output=True, text=True)
    assert rb.returncode != 0 and "FATAL ERROR: bad + fastq line entry 1" in rb.stderr
