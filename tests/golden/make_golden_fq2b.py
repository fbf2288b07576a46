#!/usr/bin/env python
"""Golden vectors for the fq2b stage, from the UNMODIFIED reference binary oracle/_ref/fq2b (built by oracle/Makefile
from /root/reference/fq2b.c): for a few seeded synthetic FASTQ pairs and whitelists (oracle/fq2b_oracle.py generates
them), the CRC32 of the .fqb it writes and the counters it prints.  Also checks the oracle restatement against it.
Run in the authoring container (needs oracle/_ref/fq2b); writes tests/golden/golden_fq2b.json."""
import json
import os
import re
import subprocess
import sys
import tempfile
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fq2b_oracle as fo  # noqa: E402

CASES = [dict(name="pairs151_wl", seed=1, n=300, l1=151, l2=151, wl=(11, 40)),
         dict(name="pairs151_nowl", seed=2, n=200, l1=151, l2=151, wl=None),
         dict(name="len16_32", seed=3, n=120, l1=16, l2=32, wl=(12, 10)),
         dict(name="len33_160_lower", seed=4, n=150, l1=33, l2=160, wl=(13, 25), lower=True),
         dict(name="single_100", seed=5, n=90, l1=100, l2=None, wl=(14, 12))]


def main():
    ref = os.path.join(ROOT, "oracle", "_ref", "fq2b")
    out = {"made_by": "oracle/_ref/fq2b (unmodified reference fq2b.c, gcc -O3) on the texts of oracle/fq2b_oracle.synth_fastq",
           "cases": {}}
    with tempfile.TemporaryDirectory() as td:
        for c in CASES:
            wl = fo.synth_whitelist(*c["wl"]) if c["wl"] else None
            f1, f2 = fo.synth_fastq(c["seed"], c["n"], c["l1"], c["l2"] or 8, wl, lower=c.get("lower", False))
            p1, p2, pw, po = (os.path.join(td, x) for x in ("r1.fq", "r2.fq", "wl.txt", "out.fqb"))
            open(p1, "wb").write(f1)
            open(p2, "wb").write(f2)
            cmd = [ref]
            if wl:
                open(pw, "w").write("".join(s + "\n" for s in wl))
                cmd += ["-10x", pw]
            cmd += ["-o", po, p1] + ([p2] if c["l2"] else [])
            r = subprocess.run(cmd, capture_output=True, text=True, check=True)
            data = open(po, "rb").read()
            m = re.search(r"(\d+) \(.*\) not matching barcodes were dropped\n(\d+) \(.*\) of those that matched were error corrected\n"
                          r"by base position:((?: \d+)+)", r.stderr)
            g = dict(crc32="%08x" % zlib.crc32(data), nbytes=len(data), stderr=r.stderr.splitlines()[-4:] if wl else r.stderr.splitlines()[-1:])
            if m:
                g.update(nBad=int(m.group(1)), nFixed=int(m.group(2)), nFixBase=[int(x) for x in m.group(3).split()])
            # the restatement must reproduce the reference byte for byte
            recs, st = fo.fq2b(f1, f2 if c["l2"] else None, [fo.pack_barcode(s) for s in wl] if wl else None)
            assert recs.astype("<u4").tobytes() == data, c["name"]
            if m:
                assert (st["nBad"], st["nFixed"], st["nFixBase"]) == (g["nBad"], g["nFixed"], g["nFixBase"]), c["name"]
            out["cases"][c["name"]] = dict(params=c, **g)
            print(c["name"], g["crc32"], g["nbytes"], g.get("nBad"), g.get("nFixed"))
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "golden_fq2b.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
