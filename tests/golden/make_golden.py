#!/usr/bin/env python
"""Generate the golden fixtures from the REFERENCE ITSELF (oracle/_ref/hash10x, compiled unmodified
from /root/reference by oracle/Makefile).  Run in the authoring container only:

    python tests/golden/make_golden.py

For each case a small FQB is written (synthetic generator or hand-built quirk records), the reference
binary builds and writes the .hash, and a compact digest of that file is stored in golden.json:
exact counters, CRC32s of hashValue / hashDepth / hashIndex / block table / ClusterHash (idx, read)
streams, and the first values.  The FQB inputs are committed too (they are tiny) so the GPU box,
which has no /root/reference, checks the same bytes.
"""
import json
import os
import subprocess
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fqbtools  # noqa: E402
import hashfile  # noqa: E402
from oracle import orc  # noqa: E402


def digest(hf):
    def crc(a):
        return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF
    return {"B": int(hf.B), "hashNumber": int(hf.hashNumber), "nBlocksMax": int(hf.nBlocksMax),
            "nHashes": int(hf.nHashes), "fileSize": int(hf.size), "depthDim": int(hf.depthDim),
            "blkDim": int(hf.blkDim),
            "crc_hashValue": crc(hf.hashValue), "crc_hashDepth": crc(hf.hashDepth),
            "crc_hashIndex": crc(hf.hashIndex), "crc_blkNRead": crc(hf.blkNRead),
            "crc_blkNHash": crc(hf.blkNHash), "crc_clusIdx": crc(hf.clusIdx), "crc_clusRead": crc(hf.clusRead),
            "hashValue_head": [int(x) for x in hf.hashValue[:6]],
            "hashDepth_head": [int(x) for x in hf.hashDepth[:6]],
            "blkNHash_head": [int(x) for x in hf.blkNHash[:8]]}


def cases():
    rng = np.random.default_rng(2024)
    p = orc.synth_params(seed=101, n_barcodes=24, pairs_min=3, pairs_max=120)
    yield "synth_small", orc.synth_fqb(p), dict(B=20)
    p = orc.synth_params(seed=102, n_barcodes=12, pairs_min=40, pairs_max=90, read_len=160)
    yield "synth_160bp_k17_w13_r5", orc.synth_fqb(p), dict(B=20, k=17, w=13, r=5)
    quirks = np.concatenate([
        fqbtools.random_records(rng, [11], [5]),
        fqbtools.const_records(12, 4, 1, 1),      # no moshes: phantom {hash 0, read 0}
        fqbtools.random_records(rng, [13], [7]),
        fqbtools.const_records(14, 3, 0, 0),      # poly-A: hash 0 moshes
        fqbtools.const_records(15, 2, 1, 1),
        fqbtools.random_records(rng, [16], [6]),  # last run: never hashed
    ])
    yield "quirks", quirks, dict(B=20)
    yield "quirks_N17", quirks, dict(B=20, N=17)
    zero = fqbtools.random_records(rng, [7, 0, 9, 0, 5, 3], [6, 4, 5, 3, 4, 2])
    yield "allA_barcode_c10", zero, dict(B=20, chunk=10)   # barcode 0 run ends on a chunk boundary
    yield "allA_barcode_c11", zero, dict(B=20, chunk=11)
    yield "single_run", fqbtools.random_records(rng, [5], [9]), dict(B=20)


def main():
    if orc.ref_binary() is None:
        sys.exit("oracle/_ref/hash10x missing: run `make -C oracle` where /root/reference exists")
    out = {}
    for name, recs, kw in cases():
        fqb = os.path.join(HERE, name + ".fqb")
        recs = np.ascontiguousarray(recs, dtype=np.uint32)
        recs.tofile(fqb)
        hp = os.path.join("/tmp", name + ".hash")
        r = orc.run_reference(fqb, hp, B=kw.get("B", 20), k=kw.get("k"), w=kw.get("w"), r=kw.get("r"),
                              N=kw.get("N"), chunk=kw.get("chunk"))
        assert r.returncode == 0, r.stderr
        hf = hashfile.parse(hp)
        d = digest(hf)
        d["params"] = kw
        d["records"] = int(recs.shape[0])
        d["stdout_created"] = [ln.strip() for ln in r.stdout.splitlines() if "created" in ln or "filled" in ln][:2]
        out[name] = d
        os.remove(hp)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote %d cases" % len(out))


if __name__ == "__main__":
    main()
