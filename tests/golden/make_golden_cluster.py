#!/usr/bin/env python
"""Golden fixtures for `--hashDepthRange` + `--cluster` (hash10x.c:528-539,738-868), generated from the REFERENCE ITSELF
(oracle/_ref/hash10x).  Run in the authoring container only:

    python tests/golden/make_golden_cluster.py

Per case a small synthetic FQB is committed and the reference binary is run twice:
  A  `--readFQB x.fqb --hashDepthRange a b -ct t --cluster c0 c1 --writeHash` - the real pipeline.  nSubCluster and
     pointToMin of every block are well defined; the subCluster bytes of entries outside the good lists are the
     reference's uninitialised malloc bytes (hash10x.c:175), so only the bytes of clustered blocks' good entries count;
  B  `--readFQB x.fqb --writeHash`, that file with bytes 6-7 of every ClusterHash zeroed, then `--readHash ...
     --hashDepthRange ... --cluster ... --writeHash`: every subCluster byte is defined.
A and B must agree on nSubCluster / pointToMin; golden_cluster.json stores CRC32s of B's nSubCluster, pointToMin (as
IEEE doubles) and subCluster bytes.  The GPU box, which has no /root/reference, checks the CUDA path against these."""
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

import hashfile  # noqa: E402
from oracle import orc  # noqa: E402

CASES = {
    # name: synthetic parameters, depth range, clusterThreshold, code range
    "cluster_small": (dict(seed=37, n_barcodes=60, pairs_min=20, pairs_max=60, genome_len=20_000, mol_len=6_000,
                           mol_per_barcode=3), 2, 13, 1, 0, 0),
    "cluster_deep": (dict(seed=41, n_barcodes=60, pairs_min=10, pairs_max=40, genome_len=6_000, mol_len=3_000,
                          mol_per_barcode=2), 3, 200, 3, 4, 50),
}


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def main():
    exe = orc.ref_binary()
    if exe is None:
        sys.exit("oracle/_ref/hash10x missing: run `make -C oracle` where /root/reference exists")
    out = {}
    for name, (sp, dmin, dmax, ct, c0, c1) in CASES.items():
        recs = np.ascontiguousarray(orc.synth_fqb(orc.synth_params(**sp)), dtype=np.uint32)
        fqb = os.path.join(HERE, name + ".fqb")
        recs.tofile(fqb)
        tail = ["--hashDepthRange", str(dmin), str(dmax), "-ct", str(ct), "--cluster", str(c0), str(c1)]
        a_hash, plain, zeroed, b_hash = ("/tmp/%s_%s.hash" % (name, s) for s in "apzb")
        r = subprocess.run([exe, "-B", "20", "--readFQB", fqb] + tail + ["--writeHash", a_hash], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        r = subprocess.run([exe, "-B", "20", "--readFQB", fqb, "--writeHash", plain], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        hf = hashfile.parse(plain)
        raw = bytearray(open(plain, "rb").read())
        base = len(raw) - 8 * hf.nHashes
        for e in range(hf.nHashes):
            raw[base + 8 * e + 6] = 0
            raw[base + 8 * e + 7] = 0
        open(zeroed, "wb").write(raw)
        r = subprocess.run([exe, "-B", "20", "--readHash", zeroed] + tail + ["--writeHash", b_hash], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        A, B = hashfile.parse(a_hash), hashfile.parse(b_hash)
        assert np.array_equal(A.blkNSub, B.blkNSub)
        assert np.array_equal(A.blkPointToMin.view(np.uint64), B.blkPointToMin.view(np.uint64))
        assert int(B.blkNSub.sum()) > 0
        out[name] = {"params": dict(B=20), "depth_range": [dmin, dmax], "clusterThreshold": ct, "codes": [c0, c1],
                     "records": int(recs.shape[0]), "nBlocksMax": int(B.nBlocksMax), "nHashes": int(B.nHashes),
                     "sub_clusters": int(B.blkNSub.sum()), "clustered_entries": int((B.clusSub > 0).sum()),
                     "crc_blkNSub": crc(B.blkNSub), "crc_pointToMin": crc(B.blkPointToMin),
                     "crc_clusSub": crc(B.clusSub), "crc_clusRaw": crc(B.clusRaw & np.uint64(0x00FFFFFFFFFFFFFF)),
                     "blkNSub_head": [int(x) for x in B.blkNSub[:10]],
                     "pointToMin_head": [float(x) for x in B.blkPointToMin[:6]]}
        for p in (a_hash, plain, zeroed, b_hash):
            os.remove(p)
    with open(os.path.join(HERE, "golden_cluster.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote %d cases" % len(out), {k: (v["records"], v["sub_clusters"]) for k, v in out.items()})


if __name__ == "__main__":
    main()
