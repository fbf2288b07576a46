#!/usr/bin/env python
"""Golden digests at the sizes bench.py TIMES, produced by the REFERENCE ITSELF (oracle/_ref/hash10x, the
unmodified reference compiled by oracle/Makefile).  Run in the authoring container only (needs /root/reference's
binary, ~60 GB of scratch disk and, for the 1 Gb set, ~25 minutes of one core):

    python tests/golden/make_golden_scale.py [--scratch DIR] name[:xN] ...

`name` is a bench.py workload (yeast, gb10th, 1gb, human8); `:xN` multiplies the barcode count by N exactly as
`bench.py --gpus N` does (weak scaling: every rank gets one workload's worth of barcodes of ONE data set).
For each case: oracle/scale_tool gen writes the FQB (the same closed-form records the device generator makes),
`hash10x -B b --readFQB f --writeHash h` builds the index on one CPU core, oracle/scale_tool digest streams the
`.hash` and the result - counters plus the position-salted sum digests of hash10x_b200/csrc/h10x_digest.h for
hashIndex, hashValue, hashDepth, blkNRead, blkNHash and the ClusterHash stream - is merged into
tests/golden/golden_scale.json next to the generator parameters.  bench.py and tests/test_gpu_scale.py compare
h10x_gpu_index_digest of the index they built against these.
"""
import argparse
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import bench  # noqa: E402  (workload table)
from oracle import orc  # noqa: E402

OUT = os.path.join(HERE, "golden_scale.json")


def case_key(name, mult):
    return name if mult == 1 else "%sx%d" % (name, mult)


def run_case(name, mult, scratch, seed=3, fifo=False):
    wl = bench.WORKLOADS[name]
    p = bench.synth_params(orc, wl, seed=seed)
    p.nBarcodes *= mult
    tool = os.path.join(ROOT, "oracle", "scale_tool")
    ref = orc.ref_binary("hash10x")
    assert ref, "oracle/_ref/hash10x is missing (make -C oracle)"
    fqb = os.path.join(scratch, "scale_%s.fqb" % case_key(name, mult))
    hsh = os.path.join(scratch, "scale_%s.hash" % case_key(name, mult))
    gen_cmd = [tool, "gen", fqb] + [str(int(x)) for x in (
        p.seed, p.genomeLen, p.nBarcodes, p.pairsMin, p.pairsMax, p.molPerBarcode, p.molLen, p.snpPeriod,
        p.errThresh, p.readLen)]
    ref_cmd = [ref, "-B", str(wl["B"]), "--readFQB", fqb, "--writeHash", hsh]
    t0 = time.time()
    if fifo:
        # the FQB never touches the disk (72 GB at human8:x8): the generator writes into a named pipe the reference reads
        # (both are strictly sequential); reference_wall_s then includes waiting for the generator
        if os.path.exists(fqb):
            os.unlink(fqb)
        os.mkfifo(fqb)
        genp = subprocess.Popen(gen_cmd, stdout=subprocess.PIPE, text=True)
        r = subprocess.run(ref_cmd, capture_output=True, text=True)
        gen_out = genp.communicate()[0]
        t_gen = t_ref = time.time() - t0
        if genp.returncode != 0:
            raise RuntimeError("generator failed on %s" % case_key(name, mult))
        gen = type("G", (), {"stdout": gen_out})()
    else:
        gen = subprocess.run(gen_cmd, check=True, capture_output=True, text=True)
        t_gen = time.time() - t0
        t0 = time.time()
        r = subprocess.run(ref_cmd, capture_output=True, text=True)
        t_ref = time.time() - t0
    os.unlink(fqb)
    if r.returncode != 0:
        raise RuntimeError("reference failed on %s: %s" % (case_key(name, mult), (r.stderr or r.stdout)[-400:]))
    d = json.loads(subprocess.run([tool, "digest", hsh], check=True, capture_output=True, text=True).stdout)
    os.unlink(hsh)
    d.update({"workload": name, "barcode_multiplier": mult, "seed": seed, "records": json.loads(gen.stdout)["records"],
              "generator": {"genome_len": int(p.genomeLen), "n_barcodes": int(p.nBarcodes), "pairs_min": int(p.pairsMin),
                            "pairs_max": int(p.pairsMax), "mol_per_barcode": int(p.molPerBarcode), "mol_len": int(p.molLen),
                            "snp_period": int(p.snpPeriod), "err_thresh": int(p.errThresh), "read_len": int(p.readLen)},
              "made_by": "oracle/_ref/hash10x -B %d --readFQB --writeHash (unmodified reference, gcc -O3, 1 core)" % wl["B"],
              "reference_wall_s": round(t_ref, 1), "generate_wall_s": round(t_gen, 1)})
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cases", nargs="+")
    ap.add_argument("--scratch", default="/tmp")
    ap.add_argument("--fifo", action="store_true", help="pipe the FQB from the generator into the reference (no FQB on disk)")
    a = ap.parse_args()
    for c in a.cases:
        name, _, m = c.partition(":x")
        mult = int(m) if m else 1
        d = run_case(name, mult, a.scratch, fifo=a.fifo)
        import fcntl
        with open(OUT + ".lock", "w") as lk:          # several generator runs may finish at the same time
            fcntl.flock(lk, fcntl.LOCK_EX)
            allg = {}
            if os.path.exists(OUT):
                with open(OUT) as f:
                    allg = json.load(f)
            allg[case_key(name, mult)] = d
            with open(OUT, "w") as f:
                json.dump(allg, f, indent=1, sort_keys=True)
                f.write("\n")
        print(case_key(name, mult), json.dumps(d))


if __name__ == "__main__":
    main()
