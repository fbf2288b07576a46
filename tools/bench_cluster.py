#!/usr/bin/env python
"""Timing of the "next" rows f1 + f2 (--hashDepthRange, --cluster) on one B200 next to the reference's own CPU
implementation (oracle/_ref/hash10x_omp, the reference's only OpenMP region is this loop, hash10x.c:1247).

  python tools/bench_cluster.py [--workload yeast] [--pairs N] [--dmin 10 --dmax 100] [--ct 5] [--no-cpu]

Not the driver's bench (bench.py measures the --readFQB build); prints one JSON line.  The CPU leg runs the
reference twice on the same FQB file (with and without --cluster) and reports the wall-clock difference."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                  # workloads and the synthetic generator parameters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="yeast", choices=sorted(bench.WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0)
    ap.add_argument("--dmin", type=int, default=10)
    ap.add_argument("--dmax", type=int, default=100)
    ap.add_argument("--ct", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--check", action="store_true", help="compare the GPU result with the reference's .hash")
    args = ap.parse_args()
    import numpy as np
    import torch
    import hash10x_b200
    from hash10x_b200 import synth as gsynth
    wl = bench.WORKLOADS[args.workload]
    p = bench.synth_params(gsynth, wl, seed=3)
    if args.pairs:
        p.nBarcodes = max(2, int(args.pairs / ((wl["pairs_min"] + wl["pairs_max"]) / 2)))
    n_rec, off = gsynth.layout(p)
    dev = torch.device("cuda", 0)
    fqb = torch.empty(n_rec * 30, dtype=torch.int32, device=dev)
    gsynth.fill_device(p, off, 0, n_rec, fqb.data_ptr())
    torch.cuda.synchronize()
    g = hash10x_b200.Hash10xGPU(B=wl["B"])
    g.build_device(fqb.data_ptr(), n_rec, torch.cuda.current_stream().cuda_stream)
    st = g.stats()
    t0 = time.perf_counter()
    _w, goff, good = g.depth_range(args.dmin, args.dmax)
    t1 = time.perf_counter()
    clus, nsub, ptm, ms_kernel = g.cluster(0, 0, args.ct)
    t2 = time.perf_counter()
    line = {"metric": "--hashDepthRange + --cluster on the resident index", "workload": wl["desc"], "pairs": int(n_rec),
            "blocks": int(st["nBlocks"]), "bins": int(st["nBins"]), "hashes": int(st["nHashes"]),
            "build_ms": st["msTotal"], "depth_range": [args.dmin, args.dmax], "clusterThreshold": args.ct,
            "good_hashes": int(good.size), "hashDepthRange_ms_incl_d2h": (t1 - t0) * 1e3,
            "cluster_kernel_ms": ms_kernel, "cluster_ms_incl_d2h": (t2 - t1) * 1e3,
            "clustered_blocks": int((nsub > 0).sum()), "sub_clusters": int(nsub.sum()),
            "clustered_entries": int((((clus >> np.uint64(48)) & np.uint64(255)) > 0).sum())}
    if not args.no_cpu:
        from oracle import orc
        exe = orc.ref_binary("hash10x_omp") or orc.ref_binary("hash10x")
        if exe:
            shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
            with tempfile.TemporaryDirectory(dir=shm) as d:
                path, out = os.path.join(d, "a.fqb"), os.path.join(d, "r.hash")
                host = torch.empty(n_rec * 30, dtype=torch.int32)
                host.copy_(fqb)
                host.numpy().tofile(path)
                base = [exe, "-B", str(wl["B"]), "--readFQB", path, "--hashDepthRange", str(args.dmin), str(args.dmax)]
                ta = time.perf_counter()
                r = subprocess.run(base, capture_output=True, text=True)
                tb = time.perf_counter()
                r2 = subprocess.run(base + ["-ct", str(args.ct), "--cluster", "0", "0"] +
                                    (["--writeHash", out] if args.check else []), capture_output=True, text=True)
                tc = time.perf_counter()
                if r.returncode or r2.returncode:
                    line["cpu"] = {"error": (r.stderr + r2.stderr)[-300:]}
                else:
                    line["cpu"] = {"binary": os.path.basename(exe), "threads": os.cpu_count(),
                                   "readFQB_plus_depthRange_s": tb - ta,
                                   "cluster_s": (tc - tb) - (tb - ta) - 0.0,
                                   "note": "wall clock of the run with --cluster minus the run without"
                                           + (" (includes --writeHash)" if args.check else "")}
                    if args.check:
                        sys.path.insert(0, os.path.join(ROOT, "tests"))
                        import hashfile
                        hf = hashfile.parse(out, keep_table=False)
                        # the reference built its own ClusterHash lists with uninitialised subCluster bytes
                        # (hash10x.c:175): compare what --cluster defines, i.e. blocks with good hashes
                        same_n = bool(np.array_equal(hf.blkNSub, nsub))
                        line["cpu"]["nSubCluster_equal"] = same_n
                        line["cpu"]["pointToMin_equal"] = bool(np.array_equal(hf.blkPointToMin.view(np.uint64),
                                                                             ptm.view(np.uint64)))
    print(json.dumps(line))
    g.close()


if __name__ == "__main__":
    main()
