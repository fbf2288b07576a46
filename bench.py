#!/usr/bin/env python
"""bench.py - read pairs/s through the `--readFQB` minhash + index build, and % of the HBM roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 1gb|yeast|human8|custom] [--impl reference]

One "step" = one complete index build (readFQB + fillHashTable, hash10x.c:1200-1205) over the
synthetic FQB of the workload.  `value` times the device-resident build (FQB already in HBM, index
left in HBM) with CUDA events; `e2e` times the same build through the C ABI's host entry point
(h10x_gpu_build_host: pinned host FQB -> H2D -> build -> D2H of every index array).  The default
N=1 workload is BASELINE.json configs[2] ("1 Gb diploid genome at 60x, ~200M read pairs, 24 GB
FQB, -B 28, 1xB200"), the largest single-GPU configuration; inputs are far larger than the 126 MB L2.

After the timed region the index that was just built is CHECKED: h10x_gpu_index_digest (device-side
position-salted sums of every array writeHashFile stores) against tests/golden/golden_scale.json, which holds the
same digests of the `.hash` the UNMODIFIED reference binary wrote for the same data set ("parity" in the line;
at N > 1 the per-rank digests add up to the digest of the stitched arrays).  At N=1 the line also carries
`weak_base` (the same build on one GPU's share of the N>1 workload, so that the 1->8 curve can be read on one
workload), and the rows that follow the build on the resident index: --hashDepthRange and --cluster.

`--impl reference` times the reference's own CPU implementation (oracle/_ref/hash10x compiled from
the unmodified reference; the oracle port if that binary is missing) on a bounded sample of the same
workload on this box's host cores.  --readFQB is single-threaded in the reference even with -DOMP.
"""
import argparse
import io
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json's metric, verbatim, in BOTH arms (the driver divides one arm's value by the other's only when
# metric, unit and direction agree)
METRIC = "read pairs/sec through --readFQB minhash+index build; % of HBM roofline"
UNIT = "read pairs/s"

WORKLOADS = {
    # name: genome, barcodes, pairs/barcode range, molecules, mol len, snp period, err rate, B
    "1gb": dict(genome_len=1_000_000_000, n_barcodes=500_000, pairs_min=300, pairs_max=500,
                mol_per_barcode=10, mol_len=50_000, snp_period=1000, err_rate=0.0005, B=28, read_len=160,
                desc="BASELINE configs[2]: synthetic 1 Gb diploid genome at 60x, ~200M read pairs, 24 GB FQB, -B 28 "
                     "(160+160 bp reads: with 151 bp the zero padding of the last packed word adds ~0.39 novel hashes "
                     "per pair and the reference itself dies with 'hashTableSize is too small' at -B 28)"),
    "yeast": dict(genome_len=12_000_000, n_barcodes=10_000, pairs_min=150, pairs_max=350,
                  mol_per_barcode=10, mol_len=50_000, snp_period=500, err_rate=0.004, B=24, read_len=151,
                  depth_range=(10, 100),
                  desc="BASELINE configs[1]: synthetic yeast-scale diploid (12 Mb, ~2.5M read pairs, 10k barcodes, 60x), -B 24"),
    "gb10th": dict(genome_len=100_000_000, n_barcodes=50_000, pairs_min=300, pairs_max=500,
                   mol_per_barcode=10, mol_len=50_000, snp_period=1000, err_rate=0.0005, B=26, read_len=160,
                   desc="one tenth of BASELINE configs[2] at the same barcode density (100 Mb genome, 50k barcodes, ~20M read "
                        "pairs, -B 26): the profiling-sized stand-in for the 1 Gb sharing structure"),
    "human8": dict(genome_len=3_100_000_000, n_barcodes=187_500, pairs_min=300, pairs_max=500,
                   mol_per_barcode=10, mol_len=50_000, snp_period=1000, err_rate=0.0005, B=30, read_len=160,
                   desc="one eighth of BASELINE configs[3]: 3.1 Gb human-scale genome, 75M read pairs per GPU, -B 30, 160+160 bp reads"),
}


def synth_params(mod, wl, seed=3):
    """mod: hash10x_b200.synth (GPU arm) or oracle.orc (CPU legs) - the same parameter struct"""
    make = getattr(mod, "make_params", None) or mod.synth_params
    return make(seed=seed, genome_len=wl["genome_len"], n_barcodes=wl["n_barcodes"],
                            pairs_min=wl["pairs_min"], pairs_max=wl["pairs_max"],
                            mol_per_barcode=wl["mol_per_barcode"], mol_len=wl["mol_len"],
                            snp_period=wl["snp_period"], err_rate=wl["err_rate"], read_len=wl.get("read_len", 151))


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md "clocks" line), read
    in-process through NVML: spawning nvidia-smi every 200 ms takes driver locks that stall
    cudaMallocAsync / stream synchronisation in the process being measured."""

    def __init__(self, gpu=0, period=0.1, enabled=True):
        super().__init__(daemon=True)
        self.gpu, self.period, self.rows, self.stop_flag = gpu, period, [], threading.Event()
        self.nv = self.h = None
        self.max_sm = None
        if not enabled:
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[gpu]) if vis and vis.split(",")[gpu].isdigit() else gpu
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, rs, time.perf_counter()))
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def window(self):
        """start of the timed region: the thread is already running (its first NVML calls are slow
        and take driver locks, so they must not land inside the timed steps)"""
        self.t_start = time.perf_counter()

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=3)
        t0 = getattr(self, "t_start", 0.0)
        self.rows = [r for r in self.rows if r[2] >= t0]
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "note": "NVML unavailable"}
        nv = self.nv
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted(nm for nm, b in bits.items() if any(r[1] & b for r in self.rows))
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_sm, "reasons": reasons,
                "samples": len(self.rows)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------- CPU legs

def cpu_sample(orc, wl, sample_pairs):
    """The first barcodes of the workload (whole barcode runs, ~sample_pairs pairs) generated on the host."""
    import numpy as np
    p = synth_params(orc, wl)
    n, off = orc.synth_layout(p)
    nb = int(np.searchsorted(off, sample_pairs, side="left"))
    nb = max(2, min(nb, p.nBarcodes))
    r1 = int(off[nb])
    return orc.synth_fqb(p, 0, r1), nb, n


def _run(cmd):
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("reference failed: " + (r.stderr or r.stdout)[-300:])
    return dt


def time_reference(orc, recs, B):
    """Seconds for the reference's --readFQB step on recs; kind 'reference' when oracle/_ref exists."""
    exe = orc.ref_binary("hash10x")
    if exe is None:
        t = orc.time_build(recs, B=B)
        return t, "port", "oracle/h10x_oracle.c (CPU restatement; oracle/_ref not built)"
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.NamedTemporaryFile(suffix=".fqb", dir=shm) as f:
        recs.tofile(f)
        f.flush()
        dt = _run([exe, "-B", str(B), "--readFQB", f.name])
    return dt, "reference", "oracle/_ref/hash10x (unmodified reference, gcc -O3), wall clock of `-B %d --readFQB`" % B


def cluster_like_for_like(orc, gpu_factory, torch, gsynth, stream):
    """--hashDepthRange + --cluster on BASELINE configs[1] (the README's yeast case, where the reference quotes
    '5 minutes'): the whole data set on the GPU and on the host's cores with the reference's -DOMP build (its only
    OpenMP region is this loop, hash10x.c:1247).  The CPU commands run on a `.hash` the reference wrote, each timed
    as the difference of two command chains."""
    wl = WORKLOADS["yeast"]
    dmin, dmax, ct = wl["depth_range"] + (5,)
    p = synth_params(gsynth, wl, seed=3)
    n, off = gsynth.layout(p)
    fq = torch.empty(n * 30, dtype=torch.int32, device="cuda:%d" % torch.cuda.current_device())
    gsynth.fill_device(p, off, 0, n, fq.data_ptr())
    torch.cuda.synchronize()
    out = {"workload": wl["desc"], "pairs": int(n), "hashDepthRange": [dmin, dmax], "clusterThreshold": ct}
    with gpu_factory(wl["B"]) as g:
        g.build_device(fq.data_ptr(), n, stream.cuda_stream)
        t0 = time.perf_counter()
        n_good = g.depth_range_device(dmin, dmax)
        t1 = time.perf_counter()
        _c, nsub, _p, ms_kernel = g.cluster(0, 0, ct, copy=False)
        out.update({"build_ms": g.stats()["msTotal"], "good_hashes": int(n_good), "depth_range_ms": (t1 - t0) * 1e3,
                    "cluster_ms": ms_kernel, "sub_clusters": int(nsub.sum())})
    exe, omp = orc.ref_binary("hash10x"), orc.ref_binary("hash10x_omp")
    if exe and omp:
        shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
        with tempfile.TemporaryDirectory(dir=shm) as d:
            path, hsh = os.path.join(d, "y.fqb"), os.path.join(d, "y.hash")
            host = torch.empty(n * 30, dtype=torch.int32)
            host.copy_(fq)
            host.numpy().tofile(path)
            t_build = _run([exe, "-B", str(wl["B"]), "--readFQB", path, "--writeHash", hsh])
            base = [omp, "-B", str(wl["B"]), "--readHash", hsh]
            t_a = _run(base)
            t_b = _run(base + ["--hashDepthRange", str(dmin), str(dmax)])
            t_c = _run(base + ["--hashDepthRange", str(dmin), str(dmax), "-ct", str(ct), "--cluster", "0", "0"])
        out["cpu_baseline"] = {"binary": "oracle/_ref/hash10x_omp (-DOMP -fopenmp)", "threads": os.cpu_count(),
                               "readFQB_writeHash_s": t_build, "hashDepthRange_s": max(t_b - t_a, 0.0),
                               "cluster_s": max(t_c - t_b, 0.0),
                               "how": "wall clock of `--readHash ... <command>` minus the same chain without the command"}
        if out["cpu_baseline"]["cluster_s"] > 0:
            out["cluster_speedup_vs_omp"] = out["cpu_baseline"]["cluster_s"] * 1e3 / max(ms_kernel, 1e-6)
    del fq
    return out


def run_reference_arm(args, wl):
    from oracle import orc
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    recs, nb, n_total = cpu_sample(orc, wl, args.cpu_pairs)
    pairs = recs.shape[0]
    times = []
    kind = how = None
    for i in range(args.warmup + args.steps):
        dt, kind, how = time_reference(orc, recs, wl["B"])
        if i >= args.warmup:
            times.append(dt)
    dt = sum(times) / len(times)
    val = pairs / dt
    full_pairs = n_total * world
    sample = ("first %d barcode runs = %d read pairs (%.2f %% of the %d pairs of the configuration) per step; %s; "
              "at this rate the whole configuration takes %.0f s on one core"
              % (nb, pairs, 100.0 * pairs / full_pairs, full_pairs, how, full_pairs / val))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": wl["desc"], "B": wl["B"], "k": 21, "w": 31,
                       "sample_pairs": int(pairs), "config_pairs": int(full_pairs),
                       "extrapolated_config_seconds": full_pairs / val},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                             "host_cores": os.cpu_count(),
                             "note": "--readFQB is single-threaded in the reference, also with -DOMP (hash10x.c:1247 is the only omp pragma)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------- GPU arm

def load_golden(key):
    try:
        with open(os.path.join(ROOT, "tests", "golden", "golden_scale.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


DIGEST_ARRAYS = ("hashIndex", "hashValue", "hashDepth", "blkNRead", "blkNHash", "clusHash")


def parity_check(g, stats, dist, torch, dev, rank, world, golden_key, cut_by_pairs):
    """Digest of the index that was just timed against the reference binary's (tests/golden/golden_scale.json)."""
    import numpy as np
    mask = (1 << 64) - 1
    if world == 1:
        dg = g.digest()
        tot = {k: dg[k] for k in DIGEST_ARRAYS}
        counts = {"nHashes": stats["nHashes"], "hashNumber": stats["nBins"] + 1, "nBlocksMax": stats["nBlocks"] + 1,
                  "nReads": stats["nRecords"]}
        extra = {"codesMissing": dg["codesMissing"], "codesUnordered": dg["codesUnordered"], "haveCodes": dg["haveCodes"]}
    else:
        info = g.dist_info(download=False)
        mine = torch.tensor([stats["nHashes"]], dtype=torch.int64, device=dev)
        allh = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        entry_base = int(sum(int(x[0]) for x in allh[:rank]))
        dg = g.digest(block_base=info["blockBase"], entry_base=entry_base, with_block_zero=(rank == 0))
        vec = [dg[k] for k in DIGEST_ARRAYS]
        t = torch.tensor(np.array(vec, dtype=np.uint64).view(np.int64), dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)        # two's complement: the sum wraps like the digest's
        summed = t.cpu().numpy().view(np.uint64)
        tot = {k: int(summed[i]) & mask for i, k in enumerate(DIGEST_ARRAYS)}
        counts = {"nHashes": int(info["nHashesGlobal"]), "nBlocksMax": int(info["nBlocksGlobal"]) + 1,
                  "nReads": int(info["nReadsGlobal"]), "hashNumber": None}
        nb = torch.tensor([stats["nBins"] + 1 if rank == 0 else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(nb, op=dist.ReduceOp.MAX)
        counts["hashNumber"] = int(nb[0])
        extra = {}
    out = {"checked": False, "digest": "sum over i of mix(mix(pos_i) ^ value_i) mod 2^64 (hash10x_b200/csrc/h10x_digest.h) of "
           + ", ".join(DIGEST_ARRAYS) + "; computed on the device by h10x_gpu_index_digest after the timed region",
           "digest_of": {k: "%016x" % v for k, v in tot.items()}}
    out.update(extra)
    gold = None if cut_by_pairs else load_golden(golden_key)
    if gold is None:
        out["reason"] = ("workload cut by --pairs: no golden" if cut_by_pairs else
                         "no golden digests for '%s' in tests/golden/golden_scale.json (the reference needs more host memory "
                         "than the authoring container has for this size)" % golden_key)
        if extra:
            out["invariants_ok"] = bool(extra["haveCodes"] and extra["codesMissing"] == 0 and extra["codesUnordered"] == 0)
        return out
    bad = [k for k in DIGEST_ARRAYS if "%016x" % tot[k] != gold["dg_" + k]]
    bad += [k for k in ("nHashes", "hashNumber", "nBlocksMax", "nReads") if counts[k] != gold[k]]
    if extra and not (extra["haveCodes"] and extra["codesMissing"] == 0 and extra["codesUnordered"] == 0):
        bad.append("codes")
    out.update({"checked": True, "ok": not bad, "mismatch": bad, "golden": "tests/golden/golden_scale.json[%s]" % golden_key,
                "golden_made_by": gold.get("made_by"), "reference_wall_s": gold.get("reference_wall_s")})
    return out


def timed_builds(g, torch, stream, build, warmup, steps, barrier):
    """W untimed + K timed device-resident builds; CUDA events on the launching stream."""
    for _ in range(warmup):
        build()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms, launches, lib_ms = {}, 0, 0.0
    e0.record(stream)
    for _ in range(steps):
        build()
        s = g.stats()
        lib_ms += s["msTotal"]
        launches += s["kernelLaunches"]
        for k_, v in s["msStage"].items():
            stage_ms[k_] = stage_ms.get(k_, 0.0) + v
    e1.record(stream)
    barrier()
    return e0.elapsed_time(e1), stage_ms, launches, lib_ms


def run_ours(args, wl, wl_name):
    import torch
    import hash10x_b200
    from hash10x_b200 import synth as gsynth
    from hash10x_b200 import shard

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def make_input(w, mult, r, nranks):
        """ONE data set (same genome); each rank a contiguous barcode range of it, generated in HBM"""
        p = synth_params(gsynth, w, seed=3)
        if args.pairs:
            mean = (w["pairs_min"] + w["pairs_max"]) / 2
            p.nBarcodes = max(2, int(args.pairs / mean))
        p.nBarcodes *= mult                  # weak scaling: the per-GPU share stays fixed
        _n, off = gsynth.layout(p)
        cut = shard.plan_shards(off, nranks)
        r0, r1 = shard.shard_records(off, cut, r)
        t = torch.empty((r1 - r0) * 30, dtype=torch.int32, device=dev)
        gsynth.fill_device(p, off, r0, r1, t.data_ptr())
        torch.cuda.synchronize()
        return t, r1 - r0

    fqb, n_rec = make_input(wl, world, rank, world)
        # FLAG_LAZY_CODES: as hash10x-b200 creates its context - the hash->code lists stay resident for --cluster and are not
    # part of what h10x_gpu_build_host copies back (no host command of the --readFQB ... --writeHash chain reads them)
    g = hash10x_b200.Hash10xGPU(B=wl["B"], device=local, flags=hash10x_b200.FLAG_LAZY_CODES)
    if world > 1:
        g.dist_init(rank, world, shard.share_unique_id(dist, rank, hash10x_b200.Hash10xGPU.dist_unique_id))

    def build_resident():
        if world > 1:
            g.build_device_dist(fqb.data_ptr(), n_rec, stream.cuda_stream)
        else:
            g.build_device(fqb.data_ptr(), n_rec, stream.cuda_stream)

    # --- value: device-resident build, CUDA events on the launching stream ---
    sampler = ClockSampler(local, enabled=not args.no_clocks)
    sampler.start()
    for _ in range(args.warmup):
        build_resident()
    barrier()
    sampler.window()
    ms, stage_ms, launches, lib_ms = timed_builds(g, torch, stream, build_resident, 0, args.steps, barrier)
    clocks = sampler.summary()
    stats = g.stats()
    ms, total_pairs = shard.job_time_and_units(dist, torch, ms, n_rec, dev)
    ms_step = ms / args.steps
    value = total_pairs / (ms_step * 1e-3)

    # --- parity of what was just timed ---
    golden_key = wl_name if world == 1 else "%sx%d" % (wl_name, world)
    parity = parity_check(g, stats, dist, torch, dev, rank, world, golden_key, bool(args.pairs))

    # --- the rows after the build, on the index still resident (N=1): --hashDepthRange, --cluster ---
    nxt = None
    if world == 1 and not args.no_next:
        dmin, dmax, ct = wl.get("depth_range", (30, 100)) + (5,)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_good = g.depth_range_device(dmin, dmax)                 # the lists stay where --cluster reads them
        t1 = time.perf_counter()
        _c, nsub, _p, ms_kernel = g.cluster(0, 0, ct, copy=False)
        t2 = time.perf_counter()
        nxt = {"hashDepthRange": [dmin, dmax], "clusterThreshold": ct, "good_hashes": int(n_good),
               "depth_range_ms": (t1 - t0) * 1e3, "cluster_ms": ms_kernel, "cluster_ms_incl_d2h": (t2 - t1) * 1e3,
               "sub_clusters": int(nsub.sum()), "clustered_blocks": int((nsub > 0).sum()),
               "note": "h10x_gpu_depth_range / h10x_gpu_cluster (hash10x.c:528-539,738-766 / 770-868) on the index the timed "
                       "build left in HBM; depth_range_ms is h10x_gpu_depth_range_device (lists left on the device, host wall clock), "
                       "cluster_ms is the kernel, cluster_ms_incl_d2h the whole call with its 11 GB of ClusterHash coming back"}

    if world > 1 and not args.no_next:
        # the same rows on N GPUs: every rank first receives hashDepth and the whole hash->code CSR (one collective),
        # then filters and clusters its own barcode blocks; times are the slowest rank's
        dmin, dmax, ct = wl.get("depth_range", (30, 100)) + (5,)
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        g.dist_global_codes()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        n_good = g.depth_range_device(dmin, dmax)
        t2 = time.perf_counter()
        _c, nsub, _p, ms_kernel = g.cluster(0, 0, ct, copy=False)
        t3 = time.perf_counter()
        tmax = torch.tensor([t1 - t0, t2 - t1, t3 - t2, ms_kernel * 1e-3], dtype=torch.float64, device=dev)
        tot = torch.tensor([int(n_good), int(nsub.sum()), int((nsub > 0).sum())], dtype=torch.int64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        nxt = {"hashDepthRange": [dmin, dmax], "clusterThreshold": ct, "good_hashes": int(tot[0]),
               "global_codes_ms": float(tmax[0]) * 1e3, "depth_range_ms": float(tmax[1]) * 1e3,
               "cluster_ms": float(tmax[3]) * 1e3, "cluster_ms_incl_d2h": float(tmax[2]) * 1e3,
               "sub_clusters": int(tot[1]), "clustered_blocks": int(tot[2]),
               "note": "h10x_gpu_dist_global_codes (collective: hashDepth + the whole hash->code CSR on every rank), then "
                       "h10x_gpu_depth_range / h10x_gpu_cluster per rank on its own barcode blocks; slowest rank's times"}

    # --- e2e: the C ABI's host entry point, H2D + build + D2H inside the timed region ---
    e2e = None
    if not args.no_e2e:
        host_t = torch.empty(n_rec * 30, dtype=torch.int32, pin_memory=True)   # pinned host FQB
        host_t.copy_(fqb)                                                      # D2H, untimed
        torch.cuda.synchronize()
        host = host_t.data_ptr()
        del fqb
        torch.cuda.empty_cache()
        e2e_steps = max(1, min(args.steps, args.e2e_steps))

        def build_e2e():
            if world > 1:
                g.build_host_dist_ptr(host, n_rec)
            else:
                g.build_host_ptr(host, n_rec, want_index=False)
        build_e2e()                                               # warm-up (pinned arena allocation)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            build_e2e()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        td = torch.tensor([dt], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        dt = float(td[0]) / e2e_steps
        s = g.stats()
        hn, H, nbm = s["nBins"] + 1, s["nHashes"], s["nBlocks"] + 1
        if world == 1:
                        d2h = (4 << wl["B"]) + 8 * hn + 4 * hn + 4 * nbm + 4 * nbm + 8 * (nbm + 1) + 8 * H
        else:       # rank 0: table + values + depths + its blocks and ClusterHash lists (hash->code parts stay resident)
            d2h = (4 << wl["B"]) + 8 * hn + 4 * hn + 4 * nbm + 4 * nbm + 8 * (nbm + 1) + 8 * H
        e2e = {"value": total_pairs / dt, "unit": UNIT, "h2d_bytes_per_step": n_rec * 120,
               "d2h_bytes_per_step": int(d2h), "ms_per_step": dt * 1e3, "steps": e2e_steps,
               "bytes_are": "rank 0's, per step",
               "api": ("h10x_gpu_build_host" if world == 1 else "h10x_gpu_build_host_dist") +
                                            " (include/h10x_gpu.h): pinned host FQB -> index arrays in pinned host memory (hashIndex, hashValue, "
                      "hashDepth, block table, ClusterHash lists = everything writeHashFile, hash10x.c:244-267, writes; the "
                      "hash->code lists stay resident, H10X_FLAG_LAZY_CODES, as in hash10x-b200); the file goes up in slabs "
                      "and the fused kernel hashes the slabs that have landed"}
        del host_t
    else:
        del fqb
    torch.cuda.empty_cache()

    # --- weak_base: at N=1, the same build on one GPU's share of the N>1 workload ---
    weak_base = None
    if world == 1 and not args.no_weak_base and not args.pairs and wl_name != "human8":
        wb = WORKLOADS["human8"]
        g.close()
        fq2, n2 = make_input(wb, 1, 0, 1)
        g = hash10x_b200.Hash10xGPU(B=wb["B"], device=local)

        def build_wb():
            g.build_device(fq2.data_ptr(), n2, stream.cuda_stream)
        ms2, st2, _l2, _ = timed_builds(g, torch, stream, build_wb, 1, 2, barrier)
        s2 = g.stats()
        par2 = parity_check(g, s2, None, torch, dev, 0, 1, "human8", False)
        weak_base = {"workload": wb["desc"], "pairs": int(n2), "B": wb["B"], "ms_per_step": ms2 / 2,
                     "value": n2 / (ms2 / 2 * 1e-3), "unit": UNIT, "steps": 2, "warmup": 1,
                     "stage_ms": {k_: v / 2 for k_, v in st2.items() if v}, "parity": par2,
                     "note": "`bench.py --gpus N` (N > 1) runs N of these shares of one human-scale data set: divide the N-GPU "
                             "value by N times this one for the same-workload weak-scaling efficiency"}
        del fq2
        torch.cuda.empty_cache()

    if rank != 0:
        g.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    # --- roofline of the dominant kernel/stage (per-stage CUDA-event times from the library) ---
    peak, peak_src = measured_peaks()
    # the dominant KERNEL group: at N > 1 `binids` is the collective stage (exchange, owner merge, agreements: dozens of small
    # kernels and copies, reported in stage_ms), not a kernel to hold against a roofline
    dom = max(((k_, v) for k_, v in stage_ms.items() if world == 1 or k_ != "binids"), key=lambda kv: kv[1])
    R, M, H, D, nB = stats["nRecords"], stats["nMoshes"], stats["nHashes"], stats["nBins"], stats["nBlocks"]
    stage_bytes = {   # algorithmic bytes of each stage per build (DESIGN.md "kernels")
        "moshes": 120 * R + 12 * M, "fused": 120 * R + 8 * H, "blocksort": 24 * M, "dedup": 8 * H + 8 * H,
        "hashsort": 32 * H, "binids": 36 * D, "entryids": 8 * H, "codes": 20 * H + 12 * D,
        "clusters": 12 * H + 8 * H + 16 * H, "table": (4 << wl["B"]) + 12 * D, "runs": 4 * R * 3, "other": 0}
    dom_ms = dom[1] / args.steps
    dom_bytes = stage_bytes.get(dom[0], 0)
    ach = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic = alu = None
    try:        # DRAM bytes and ALU-pipe instructions of the fused kernel from the committed ncu captures, per read pair
        with open(os.path.join(ROOT, "profiles", "r02_fused_kernel.json")) as f:
            tj = json.load(f)
        if dom[0] == "fused":
            traffic = tj["dram_bytes_per_pair"] * R
            alu = tj
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom[0], "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                "traffic_source": "profiles/r02_fused_kernel.json (ncu dram__bytes_read+write of the two fused launches at this workload) x pairs" if traffic else None,
                "ms_per_launch_group": dom_ms, "share_of_step": dom_ms / ms_step}
    rooflines = [roofline]
    if alu is not None and alu.get("fmaheavy_busy_ns_per_pair"):
        # the fused kernel's real bound: the half-rate integer multiply ("fmaheavy") pipe.  busy = the pipe's active time per
        # read pair from the committed ncu capture (pct of cycles active x duration / pairs, whole GPU); it does not depend on
        # the run, so busy x pairs / live stage time is the fraction of that pipe's peak this run reached
        busy_ms = alu["fmaheavy_busy_ns_per_pair"] * R * 1e-6
        rooflines.append({"bound": "int_mul_pipe (fmaheavy)", "kernel": "fused", "achieved": busy_ms, "peak": dom_ms,
                          "unit": "ms of multiply-pipe work per ms of kernel", "frac": busy_ms / dom_ms,
                          "busy_ns_per_pair": alu["fmaheavy_busy_ns_per_pair"], "source": alu.get("pipe_capture"),
                          "note": "SURVEY 8d expected HBM to bind this stage; ncu shows the integer multiply pipe does (fmaheavy %s %%, "
                                  "ALU %s %%, issue %s %%, DRAM %s %%): its 120 R + 8 M bytes would take ~6 ms at the HBM peak"
                                  % (alu.get("fmaheavy_pct"), alu.get("alu_pct"), alu.get("issue_pct"), alu.get("dram_pct"))})
        roofline["note"] = "bound by the integer multiply pipe, not HBM: see roofline_int_pipe"
    pipe_ach = stats["algorithmicBytes"] / (ms_step * 1e-3) / 1e9
    pipeline = {"bound": "hbm", "algorithmic_bytes": stats["algorithmicBytes"],
                "bytes_per_pair": stats["algorithmicBytes"] / max(1, R), "achieved": pipe_ach, "peak": peak,
                "unit": "GB/s", "frac": pipe_ach / peak,
                "formula": "120 R + 12 H + 12 D + 4*2^B + 32 (nB+1)  (SURVEY.md 8d)"}

    # --- cpu baseline: the reference's own CPU path on a bounded sample, rank 0, N=1 only ---
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import orc          # the only use of oracle/ in this arm: the timed CPU baseline
        recs, nb, n_total = cpu_sample(orc, wl, args.cpu_pairs)
        dt, kind, how = time_reference(orc, recs, wl["B"])
        cpu = {"value": recs.shape[0] / dt, "unit": UNIT, "cores": 1, "kind": kind,
               "sample": "first %d barcode runs = %d read pairs (%.2f %%) of the same workload, %.1f s; %s; the whole "
                         "configuration would take %.0f s at this rate"
                         % (nb, recs.shape[0], 100.0 * recs.shape[0] / n_total, dt, how, n_total * dt / recs.shape[0]),
               "host_cores": os.cpu_count()}
        if nxt is not None:
            nxt["like_for_like"] = cluster_like_for_like(
                orc, lambda B: hash10x_b200.Hash10xGPU(B=B, device=local), torch, gsynth, stream)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": wl["desc"] + (" (cut to %d pairs per GPU by --pairs)" % n_rec if args.pairs else ""),
                       "pairs_per_gpu": n_rec, "barcodes_per_gpu": int(nB), "B": wl["B"], "k": 21, "w": 31,
                       "l2": "inputs (%.1f GB per GPU) are larger than the 126 MB L2; no flush needed" % (n_rec * 120 / 1e9),
                       "parallelism": "1 GPU" if world == 1 else
                       "%d ranks: barcode-range shards, every rank groups its own entries with the hand-written tail; the "
                       "rank-distinct (hash, depth, first block) triples go to hash-range owners (ranges cut at the quantiles of the "
                       "mosh density) by stores / copy engines into peer memory over NVLink (NCCL for the small collectives and as "
                       "fallback); owners merge the sorted runs and hand out the reference's bin ids; hashValue / hashDepth / "
                       "hashIndex on rank 0" % world},
            "roofline": roofline, "pipeline_roofline": pipeline, "cpu_baseline": cpu, "e2e": e2e, "parity": parity,
            "gpu_launches": int(launches), "clocks": clocks, "lib_ms_per_step": lib_ms / args.steps,
            "stage_ms": {k_: v / args.steps for k_, v in stage_ms.items() if v},
            "tail": "hand-written (h10x_tail.cuh)" if stats.get("tailPath") == 2 else "library radix sort (round-1 tail)",
            "counts": {"pairs": R, "moshes": M, "block_unique_hashes": H, "bins": D, "blocks": nB,
                       "peak_device_bytes": stats["peakDeviceBytes"]}}
    if len(rooflines) > 1:
        line["roofline_int_pipe"] = rooflines[1]
    if nxt:
        line["next_rows"] = nxt
    if weak_base:
        line["weak_base"] = weak_base
    print(json.dumps(line))
    g.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="cut the workload to about this many pairs per GPU")
    ap.add_argument("--cpu-pairs", type=int, default=2_000_000, help="size of the CPU baseline sample")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-next", action="store_true", help="skip --hashDepthRange / --cluster on the resident index")
    ap.add_argument("--no-weak-base", action="store_true", help="skip the human8 (one GPU's share of the N>1 workload) build at N=1")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample NVML clocks during the timed region")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    name = args.workload or ("1gb" if world == 1 else "human8")
    wl = WORKLOADS[name]
    # stdout carries the ONE JSON line and nothing else: libraries that write to file descriptor 1 themselves
    # (NCCL prints "NCCL version ..." there when NCCL_DEBUG is set) are sent to stderr for the duration of the run
    sys.stdout.flush()
    real_out = os.dup(1)
    os.dup2(2, 1)
    captured = io.StringIO()
    py_out, sys.stdout = sys.stdout, captured
    try:
        if args.impl == "reference":
            run_reference_arm(args, wl)
        else:
            run_ours(args, wl, name)
    finally:
        sys.stdout = py_out
        sys.stdout.flush()
        os.dup2(real_out, 1)
        os.close(real_out)
    sys.stdout.write(captured.getvalue())
    sys.stdout.flush()


if __name__ == "__main__":
    main()
