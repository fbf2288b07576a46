#!/usr/bin/env python
"""bench.py - read pairs/s through the `--readFQB` minhash + index build, and % of the HBM roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 1gb|yeast|human8|custom] [--impl reference]

One "step" = one complete index build (readFQB + fillHashTable, hash10x.c:1200-1205) over the
synthetic FQB of the workload.  `value` times the device-resident build (FQB already in HBM, index
left in HBM) with CUDA events; `e2e` times the same build through the C ABI's host entry point
(h10x_gpu_build_host: pinned host FQB -> H2D -> build -> D2H of every index array).  The default
N=1 workload is BASELINE.json configs[2] ("1 Gb diploid genome at 60x, ~200M read pairs, 24 GB
FQB, -B 28, 1xB200"), the largest single-GPU configuration; inputs are far larger than the 126 MB L2.

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

WORKLOADS = {
    # name: genome, barcodes, pairs/barcode range, molecules, mol len, snp period, err rate, B
    "1gb": dict(genome_len=1_000_000_000, n_barcodes=500_000, pairs_min=300, pairs_max=500,
                mol_per_barcode=10, mol_len=50_000, snp_period=1000, err_rate=0.0005, B=28, read_len=160,
                desc="BASELINE configs[2]: synthetic 1 Gb diploid genome at 60x, ~200M read pairs, 24 GB FQB, -B 28 "
                     "(160+160 bp reads: with 151 bp the zero padding of the last packed word adds ~0.39 novel hashes "
                     "per pair and the reference itself dies with 'hashTableSize is too small' at -B 28)"),
    "yeast": dict(genome_len=12_000_000, n_barcodes=10_000, pairs_min=150, pairs_max=350,
                  mol_per_barcode=10, mol_len=50_000, snp_period=500, err_rate=0.004, B=24, read_len=151,
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
    return orc.synth_fqb(p, 0, r1), nb


def time_reference(orc, recs, B, threads_note=True):
    """Seconds for the reference's --readFQB step on recs; kind 'reference' when oracle/_ref exists."""
    exe = orc.ref_binary("hash10x")
    if exe is None:
        t = orc.time_build(recs, B=B)
        return t, "port", "oracle/h10x_oracle.c (CPU restatement; oracle/_ref not built)"
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.NamedTemporaryFile(suffix=".fqb", dir=shm) as f:
        recs.tofile(f)
        f.flush()
        t0 = time.perf_counter()
        r = subprocess.run([exe, "-B", str(B), "--readFQB", f.name], capture_output=True, text=True)
        dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("reference failed: " + r.stderr[-300:])
    return dt, "reference", "oracle/_ref/hash10x (unmodified reference, gcc -O3), wall clock of `-B %d --readFQB`" % B


def run_reference_arm(args, wl):
    from oracle import orc
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    recs, nb = cpu_sample(orc, wl, args.cpu_pairs)
    pairs = recs.shape[0]
    times = []
    kind = sample = None
    for i in range(args.warmup + args.steps):
        dt, kind, how = time_reference(orc, recs, wl["B"])
        if i >= args.warmup:
            times.append(dt)
    dt = sum(times) / len(times)
    val = pairs / dt
    sample = "first %d barcode runs = %d read pairs of the workload per step; %s" % (nb, pairs, how)
    line = {"impl": "reference", "metric": "read pairs/sec through --readFQB minhash+index build",
            "value": val, "unit": "read pairs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": wl["desc"], "B": wl["B"], "k": 21, "w": 31},
            "cpu_baseline": {"value": val, "unit": "read pairs/s", "cores": 1, "kind": kind, "sample": sample,
                             "host_cores": os.cpu_count(),
                             "note": "--readFQB is single-threaded in the reference, also with -DOMP (hash10x.c:1247 is the only omp pragma)"},
            "e2e": {"value": val, "unit": "read pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------- GPU arm

def run_ours(args, wl):
    import ctypes as C
    import numpy as np
    import torch
    import hash10x_b200
    from hash10x_b200 import synth as gsynth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    # --- synthetic FQB generated in HBM: ONE data set (same genome), each rank a contiguous barcode range ---
    from hash10x_b200 import shard
    p = synth_params(gsynth, wl, seed=3)
    if args.pairs:
        mean = (wl["pairs_min"] + wl["pairs_max"]) / 2
        p.nBarcodes = max(2, int(args.pairs / mean))
    p.nBarcodes *= world                     # weak scaling: the per-GPU share stays fixed
    _n_total, off = gsynth.layout(p)
    cut = shard.plan_shards(off, world)
    r0, r1 = shard.shard_records(off, cut, rank)
    n_rec = r1 - r0
    fqb = torch.empty(n_rec * 30, dtype=torch.int32, device=dev)
    gsynth.fill_device(p, off, r0, r1, fqb.data_ptr())
    torch.cuda.synchronize()

    g = hash10x_b200.Hash10xGPU(B=wl["B"], device=local)
    stream = torch.cuda.current_stream()
    if world > 1:
        g.dist_init(rank, world, shard.share_unique_id(dist, rank, hash10x_b200.Hash10xGPU.dist_unique_id))

    def build_resident():
        if world > 1:
            g.build_device_dist(fqb.data_ptr(), n_rec, stream.cuda_stream)
        else:
            g.build_device(fqb.data_ptr(), n_rec, stream.cuda_stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # --- value: device-resident build, CUDA events on the launching stream ---
    sampler = ClockSampler(local, enabled=not args.no_clocks)
    sampler.start()
    for _ in range(args.warmup):
        build_resident()
    barrier()
    sampler.window()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = {}
    launches = 0
    e0.record(stream)
    lib_ms = 0.0
    for _ in range(args.steps):
        t_py = time.perf_counter()
        build_resident()
        if os.environ.get("H10X_TRACE"):
            print("py-trace build call %.3f ms" % ((time.perf_counter() - t_py) * 1e3), file=sys.stderr)
        s = g.stats()
        lib_ms += s["msTotal"]
        launches += s["kernelLaunches"]
        for k_, v in s["msStage"].items():
            stage_ms[k_] = stage_ms.get(k_, 0.0) + v
    e1.record(stream)
    barrier()
    clocks = sampler.summary()
    ms = e0.elapsed_time(e1)
    stats = g.stats()
    ms, total_pairs = shard.job_time_and_units(dist, torch, ms, n_rec, dev)
    ms_step = ms / args.steps
    value = total_pairs / (ms_step * 1e-3)

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
            d2h = (4 << wl["B"]) + 8 * hn + 4 * hn + 4 * nbm + 4 * nbm + 8 * (nbm + 1) + 8 * H + 8 * (hn + 1) + 4 * H
        else:       # rank 0: table + values + depths + its blocks and ClusterHash lists (hash->code parts stay resident)
            d2h = (4 << wl["B"]) + 8 * hn + 4 * hn + 4 * nbm + 4 * nbm + 8 * (nbm + 1) + 8 * H
        e2e = {"value": total_pairs / dt, "unit": "read pairs/s", "h2d_bytes_per_step": n_rec * 120,
               "d2h_bytes_per_step": int(d2h), "ms_per_step": dt * 1e3, "steps": e2e_steps,
               "bytes_are": "rank 0's, per step",
               "api": ("h10x_gpu_build_host" if world == 1 else "h10x_gpu_build_host_dist") +
                      " (include/h10x_gpu.h): pinned host FQB -> index arrays in pinned host memory"}
        del host_t

    if rank != 0:
        g.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    # --- roofline of the dominant kernel/stage (per-stage CUDA-event times from the library) ---
    peak, peak_src = measured_peaks()
    dom = max(stage_ms.items(), key=lambda kv: kv[1])
    R, M, H, D, nB = stats["nRecords"], stats["nMoshes"], stats["nHashes"], stats["nBins"], stats["nBlocks"]
    stage_bytes = {   # algorithmic bytes of each stage per build (DESIGN.md "kernels")
        "moshes": 120 * R + 12 * M, "fused": 120 * R + 8 * H, "blocksort": 24 * M, "dedup": 8 * H + 12 * H,
        "hashsort": 24 * H, "binids": 4 * H + 40 * D, "entryids": 8 * H, "codes": 20 * H + 12 * D,
        "clusters": 12 * H + 8 * H, "table": (4 << wl["B"]) + 12 * D, "runs": 4 * R * 3, "other": 0}
    dom_ms = dom[1] / args.steps
    dom_bytes = stage_bytes.get(dom[0], 0)
    ach = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic = None
    try:        # DRAM bytes of the fused kernel from the committed ncu capture, scaled per read pair
        with open(os.path.join(ROOT, "profiles", "r01_fused_traffic.json")) as f:
            tj = json.load(f)
        if dom[0] == "fused":
            traffic = tj["dram_bytes_per_pair"] * R
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom[0], "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                "traffic_source": "profiles/r01_fused_traffic.json (ncu dram__bytes_read+write at 50M pairs) x pairs" if traffic else None,
                "note": "the fused kernel is bound by the integer pipes (ncu: ALU 59 %, FMA/IMAD 42 %, issue 74 %, DRAM 7 %), see profiles/README.md",
                "ms_per_launch_group": dom_ms, "share_of_step": dom_ms / ms_step}
    pipe_ach = stats["algorithmicBytes"] / (ms_step * 1e-3) / 1e9
    pipeline = {"bound": "hbm", "algorithmic_bytes": stats["algorithmicBytes"],
                "bytes_per_pair": stats["algorithmicBytes"] / max(1, R), "achieved": pipe_ach, "peak": peak,
                "unit": "GB/s", "frac": pipe_ach / peak,
                "formula": "120 R + 12 H + 12 D + 4*2^B + 32 (nB+1)  (SURVEY.md 8d)"}

    # --- cpu baseline: the reference's own CPU path on a bounded sample, rank 0, N=1 only ---
    cpu = None
    if world == 1 and not args.no_cpu:
        from oracle import orc          # the only use of oracle/ in this arm: the timed CPU baseline
        recs, nb = cpu_sample(orc, wl, args.cpu_pairs)
        dt, kind, how = time_reference(orc, recs, wl["B"])
        cpu = {"value": recs.shape[0] / dt, "unit": "read pairs/s", "cores": 1, "kind": kind,
               "sample": "first %d barcode runs = %d read pairs of the same workload, %.1f s; %s"
                         % (nb, recs.shape[0], dt, how), "host_cores": os.cpu_count()}

    line = {"metric": "read pairs/sec through --readFQB minhash+index build; % of HBM roofline",
            "value": value, "unit": "read pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": wl["desc"] + (" (cut to %d pairs per GPU by --pairs)" % n_rec if args.pairs else ""),
                       "pairs_per_gpu": n_rec, "barcodes_per_gpu": int(nB), "B": wl["B"], "k": 21, "w": 31,
                       "l2": "inputs (%.1f GB per GPU) are larger than the 126 MB L2; no flush needed" % (n_rec * 120 / 1e9),
                       "parallelism": "1 GPU" if world == 1 else
                       "%d ranks: barcode-range shards, NCCL all-to-all-v of rank-distinct hashes to hash-range owners, "
                       "global bin ids; hashValue/hashDepth/hashIndex on rank 0" % world},
            "roofline": roofline, "pipeline_roofline": pipeline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks, "lib_ms_per_step": lib_ms / args.steps,
            "stage_ms": {k_: v / args.steps for k_, v in stage_ms.items() if v},
            "counts": {"pairs": R, "moshes": M, "block_unique_hashes": H, "bins": D, "blocks": nB,
                       "peak_device_bytes": stats["peakDeviceBytes"]}}
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
            run_ours(args, wl)
    finally:
        sys.stdout = py_out
        sys.stdout.flush()
        os.dup2(real_out, 1)
        os.close(real_out)
    sys.stdout.write(captured.getvalue())
    sys.stdout.flush()


if __name__ == "__main__":
    main()
