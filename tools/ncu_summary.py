#!/usr/bin/env python
"""Summaries of ncu output kept under profiles/ (the .ncu-rep files themselves stay in gpurun_out/).
  tools/ncu_summary.py launches gpurun_out/r01_launches.csv      -> per-kernel time shares of one build
  tools/ncu_summary.py raw gpurun_out/r01_fused.ncu-rep           -> key metrics per profiled launch"""
import collections
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "lts__t_bytes.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(path, skip_first_build=True):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H, data = rows[h], rows[h + 1:]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    recs = [(r[ki], float(r[vi].replace(",", "")), r[ui]) for r in data if len(r) > vi]
    # keep the last build only: everything after the last k_run_flags launch
    starts = [i for i, r in enumerate(recs) if r[0].startswith("k_run_flags")]
    if starts:
        recs = recs[starts[-1]:]
    agg = collections.OrderedDict()
    for n, v, u in recs:
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        key = n.split("(")[0][:90]
        agg.setdefault(key, [0, 0.0])
        agg[key][0] += 1
        agg[key][1] += v * scale
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | ms (ncu, serialised, cold cache) | share |")
    print("|---|---:|---:|---:|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f%% |" % (n, c, t, 100 * t / tot))
    print("| **total** | %d | %.3f | |" % (sum(v[0] for v in agg.values()), tot))


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H = rows[0]
    units = rows[1]
    for r in rows[2:]:
        print("### %s" % r[H.index("Kernel Name")][:100])
        for m in METRICS:
            if m in H:
                print("- %s = %s %s" % (m, r[H.index(m)], units[H.index(m)]))
        print()


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
