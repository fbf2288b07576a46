#!/usr/bin/env python
"""Aggregate an ncu source page (cuda,sass view) by CUDA source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass [--kernel-name regex:...] | tools/ncu_lines.py [N]"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
for h in hi:
    H = rows[h]
    name = rows[h - 1][1] if h >= 1 and len(rows[h - 1]) > 1 else ""
    fpath = rows[h - 2][1] if h >= 2 and len(rows[h - 2]) > 1 else ""
    smp, ie, ti = H.index("# Samples"), H.index("Instructions Executed"), H.index("Thread Instructions Executed")
    body = []
    for r in rows[h + 1:]:
        if not r or r[0] in ("File Path", "Function Name", "Line No"):
            break
        if len(r) > ti and r[0].isdigit() and r[2] == "-":      # source-line summary rows
            body.append((int(r[smp] or 0), int(r[ie] or 0), int(r[ti] or 0), int(r[0]), r[1]))
    ts, tinst = sum(b[0] for b in body) or 1, sum(b[1] for b in body) or 1
    print("== %s :: %s  samples=%d warp-inst=%d thread-eff=%.1f" % (fpath.split("/")[-1], name[:60], ts, tinst,
          sum(b[2] for b in body) / tinst))
    for s, i, t, ln, txt in sorted(body, reverse=True)[:top]:
        print("%5.1f%% smp %5.1f%% inst eff %4.1f  L%-4d %s" % (100 * s / ts, 100 * i / tinst, t / max(i, 1), ln, txt.strip()[:100]))
