#!/usr/bin/env python
"""Per-launch time and DRAM bytes from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv` log: usage launch_summary.py launches.csv [min_ms]"""
import csv
import sys
from collections import OrderedDict

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
agg = OrderedDict()
for row in csv.DictReader(lines):
    k = (row['ID'], row['Kernel Name'][:72])
    agg.setdefault(k, {})[row['Metric Name']] = (float(row['Metric Value'].replace(',', '')), row['Metric Unit'])
tot = 0.0
for (i, n), m in agg.items():
    t, tu = m.get('gpu__time_duration.sum', (0, 'ns'))
    ms = t / 1e6 if tu == 'ns' else (t / 1e3 if tu == 'us' else (t if tu == 'ms' else t * 1e3))
    gb = lambda x: x[0] * {'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1.0}.get(x[1], 0)
    rd, wr = gb(m.get('dram__bytes_read.sum', (0, 'byte'))), gb(m.get('dram__bytes_write.sum', (0, 'byte')))
    if 'k_synth' not in n:
        tot += ms
    if ms > thr:
        print('%8.3f ms  R %6.2f GB  W %6.2f GB  %5.2f TB/s  %s' % (ms, rd, wr, (rd + wr) / ms if ms else 0, n))
print('total without the generator: %.1f ms' % tot)
