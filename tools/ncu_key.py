#!/usr/bin/env python
"""Key per-kernel numbers from `ncu -i X.ncu-rep --page raw --csv` output (file argument): duration, DRAM bytes and
throughput, issue utilisation, occupancy, the leading stall reasons."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum']
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for d in data:
    print('---- %s' % d[hdr.index('Kernel Name')][:110])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print('  %-62s %16s %s' % (w, d[i], units[i]))
    st = sorted(((float(d[hdr.index(h)].replace(',', '') or 0), h) for h in stall), reverse=True)[:5]
    print('  stalls: ' + ', '.join('%s %.2f' % (h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')], v) for v, h in st))
