#!/usr/bin/env python
"""Compact one-line-per-launch table from tools/ncu_summary.py's csv.  Usage: ncu_table.py sum.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = [x.split(' [')[0] for x in rows[0]]
short = {'gpu__time_duration.sum': 'us', 'dram__bytes_read.sum': 'rdMB', 'dram__bytes_write.sum': 'wrMB',
         'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram%', 'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm%',
         'sm__warps_active.avg.pct_of_peak_sustained_active': 'occ%', 'launch__registers_per_thread': 'regs',
         'smsp__inst_executed.sum': 'Minst', 'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue%',
         'l1tex__t_sector_hit_rate.pct': 'l1hit', 'lts__t_sector_hit_rate.pct': 'l2hit'}
for k in h[1:]:
    if 'stalled' in k and '_per_issue' in k:
        short[k] = k.split('stalled_')[1].split('_per')[0][:9]
cols = [i for i, x in enumerate(h) if x in short]
print("%-26s" % "kernel" + " ".join("%9s" % short[h[i]][:9] for i in cols))
for r in rows[1:]:
    def f(i):
        x = r[i]
        try:
            v = float(x.replace(',', ''))
            if short[h[i]] == 'Minst':
                v /= 1e6
            return "%9.2f" % v
        except ValueError:
            return "%9s" % x[:9]
    name = r[0].replace('<unnamed>::', '').replace('void ', '')
    print("%-26s" % name[:26] + " ".join(f(i) for i in cols))
