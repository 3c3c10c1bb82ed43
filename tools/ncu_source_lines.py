#!/usr/bin/env python
"""Per-SOURCE-LINE executed-instruction profile of a kernel from `ncu --page source --csv --print-source cuda,sass` output
(needs -lineinfo + --import-source on).  Usage: ncu_source_lines.py file.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hdr = rows[hi]
ie, te, smp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
wf, wfi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
data = []
for r in rows[hi + 1:]:
    if not r or not r[0].strip().isdigit():
        continue
    try:
        data.append((int(r[0]), r[1], int(r[ie]), int(r[te]), int(r[smp]), int(r[wf] or 0), int(r[wfi] or 0)))
    except ValueError:
        pass
tot = sum(d[2] for d in data) or 1
st = sum(d[4] for d in data) or 1
print("total warp instructions %d, thread instructions %d" % (tot, sum(d[3] for d in data)))
for d in sorted(data, key=lambda d: -d[2])[:top]:
    print("%4d %5.2f%% smp %5.2f%% thr/inst %4.1f smem wavefronts %6.1fM (ideal %6.1fM) | %s"
          % (d[0], 100 * d[2] / tot, 100 * d[4] / st, d[3] / max(d[2], 1), d[5] / 1e6, d[6] / 1e6, d[1].strip()[:100]))
