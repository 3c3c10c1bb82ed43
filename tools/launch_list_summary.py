#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics ... --csv` launch list: launch_list_summary.py in.csv out.txt "title" """
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]; ki = h.index('Kernel Name'); mi = h.index('Metric Name'); vi = h.index('Metric Value'); ii = h.index('ID'); ui = h.index('Metric Unit')
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki].split('(')[0].split('::')[-1]), {})[r[mi]] = (float(r[vi].replace(',', '')), r[ui])
agg = collections.OrderedDict()
for (i, k), m in d.items():
    a = agg.setdefault(k, collections.defaultdict(list))
    for mk, (v, u) in m.items():
        if mk == 'gpu__time_duration.sum': v = v * {'ns': 1e-3, 'us': 1, 'ms': 1e3}.get(u, 1)
        if mk.startswith('dram__bytes'): v = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        a[mk].append(v)
out = ["%-44s %5s %9s %9s %7s %7s %7s %7s %10s" % (sys.argv[3][:44], "n", "avg us", "Kinst", "issue%", "alu%", "fmaH%", "xu%", "dramKB")]
for k, a in agg.items():
    n = len(a['gpu__time_duration.sum']); f = lambda key: sum(a[key]) / max(len(a[key]), 1)
    out.append("%-44s %5d %9.2f %9.1f %7.1f %7.1f %7.1f %7.1f %10.1f" % (k[:44], n, f('gpu__time_duration.sum'), f('smsp__inst_executed.sum') / 1e3,
               f('smsp__issue_active.avg.pct_of_peak_sustained_active'), f('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'),
               f('sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed'), f('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'),
               (f('dram__bytes_read.sum') + f('dram__bytes_write.sum')) / 1e3))
open(sys.argv[2], 'w').write("\n".join(out) + "\n")
print("\n".join(out))
