#!/bin/bash
# per-kernel durations of the pair-batch matching kernels (config 2 / 3 legs of bench.py) under ncu: pairs_launch_list.sh out.csv
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"bf_|window_|grid_build|stereo_|search_batch" -c ${2:-40} --csv --log-file $1 python bench.py --no-hamming --no-cpu --e2e-callers 1 --steps 2 --warmup 1 > /dev/null 2>&1
python - "$1" <<PY
import csv,collections,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); ii=h.index("ID"); gi=h.index("Grid Size"); bi=h.index("Block Size")
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki].split("(")[0].split("::")[-1], r[gi], r[bi]),{})[r[mi]]=float(r[vi].replace(",",""))
seen=set()
for k,m in d.items():
    if (k[1],k[2]) in seen: continue
    seen.add((k[1],k[2]))
    print("%-34s grid %-16s block %-12s %8.1f us %8.2f Minst issue %5.1f%%"%(k[1],k[2],k[3],m["gpu__time_duration.sum"]/1e3,m["smsp__inst_executed.sum"]/1e6,m["smsp__issue_active.avg.pct_of_peak_sustained_active"]))
PY
