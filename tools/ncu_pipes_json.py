#!/usr/bin/env python
"""Pipe / issue utilisation per STAGE from an .ncu-rep (one bench step, `ncu --set full`), duration-weighted over the stage's
launches -> JSON that bench.py attaches to roofline.stages[*].pipes.  Usage: ncu_pipes_json.py rep out.json"""
import csv
import io
import json
import subprocess
import sys

STAGE = [("pyr_", "pyramid"), ("fast_strip", "fast_cells"), ("quadtree", "quadtree"), ("blur_kernel", "blur"), ("orient_describe", "orient_describe"),
         ("hamming_top2", "hamming_top2")]
M = {"alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
     "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
     "xu_pipe_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
     "lsu_wavefronts_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
     "issue_slots_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active"}
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
kn, tm = h.index("Kernel Name"), h.index("gpu__time_duration.sum")
acc = {}
for r in rows[2:]:
    st = next((s for p, s in STAGE if p in r[kn]), None)
    if st is None:
        continue
    t = float(r[tm].replace(",", ""))
    a = acc.setdefault(st, {"t": 0.0, "launches": 0})
    a["t"] += t
    a["launches"] += 1
    for k, m in M.items():
        if m in h and r[h.index(m)]:
            a[k] = a.get(k, 0.0) + t * float(r[h.index(m)].replace(",", ""))
out = {}
for st, a in acc.items():
    out[st] = {k: round(a[k] / a["t"], 1) for k in M if k in a}
    out[st]["launches"] = a["launches"]
    lim = max(((k, v) for k, v in out[st].items() if k.endswith("_pct") and k not in ("occupancy_pct",)), key=lambda kv: kv[1])
    out[st]["most_utilised"] = lim[0]
json.dump({"source": "ncu --set full --clock-control none, one bench step (%s)" % sys.argv[1], "stages": out}, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
