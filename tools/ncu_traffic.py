#!/usr/bin/env python
"""DRAM traffic per stage and per frame from one `ncu --set full` capture of one bench step (summary csv made by
tools/ncu_summary.py).  Writes profiles/<name>_traffic.json, which bench.py reads to fill roofline.traffic.
Usage: ncu_traffic.py sum.csv frames out.json"""
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
STAGE = [("pyr_", "pyramid"), ("fast_strip", "fast_cells"), ("quadtree", "quadtree"), ("blur_kernel", "blur"),
         ("orient_describe", "orient_describe")]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    frames = int(sys.argv[2])
    hdr = rows[0]

    def col(prefix):
        i = [k for k, x in enumerate(hdr) if x.startswith(prefix)][0]
        return i, UNIT[re.search(r"\[(.*)\]", hdr[i]).group(1)]
    (it, ut), (ir, ur), (iw, uw) = col("gpu__time_duration.sum"), col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
    out = {}
    for r in rows[1:]:
        st = next((s for p, s in STAGE if p in r[0]), None)
        if st is None:
            continue
        o = out.setdefault(st, {"dram_bytes_per_frame": 0.0, "ncu_seconds": 0.0, "launches": 0})
        o["dram_bytes_per_frame"] += (float(r[ir].replace(",", "")) * ur + float(r[iw].replace(",", "")) * uw) / frames
        o["ncu_seconds"] += float(r[it].replace(",", "")) * ut
        o["launches"] += 1
    tot = sum(o["ncu_seconds"] for o in out.values())
    for o in out.values():
        o["share_of_step"] = o["ncu_seconds"] / tot
    json.dump({"frames_per_launch": frames, "source": "ncu --set full --clock-control none, one bench step", "stages": out},
              open(sys.argv[3], "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
