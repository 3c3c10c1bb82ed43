#!/usr/bin/env python
"""Executed-instruction profile of one kernel of an .ncu-rep along its SASS, in blocks of N instructions.
Usage: ncu_sass_regions.py rep kernel_regex [block]"""
import collections
import csv
import io
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
blk = int(sys.argv[3]) if len(sys.argv) > 3 else 50
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# several launches may match: keep the first table only
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
rows = rows[hdr_i[0]:(hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows))]
hdr = rows[0]
ie, te, src = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Source")
smp = hdr.index("# Samples") if "# Samples" in hdr else None
data = []
for r in rows[1:]:
    try:
        data.append((int(r[ie]), int(r[te]), r[src], int(r[smp]) if smp is not None and r[smp] else 0))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
stot = sum(d[3] for d in data) or 1
print("total warp inst %d, thread inst %d, SASS lines %d" % (tot, sum(d[1] for d in data), len(data)))
for i in range(0, len(data), blk):
    b = data[i:i + blk]
    s = sum(d[0] for d in b)
    t = sum(d[1] for d in b)
    ops = collections.Counter((d[2].split()[1] if d[2].startswith('@') else d[2].split()[0]) for d in b if d[2])
    print("%5d inst%6.2f%% samples%6.2f%% thr/inst=%5.1f  %s" % (i, 100 * s / tot, 100 * sum(d[3] for d in b) / stot, t / max(s, 1),
                                                        dict(ops.most_common(6))))
