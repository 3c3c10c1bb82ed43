#!/usr/bin/env python
"""Opcode histogram per kernel of an object / library: sass_hist.py file kernel_substring"""
import collections, re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
c = collections.defaultdict(collections.Counter); name = None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = m.group(1); continue
    m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
    if m and name:
        c[name][m.group(2)] += 1
for k, v in c.items():
    if len(sys.argv) < 3 or sys.argv[2] in k:
        print(k[-90:], sum(v.values())); print("  ", v.most_common(28))
