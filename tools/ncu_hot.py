#!/usr/bin/env python
"""Top stall-sample instructions of a kernel: reads `ncu -i X --page source --csv` (SASS view).  usage: ncu_hot.py rep [n]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"] + (["--kernel-name", sys.argv[3]] if len(sys.argv) > 3 else []), capture_output=True, text=True).stdout
lines = out.splitlines()
# several kernels may be concatenated: split on "Kernel Name" lines
blocks, cur = [], None
for l in lines:
    if l.startswith('"Kernel Name"'):
        cur = [l]; blocks.append(cur)
    elif cur is not None:
        cur.append(l)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for b in blocks:
    print("=====", b[0][:160])
    rows = list(csv.DictReader(b[1:]))
    tot = sum(int(r["# Samples"] or 0) for r in rows)
    rows_s = sorted(rows, key=lambda r: -int(r["# Samples"] or 0))[:n]
    for r in rows_s:
        stalls = {k[6:]: int(v) for k, v in r.items() if k.startswith("stall_") and "Not Issued" not in k and v and int(v) > 0}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
        print("%5.1f%%  %-90s %s" % (100.0 * int(r["# Samples"] or 0) / max(tot, 1), r["Source"][:90], top))
