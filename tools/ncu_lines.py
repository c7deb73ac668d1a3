#!/usr/bin/env python
"""Stall samples, shared-memory wavefronts and executed instructions of one kernel aggregated per SOURCE LINE
(reads `ncu -i X --page source --csv --print-source cuda,sass`; needs a capture taken with --import-source on and a
library built with -lineinfo).   usage: ncu_lines.py file.ncu-rep source.cu [n]"""
import collections
import csv
import subprocess
import sys


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def main():
    rep, src_path = sys.argv[1], sys.argv[2]
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[hi]
    li, si, ii = hdr.index("Line No"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    wi = hdr.index("L1 Wavefronts Shared")
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= max(si, ii, wi) or not r[li].isdigit():
            continue
        a = agg[int(r[li])]
        a[0] += num(r[si])
        a[1] += num(r[ii])
        a[2] += num(r[wi])
    src = open(src_path).read().splitlines()
    tot_s, tot_i = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
    print("%s: %d stall samples, %.1f M warp instructions; lines of %s (other files' lines, e.g. inlined helpers, show the "
          "number only)" % (rows[1][1][:80] if len(rows) > 1 and len(rows[1]) > 1 else rep, tot_s, tot_i / 1e6, src_path))
    print("%7s %6s %12s %12s  %s" % ("samples", "line", "warp instr", "smem wavefr", "source"))
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
        text = src[ln - 1].strip()[:110] if 0 < ln <= len(src) else ""
        print("%6.1f%% %6d %12d %12d  %s" % (100 * a[0] / max(tot_s, 1), ln, a[1], a[2], text))


if __name__ == "__main__":
    main()
