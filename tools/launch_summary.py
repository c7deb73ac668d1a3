#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one step (between two conv1_1 forward launches),
per-kernel totals and the tensor-core conv launches in order.   usage: launch_summary.py launches.csv [step_index]"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    m = re.search(r"szn::(\w+)(<[^>]*>)?", name)
    if m:
        return m.group(1) + (m.group(2) or "")
    return re.sub(r"\(.*", "", name)[:70]


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else -1
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    starts = [i for i, r in enumerate(rows) if re.search(r"conv1_1_(fwd|tc)_kernel", r["Kernel Name"])]
    if not starts:
        sys.exit("no conv1_1 forward launch in %s: cannot find a step" % path)
    starts.append(len(rows))
    k = which if which >= 0 else len(starts) - 2
    step = rows[starts[k]:starts[k + 1]]
    # bench.py measures a cuBLAS TF32 GEMM peak after the timed steps: not part of a step
    probe = []
    first = next((i for i, r in enumerate(step) if re.search(r"cutlass|cublas|gemm", r["Kernel Name"], re.I)), None)
    if first is not None:
        last_own = max(i for i, r in enumerate(step[:first]) if "szn::" in r["Kernel Name"])
        step, probe = step[:last_own + 1], step[last_own + 1:]
    tot = sum(float(r["Metric Value"]) for r in step)
    print("step %d of %d: %d launches, %.3f ms total (serialised, cold-cache: compare shares)" %
          (k, len(starts) - 1, len(step), tot / 1e6))
    agg = OrderedDict()
    for r in step:
        d = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
        d[0] += 1
        d[1] += float(r["Metric Value"])
    print("%-64s %5s %10s %7s" % ("kernel", "n", "ms", "share"))
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-64s %5d %10.3f %6.1f%%" % (name, n, t / 1e6, 100 * t / tot))
    if probe:
        print("(%d library launches of bench.py's in-run TF32 GEMM peak probe, after the step's last kernel, left out)" % len(probe))
    print("\ntensor-core conv launches in order:")
    for r in step:
        if "umma_conv_kernel" in r["Kernel Name"]:
            print("  %-44s grid %-14s %9.3f ms" % (short(r["Kernel Name"]), r["Grid Size"], float(r["Metric Value"]) / 1e6))


if __name__ == "__main__":
    main()
