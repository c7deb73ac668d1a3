#!/usr/bin/env python
"""DRAM traffic of the tensor-core conv family from an ncu CSV log -> profiles/r02_umma_traffic.json, stamped with the hash of
the libszn.so it was taken from (bench.py refuses to call it `same build` otherwise).
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:umma_conv_kernel -s 135 -c 45 \
      --csv --log-file gpurun_out/r02_umma_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e
  python tools/ncu_traffic.py gpurun_out/r02_umma_traffic.csv <config> <precision> > profiles/r02_umma_traffic.json"""
import csv
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path, config, prec = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = list(csv.DictReader([l for l in open(path) if not l.startswith("==")]))
per = {}
for r in rows:
    k = (r["ID"], r["Kernel Name"])
    d = per.setdefault(k, {})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ns": 1.0, "us": 1e3, "ms": 1e6}.get(unit, 1.0)
    d[r["Metric Name"]] = v * scale
fam = {"fwd_dgrad": [0, 0.0, 0.0], "wgrad": [0, 0.0, 0.0]}
for (_, name), d in per.items():
    g = fam["wgrad"] if "(int)2" in name or ", 2," in name else fam["fwd_dgrad"]
    g[0] += 1
    g[1] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    g[2] += d.get("gpu__time_duration.sum", 0.0)


def build_id(path):
    """include/szn_build.h: hash of the kernel sources the library was built from (None for a library older than that)."""
    import ctypes
    try:
        lib = ctypes.CDLL(path)
        lib.szn_build_id.restype = ctypes.c_char_p
        return lib.szn_build_id().decode()
    except Exception:
        return None


n = sum(g[0] for g in fam.values())
tot = sum(g[1] for g in fam.values())
so = os.path.join(ROOT, "zeroshotsemanticsegmentation_b200", "libszn.so")
print(json.dumps({
    "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:umma_conv_kernel "
              "-s 135 -c 45 python bench.py --steps 1 --warmup 3 (one step = 45 launches)",
    "config": config, "precision": prec, "so_sha256_16": hashlib.sha256(open(so, "rb").read()).hexdigest()[:16],
    "build_id": build_id(so),
    "per_kernel": {k: {"launches": g[0], "dram_bytes": g[1], "ns": g[2], "dram_bytes_per_launch": g[1] / max(g[0], 1)}
                   for k, g in fam.items()},
    "umma_family_dram_bytes_per_launch": tot / max(n, 1), "umma_family_dram_bytes_per_step": tot, "launches": n}, indent=1))
