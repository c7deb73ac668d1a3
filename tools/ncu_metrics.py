#!/usr/bin/env python
"""Print the roofline-relevant metrics of every kernel in an .ncu-rep (reads `ncu -i X --page raw --csv`).
usage: ncu_metrics.py file.ncu-rep [extra-metric-substring ...]"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "gpc__cycles_elapsed.avg.per_second",
        "lts__t_sector_hit_rate.pct", "launch__waves_per_multiprocessor"]
SUB = ["pipe_tensor", "l1tex__m_xbar2l1tex_read_bytes", "smem", "lts__t_sectors_srcunit_tex_op_read.sum",
       "sm__inst_executed_pipe_uniform", "tmem", "utcmma", "l1tex__data_bank"]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    subs = SUB + sys.argv[2:]
    idx = [i for i, h in enumerate(hdr) if h in WANT or any(s in h for s in subs)]
    for r in rows[2:]:
        print("---")
        for i in idx:
            print("  %-78s %-10s %s" % (hdr[i], units[i], r[i][:100]))


if __name__ == "__main__":
    main()
