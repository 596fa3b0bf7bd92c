#!/usr/bin/env python
"""Summarise an ncu report (CPU side): per-launch key metrics of the NI kernels -> CSV on stdout.
usage: scripts/summarize_ncu.py gpurun_out/c2_full.ncu-rep > profiles/r01_c2_full_summary.csv"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second"]

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
idx = [hdr.index(w) for w in WANT if w in hdr]
w = csv.writer(sys.stdout)
w.writerow([f"{hdr[i]} [{units[i]}]" if units[i] else hdr[i] for i in idx])
for r in rows[2:]:
    w.writerow([r[i][:60] for i in idx])
