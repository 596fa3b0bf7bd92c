#!/usr/bin/env python
"""Summarise an ncu capture (CPU side): per-launch key metrics of the NI kernels -> CSV on stdout.
Input: a .ncu-rep (read with `ncu -i ... --page raw --csv`) or the raw CSV that scripts/r02_profile.sh exports on the box.
usage: scripts/summarize_ncu.py gpurun_out/r02_c3_full_raw.csv > profiles/r02_c3_lean_summary.csv"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "sm__cycles_elapsed.avg.per_second", "dram__cycles_elapsed.avg.per_second"]

src = sys.argv[1]
if src.endswith(".csv"):
    out = open(src).read()
else:
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(l for l in out.splitlines() if not l.startswith("==")))
hdr, units = rows[0], rows[1]
idx = [hdr.index(w) for w in WANT if w in hdr]
w = csv.writer(sys.stdout)
w.writerow([f"{hdr[i]} [{units[i]}]" if units[i] else hdr[i] for i in idx])
for r in rows[2:]:
    w.writerow([r[i][:70] for i in idx])
