#!/usr/bin/env python
"""CPU side: turn what scripts/r02_profile_final.sh left in gpurun_out/ into the committed summaries under profiles/
(launch lists with shares, per-launch ncu tables, details pages, source hotspots) and refresh profiles/traffic.json with the
hash of the kernel sources the capture was taken with.   python scripts/refresh_profiles.py"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
O, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def run(script, src, dst, *extra):
    if os.path.isfile(src):
        with open(dst, "w") as f:
            subprocess.run([sys.executable, os.path.join(ROOT, "scripts", script), src, *extra], stdout=f, check=True)


for cfg in ("c2", "c3"):
    run("summarize_launches.py", f"{O}/r02_{cfg}_launches.csv", f"{P}/r02_{cfg}_launches.csv")
for src, dst in (("c2_full", "c2_lean"), ("c3_full", "c3_lean"), ("c4_markov_full", "c4_markov_lean"), ("pixel_lean", "pixel_lean"), ("pixel_generic", "pixel_generic")):
    run("summarize_ncu.py", f"{O}/r02_{src}_raw.csv", f"{P}/r02_{dst}_summary.csv")
for n in ("r02_c3_step7_details.txt", "r02_c4_markov_details.txt", "r02_pytest_gpu.log", "r02_bench_reference_n1.json"):
    if os.path.isfile(f"{O}/{n}"):
        shutil.copy(f"{O}/{n}", f"{P}/{n}")
run("source_hotspots.py", f"{O}/r02_c3_step7_source.csv", f"{P}/r02_c3_source_hotspots.txt", "0")
if os.path.isfile(f"{O}/r02_bench_n1.json"):
    shutil.copy(f"{O}/r02_bench_n1.json", f"{P}/r02_bench_n1.json")

import bench  # noqa: E402


def mean_traffic(name):
    rows = list(csv.reader(open(f"{P}/r02_{name}_summary.csv")))
    h = rows[0]
    ir = [i for i, x in enumerate(h) if x.startswith("dram__bytes_read.sum [")][0]
    iw = [i for i, x in enumerate(h) if x.startswith("dram__bytes_write.sum [")][0]
    mul = lambda u: {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    ur, uw = h[ir].split("[")[1].rstrip("]"), h[iw].split("[")[1].rstrip("]")
    tot = [float(r[ir]) * mul(ur) + float(r[iw]) * mul(uw) for r in rows[1:]]
    return sum(tot) / len(tot), len(tot)


sha = bench.kernel_source_sha()
t = {}
for cfg, name in (("c2", "c2_lean"), ("c3", "c3_lean")):
    m, k = mean_traffic(name)
    t[f"{cfg}:stored"] = {"dram_bytes_per_launch": m, "launches": k, "kernel_source_sha": sha,
                          "source": f"profiles/r02_{name}_summary.csv (ncu --set full, one trajectory, mean of dram__bytes_read.sum + dram__bytes_write.sum)"}
t["_note"] = ("bench.py reports roofline.traffic only when kernel_source_sha matches the specialised step kernels' sources in the tree (csrc/ni_step_lean*, "
              "ni_common.cuh); DRAM reads equal the algorithmic read bytes; writes are under-counted by what is still dirty in the 126 MB L2 when the kernel ends")
json.dump(t, open(f"{P}/traffic.json", "w"), indent=1)
print("kernel_source_sha", sha, {k: round(v["dram_bytes_per_launch"]) for k, v in t.items() if k != "_note"})
