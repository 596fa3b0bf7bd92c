#!/usr/bin/env python
"""Top stall locations (SASS) of one ni_step launch from an ncu report with --import-source on, or from the
`--page source --csv` export of one launch that scripts/r02_profile.sh makes on the box.
usage: scripts/source_hotspots.py report.ncu-rep|source.csv [launch_index] > profiles/...txt"""
import csv
import subprocess
import sys

rep, which = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 7
if rep.endswith(".csv"):
    txt = open(rep).read()
else:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:ni_step"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(txt.splitlines()):
    if r and r[0] == "Kernel Name":
        cur = []
        blocks.append(cur)
        print("kernel:", r[1][:150])
        continue
    if cur is not None and r:
        cur.append(r)
b = blocks[min(which, len(blocks) - 1)]
hdr, data = b[0], b[1:]
i_src, i_s = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[i_s] or 0) for r in data) or 1
print(f"ncu --set full, source page (SASS), launch #{which} of the capture: {len(data)} SASS instructions, {tot} stall samples")
print("samples  share  instruction")
for r in sorted(data, key=lambda r: -int(r[i_s] or 0))[:14]:
    print(f"{int(r[i_s] or 0):7d}  {100 * int(r[i_s] or 0) / tot:5.1f}%  {r[i_src].strip()}")
grp = lambda pat: 100 * sum(int(r[i_s] or 0) for r in data if any(p in r[i_src] for p in pat)) / tot
print(f"samples at LDG {grp(['LDG']):.1f}%, at FFMA/FMUL (first consumers of loaded data) {grp(['FFMA', 'FMUL']):.1f}%, at STG {grp(['STG']):.1f}%, at LDC/ULDC/LDCU (parameter tables) {grp(['LDC']):.1f}%")
