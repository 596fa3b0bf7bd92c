#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> compact CSV with per-kernel shares."""
import csv
import sys

rows = list(csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("==")))
w = csv.writer(sys.stdout)
w.writerow(["id", "kernel", "grid", "block", "gpu__time_duration.sum [us]"])
tot = {}
for r in rows:
    k = r["Kernel Name"].split("(")[0].replace("void <unnamed>::", "")
    t = float(r["Metric Value"].replace(",", ""))
    if r["Metric Unit"] in ("ns", "nsecond"):
        t /= 1000
    w.writerow([r["ID"], k, r["Grid Size"], r["Block Size"], round(t, 3)])
    tot[k] = tot.get(k, 0) + t
s = sum(tot.values())
for k, v in tot.items():
    w.writerow(["share", k, "", "", f"{v:.1f} us = {100 * v / s:.1f}%"])
