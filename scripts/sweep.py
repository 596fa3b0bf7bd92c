#!/usr/bin/env python
"""Run bench.py over kernel variants / options on the GPU box and print one compact line each.
usage: scripts/sweep.py [--lib path.so ...]   (each --lib is an alternative build of libni_b200.so)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNS = [
    ("c2", ["--variant", "1"]),
    ("c2", ["--variant", "2"]),
    ("c2", ["--variant", "2", "--opt", "tma_warps=4"]),
    ("c2", ["--variant", "2", "--opt", "tma_warps=12"]),
    ("c2", ["--variant", "2", "--opt", "tma_warps=16"]),
    ("c2", ["--variant", "2", "--opt", "tma_warps=8", "--opt", "tma_ctas_per_sm=2"]),
    ("c2", ["--variant", "2", "--opt", "tma_warps=4", "--opt", "tma_ctas_per_sm=2"]),
    ("c2", ["--variant", "2", "--opt", "tma_warps=16", "--opt", "tma_smem_kb=120"]),
    ("c3", ["--variant", "1"]),
    ("c3", ["--variant", "2"]),
    ("c3", ["--variant", "2", "--opt", "tma_warps=12"]),
    ("c3", ["--variant", "2", "--opt", "tma_warps=16"]),
    ("c3", ["--variant", "2", "--opt", "tma_warps=8", "--opt", "tma_ctas_per_sm=2"]),
]
libs = [a for i, a in enumerate(sys.argv) if i > 0 and sys.argv[i - 1] == "--lib"] or [None]
if "--tma2" in sys.argv:
    RUNS = [("c2", ["--variant", "1"]), ("c3", ["--variant", "1"])]
    for cfg in ("c2", "c3"):
        for w in (8, 12, 16):
            for c in (1, 2):
                RUNS.append((cfg, ["--variant", "2", "--opt", f"tma_warps={w}", "--opt", f"tma_ctas_per_sm={c}"]))
        RUNS.append((cfg, ["--variant", "2", "--opt", "tma_warps=16", "--opt", "tma_ctas_per_sm=2", "--opt", "tma_smem_kb=226"]))
        RUNS.append((cfg, ["--variant", "2", "--opt", "tma_warps=8", "--opt", "tma_ctas_per_sm=3"]))
if "--tma3" in sys.argv:
    RUNS = [("c2", ["--variant", "1"]), ("c3", ["--variant", "1"])]
    for cfg in ("c2", "c3"):
        for w, c in ((6, 1), (8, 1), (12, 1), (4, 2), (6, 2), (7, 2), (4, 3)):
            RUNS.append((cfg, ["--variant", "2", "--opt", f"tma_warps={w}", "--opt", f"tma_ctas_per_sm={c}"]))
if "--pdl" in sys.argv:
    RUNS = [(c, ["--opt", f"pdl={v}"] + e) for c, e in (("c2", []), ("c3", []), ("c4", []), ("c5", []), ("c5", ["--batch", "4"])) for v in (0, 1)]
if "--policy" in sys.argv:  # step-kernel load flavour forced L2-friendly (1) / streaming (2) / auto (0) on every shape
    shapes = [("c2", []), ("c3", []), ("c4", ["--markov", "0"]), ("c4", []), ("c5", []), ("c5", ["--markov", "0"]), ("c5s", []),
              ("c5", ["--batch", "256", "--markov", "0"])]
    RUNS = [(c, e + ["--opt", f"load_policy={v}"]) for c, e in shapes for v in (1, 2, 0)]
if "--ldg-only" in sys.argv:
    RUNS = [("c2", ["--variant", "1"]), ("c3", ["--variant", "1"])]
if "--cfgs" in sys.argv:  # --cfgs c2,c3,c5: the direct-load kernel on the named configs
    RUNS = [(c, ["--variant", "1"]) for c in sys.argv[sys.argv.index("--cfgs") + 1].split(",")]
out = []
for lib in libs:
    for cfg, extra in RUNS:
        env = dict(os.environ)
        if lib:
            env["NI_B200_LIB"] = os.path.abspath(lib)
        steps = {"c2": "1500", "c3": "300", "c4": "50", "c5": "300", "c5s": "300"}.get(cfg, "100")
        if cfg == "c4" and "0" in extra[:2]:
            steps = "4"   # dense DDPM-250 rows: 150 ms per trajectory
        if "256" in extra:
            steps = "50"
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--config", cfg, "--steps", steps, "--warmup", "20", "--no-cpu-baseline", "--no-e2e"] + extra
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=120)
        except subprocess.TimeoutExpired:
            print(f"{lib} {cfg} {extra} TIMEOUT (120 s) -- aborting the sweep", flush=True)
            break
        try:
            j = json.loads(r.stdout.strip().splitlines()[-1])
            line = f"{os.path.basename(lib) if lib else 'default':28s} {cfg} {' '.join(extra):42s} ms/traj={j['ms_per_step']:.4f} GB/s={j['roofline']['achieved']:.0f} frac={j['roofline']['frac']:.3f} sm_mhz={j['clocks']['sm_mhz']}"
        except Exception as e:  # noqa: BLE001
            line = f"{lib} {cfg} {extra} FAILED: {r.stderr[-400:]}"
        print(line, flush=True)
        out.append(line)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "sweep.txt"), "a").write("\n".join(out) + "\n")
