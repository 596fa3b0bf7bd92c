#!/usr/bin/env python
"""One-process sweep of the TMA-staged step kernel's knobs against the direct-load kernels on the BASELINE shapes
(decide TMA on evidence).  Prints one line per (shape, variant, options): ms per trajectory, GB/s.

    python scripts/tma_sweep.py [--shapes c2,c3,c5_256,c4] [--quick]
"""
import argparse
import itertools
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="c2,c3,c5_256,c4")
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    import torch
    from naturaldiffusion_b200 import _lib
    args = types.SimpleNamespace(no_numa=True)
    ctx = bench.Ctx(args, 0, 1, 0)
    shapes = {"c2": ("c2", 0, "auto"), "c3": ("c3", 0, "auto"), "c5_256": ("c5", 256, "0"), "c4": ("c4", 0, "0"), "c5s_64": ("c5s", 64, "0")}
    defaults = dict(tma_tile_kb=2, tma_l2_hint=0, tma_dynamic=0, tma_warps=8, tma_ctas_per_sm=2, tma_max_stages=32)
    grid = []
    for tile, hint, dyn in itertools.product((2, 4), (0, 1, 2), (0, 1)):
        grid.append(dict(tma_tile_kb=tile, tma_l2_hint=hint, tma_dynamic=dyn))
    grid += [dict(tma_warps=16), dict(tma_ctas_per_sm=1, tma_warps=16), dict(tma_tile_kb=4, tma_warps=16, tma_dynamic=1), dict(tma_tile_kb=4, tma_ctas_per_sm=1, tma_warps=16, tma_dynamic=1),
             dict(tma_l2_hint=3), dict(tma_tile_kb=4, tma_l2_hint=3, tma_dynamic=1)]
    if a.quick:
        grid = grid[:4]
    for name in a.shapes.split(","):
        cfg, b, mk = shapes[name]
        w = bench.build_workload(ctx, cfg, b, mk, "stored")
        est = w["bytes_per_traj"] / 6.5e12 * 1e3
        block = max(1, int(60 / est))
        rows = []
        for variant, opts in [(0, {}), (1, {})] + [(2, o) for o in grid]:
            _lib.set_option("variant", variant)
            for k, v in {**defaults, **opts}.items():
                _lib.set_option(k, v)
            w["sampler"]._launch_cache.clear(); w["sampler"]._launches = None
            try:
                ms, _, _ = bench.time_resident(ctx, w, 3, 2, block, graph=True)
            except Exception as e:  # noqa: BLE001
                print(json.dumps({"shape": name, "variant": variant, "opts": opts, "error": str(e)[:200]}), flush=True)
                continue
            per = ms / (3 * block)
            row = {"shape": name, "variant": {0: "lean", 1: "generic", 2: "tma"}[variant], "opts": opts, "ms_per_trajectory": round(per, 5),
                   "gbs": round(w["bytes_per_traj"] / (per * 1e-3) / 1e9, 1)}
            rows.append(row)
            print(json.dumps(row), flush=True)
        best = max((r for r in rows if r["variant"] == "tma"), key=lambda r: r["gbs"], default=None)
        print(json.dumps({"shape": name, "best_tma": best, "lean": rows[0]["gbs"], "generic": rows[1]["gbs"]}), flush=True)
        del w
        torch.cuda.empty_cache()
    _lib.set_option("variant", 0)


if __name__ == "__main__":
    main()
