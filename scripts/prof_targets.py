#!/usr/bin/env python
"""Small launch sequences for ncu (one kernel family each; run under `ncu -k regex:...`):
    pixel   C2 trajectories ending in the fused uint8 output stage (lean PIX instantiation; `--variant 1` = generic byte stores)
    wsum    stand-alone ni_weighted_sum: 8 fp32 sources and 4 fp64 sources of 201 MB (the function-level drop-ins)
    fid     ni_fid_accumulate on [8192, 2048] fp32 activations
    normal  ni_philox_normal on a 201 MB tensor
Also prints CUDA-event timings of the same launches (not taken under the profiler when run bare)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import naturaldiffusion_b200 as ni  # noqa: E402
from naturaldiffusion_b200 import _lib  # noqa: E402
from naturaldiffusion_b200.ops import philox_normal, weighted_sum_tensors  # noqa: E402
from naturaldiffusion_b200.sampler import NaturalInferenceSampler  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("what", choices=["pixel", "wsum", "fid", "normal"])
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda:0")
_lib.set_option("variant", a.variant)


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = {"what": a.what, "variant": a.variant}
if a.what == "pixel":
    W = os.path.join(ROOT, "naturaldiffusion_b200", "data", "weights", "step_10_weight_42.npz")
    t = ni.CoeffTriple.from_npz(W)
    B = 4096
    s = NaturalInferenceSampler(t, ni.io_score_vp(t.node), B, (3, 32, 32), device=dev, seed=888, advance=0)
    o = philox_normal((B, 3, 32, 32), seed=1, tensor_id=9, device=dev)
    noise = philox_normal((B, 3, 32, 32), seed=1, tensor_id=0, device=dev)
    pix = torch.empty(B, 32, 32, 3, dtype=torch.uint8, device=dev)
    s.sample(lambda x, k: o, noise=noise, pixels_out=pix)
    st = torch.cuda.current_stream().cuda_stream
    ms = timed(lambda: s.step(9, o, st), a.reps * 10)
    units = s.plan.units(9, 1) - 1 + 0.25  # x_K is not written; the uint8 image is (1/4 of a tensor)
    out.update(ms_last_step=ms, gbs=units * s.numel * 4 / ms / 1e6, units=units)
elif a.what == "wsum":
    n = 16384 * 3072
    for dt, k, odt in ((torch.float32, 8, torch.float32), (torch.float64, 4, torch.float32), (torch.float16, 16, torch.float16), (torch.float32, 2, torch.float32)):
        xs = [torch.randn(n, device=dev, dtype=torch.float32).to(dt) for _ in range(k)]
        dst = torch.empty(n, device=dev, dtype=odt)
        cs = [0.1 * (i + 1) for i in range(k)]
        ms = timed(lambda: weighted_sum_tensors(cs, xs, out=dst), a.reps)
        byt = n * (k * xs[0].element_size() + dst.element_size())
        out[f"{str(dt).split('.')[-1]}x{k}"] = dict(ms=ms, gbs=byt / ms / 1e6)
        del xs, dst
elif a.what == "fid":
    from naturaldiffusion_b200.fid import FidAccumulator
    m, d = 8192, 2048
    x = torch.randn(m, d, device=dev)
    acc = FidAccumulator(dim=d, device=dev)
    ms = timed(lambda: acc.update(x), a.reps)
    x64 = x.double()
    S = torch.zeros(d, d, dtype=torch.float64, device=dev)
    ms_torch = timed(lambda: S.addmm_(x.double().t(), x.double()), a.reps)
    flops = m * d * (d + 1.0)  # the symmetric half, multiply + add
    out.update(m=m, d=d, ms=ms, tflops_symmetric=flops / ms / 1e9, tflops_full_equivalent=2.0 * m * d * d / ms / 1e9,
               torch_fp64_addmm_ms=ms_torch, torch_tflops=2.0 * m * d * d / ms_torch / 1e9)
else:
    n = 16384 * 3072
    dst = torch.empty(n, device=dev)
    ms = timed(lambda: philox_normal((n,), seed=3, tensor_id=1, out=dst), a.reps)
    out.update(ms=ms, gbs=n * 4 / ms / 1e6, gsamples_per_s=n / ms / 1e6)
print(json.dumps(out))
