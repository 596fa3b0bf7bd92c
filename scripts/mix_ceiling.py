#!/usr/bin/env python
"""Bandwidth ceiling of this GPU as a function of the read:write mix, measured with our own stand-alone weighted-sum
kernel (n reads : 1 write, 201 MB fp32 tensors, larger than L2) next to torch's copy.  Puts the fused step's GB/s in
context: C2 moves 46 reads : 19 writes, C3 91 : 29, dense rows are almost pure reads."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from naturaldiffusion_b200.ops import weighted_sum_tensors  # noqa: E402

dev = torch.device("cuda:0")
N = 16384 * 3072
res = {}
for n in (1, 2, 3, 4, 6, 8, 12, 16):
    xs = [torch.randn(N, device=dev) for _ in range(n)]
    out = torch.empty(N, device=dev)
    cs = [1.0 / (i + 1) for i in range(n)]
    for _ in range(5):
        weighted_sum_tensors(cs, xs, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 50
    e0.record()
    for _ in range(reps):
        weighted_sum_tensors(cs, xs, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res[f"{n}r:1w"] = round((n + 1) * N * 4 / (ms * 1e-3) / 1e9)
    del xs, out
a, b = torch.randn(N, device=dev), torch.empty(N, device=dev)
for _ in range(5):
    b.copy_(a)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    b.copy_(a)
e1.record()
torch.cuda.synchronize()
res["torch_copy_1r:1w"] = round(2 * N * 4 / (e0.elapsed_time(e1) / 50 * 1e-3) / 1e9)
print(json.dumps({"GB/s by mix (201 MB fp32 tensors)": res}))
