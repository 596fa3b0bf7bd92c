#!/usr/bin/env python
"""A compact pass over every kernel of libni_b200.so on small shapes, for compute-sanitizer (memcheck / racecheck /
initcheck):   compute-sanitizer --tool memcheck python scripts/sanitize_targets.py
Covers: specialised step kernels at 128 and 256 bits (exact row shapes, runtime loop, strided outputs, generated noise kept and
not, low-precision copy, per-sample norms, the uint8 output stage with its shared-memory staging), the generic kernel (scalar and
vector), the TMA-staged kernel (2/4 KB tiles, L2 hints, dynamic tile claiming), weighted sum (fp32/fp16/fp64 sources), Philox
with host and device offsets, the device counter, the pixel kernel, the FID rank-k update (cp.async double buffer + DMMA)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import naturaldiffusion_b200 as ni  # noqa: E402
from naturaldiffusion_b200 import _lib, generators  # noqa: E402
from naturaldiffusion_b200.coeffs import ddim_x0_coeffs, io_eps_cfg  # noqa: E402
from naturaldiffusion_b200.fid import FidAccumulator  # noqa: E402
from naturaldiffusion_b200.ops import fused_step, philox_normal, to_pixel_u8, weighted_sum_tensors  # noqa: E402
from naturaldiffusion_b200.sampler import NaturalInferenceSampler  # noqa: E402

dev = "cuda:0"
g = torch.Generator().manual_seed(0)
mk = lambda dt, *s: torch.randn(*s, generator=g).to(dt).to(dev)
n = 0
for dt in (torch.float32, torch.float16, torch.bfloat16):
    for shape, cout, ncond in (((3, 4, 32, 32), 8, 2), ((5, 3, 8, 8), 3, 1), ((2, 3, 5, 7), 3, 1)):
        B, C, H, W = shape
        for nt in (0, 3, 9):
            kw = dict(x_in=mk(dt, *shape), outs=[mk(dt, B, cout, H, W) for _ in range(ncond)], a=1.1, b=[-0.5, 0.2][:ncond], c_x0=0.7, c_xin=0.1,
                      terms=[(0.1 * (i + 1), mk(dt, *shape)) for i in range(nt)], gens=[(0.3, 5), (0.1, 6)][: 1 + nt % 2], seed=3, keep_gen=[True, False],
                      per_sample=C * H * W, out_sample_stride=cout * H * W, want_sumsq=True, lp_dtype=torch.bfloat16 if dt == torch.float32 else None)
            for variant, opts in ((0, dict(wide=1)), (0, dict(wide=0)), (1, {}), (2, dict(tma_tile_kb=2, tma_dynamic=0)), (2, dict(tma_tile_kb=4, tma_dynamic=1, tma_l2_hint=1))):
                _lib.set_option("variant", variant)
                for k, v in opts.items():
                    _lib.set_option(k, v)
                fused_step(**kw)
                n += 1
# the TMA ring WRAPPING (2 stages, ~3.5 tiles per CTA): stage buffers and claimed-index slots are rewritten while consumers run
big = (64, 4, 32, 32)
kwb = dict(x_in=mk(torch.float32, *big), outs=[mk(torch.float32, *big)], a=1.1, b=[-0.5], c_x0=0.7, terms=[(0.1 * (i + 1), mk(torch.float32, *big)) for i in range(3)],
           gens=[(0.2, 4)], seed=5, keep_gen=[True], per_sample=4096, out_sample_stride=4096, want_sumsq=True)
_lib.set_option("variant", 2)
_lib.set_option("tma_max_stages", 2); _lib.set_option("tma_ctas_per_sm", 1); _lib.set_option("tma_tile_kb", 2)
for dyn in (0, 1):
    _lib.set_option("tma_dynamic", dyn)
    fused_step(**kwb)
    n += 1
_lib.set_option("tma_max_stages", 32); _lib.set_option("tma_ctas_per_sm", 2); _lib.set_option("tma_dynamic", 0)
_lib.set_option("variant", 0)
# samplers: CIFAR matrix with the uint8 stage (PIX kernel), DDPM first-order with in-kernel noise + graph replay
t = ni.CoeffTriple.from_npz(os.path.join(ROOT, "naturaldiffusion_b200", "data", "weights", "step_10_weight_42.npz"))
s = NaturalInferenceSampler(t, ni.io_score_vp(t.node), 8, (3, 32, 32), device=dev, seed=1, track_sumsq=False)
den = lambda x, k: torch.tanh(x) * 0.5
pix = torch.empty(8, 32, 32, 3, dtype=torch.uint8, device=dev)
s.sample(den, pixels_out=pix)
s.sample(den)
td = generators.ddpm_triple(12)
c1, c2, _ = ddim_x0_coeffs(12)
sd = NaturalInferenceSampler(td, io_eps_cfg(c1, c2, 4.0), 4, (4, 16, 16), device=dev, seed=2)
den2 = lambda x, k: (torch.tanh(x).repeat(1, 2, 1, 1), torch.sin(x).repeat(1, 2, 1, 1))
sd.sample(den2)
sd.capture(den2)
sd.replay(); sd.replay()
# host pipeline
outs = [torch.empty(8, 32, 32, 3, dtype=torch.uint8).pin_memory() for _ in range(3)]
s.sample_host_many(den, None, outs, pixels=True, graph=True)
# stand-alone kernels
for dt, odt in ((torch.float32, torch.float32), (torch.float16, torch.float16), (torch.float64, torch.float32)):
    xs = [mk(torch.float32, 1000).to(dt) for _ in range(11)]
    weighted_sum_tensors([0.1] * 11, xs, out_dtype=odt)
    weighted_sum_tensors([0.1] * 3, [x[1:] for x in xs[:3]], out_dtype=odt)  # misaligned -> scalar path
ctr = torch.tensor([6], dtype=torch.int64, device=dev)
philox_normal((1001,), seed=1, tensor_id=2, device=dev)
philox_normal((1024,), seed=1, tensor_id=2, elem_offset=1, elem_offset_dev=ctr, dtype=torch.float16, device=dev)
_lib.check(_lib.lib().ni_counter_add(ctr.data_ptr(), 4, torch.cuda.current_stream().cuda_stream))
to_pixel_u8(mk(torch.float32, 3, 3, 8, 8))
for m, d in ((100, 64), (33, 100), (257, 192)):
    FidAccumulator(dim=d, device=dev).update(mk(torch.float32, m, d))
torch.cuda.synchronize()
print(f"sanitize_targets: {n} fused-step launches + samplers + stand-alone kernels done; launches={ni.launch_count()}")
