#!/usr/bin/env python
"""The reference's SD3 driver (`sd_natural_inference_tx`, src/SD3NaturalInference.py:171-245) on the B200 path:
28-step flow matching on a 16 x 128 x 128 latent with classifier-free guidance 7, the weight table read from the
reference's csv, fp16 state like the reference; one fused `ni_step` per step (input mix sigma*noise + (1-sigma)*avg,
x0 = x - sigma*v, CFG on the x0's, weighted history sum), the VAE un-scaling folded into the last step.

Offline there is no diffusers pipeline and no checkpoint: the transformer is the SD3-medium-shaped MMDiT stand-in
(random init) and the prompt embeddings are random tensors of the pipeline's shapes, so the latents are meaningless;
what runs is the exact data flow.  `--small` uses a small instance of the same architecture on a 32 x 32 latent.

  python examples/sd3_pipeline.py --small
  python examples/sd3_pipeline.py --table sd3_step_28_weight_sharp.csv --batch 4
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from naturaldiffusion_b200 import presets  # noqa: E402
from naturaldiffusion_b200.denoisers import MMDiT, mmdit_sd3_medium  # noqa: E402

WEIGHTS = os.path.join(ROOT, "naturaldiffusion_b200", "data", "weights")


@torch.no_grad()
def run(table="sd3_step_28_weight.csv", batch=4, small=False, seed=10, quiet=False):
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    if small:
        latent, ctx_len, ctx_dim, pooled_dim = (16, 32, 32), 7, 32, 16
        net = MMDiT(dim=64, depth=2, heads=4, ctx_dim=ctx_dim, pooled_dim=pooled_dim, max_grid=32)
    else:
        latent, ctx_len, ctx_dim, pooled_dim = (16, 128, 128), 333, 4096, 2048
        net = mmdit_sd3_medium()
    net = net.to(dev).half().eval()
    mk = lambda *s: torch.randn(*s, device=dev).half()
    ctx, pooled, neg_ctx, neg_pooled = mk(batch, ctx_len, ctx_dim), mk(batch, pooled_dim), mk(batch, ctx_len, ctx_dim), mk(batch, pooled_dim)
    # latents = latents / vae.scaling_factor + vae.shift_factor (src/SD3NaturalInference.py:238; SD3 VAE: 1.5305, 0.0609)
    sampler, wrap = presets.sd3(os.path.join(WEIGHTS, table), batch, latent=latent, seed=seed, device=dev,
                                final_scale=1.0 / 1.5305, final_bias=0.0609)
    den = wrap(net, ctx, pooled, neg_ctx, neg_pooled)
    sampler.sample(den)  # warm-up (allocations, cuBLAS handles)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    z = sampler.sample(den)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    p = sampler.plan
    if not quiet:
        print(f"{table}: {p.K} steps, batch {batch} x {latent} fp16, {'first-order path' if p.markov else f'{p.n_x0_slots} x0 slots'}, "
              f"{p.total_units(2)} tensor transfers per trajectory; {dt * 1e3:.1f} ms per trajectory incl. the random-init MMDiT "
              f"({sum(q.numel() for q in net.parameters()) / 1e6:.0f} M params); finite: {bool(torch.isfinite(z).all())}")
    return z


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--table", default="sd3_step_28_weight.csv", choices=["sd3_step_28_weight.csv", "sd3_step_28_weight_sharp.csv"])
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--small", action="store_true")
    a = ap.parse_args()
    run(a.table, a.batch, a.small)
