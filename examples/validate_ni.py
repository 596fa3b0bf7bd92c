#!/usr/bin/env python
"""The reference's consistency check (`compare_output_tx`, src/ValidateNaturalInference.py:375-391) on the B200 path:
the ORIGINAL first-order sampler (DDPM ancestral / DDIM) and Natural Inference with the matching coefficient matrix
produce the same samples from the same noise.

Both sides run on the fused `ni_step` kernel with the same denoiser and the same in-kernel Philox noise:
  original sampler  = the first-order path (`markov=True`): every step reads only the model output(s) and x_k,
                      x_{k+1} = c_k x_k + A[k,k] x0_k + B[k,k+1] eps_{k+1}  -- the sampler's own recurrence;
  Natural Inference = the dense rows of the matrix (`markov=False`): x_{k+1} = sum_j A[k,j] x0_j + sum_j B[k,j] eps_j
                      over the stored history ring.
Offline there are no checkpoints: the denoiser is random-init (last layer re-initialised so it is not the zero map).

  python examples/validate_ni.py                       # config C1: DDIM 10 steps, NCSN++ 61.8 M, batch 64
  python examples/validate_ni.py --alg ddpm --steps 250 --model dit --batch 16 --cfg 4.0      # config C4 shapes
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from naturaldiffusion_b200 import generators  # noqa: E402
from naturaldiffusion_b200.adapters import dit_cfg_denoiser  # noqa: E402
from naturaldiffusion_b200.coeffs import ddim_x0_coeffs, io_eps_cfg  # noqa: E402
from naturaldiffusion_b200.sampler import NaturalInferenceSampler  # noqa: E402


@torch.no_grad()
def run(alg="ddim", steps=10, batch=64, model="ncsnpp", cfg=None, seed=0, small=False, quiet=False):
    dev = torch.device("cuda", 0)
    triple = generators.ddim_triple(steps) if alg == "ddim" else generators.ddpm_triple(steps)
    c1, c2, _ = ddim_x0_coeffs(steps)
    torch.manual_seed(0)
    if model == "ncsnpp":
        from naturaldiffusion_b200.denoisers import NCSNppVP
        net = (NCSNppVP(nf=32, num_res_blocks=1) if small else NCSNppVP()).reinit_output(std=0.02).to(dev).eval()
        shape, cfg = (3, 32, 32), None
        den = lambda z, k: net(z, torch.full((z.shape[0],), float(triple.node[k, 0]), device=dev))  # eps-model, discrete label
    else:
        from naturaldiffusion_b200.denoisers import DiT, dit_xl_2
        net = (DiT(dim=64, depth=2, heads=4) if small else dit_xl_2()).reinit_output(std=0.05).to(dev).eval()
        shape, cfg = (4, 32, 32), (4.0 if cfg is None else cfg)
        labels = torch.arange(batch, device=dev) % 1000
        den = dit_cfg_denoiser(net, triple.node, labels)
    io = io_eps_cfg(c1, c2, cfg)
    out = {}
    for name, markov in (("original", True), ("natural_inference", False)):
        s = NaturalInferenceSampler(triple, io, batch, shape, device=dev, seed=seed, markov=markov)
        out[name] = s.sample(den).clone()
        out[name + "_units"] = s.plan.total_units(1 if cfg is None else 2)
        del s
    a, b = out["original"].double(), out["natural_inference"].double()
    rel = float((a - b).abs().max() / a.norm())
    if not quiet:
        print(f"{alg.upper()} {steps} steps, {model} ({sum(p.numel() for p in net.parameters()) / 1e6:.1f} M params, random init), batch {batch}: "
              f"max|original - NI| / ||x||_2 = {rel:.3e} (tolerance 1e-5); tensor transfers per trajectory: original "
              f"{out['original_units']}, NI dense rows {out['natural_inference_units']}")
    return rel


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--alg", default="ddim", choices=["ddim", "ddpm"])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--model", default="ncsnpp", choices=["ncsnpp", "dit"])
    ap.add_argument("--cfg", type=float, default=None)
    ap.add_argument("--small", action="store_true", help="small instance of the same architecture (quick check)")
    a = ap.parse_args()
    rel = run(a.alg, a.steps, a.batch, a.model, a.cfg, small=a.small)
    sys.exit(0 if rel <= 1e-5 else 1)
