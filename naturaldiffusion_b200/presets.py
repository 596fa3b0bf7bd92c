"""One-line constructors for the three reference scripts' configurations (loop-level drop-in).

    sampler, wrap = presets.cifar(weight_path, batch=500)            # src/CIFAR10NaturalInference.py:241-317
    images = sampler.sample(wrap(score_model), pixels_out=buf)

    sampler, wrap = presets.dit("results/ddpm/ddpm_sympy_024.npz", batch=8)   # src/ValidateNaturalInference.py:311-372
    z = sampler.sample(wrap(dit_model, class_labels))

    sampler, wrap = presets.sd3("weights/sd3_step_28_weight.csv", batch=4)    # src/SD3NaturalInference.py:171-245
    x0 = sampler.sample(wrap(transformer, prompt_embeds, pooled, negative_embeds, negative_pooled))
"""
from __future__ import annotations

import torch

from . import adapters
from .coeffs import CoeffTriple, ddim_x0_coeffs, flow_match_sigmas, io_eps_cfg, io_score_vp, io_velocity_cfg
from .sampler import NaturalInferenceSampler


def cifar(weight_path, batch: int, *, device="cuda", seed: int = 888, **kw):
    """VP score model on 3x32x32 (deps/score_sde_pytorch NCSN++); `wrap(model)` evaluates model(x, 999*t)."""
    triple = CoeffTriple.from_npz(weight_path)
    s = NaturalInferenceSampler(triple, io_score_vp(triple.node), batch, (3, 32, 32), device=device, seed=seed, **kw)
    return s, (lambda model, autocast_dtype=None: adapters.ncsnpp_denoiser(model, triple.node, autocast_dtype))


def dit(matrix, batch: int, *, cfg_scale: float = 4.0, latent=(4, 32, 32), device="cuda", seed: int = 0, **kw):
    """DiT + classifier-free guidance on the discrete VP grid; `matrix` is an npz path or a CoeffTriple
    (results/{ddpm,ddim}/*.npz or generators.ddpm_triple / ddim_triple); final latent scaled by 1/0.18215 when
    `vae_scale=True` is passed."""
    triple = matrix if isinstance(matrix, CoeffTriple) else CoeffTriple.from_npz(matrix)
    if kw.pop("vae_scale", False):
        kw.setdefault("final_scale", 1.0 / 0.18215)
    c1, c2, _ = ddim_x0_coeffs(triple.K)
    s = NaturalInferenceSampler(triple, io_eps_cfg(c1, c2, cfg_scale), batch, latent, device=device, seed=seed, **kw)
    return s, (lambda model, class_labels, null_class=1000: adapters.dit_cfg_denoiser(model, triple.node, class_labels, null_class))


def sd3(csv_path, batch: int, *, cfg_scale: float = 7.0, latent=(16, 128, 128), dtype=torch.float16, device="cuda", seed: int = 10,
        sigmas=None, **kw):
    """SD3 flow matching with a csv weight table and the FlowMatchEuler sigma grid (28 steps, shift 3)."""
    from .coeffs import load_weight_csv
    W = load_weight_csv(csv_path)
    sig = flow_match_sigmas(W.shape[0]) if sigmas is None else sigmas
    triple = CoeffTriple.from_sd3_table(W, sig, name=str(csv_path))
    s = NaturalInferenceSampler(triple, io_velocity_cfg(sig, cfg_scale), batch, latent, device=device, dtype=dtype, seed=seed, **kw)
    return s, (lambda model, ctx, pooled, neg_ctx, neg_pooled: adapters.mmdit_cfg_denoiser(model, sig, ctx, pooled, neg_ctx, neg_pooled))
