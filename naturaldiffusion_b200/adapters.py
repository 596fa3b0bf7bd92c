"""Denoiser adapters: wrap the three model interfaces the reference drives into the sampler's
``denoiser(x, k) -> out | (out0, out1)`` protocol.  Pure torch glue (the denoiser forward stays torch)."""
from __future__ import annotations

from typing import Optional, Sequence

import torch


def ncsnpp_denoiser(model, node, autocast_dtype: Optional[torch.dtype] = None):
    """CIFAR loop: ``score_fn(x, vec_t)`` evaluates ``model(x, labels = 999 * t)``
    (deps/score_sde_pytorch/models/utils.py:150-151); the -1/std(t) and (sigma^2, 1/alpha) scalings are folded into
    the fused step by ``coeffs.io_score_vp``.  node = triple.node (column 0 = t)."""
    ts = [float(t) for t in node[:, 0]]
    cache = {}

    def den(x, k):
        key = (k, x.shape[0], x.device)
        if key not in cache:
            cache[key] = torch.full((x.shape[0],), ts[k], device=x.device, dtype=torch.float32) * 999
        if autocast_dtype is not None:
            with torch.autocast("cuda", dtype=autocast_dtype):
                return model(x, cache[key])
        return model(x, cache[key])

    return den


def dit_cfg_denoiser(model, node, class_labels: torch.Tensor, null_class: int = 1000, batched: bool = True):
    """Validate loop: ``forward_cfg`` (src/ValidateNaturalInference.py:185-195) runs the model on the class labels and
    on the null class; timestep = int(node[k,0]) (:350).  Returns the two full 8-channel outputs -- the fused step
    reads channels [:4] through its sample stride and applies the CFG mix.  batched=True evaluates both in one
    forward of 2B samples (same values, one launch sequence); the halves are contiguous views."""
    steps = [int(t) for t in node[:, 0]]

    def den(z, k):
        B = z.shape[0]
        t = torch.full((B,), steps[k], dtype=torch.int32, device=z.device)
        nul = torch.full_like(class_labels, null_class)
        if batched:
            out = model(torch.cat([z, z]), torch.cat([t, t]), torch.cat([class_labels, nul]))
            return out[:B], out[B:]
        return model(z, t, class_labels), model(z, t, nul)

    return den


def mmdit_cfg_denoiser(model, sigmas: Sequence[float], context, pooled, neg_context, neg_pooled, batched: bool = True):
    """SD3 loop: ``pipe.transformer(hidden_states, timestep = 1000*sigma, encoder_hidden_states, pooled_projections)``
    on the prompt and on the negative prompt (src/SD3NaturalInference.py:210-213).  Returns (v_text, v_null)."""
    ts = [1000.0 * float(s) for s in sigmas]

    def den(x, k):
        B = x.shape[0]
        t = torch.full((B,), ts[k], device=x.device, dtype=torch.float32)
        if batched:
            out = model(torch.cat([x, x]), torch.cat([t, t]), torch.cat([context, neg_context]), torch.cat([pooled, neg_pooled]))
            return out[:B], out[B:]
        return model(x, t, context, pooled), model(x, t, neg_context, neg_pooled)

    return den
