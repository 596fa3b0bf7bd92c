"""Small deterministic stand-ins for the denoisers (test scaffolding shared by the golden-vector
generator and the tests).  Pure torch, same code on CPU and CUDA; non-linear so that an error in
any step propagates to the next denoiser call."""
from __future__ import annotations

import torch


def _mix(channels: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(channels, channels, generator=g) * (1.0 / channels ** 0.5)


class ToyEps:
    """eps/score/velocity-like net: h(x, t) = tanh(W x) * (1 + t_scale * t) + 0.1 * x, applied per pixel."""

    def __init__(self, channels: int, seed: int = 7, t_scale: float = 1e-3, out_channels=None):
        self.W = _mix(channels, seed)
        self.W2 = _mix(channels, seed + 1)
        self.t_scale = t_scale
        self.channels = channels
        self.out_channels = out_channels or channels

    def __call__(self, x: torch.Tensor, t, variant: int = 0) -> torch.Tensor:
        W = (self.W if variant == 0 else self.W2).to(device=x.device, dtype=torch.float32)
        xf = x.to(torch.float32)
        if not torch.is_tensor(t):
            t = torch.full((x.shape[0],), float(t), device=x.device)
        h = torch.tanh(torch.einsum("oc,bchw->bohw", W, xf)) * (1.0 + self.t_scale * t.to(torch.float32).view(-1, 1, 1, 1)) + 0.1 * xf
        h = h.to(x.dtype)
        if self.out_channels != self.channels:  # DiT-like: extra (ignored) channels after the eps channels
            pad = torch.zeros(x.shape[0], self.out_channels - self.channels, *x.shape[2:], device=x.device, dtype=x.dtype) + 3.0
            h = torch.cat([h, pad], dim=1)
        return h.contiguous()


class ToyVPDenoiser:
    """A bounded, non-linear eps-predictor for the discrete VP schedule (betas linspace(1e-4, 0.02, 1000)):
    eps(z, t) = sqrt(1 - abar_t) * z + 0.1 * tanh(W z).  The linear part is the optimal predictor for unit-variance
    data, so long trajectories (DDPM-250) stay O(1) like a trained net's; the tanh part makes errors propagate."""

    def __init__(self, channels: int, seed: int = 23, out_channels=None):
        self.W = [_mix(channels, seed), _mix(channels, seed + 1)]
        betas = torch.linspace(0.0001, 0.02, 1000, dtype=torch.float64)
        self.sigma = torch.sqrt(1.0 - torch.cumprod(1.0 - betas, 0)).to(torch.float32)
        self.channels = channels
        self.out_channels = out_channels or channels

    def __call__(self, z: torch.Tensor, t: int, variant: int = 0) -> torch.Tensor:
        W = self.W[variant].to(z.device)
        h = float(self.sigma[max(int(t), 0)]) * z + 0.1 * torch.tanh(torch.einsum("oc,bchw->bohw", W, z))
        if self.out_channels != self.channels:
            h = torch.cat([h, torch.full((z.shape[0], self.out_channels - self.channels, *z.shape[2:]), 3.0, device=z.device)], dim=1)
        return h.contiguous()
