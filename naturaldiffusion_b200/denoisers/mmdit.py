"""SD3-medium-shaped MMDiT stand-in (random init).  The reference drives
`StableDiffusion3Pipeline.transformer` from diffusers (src/SD3NaturalInference.py:175-176, 210-213), which is not in
the reference tree and not installed offline; this module reproduces its SHAPE CONTRACT from the public model card so
the SD3 loop can be driven end to end: 24 joint blocks, 24 heads x 64 = hidden 1536, patch 2 on a 16-channel latent
(128x128 -> 4096 image tokens), text context [B, 333, 4096] projected to 1536, pooled text [B, 2048] added to the
timestep embedding, joint attention over text+image tokens with separate weights per stream, adaLN modulation.
I/O: forward(x[B,16,H,W], timestep[B] = 1000*sigma, context[B,L,4096], pooled[B,2048]) -> v[B,16,H,W]."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _mod(x, shift, scale):
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


class _Stream(nn.Module):
    """per-modality weights of one joint block"""

    def __init__(self, dim, last=False):
        super().__init__()
        self.last = last
        self.ada = nn.Linear(dim, (2 if last else 6) * dim)
        self.n1 = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.qkv = nn.Linear(dim, 3 * dim)
        if not last:
            self.proj = nn.Linear(dim, dim)
            self.n2 = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
            self.fc1, self.fc2 = nn.Linear(dim, 4 * dim), nn.Linear(4 * dim, dim)


class _JointBlock(nn.Module):
    def __init__(self, dim, heads, context_last=False):
        super().__init__()
        self.heads = heads
        self.x, self.c = _Stream(dim), _Stream(dim, last=context_last)

    def forward(self, x, c, vec):
        B, Nx, D = x.shape
        mx = self.x.ada(F.silu(vec)).chunk(6, dim=1)
        mc = self.c.ada(F.silu(vec)).chunk(2 if self.c.last else 6, dim=1)
        qkv = torch.cat([self.c.qkv(_mod(self.c.n1(c), mc[0], mc[1])), self.x.qkv(_mod(self.x.n1(x), mx[0], mx[1]))], dim=1)
        q, k, v = qkv.view(B, -1, 3, self.heads, D // self.heads).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, -1, D)
        ac, ax = a[:, : c.shape[1]], a[:, c.shape[1]:]
        x = x + mx[2].unsqueeze(1) * self.x.proj(ax)
        x = x + mx[5].unsqueeze(1) * self.x.fc2(F.gelu(self.x.fc1(_mod(self.x.n2(x), mx[3], mx[4])), approximate="tanh"))
        if not self.c.last:
            c = c + mc[2].unsqueeze(1) * self.c.proj(ac)
            c = c + mc[5].unsqueeze(1) * self.c.fc2(F.gelu(self.c.fc1(_mod(self.c.n2(c), mc[3], mc[4])), approximate="tanh"))
        return x, c


class MMDiT(nn.Module):
    def __init__(self, in_ch=16, patch=2, dim=1536, depth=24, heads=24, ctx_dim=4096, pooled_dim=2048, max_grid=192):
        super().__init__()
        self.patch, self.in_ch, self.dim, self.max_grid = patch, in_ch, dim, max_grid
        self.embed = nn.Conv2d(in_ch, dim, patch, stride=patch)
        self.pos = nn.Parameter(torch.randn(1, max_grid * max_grid, dim) * 0.02)
        self.t_mlp = nn.Sequential(nn.Linear(256, dim), nn.SiLU(), nn.Linear(dim, dim))
        self.y_mlp = nn.Sequential(nn.Linear(pooled_dim, dim), nn.SiLU(), nn.Linear(dim, dim))
        self.ctx = nn.Linear(ctx_dim, dim)
        self.blocks = nn.ModuleList(_JointBlock(dim, heads, context_last=(i == depth - 1)) for i in range(depth))
        self.n_out = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.ada_out = nn.Linear(dim, 2 * dim)
        self.lin_out = nn.Linear(dim, patch * patch * in_ch)

    def _pos(self, gh, gw):  # centre crop of the learned table, as SD3 does
        g = self.max_grid
        top, left = (g - gh) // 2, (g - gw) // 2
        return self.pos.view(1, g, g, self.dim)[:, top:top + gh, left:left + gw].reshape(1, gh * gw, self.dim)

    def forward(self, x, timestep, context, pooled):
        B, C, H, W = x.shape
        gh, gw = H // self.patch, W // self.patch
        dt = self.embed.weight.dtype
        half = 128
        freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=x.device) / half)
        a = timestep.float()[:, None] * freqs[None]
        vec = self.t_mlp(torch.cat([a.cos(), a.sin()], dim=-1).to(dt)) + self.y_mlp(pooled.to(dt))
        h = self.embed(x.to(dt)).flatten(2).transpose(1, 2) + self._pos(gh, gw).to(dt)
        c = self.ctx(context.to(dt))
        for blk in self.blocks:
            h, c = blk(h, c, vec)
        sh, sc = self.ada_out(F.silu(vec)).chunk(2, dim=1)
        h = self.lin_out(_mod(self.n_out(h), sh, sc))
        p = self.patch
        return h.view(B, gh, gw, p, p, C).permute(0, 5, 1, 3, 2, 4).reshape(B, C, H, W)


def mmdit_sd3_medium(**kw) -> MMDiT:
    return MMDiT(in_ch=16, patch=2, dim=1536, depth=24, heads=24, **kw)
