"""DiT (Peebles & Xie) at the XL/2 size the reference instantiates
(src/ValidateNaturalInference.py:336 -> deps/DiT/models.py:149-253,333-334): depth 28, hidden 1152, 16 heads,
patch 2 on a 4x32x32 latent, class-conditional (1000 classes + 1 null row), learn_sigma -> 8 output channels of
which the path reads [:4].  adaLN-Zero blocks.  Written from the paper's description with stock torch modules
(no timm); ~675 M parameters.  I/O: forward(z[B,4,32,32], t[B], y[B]) -> [B,8,32,32]."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _modulate(x, shift, scale):
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


class _Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, D = x.shape
        q, k, v = self.qkv(x).view(B, N, 3, self.heads, D // self.heads).permute(2, 0, 3, 1, 4)
        return self.proj(F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, D))


class _Block(nn.Module):
    def __init__(self, dim, heads, mlp_ratio=4.0):
        super().__init__()
        self.n1 = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.attn = _Attention(dim, heads)
        self.n2 = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        hid = int(dim * mlp_ratio)
        self.fc1, self.fc2 = nn.Linear(dim, hid), nn.Linear(hid, dim)
        self.ada = nn.Linear(dim, 6 * dim)

    def forward(self, x, c):
        s1, sc1, g1, s2, sc2, g2 = self.ada(F.silu(c)).chunk(6, dim=1)
        x = x + g1.unsqueeze(1) * self.attn(_modulate(self.n1(x), s1, sc1))
        return x + g2.unsqueeze(1) * self.fc2(F.gelu(self.fc1(_modulate(self.n2(x), s2, sc2)), approximate="tanh"))


class DiT(nn.Module):
    def __init__(self, input_size=32, patch=2, in_ch=4, dim=1152, depth=28, heads=16, num_classes=1000, learn_sigma=True, freq_dim=256):
        super().__init__()
        self.patch, self.in_ch, self.out_ch, self.freq_dim = patch, in_ch, in_ch * (2 if learn_sigma else 1), freq_dim
        self.grid = input_size // patch
        self.embed = nn.Conv2d(in_ch, dim, patch, stride=patch)
        self.t_mlp = nn.Sequential(nn.Linear(freq_dim, dim), nn.SiLU(), nn.Linear(dim, dim))
        self.y_embed = nn.Embedding(num_classes + 1, dim)  # last row = the null class used for CFG
        self.register_buffer("pos", self._sincos_2d(dim, self.grid), persistent=False)
        self.blocks = nn.ModuleList(_Block(dim, heads) for _ in range(depth))
        self.n_out = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.ada_out = nn.Linear(dim, 2 * dim)
        self.lin_out = nn.Linear(dim, patch * patch * self.out_ch)
        for blk in self.blocks:  # adaLN-Zero
            nn.init.zeros_(blk.ada.weight); nn.init.zeros_(blk.ada.bias)
        for lin in (self.ada_out, self.lin_out):
            nn.init.zeros_(lin.weight); nn.init.zeros_(lin.bias)

    def reinit_output(self, std=0.02, seed=0):
        """the published init zeroes the final layer and all adaLN gates -> a random-init DiT outputs exactly 0;
        re-initialise them so the trajectory depends on the network"""
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            self.lin_out.weight.copy_(torch.randn(self.lin_out.weight.shape, generator=g) * std)
            for blk in self.blocks:
                blk.ada.weight.copy_(torch.randn(blk.ada.weight.shape, generator=g) * std)
        return self

    @staticmethod
    def _sincos_2d(dim, grid):
        def one(pos, d):
            om = 1.0 / 10000 ** (torch.arange(d // 2, dtype=torch.float64) / (d / 2))
            out = pos.reshape(-1).double()[:, None] * om[None]
            return torch.cat([out.sin(), out.cos()], dim=1)
        gh, gw = torch.meshgrid(torch.arange(grid), torch.arange(grid), indexing="ij")
        return torch.cat([one(gw, dim // 2), one(gh, dim // 2)], dim=1).float().unsqueeze(0)

    def t_embed(self, t):
        half = self.freq_dim // 2
        freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
        a = t.float()[:, None] * freqs[None]
        return self.t_mlp(torch.cat([a.cos(), a.sin()], dim=-1).to(self.t_mlp[0].weight.dtype))

    def forward(self, z, t, y):
        B = z.shape[0]
        x = self.embed(z).flatten(2).transpose(1, 2) + self.pos.to(z.dtype)
        c = self.t_embed(t) + self.y_embed(y)
        for blk in self.blocks:
            x = blk(x, c)
        sh, sc = self.ada_out(F.silu(c)).chunk(2, dim=1)
        x = self.lin_out(_modulate(self.n_out(x), sh, sc))
        p, g, o = self.patch, self.grid, self.out_ch
        return x.view(B, g, g, p, p, o).permute(0, 5, 1, 3, 2, 4).reshape(B, o, g * p, g * p)


def dit_xl_2(**kw) -> DiT:
    return DiT(dim=1152, depth=28, heads=16, patch=2, **kw)
