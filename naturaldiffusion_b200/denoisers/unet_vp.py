"""NCSN++ for the VP SDE at the reference's CIFAR-10 configuration
(deps/score_sde_pytorch/configs/vp/cifar10_ddpmpp_continuous.py:41-64 with models/ncsnpp.py:35-381):
nf 128, ch_mult (1,2,2,2), 4 residual blocks per level, self-attention at 16x16, BigGAN residual blocks with
average-pool / nearest-neighbour resampling (fir=False), sinusoidal time embedding, skip rescale 1/sqrt(2),
GroupNorm(min(C/4,32)), swish, no progressive paths.  Written from that description in plain modern torch
(scaled_dot_product_attention, avg_pool2d, interpolate); 61.8 M parameters like the reference model.
I/O contract: forward(x[B,3,32,32], labels[B] = 999*t) -> h[B,3,32,32] (the eps-prediction the score wrapper
divides by -std(t), models/utils.py:150-159)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

_RSQRT2 = 1.0 / math.sqrt(2.0)


def _gn(ch: int) -> nn.GroupNorm:
    return nn.GroupNorm(min(ch // 4, 32), ch, eps=1e-6)


def _scaled_init_(w: torch.Tensor, scale: float):
    """variance-scaling (fan_avg, uniform), scale 0 -> 1e-10 as in the reference's default_init"""
    scale = 1e-10 if scale == 0 else scale
    fan_in = w[0].numel()
    fan_out = w.shape[0] * (w[0][0].numel() if w.dim() > 2 else 1)
    bound = math.sqrt(3.0 * scale / ((fan_in + fan_out) / 2.0))
    with torch.no_grad():
        w.uniform_(-bound, bound)


class _Conv(nn.Conv2d):
    def __init__(self, cin, cout, k, init_scale=1.0):
        super().__init__(cin, cout, k, padding=k // 2)
        _scaled_init_(self.weight, init_scale)
        nn.init.zeros_(self.bias)


class ResBlock(nn.Module):
    def __init__(self, cin, cout=None, temb_dim=512, up=False, down=False, dropout=0.1):
        super().__init__()
        cout = cout or cin
        self.up, self.down = up, down
        self.norm0, self.conv0 = _gn(cin), _Conv(cin, cout, 3)
        self.temb = nn.Linear(temb_dim, cout)
        _scaled_init_(self.temb.weight, 1.0)
        nn.init.zeros_(self.temb.bias)
        self.norm1, self.drop, self.conv1 = _gn(cout), nn.Dropout(dropout), _Conv(cout, cout, 3, init_scale=0.0)
        self.skip = _Conv(cin, cout, 1) if (cin != cout or up or down) else None

    def forward(self, x, temb):
        h = F.silu(self.norm0(x))
        if self.up:
            h, x = F.interpolate(h, scale_factor=2, mode="nearest"), F.interpolate(x, scale_factor=2, mode="nearest")
        elif self.down:
            h, x = F.avg_pool2d(h, 2), F.avg_pool2d(x, 2)
        h = self.conv0(h) + self.temb(F.silu(temb))[:, :, None, None]
        h = self.conv1(self.drop(F.silu(self.norm1(h))))
        if self.skip is not None:
            x = self.skip(x)
        return (x + h) * _RSQRT2


class AttnBlock(nn.Module):
    """single-head self-attention over the H*W positions, C-dimensional keys"""

    def __init__(self, ch):
        super().__init__()
        self.norm = _gn(ch)
        self.q, self.k, self.v = (nn.Linear(ch, ch) for _ in range(3))
        self.o = nn.Linear(ch, ch)
        for lin, sc in ((self.q, 0.1), (self.k, 0.1), (self.v, 0.1), (self.o, 0.0)):
            _scaled_init_(lin.weight, sc)
            nn.init.zeros_(lin.bias)

    def forward(self, x):
        B, C, H, W = x.shape
        t = self.norm(x).flatten(2).transpose(1, 2)  # [B, HW, C]
        a = F.scaled_dot_product_attention(self.q(t).unsqueeze(1), self.k(t).unsqueeze(1), self.v(t).unsqueeze(1)).squeeze(1)
        h = self.o(a).transpose(1, 2).reshape(B, C, H, W)
        return (x + h) * _RSQRT2


class NCSNppVP(nn.Module):
    def __init__(self, channels=3, image_size=32, nf=128, ch_mult=(1, 2, 2, 2), num_res_blocks=4, attn_resolutions=(16,), dropout=0.1):
        super().__init__()
        self.nf = nf
        td = nf * 4
        self.t0, self.t1 = nn.Linear(nf, td), nn.Linear(td, td)
        for lin in (self.t0, self.t1):
            _scaled_init_(lin.weight, 1.0)
            nn.init.zeros_(lin.bias)
        self.conv_in = _Conv(channels, nf, 3)
        res = image_size
        self.down = nn.ModuleList()
        skip_ch, cin = [nf], nf
        for lvl, mult in enumerate(ch_mult):
            for _ in range(num_res_blocks):
                blk = nn.ModuleList([ResBlock(cin, nf * mult, td, dropout=dropout)])
                cin = nf * mult
                if res in attn_resolutions:
                    blk.append(AttnBlock(cin))
                self.down.append(blk)
                skip_ch.append(cin)
            if lvl != len(ch_mult) - 1:
                self.down.append(nn.ModuleList([ResBlock(cin, temb_dim=td, down=True, dropout=dropout)]))
                skip_ch.append(cin)
                res //= 2
        self.mid = nn.ModuleList([ResBlock(cin, temb_dim=td, dropout=dropout), AttnBlock(cin), ResBlock(cin, temb_dim=td, dropout=dropout)])
        self.up = nn.ModuleList()
        for lvl, mult in reversed(list(enumerate(ch_mult))):
            for _ in range(num_res_blocks + 1):
                self.up.append(nn.ModuleList([ResBlock(cin + skip_ch.pop(), nf * mult, td, dropout=dropout)]))
                cin = nf * mult
            if res in attn_resolutions:
                self.up[-1].append(AttnBlock(cin))
            if lvl != 0:
                self.up.append(nn.ModuleList([ResBlock(cin, temb_dim=td, up=True, dropout=dropout)]))
                res *= 2
        assert not skip_ch
        self._n_skip_consumers = [isinstance(b[0], ResBlock) and not b[0].up for b in self.up]
        self.norm_out, self.conv_out = _gn(cin), _Conv(cin, channels, 3, init_scale=0.0)

    def reinit_output(self, std=0.02, seed=0):
        """random-init nets end in a ~zero conv (init_scale=0): give the last layer real weights so that a
        trajectory actually depends on the network"""
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            self.conv_out.weight.copy_(torch.randn(self.conv_out.weight.shape, generator=g) * std)
        return self

    def time_embedding(self, labels):
        half = self.nf // 2
        freq = torch.exp(torch.arange(half, dtype=torch.float32, device=labels.device) * (-math.log(10000.0) / (half - 1)))
        ang = labels.float()[:, None] * freq[None, :]
        return torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)

    def forward(self, x, labels):
        temb = self.t1(F.silu(self.t0(self.time_embedding(labels))))
        hs = [self.conv_in(x)]
        for blk in self.down:
            h = blk[0](hs[-1], temb)
            for extra in blk[1:]:
                h = extra(h)
            hs.append(h)
        h = hs[-1]
        h = self.mid[2](self.mid[1](self.mid[0](h, temb)), temb)
        for blk, takes_skip in zip(self.up, self._n_skip_consumers):
            h = blk[0](torch.cat([h, hs.pop()], dim=1) if takes_skip else h, temb)
            for extra in blk[1:]:
                h = extra(h)
        assert not hs
        return self.conv_out(F.silu(self.norm_out(h)))
