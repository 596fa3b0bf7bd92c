"""FID statistics across ranks: the one collective of the north-star.

The reference computes ``mu = np.mean(act, 0)``, ``sigma = np.cov(act, rowvar=False)`` over all 50 000 Inception
pool3 activations on one host (src/CIFAR10NaturalInference.py:73-86) and calls pytorch_fid's
``calculate_frechet_distance``.  With sampling sharded by batch, each rank accumulates the sufficient statistics
(n, sum x, sum x x^T) of its own activations on its own GPU in fp64 -- a 2048-wide rank-k update, a plain library
GEMM -- and ONE all-reduce (NCCL over NVLink on GPU tensors, gloo on CPU tensors) merges them; mean/covariance and the
Frechet distance are then a host-side O(d^3) step exactly as in the reference.  34 MB per evaluation, not per step.
"""
from __future__ import annotations

import numpy as np
import torch


class FidAccumulator:
    def __init__(self, dim: int = 2048, device="cuda", host_logic_only: bool = False):
        """device: a CUDA device -- the statistics live on that GPU and `update` is the fp64 tensor-core kernel.  There is no
        silent CPU path: a CPU device is refused unless `host_logic_only=True` is passed explicitly, which exists for the
        world-size-2 gloo tests of the merge / finalize logic (tests/test_multirank_cpu.py) and computes with plain torch."""
        self.dim = dim
        self.device = torch.device(device)
        if self.device.type != "cuda" and not host_logic_only:
            from ._lib import NiError
            raise NiError("FidAccumulator needs a CUDA device; there is no CPU fallback (host_logic_only=True is for the gloo tests)")
        # one flat fp64 buffer so the merge is a single all-reduce: [n | sum x (d) | sum x x^T (d*d)]
        self.buf = torch.zeros(1 + dim + dim * dim, dtype=torch.float64, device=self.device)

    @property
    def n(self) -> float:
        """sample count so far (reads the device buffer: synchronises; call it after the loop, not inside it)"""
        return float(self.buf[0])

    @torch.no_grad()
    def update(self, feats: torch.Tensor):
        """feats: [m, dim] activations of this rank's samples.  On a CUDA accumulator this is ONE call of
        `ni_fid_accumulate` (csrc/ni_fid.cu: fp32 activations widened in registers, fp64 tensor-core rank-k update, column
        sums, count) on the current stream -- no fp64 copy of the activations, no sync.  A CPU accumulator (the gloo
        host-logic tests) uses plain torch."""
        if feats.dim() != 2 or feats.shape[1] != self.dim:
            raise ValueError(f"expected [m, {self.dim}] features, got {tuple(feats.shape)}")
        if self.device.type == "cuda":
            from . import _lib
            f = feats.to(self.device, torch.float32)
            if f.stride(1) != 1 or f.stride(0) % 4 != 0 or f.data_ptr() % 16 != 0:
                f = f.contiguous()
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().ni_fid_accumulate(f.data_ptr(), f.shape[0], self.dim, f.stride(0) if f.shape[0] > 1 else self.dim,
                                                        self.buf.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream), "ni_fid_accumulate")
            return self
        f = feats.to(self.device, torch.float64)
        self.buf[0] += f.shape[0]
        self.buf[1:1 + self.dim] += f.sum(0)
        self.buf[1 + self.dim:].view(self.dim, self.dim).addmm_(f.t(), f)
        return self

    def all_reduce(self, group=None):
        """sum the statistics over all ranks (no-op without an initialised process group)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=group)
        return self

    def finalize(self):
        """(mu, sigma) as float64 numpy; sigma is the unbiased covariance, like np.cov(act, rowvar=False)."""
        n = self.buf[0]
        s = self.buf[1:1 + self.dim]
        ss = self.buf[1 + self.dim:].view(self.dim, self.dim)
        mu = s / n
        sigma = (ss - torch.outer(s, s) / n) / (n - 1)
        return mu.cpu().numpy(), sigma.cpu().numpy()


def frechet_distance(mu1, sigma1, mu2, sigma2, eps: float = 1e-6) -> float:
    """d^2 = |mu1-mu2|^2 + Tr(S1 + S2 - 2 sqrt(S1 S2)).  pytorch_fid's `calculate_frechet_distance` (third party,
    absent offline; published algorithm restated): scipy sqrtm of the product, eps*I regularisation if the product is
    near-singular, imaginary round-off discarded."""
    from scipy import linalg
    mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
    sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
    diff = mu1 - mu2
    covmean = linalg.sqrtm(sigma1.dot(sigma2))  # (scipy >= 1.16 dropped the `disp` flag pytorch_fid passes)
    if not np.isfinite(covmean).all():
        off = np.eye(sigma1.shape[0]) * eps
        covmean = linalg.sqrtm((sigma1 + off).dot(sigma2 + off))
    if np.iscomplexobj(covmean):
        if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
            raise ValueError("imaginary component in sqrtm: %g" % np.max(np.abs(covmean.imag)))
        covmean = covmean.real
    return float(diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean))


@torch.no_grad()
def accumulate_images(acc: FidAccumulator, images_u8_nhwc: torch.Tensor, feature_fn, batch_size: int = 500):
    """Reference flow of `get_activation` (src/CIFAR10NaturalInference.py:44-70) without the host round trip:
    uint8 NHWC images (as `ni_to_pixel_u8` emits them) -> float NCHW in [0,1] -> feature_fn -> [m, dim] -> statistics."""
    for i in range(0, images_u8_nhwc.shape[0], batch_size):
        b = images_u8_nhwc[i:i + batch_size].to(acc.device, torch.float32).div_(255).permute(0, 3, 1, 2)
        f = feature_fn(b)
        if f.dim() == 4:
            f = f.mean(dim=(2, 3))
        acc.update(f)
    return acc


def inception_pool3_standin(device="cuda", dtype=torch.float32):
    """Inception-shaped feature extractor for offline runs: torchvision's InceptionV3 architecture (what pytorch_fid's
    `InceptionV3([3])` wraps, src/CIFAR10NaturalInference.py:52-70) with RANDOM weights -- the FID checkpoint cannot be
    downloaded here -- returning the 2048-d pool3 activations.  Input: float NCHW in [0, 1]; resized to 299x299 (bilinear)
    and scaled to [-1, 1] like pytorch_fid's resize_input / normalize_input.  Scaffolding: it gives the evaluation flow its
    real shapes and cost (5.7 GFLOP per image), not meaningful FID values."""
    from torchvision.models import inception_v3
    torch.manual_seed(0)
    net = inception_v3(weights=None, aux_logits=False, init_weights=False)
    net.fc = torch.nn.Identity()
    net = net.to(device=device, dtype=dtype).eval()

    @torch.no_grad()
    def features(x):
        x = torch.nn.functional.interpolate(x.to(dtype), size=(299, 299), mode="bilinear", align_corners=False)
        return net(2.0 * x - 1.0).float()

    return features
