"""GPU: FID sufficient statistics (csrc/ni_fid.cu -- fp64 tensor-core rank-k update) against numpy in fp64,
the reference's own arithmetic for this step (np.mean / np.cov, src/CIFAR10NaturalInference.py:73-86)."""
import numpy as np
import pytest
import torch

from naturaldiffusion_b200.fid import FidAccumulator, accumulate_images, frechet_distance

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = "cuda:0"


@pytest.mark.parametrize("m,d", [(1000, 256), (777, 2048), (33, 100), (1, 64), (4096, 192), (50, 4)])
def test_statistics_match_numpy_fp64(m, d):
    g = torch.Generator().manual_seed(m + d)
    x = (torch.randn(m, d, generator=g) * 3 + 0.5).to(DEV)
    acc = FidAccumulator(dim=d, device=DEV).update(x)
    xd = x.double().cpu().numpy()
    buf = acc.buf.cpu().numpy()
    assert buf[0] == m
    assert np.allclose(buf[1:1 + d], xd.sum(0), rtol=1e-12, atol=1e-9)
    S = buf[1 + d:].reshape(d, d)
    ref = xd.T @ xd
    assert np.abs(S - ref).max() <= 1e-12 * np.abs(ref).max()
    assert np.array_equal(S, S.T)  # mirrored tiles are the same numbers


def test_accumulation_over_batches_strided_input_and_finalize():
    d = 128
    g = torch.Generator().manual_seed(3)
    big = torch.randn(900, d + 32, generator=g).to(DEV)
    parts = [big[:400, :d], big[400:, :d]]  # row pitch 160 != d: read in place
    acc = FidAccumulator(dim=d, device=DEV)
    for p in parts:
        acc.update(p)
    acc.update(big[:0, :d])  # an empty batch changes nothing
    mu, sigma = acc.finalize()
    xd = big[:, :d].double().cpu().numpy()
    assert np.allclose(mu, xd.mean(0), rtol=1e-10, atol=1e-12) and np.allclose(sigma, np.cov(xd, rowvar=False), rtol=1e-9, atol=1e-11)
    # deterministic: same inputs, same bits
    acc2 = FidAccumulator(dim=d, device=DEV)
    for p in parts:
        acc2.update(p)
    assert torch.equal(acc.buf[1 + d:], acc2.buf[1 + d:])
    assert abs(frechet_distance(mu, sigma, mu, sigma)) < 1e-6
    # half-precision activations are widened by torch first, then take the same kernel
    acc3 = FidAccumulator(dim=d, device=DEV).update(big[:, :d].half())
    assert np.allclose(acc3.finalize()[0], big[:, :d].half().double().cpu().numpy().mean(0), rtol=1e-10, atol=1e-12)


def test_image_to_statistics_flow():
    """uint8 NHWC images (what the fused output stage emits) -> features -> statistics, no host round trip"""
    g = torch.Generator().manual_seed(0)
    imgs = torch.randint(0, 256, (300, 32, 32, 3), generator=g, dtype=torch.uint8).to(DEV)
    P = torch.randn(3, 64, generator=g).to(DEV)
    feat = lambda b: torch.einsum("bchw,cf->bf", b, P) / 1024.0
    acc = accumulate_images(FidAccumulator(dim=64, device=DEV), imgs, feat, batch_size=128)
    ref = feat(imgs.float().div(255).permute(0, 3, 1, 2)).double().cpu().numpy()
    mu, sigma = acc.finalize()
    assert acc.n == 300 and np.allclose(mu, ref.mean(0), rtol=1e-9) and np.allclose(sigma, np.cov(ref, rowvar=False), rtol=1e-7, atol=1e-12)


def test_fid_accumulate_rejects_bad_arguments():
    import naturaldiffusion_b200 as ni
    from naturaldiffusion_b200 import _lib
    L = _lib.lib()
    x = torch.randn(16, 64, device=DEV)
    buf = torch.zeros(1 + 64 + 64 * 64, dtype=torch.float64, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    assert L.ni_fid_accumulate(x.data_ptr(), 16, 64, 32, buf.data_ptr(), st) == -1       # ld < d
    assert L.ni_fid_accumulate(x.data_ptr(), 16, 62, 64, buf.data_ptr(), st) == -1       # d not a multiple of 4
    assert L.ni_fid_accumulate(x.data_ptr() + 4, 16, 64, 64, buf.data_ptr(), st) == -1   # misaligned activations
    assert L.ni_fid_accumulate(None, 16, 64, 64, buf.data_ptr(), st) == -1 and b"ni_fid_accumulate" in L.ni_last_error()
    assert L.ni_fid_accumulate(x.data_ptr(), 0, 64, 64, buf.data_ptr(), st) == 0 and float(buf.abs().sum()) == 0.0
    with pytest.raises(ValueError):
        FidAccumulator(dim=64, device=DEV).update(torch.randn(4, 32, device=DEV))
