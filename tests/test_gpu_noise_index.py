"""GPU tests of the noise index (round 2): every trajectory that draws noise in-kernel moves a DEVICE counter forward,
so successive calls / batches / graph replays get new noise like the reference's torch.randn / randn_like
(src/ValidateNaturalInference.py:345,359), while any sharding of a run still draws the single-GPU tensors."""
import os

import numpy as np
import pytest
import torch

import naturaldiffusion_b200 as ni
from naturaldiffusion_b200 import generators
from naturaldiffusion_b200.coeffs import CoeffTriple, ddim_x0_coeffs, io_eps_cfg, io_score_vp
from naturaldiffusion_b200.ops import philox_normal, to_pixel_u8
from naturaldiffusion_b200.sampler import NaturalInferenceSampler
from oracle import philox

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = "cuda:0"
den = lambda x, k: torch.tanh(0.7 * x) * (1.0 + 0.01 * k) + 0.1 * x


def _ddpm(batch, K=12, **kw):
    triple = generators.ddpm_triple(K)
    c1, c2, _ = ddim_x0_coeffs(K)
    return NaturalInferenceSampler(triple, io_eps_cfg(c1, c2, None), batch, (4, 16, 16), device=DEV, seed=7, **kw)


def test_two_batches_draw_different_fresh_noise_even_with_supplied_initial_noise():
    """with the initial noise supplied by the caller the per-step fresh noise eps_1..K must still differ from
    batch to batch (DDPM: B[k,k+1] != 0)"""
    s = _ddpm(8)
    noise = torch.randn(8, 4, 16, 16, device=DEV)
    a = s.sample(den, noise=noise).clone()
    b = s.sample(den, noise=noise).clone()
    assert not torch.equal(a, b)
    s.set_sample_offset(0)
    assert torch.equal(s.sample(den, noise=noise), a)
    assert torch.equal(s.sample(den, noise=noise), b)
    # and without supplied noise
    c = s.sample(den).clone()
    assert not torch.equal(s.sample(den), c)


def test_batches_of_one_sampler_equal_one_big_run():
    """three calls of a batch-8 sampler == one batch-24 sampler (global sample index), stochastic matrix"""
    small, big = _ddpm(8), _ddpm(24)
    parts = torch.cat([small.sample(den).clone() for _ in range(3)])
    assert torch.equal(parts, big.sample(den))
    # two ranks interleaving batches: advance = G*B
    r0, r1 = _ddpm(4, sample_offset=0, advance=8), _ddpm(4, sample_offset=4, advance=8)
    inter = []
    for _ in range(3):
        inter += [r0.sample(den).clone(), r1.sample(den).clone()]
    assert torch.equal(torch.cat(inter), big.sample(den) if False else _ddpm(24).sample(den))


def test_graph_replays_draw_new_noise_and_match_eager_calls():
    s, e = _ddpm(8), _ddpm(8)
    s.capture(den)
    outs = [s.replay().clone() for _ in range(3)]
    assert not torch.equal(outs[0], outs[1])
    for o in outs:
        assert torch.equal(o, e.sample(den))
    assert s.elem_offset == e.elem_offset == 3 * 8 * 4 * 16 * 16
    s.set_sample_offset(8)   # rewind: the graph stays valid (the counter is device state, not a baked-in parameter)
    assert torch.equal(s.replay(), outs[1])


def test_misaligned_device_offset_takes_the_shifted_philox_path():
    """an offset that is not a multiple of 4 cannot be rejected on the host when it lives on the device: the kernel
    draws one more Philox group and shifts"""
    ctr = torch.tensor([6], dtype=torch.int64, device=DEV)
    got = philox_normal((1024,), seed=11, tensor_id=2, elem_offset=1, elem_offset_dev=ctr, device=DEV).cpu().numpy()
    ref = philox.normal((1024,), seed=11, tensor_id=2, elem_offset=7)
    assert np.abs(got - ref).max() < 6e-6
    h = philox_normal((1024,), seed=11, tensor_id=2, elem_offset=1, elem_offset_dev=ctr, dtype=torch.float16, device=DEV).float().cpu().numpy()
    assert np.abs(h - ref).max() < 4e-3


def test_sample_host_with_regenerated_eps0(weights_dir):
    """sample_host with eps0='regen' staged the host noise in the X ping-pong buffer that step 1 overwrites"""
    triple = CoeffTriple.from_npz(os.path.join(weights_dir, "step_10_weight_42.npz"))
    mk = lambda mode: NaturalInferenceSampler(triple, io_score_vp(triple.node), 64, (3, 32, 32), device=DEV, seed=888, eps0=mode)
    g = torch.Generator().manual_seed(4)
    nh = torch.randn(64, 3, 32, 32, generator=g).pin_memory()
    ref = to_pixel_u8(mk("stored").sample(den, noise=nh.to(DEV))).cpu()
    for mode in ("stored", "regen"):
        o = torch.empty(64, 32, 32, 3, dtype=torch.uint8).pin_memory()
        mk(mode).sample_host(den, nh, o, pixels=True)
        torch.cuda.synchronize()
        assert torch.equal(o, ref), mode


@pytest.mark.parametrize("pixels", [True, False])
def test_graphed_host_pipeline_equals_the_launch_by_launch_pipeline(weights_dir, pixels):
    """sample_host_many(graph=True): one captured graph per staging-buffer parity, replayed per batch; host noise and
    device noise variants both equal the ungraphed pipeline, byte for byte"""
    triple = CoeffTriple.from_npz(os.path.join(weights_dir, "step_10_weight_42.npz"))
    mk = lambda: NaturalInferenceSampler(triple, io_score_vp(triple.node), 128, (3, 32, 32), device=DEV, seed=888)
    g = torch.Generator().manual_seed(5)
    noises = [torch.randn(128, 3, 32, 32, generator=g).pin_memory() for _ in range(5)]
    new_out = lambda: [(torch.empty(128, 32, 32, 3, dtype=torch.uint8) if pixels else torch.empty(128, 3, 32, 32)).pin_memory() for _ in range(5)]
    for nh in (noises, None):
        a, b = new_out(), new_out()
        s1, s2 = mk(), mk()
        s1.sample_host_many(den, nh, a, pixels=pixels, first_sample=256)
        n0 = ni.launch_count()
        s2.sample_host_many(den, nh, b, pixels=pixels, first_sample=256, graph=True)  # captures (launches through the ABI) ...
        s2.sample_host_many(den, nh, b, pixels=pixels, first_sample=256, graph=True)  # ... then replays only
        torch.cuda.synchronize()
        for x, y in zip(a, b):
            assert torch.equal(x, y)
        if nh is None:
            assert not torch.equal(a[0], a[1])


def test_device_counter_arguments_are_validated():
    from naturaldiffusion_b200 import _lib
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    buf = torch.zeros(4, dtype=torch.int64, device=DEV)
    dst = torch.empty(64, device=DEV)
    assert L.ni_counter_add(None, 4, st) == -1
    assert L.ni_counter_add(buf.data_ptr() + 4, 4, st) == -1
    assert L.ni_philox_normal_at(dst.data_ptr(), 64, 0, 1, 0, 0, buf.data_ptr() + 4, st) == -1
    assert L.ni_counter_add(buf.data_ptr(), 5, st) == 0 and L.ni_counter_add(buf.data_ptr(), 7, st) == 0
    torch.cuda.synchronize()
    assert int(buf[0]) == 12


def test_graphed_host_pipeline_with_a_stochastic_matrix_advances_fresh_noise():
    """host noise in + DDPM fresh noise drawn in-kernel: every batch of the graphed pipeline draws its own eps_1..K (device
    counter advanced inside the graph) and equals the launch-by-launch pipeline"""
    g = torch.Generator().manual_seed(9)
    noises = [torch.randn(8, 4, 16, 16, generator=g).pin_memory() for _ in range(4)]
    new_out = lambda: [torch.empty(8, 4, 16, 16).pin_memory() for _ in range(4)]
    a, b = new_out(), new_out()
    _ddpm(8).sample_host_many(den, noises, a, first_sample=16)
    s = _ddpm(8)
    s.sample_host_many(den, noises, b, first_sample=16, graph=True)
    s.sample_host_many(den, noises, b, first_sample=16, graph=True)
    torch.cuda.synchronize()
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    # same host noise in batches 0 and 1 would still give different samples: the fresh noise differs
    same = [noises[0], noises[0]]
    o = new_out()[:2]
    _ddpm(8).sample_host_many(den, same, o)
    torch.cuda.synchronize()
    assert not torch.equal(o[0], o[1])


def test_in_kernel_noise_passes_a_kolmogorov_smirnov_test():
    from scipy import stats
    z = philox_normal((1 << 20,), seed=12345, tensor_id=77, device=DEV).cpu().numpy().astype(np.float64)
    ks = stats.kstest(z, "norm")
    assert ks.statistic < 2.5e-3, ks
    assert abs(stats.skew(z)) < 0.01 and abs(stats.kurtosis(z)) < 0.02
    # independence across tensor ids and neighbouring elements
    z2 = philox_normal((1 << 20,), seed=12345, tensor_id=78, device=DEV).cpu().numpy().astype(np.float64)
    assert abs(np.corrcoef(z, z2)[0, 1]) < 5e-3 and abs(np.corrcoef(z[:-1], z[1:])[0, 1]) < 5e-3
