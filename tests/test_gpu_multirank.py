"""GPU, world_size 2 over NCCL (skipped on a 1-GPU box): sharded sampling equals the single-GPU run bit for bit and
needs no collective; FID statistics merge with one NCCL all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
WEIGHTS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "naturaldiffusion_b200", "data", "weights")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _den(x, k):
    return torch.tanh(0.7 * x) * (1.0 + 0.01 * k) + 0.1 * x


def _worker(rank, world, port, tmp):
    import naturaldiffusion_b200 as ni
    from naturaldiffusion_b200.fid import FidAccumulator
    from naturaldiffusion_b200.sampler import NaturalInferenceSampler, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        triple = ni.CoeffTriple.from_npz(os.path.join(WEIGHTS, "step_10_weight_42.npz"))
        B = 512
        lo, hi = shard_range(B, rank, world)
        s = NaturalInferenceSampler(triple, ni.io_score_vp(triple.node), hi - lo, (3, 32, 32), device=dev, seed=888, sample_offset=lo)
        x = s.sample(_den)
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x.contiguous())  # test-only gather; the sampling path itself used no collective
        feats = torch.from_numpy(np.load(os.path.join(tmp, "feats.npy"))).to(dev)
        flo, fhi = shard_range(feats.shape[0], rank, world)
        acc = FidAccumulator(dim=feats.shape[1], device=dev).update(feats[flo:fhi]).all_reduce()
        mu, sigma = acc.finalize()
        # host-copy routing (hostutil.choose_host_relay): the probe runs on any box; "force" relays every rank's host copies
        # through the other GPU over NVLink -- same bytes out as the direct route, noise in and images out
        from naturaldiffusion_b200.hostutil import choose_host_relay
        auto, info = choose_host_relay(rank, world, dev, mode="auto")
        assert set(auto) == {"d2h", "bidir"} and "error" not in info, info
        forced, finfo = choose_host_relay(rank, world, dev, mode="force")
        assert forced == {"d2h": 1 - rank, "bidir": 1 - rank}, (forced, finfo)
        g = torch.Generator().manual_seed(11 + rank)
        noises = [torch.randn(hi - lo, 3, 32, 32, generator=g).pin_memory() for _ in range(5)]
        outs = {}
        for route in (None, forced["bidir"]):
            s.set_host_relay(route)
            for with_noise in (True, False):
                for graph in (False, True):
                    o = [torch.empty(hi - lo, 32, 32, 3, dtype=torch.uint8).pin_memory() for _ in range(5)]
                    s.sample_host_many(_den, noises if with_noise else None, o, pixels=True, first_sample=lo, graph=graph)
                    torch.cuda.synchronize()
                    outs[(route is not None, with_noise, graph)] = torch.stack(o)
        for with_noise in (True, False):
            for graph in (False, True):
                assert torch.equal(outs[(True, with_noise, graph)], outs[(False, with_noise, graph)]), (rank, with_noise, graph)
        s.set_host_relay(None)
        if rank == 0:
            torch.save(torch.cat(parts).cpu(), os.path.join(tmp, "sharded.pt"))
            np.save(os.path.join(tmp, "mu.npy"), mu)
            np.save(os.path.join(tmp, "sigma.npy"), sigma)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_sampling_and_fid_allreduce(tmp_path):
    import naturaldiffusion_b200 as ni
    from naturaldiffusion_b200.sampler import NaturalInferenceSampler
    rng = np.random.default_rng(0)
    feats = rng.standard_normal((1000, 256)).astype(np.float32)
    np.save(tmp_path / "feats.npy", feats)
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    triple = ni.CoeffTriple.from_npz(os.path.join(WEIGHTS, "step_10_weight_42.npz"))
    s = NaturalInferenceSampler(triple, ni.io_score_vp(triple.node), 512, (3, 32, 32), device="cuda:0", seed=888)
    full = s.sample(_den).cpu()
    assert torch.equal(torch.load(tmp_path / "sharded.pt"), full)
    assert np.abs(np.load(tmp_path / "mu.npy") - feats.astype(np.float64).mean(0)).max() < 1e-12
    assert np.abs(np.load(tmp_path / "sigma.npy") - np.cov(feats.astype(np.float64), rowvar=False)).max() < 1e-10


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_host_relay_route_is_bit_identical_single_process():
    """one process, sampler on cuda:0, host copies relayed through cuda:1 (set_host_relay): latent and pixel outputs equal the
    direct route; an invalid peer is refused"""
    import naturaldiffusion_b200 as ni
    from naturaldiffusion_b200.sampler import NaturalInferenceSampler
    triple = ni.CoeffTriple.from_npz(os.path.join(WEIGHTS, "step_15_weight_173.npz"))
    s = NaturalInferenceSampler(triple, ni.io_score_vp(triple.node), 96, (3, 32, 32), device="cuda:0", seed=3, advance=0)
    g = torch.Generator().manual_seed(2)
    noises = [torch.randn(96, 3, 32, 32, generator=g).pin_memory() for _ in range(7)]
    res = {}
    for route in (None, 1):
        s.set_host_relay(route)
        lat = [torch.empty(96, 3, 32, 32).pin_memory() for _ in range(7)]
        s.sample_host_many(_den, noises, lat)
        torch.cuda.synchronize()
        res[route] = torch.stack(lat)
    assert torch.equal(res[None], res[1])
    ref = torch.stack([s.sample(_den, noise=n.to("cuda:0")).cpu() for n in noises])
    assert torch.equal(res[1], ref)
    with pytest.raises(ni.NiError):
        s.set_host_relay(0)
    with pytest.raises(ni.NiError):
        s.set_host_relay(torch.cuda.device_count())
