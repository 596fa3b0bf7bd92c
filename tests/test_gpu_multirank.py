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
