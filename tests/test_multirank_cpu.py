"""CPU, world_size 2 over gloo: the host-side N>1 logic -- batch sharding with global Philox offsets (no collective
on the sampling path) and the all-reduce of FID statistics (the one collective of the north-star)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from naturaldiffusion_b200.fid import FidAccumulator, frechet_distance
from naturaldiffusion_b200.sampler import shard_range
from oracle import philox


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # --- FID statistics: every rank holds a different shard of the activations
        feats = np.load(os.path.join(tmp, "feats.npy"))
        lo, hi = shard_range(feats.shape[0], rank, world)
        acc = FidAccumulator(dim=feats.shape[1], device="cpu", host_logic_only=True)
        for i in range(lo, hi, 37):  # ragged micro-batches
            acc.update(torch.from_numpy(feats[i:min(hi, i + 37)]).float())
        acc.all_reduce()
        mu, sigma = acc.finalize()
        np.save(os.path.join(tmp, f"mu{rank}.npy"), mu)
        np.save(os.path.join(tmp, f"sigma{rank}.npy"), sigma)
        # --- sharded noise: rank r draws samples [lo, hi) of the global batch, keyed by the global element index
        B, per = 10, 48
        lo, hi = shard_range(B, rank, world)
        mine = philox.normal((hi - lo, per), seed=888, tensor_id=2, elem_offset=lo * per)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            np.save(os.path.join(tmp, "noise.npy"), np.concatenate(gathered))
    finally:
        dist.destroy_process_group()


def test_two_ranks_fid_allreduce_and_sharded_noise(tmp_path):
    rng = np.random.default_rng(0)
    n, d = 301, 24
    feats = (rng.standard_normal((n, d)) @ rng.standard_normal((d, d)) + rng.standard_normal(d)).astype(np.float32)
    np.save(tmp_path / "feats.npy", feats)
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    # reference arithmetic (src/CIFAR10NaturalInference.py:78-80)
    mu_ref, sigma_ref = np.mean(feats.astype(np.float64), axis=0), np.cov(feats.astype(np.float64), rowvar=False)
    for r in range(2):
        assert np.abs(np.load(tmp_path / f"mu{r}.npy") - mu_ref).max() < 1e-12
        assert np.abs(np.load(tmp_path / f"sigma{r}.npy") - sigma_ref).max() < 1e-10
    assert np.array_equal(np.load(tmp_path / "noise.npy"), philox.normal((10, 48), seed=888, tensor_id=2))


def test_frechet_distance_known_answers():
    rng = np.random.default_rng(1)
    d = 16
    a = rng.standard_normal((d, d)); s1 = a @ a.T + np.eye(d)
    mu1, mu2 = rng.standard_normal(d), rng.standard_normal(d)
    assert abs(frechet_distance(mu1, s1, mu1, s1)) < 1e-6
    # commuting covariances: closed form |dmu|^2 + sum (sqrt(l1) - sqrt(l2))^2
    l1, l2 = rng.uniform(0.5, 2, d), rng.uniform(0.5, 2, d)
    q, _ = np.linalg.qr(rng.standard_normal((d, d)))
    want = np.sum((mu1 - mu2) ** 2) + np.sum((np.sqrt(l1) - np.sqrt(l2)) ** 2)
    got = frechet_distance(mu1, q @ np.diag(l1) @ q.T, mu2, q @ np.diag(l2) @ q.T)
    assert abs(got - want) < 1e-8
    assert frechet_distance(mu1, s1, mu2, s1) == pytest.approx(np.sum((mu1 - mu2) ** 2), abs=1e-6)


def test_accumulator_single_process_matches_numpy():
    rng = np.random.default_rng(2)
    feats = rng.standard_normal((500, 32)).astype(np.float32)
    acc = FidAccumulator(dim=32, device="cpu", host_logic_only=True)
    acc.update(torch.from_numpy(feats[:123])).update(torch.from_numpy(feats[123:])).all_reduce()
    mu, sigma = acc.finalize()
    assert np.abs(mu - feats.astype(np.float64).mean(0)).max() < 1e-12
    assert np.abs(sigma - np.cov(feats.astype(np.float64), rowvar=False)).max() < 1e-10


def test_fid_accumulator_refuses_a_silent_cpu_path():
    from naturaldiffusion_b200 import NiError
    with pytest.raises(NiError):
        FidAccumulator(dim=8, device="cpu")
