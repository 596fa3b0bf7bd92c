#!/usr/bin/env python
"""The reference's CIFAR-10 driver (`natural_inference_tx`, src/CIFAR10NaturalInference.py:241-317) on the B200 path:
N samples in batches, sharded by batch over the ranks (one process per GPU, no collective while sampling), fused
`ni_step` per step, uint8 NHWC images from the last step, Inception-style features -> FID statistics on each GPU ->
ONE all-reduce -> Frechet distance on the host.

Offline there is no checkpoint, no pytorch_fid Inception and no weights/cifar10_mu_sigma.npz, so the score model is a
random-init NCSN++ and the feature extractor a small random conv net: the printed "FID" only exercises the plumbing.

  python examples/cifar10_pipeline.py --samples 2000 --batch 500
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/cifar10_pipeline.py --samples 50000
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import naturaldiffusion_b200 as ni  # noqa: E402
from naturaldiffusion_b200.adapters import ncsnpp_denoiser  # noqa: E402
from naturaldiffusion_b200.denoisers import NCSNppVP  # noqa: E402
from naturaldiffusion_b200.fid import FidAccumulator, accumulate_images, frechet_distance  # noqa: E402
from naturaldiffusion_b200.sampler import NaturalInferenceSampler, shard_range  # noqa: E402


def feature_net(dim, device):
    g = torch.Generator().manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 64, 3, 2, 1), torch.nn.ReLU(), torch.nn.Conv2d(64, dim, 3, 2, 1), torch.nn.ReLU(),
                              torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten())
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    return net.to(device).eval()


def run(samples=2000, batch=500, weights="step_10_weight_42.npz", feat_dim=256, seed=888, small_model=False, quiet=False, features="small"):
    rank, world, lr = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group("nccl", device_id=dev)
    triple = ni.CoeffTriple.from_npz(os.path.join(ROOT, "naturaldiffusion_b200", "data", "weights", weights))
    torch.manual_seed(0)
    model = (NCSNppVP(nf=32, num_res_blocks=1) if small_model else NCSNppVP()).reinit_output().to(dev).eval()
    den = ncsnpp_denoiser(model, triple.node)
    if features == "inception":   # the reference's evaluation shape: 2048-d pool3 activations of an InceptionV3 (random-init offline)
        from naturaldiffusion_b200.fid import inception_pool3_standin
        feat_dim = 2048
        feats = inception_pool3_standin(dev, torch.bfloat16)
    else:
        feats = feature_net(feat_dim, dev)
    lo, hi = shard_range(samples, rank, world)                      # this rank's samples [lo, hi) of the global run
    acc = FidAccumulator(dim=feat_dim, device=dev)
    t0 = time.perf_counter()
    samplers = {}
    for start in range(lo, hi, batch):
        b = min(batch, hi - start)
        if b not in samplers:
            samplers[b] = NaturalInferenceSampler(triple, ni.io_score_vp(triple.node), b, (3, 32, 32), device=dev, seed=seed)
        s = samplers[b]
        s.set_sample_offset(start)                                    # global Philox index: same images for any batching / sharding
        pix = torch.empty((b, 32, 32, 3), dtype=torch.uint8, device=dev)
        s.sample(den, pixels_out=pix)                                 # K fused steps, last one emits uint8 NHWC
        accumulate_images(acc, pix, feats, batch_size=b)
    torch.cuda.synchronize()
    t_ar = time.perf_counter()
    acc.all_reduce()                                                   # the only collective
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    ar_ms = (time.perf_counter() - t_ar) * 1e3
    mu, sigma = acc.finalize()
    fid = float("nan")
    if rank == 0:                                                      # the O(d^3) host step runs once, like in the reference
        rng = np.random.default_rng(0)                                 # stand-in for weights/cifar10_mu_sigma.npz
        ref = rng.standard_normal((4 * feat_dim, feat_dim)) * 0.05 + mu
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")                            # random-init features: near-singular covariances
            fid = frechet_distance(np.mean(ref, 0), np.cov(ref, rowvar=False), mu, sigma)
    if rank == 0 and not quiet:
        print(f"{samples} samples on {world} GPU(s) in {dt:.2f} s ({samples / dt:.0f} samples/s incl. random-init NCSN++); "
              f"features={features} ({feat_dim}-d); all-reduce of {acc.buf.numel() * 8 / 1e6:.1f} MB statistics {ar_ms:.2f} ms; "
              f"n={int(acc.n)}; plumbing-only FID vs synthetic statistics = {fid:.4f}", flush=True)
    return dict(n=acc.n, mu=mu, sigma=sigma, fid=fid, seconds=dt)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=2000)
    ap.add_argument("--batch", type=int, default=500)
    ap.add_argument("--weights", default="step_10_weight_42.npz")
    ap.add_argument("--small-model", action="store_true")
    ap.add_argument("--features", default="small", choices=["small", "inception"], help="inception: torchvision InceptionV3 (random init), 2048-d pool3")
    a = ap.parse_args()
    run(a.samples, a.batch, a.weights, small_model=a.small_model, features=a.features)
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()
