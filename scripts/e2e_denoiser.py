#!/usr/bin/env python
"""End-to-end sampling throughput WITH a real (random-init) torch denoiser, next to the reference's eager loop on the
same GPU.  Not the bench headline (the denoiser forward is torch and dominates); it shows what the fused step buys in
a full pipeline and is the number that scales across GPUs.

  python scripts/e2e_denoiser.py --config c2 --batch 1024 --reps 3 [--autocast bf16] [--gpus N under torchrun]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import naturaldiffusion_b200 as ni  # noqa: E402
from naturaldiffusion_b200.adapters import ncsnpp_denoiser  # noqa: E402
from naturaldiffusion_b200.denoisers import NCSNppVP  # noqa: E402
from naturaldiffusion_b200.sampler import NaturalInferenceSampler  # noqa: E402

W = os.path.join(ROOT, "naturaldiffusion_b200", "data", "weights")


def reference_eager_loop(A, B, node, model, noise):
    """the reference's CIFAR loop verbatim in structure (oracle restatement) on the GPU: fp64 history, one torch kernel per op"""
    from oracle import ni_oracle as O
    score_fn = O.make_vp_score_fn(lambda x, labels: model(x, labels))
    ts = node[:, 0]
    seq, x = [], noise
    for kk in range(len(ts) - 1):
        vec_t = ts[kk] * torch.ones(x.shape[0], device=x.device)
        score = score_fn(x, vec_t.float())
        x64, s64 = x.double(), score.double()
        e = torch.tensor(node[kk, 2], dtype=torch.float64, device=x.device)
        a = torch.tensor(node[kk, 1], dtype=torch.float64, device=x.device)
        seq.append((s64 * e ** 2 + x64) / a)
        out = torch.zeros_like(seq[0])
        for ii, x0 in enumerate(seq):
            out += x0 * A[kk][ii]
        x = out.float() + B[kk, 0] * noise
    return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--autocast", default="none", choices=["none", "bf16", "fp16"])
    args = ap.parse_args()
    rank, world, lr = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    fname = {"c2": "step_10_weight_42.npz", "c3": "step_15_weight_173.npz"}[args.config]
    triple = ni.CoeffTriple.from_npz(os.path.join(W, fname))
    torch.manual_seed(0)
    model = NCSNppVP().reinit_output().to(dev).eval()
    ac = {"none": None, "bf16": torch.bfloat16, "fp16": torch.float16}[args.autocast]
    den = ncsnpp_denoiser(model, triple.node, autocast_dtype=ac)
    s = NaturalInferenceSampler(triple, ni.io_score_vp(triple.node), args.batch, (3, 32, 32), device=dev, seed=888, sample_offset=rank * args.batch)

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    with torch.no_grad():
        ms_ours = timed(lambda: s.sample(den), args.reps)
        ms_model = timed(lambda: [den(s._X[0].view(s.full_shape()), k) for k in range(triple.K)], args.reps)
        res = {"config": args.config, "batch_per_gpu": args.batch, "n_gpus": world, "K": triple.K, "autocast": args.autocast,
               "ours_ms_per_trajectory": ms_ours, "ours_samples_per_s": world * args.batch / (ms_ours * 1e-3),
               "denoiser_only_ms": ms_model, "update_share_ours": max(0.0, 1 - ms_model / ms_ours)}
        # true end to end with HOST buffers: pinned fp32 noise in (H2D), K x (NCSN++ forward + fused step), uint8 images out (D2H)
        shape = s.full_shape()
        nb = 3
        noise_h = [torch.randn(shape).pin_memory() for _ in range(2)]
        out_h = [torch.empty((args.batch, 32, 32, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
        ms_host = timed(lambda: s.sample_host_many(den, [noise_h[i % 2] for i in range(nb)], [out_h[i % 2] for i in range(nb)], pixels=True), 1) / nb
        res.update(e2e_host_ms_per_batch=ms_host, e2e_host_samples_per_s=world * args.batch / (ms_host * 1e-3),
                   e2e_host_bytes_per_batch={"h2d": noise_h[0].numel() * 4, "d2h": out_h[0].numel()})
        if world == 1:
            from naturaldiffusion_b200.ops import philox_normal
            noise = philox_normal(s.full_shape(), seed=888, tensor_id=0, device=dev)
            plain = (lambda x, labels: model(x, labels)) if ac is None else (lambda x, labels: torch.autocast("cuda", dtype=ac)(model)(x, labels).float())
            ms_ref = timed(lambda: reference_eager_loop(triple.A, triple.B, triple.node, plain, noise), args.reps)
            res.update(reference_eager_gpu_ms=ms_ref, reference_eager_gpu_samples_per_s=args.batch / (ms_ref * 1e-3),
                       update_ms_reference_eager=ms_ref - ms_model, update_ms_ours=ms_ours - ms_model)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
