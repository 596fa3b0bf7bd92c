#!/usr/bin/env python
"""bench.py -- Natural Inference update throughput on B200 (contract: see the task's bench section).

Workload (BASELINE.json configs[1], "C2"): CIFAR-10 32x32 Natural Inference with the reference's
step_10_weight_42 coefficient matrix, batch 4096 per GPU, fp32 state.  One bench "step" = one full
K=10-step NI trajectory over one batch: 10 fused `ni_step` launches, 65 tensor-sized HBM transfers
(3.27 GB algorithmic).  The denoiser is the *null denoiser* of SURVEY 8d (a pre-generated N(0,1)
model-output tensor re-read from HBM every step), so the timed region contains only our kernels:
the denoiser forward stays torch and is not what this repo accelerates.

  value     samples/s, inputs resident in HBM, whole job over all ranks (weak scaling: 4096 samples/GPU)
  e2e       same metric through NaturalInferenceSampler.sample_host(): pinned host noise -> H2D -> 10 steps
            -> fused uint8 pixel stage -> D2H, copies inside the timed region
  roofline  algorithmic bytes per ni_step launch / mean launch duration (CUDA events over the timed region)
  cpu_baseline / --impl reference: the oracle's torch-CPU restatement of the reference loop
            (src/CIFAR10NaturalInference.py:219-238,294-304) on this box's host cores -- the reference is
            pure Python/torch, there is nothing to compile into oracle/_ref.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WEIGHTS = os.path.join(ROOT, "tests", "golden", "reference_weights")
CONFIGS = {
    # name: (matrix, per-GPU batch, sample shape, model outputs m, model-output channels, state dtype)
    "c2": ("step_10_weight_42.npz", 4096, (3, 32, 32), 1, 3, "f32"),           # BASELINE configs[1] -- the bench headline
    "c3": ("step_15_weight_173.npz", 16384, (3, 32, 32), 1, 3, "f32"),         # configs[2]
    "c4": ("ddpm_250 (generated)", 1024, (4, 32, 32), 2, 8, "f32"),            # configs[3]: DiT-XL/2 shapes, CFG, 8-channel outputs
    "c5": ("sd3_step_28_weight.csv", 64, (16, 128, 128), 2, 16, "f16"),        # configs[4]: SD3 shapes, fp16 state
    "c5s": ("sd3_step_28_weight_sharp.csv", 64, (16, 128, 128), 2, 16, "f16"),
}
METRIC = "NI-update samples/sec (HBM GB/s and % peak in `roofline`)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed trajectories (default: ~2 s worth for ours, 10 for --impl reference)")
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--eps0", default="stored", choices=["stored", "regen"])
    ap.add_argument("--variant", type=int, default=0, help="0 auto, 1 direct-load kernel, 2 TMA-staged kernel")
    ap.add_argument("--opt", action="append", default=[], help="ni_set_option name=value (tuning)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cold", action="store_true", help="run the cold-L2 per-launch check on c4/c5 too (default: c2/c3 only)")
    ap.add_argument("--no-cold", action="store_true", help="skip the cold-L2 per-launch check (keeps profiler launch lists to the timed region)")
    ap.add_argument("--no-numa", action="store_true", help="do not pin the process to the GPU's NUMA node")
    ap.add_argument("--eager-comparator", action="store_true", help="also time the reference's eager torch loop on this GPU (c2/c3)")
    ap.add_argument("--markov", default="auto", choices=["auto", "0", "1"], help="first-order fast path (c4/c5): auto|0|1")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 10 if args.impl == "reference" else {"c2": 4000, "c3": 500, "c4": 8, "c5": 40, "c5s": 40}[args.config]
    if args.warmup is None:
        args.warmup = 1 if args.impl == "reference" else (20 if args.config in ("c2", "c3") else 3)
    return args


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference loop, null denoiser
# ----------------------------------------------------------------------------------------------
def make_cpu_trajectory(cfg, batch):
    """Reference arithmetic (oracle restatement, reference dtypes) for one batch with the null denoiser.
    c2/c3: src/CIFAR10NaturalInference.py:294-304; c4: src/ValidateNaturalInference.py:349-366;
    c5: src/SD3NaturalInference.py:201-221 (in fp32: fp16 elementwise on a CPU is not representative)."""
    import torch
    from oracle import ni_oracle as O
    fname, _, shape, m, cout, _ = CONFIGS[cfg]
    g = torch.Generator().manual_seed(888)
    noise = torch.randn((batch,) + shape, generator=g)
    outs = [torch.randn((batch, cout) + shape[1:], generator=g) for _ in range(m)]
    if cfg in ("c2", "c3"):
        A, B, node = O.load_triple(os.path.join(WEIGHTS, fname))
        score_fn = O.make_vp_score_fn(lambda x, labels: outs[0])
        return lambda: O.cifar_ni_loop(A, B, node, score_fn, noise)[0]
    if cfg == "c4":
        A, B, node = O.ddpm_triple(250)
        fresh = [torch.randn((batch,) + shape, generator=g) for _ in range(4)]
        fresh = [fresh[k % 4] for k in range(250)]  # distinct values are irrelevant for timing; 250 tensors would only cost RAM
        eps_model = lambda z, t: (outs[0][:, :4], outs[1][:, :4])
        return lambda: O.validate_ni_loop(A, B, node, eps_model, noise, fresh)[0]
    W = O.load_sd3_csv(os.path.join(WEIGHTS, fname))
    sig = O.sd3_sigmas()
    return lambda: O.sd3_ni_loop(W, sig, lambda x, k: (outs[0], outs[1]), noise)[0]


def time_cpu(cfg, batch, reps, warm):
    fn = make_cpu_trajectory(cfg, batch)
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return ts


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args, rank):
    """`--impl reference`: rank 0 alone times the reference's CPU path; other ranks exit."""
    if rank != 0:
        return
    import torch
    cores = torch.get_num_threads()
    fname, full_batch, shape, m, cout, _ = CONFIGS[args.config]
    # bounded sample: the full per-GPU batch when (steps+warmup) trajectories of it fit in about two minutes,
    # otherwise the largest batch that does (time is super-linear in the batch once tensors leave the CPU caches)
    pb = min(full_batch, 64)
    probe = time_cpu(args.config, pb, 1, 1)[0]
    budget = 120.0 / max(1, args.steps + args.warmup)
    batch = int(max(8, min(full_batch, pb * budget / probe / 4)))
    if batch >= 64:
        batch -= batch % 64
    ts = time_cpu(args.config, batch, args.steps, args.warmup)
    total = sum(ts)
    val = batch * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 history / f32 state (reference dtypes)", "data": "synthetic",
        "config": {"workload": f"{args.config}: {fname} NI update-only, null denoiser, sample of {batch} of {full_batch} samples per step",
                   "shape": [batch] + list(shape)},
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{batch}-sample batches x {args.steps} trajectories; torch {torch.__version__} CPU; {cpu_model()}"},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), [v.strip() for v in ln.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows]
        sm, mx, reasons, pw = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(cfg, eps0):
    """DRAM bytes per ni_step launch from the committed ncu --set full capture, if one matches this workload."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(f"{cfg}:{eps0}")
    except Exception:
        return None


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import naturaldiffusion_b200 as ni
    from naturaldiffusion_b200.sampler import NaturalInferenceSampler

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from naturaldiffusion_b200.hostutil import bind_to_gpu_numa_node
    all_cpus = os.sched_getaffinity(0)
    numa_cpus = None if args.no_numa else bind_to_gpu_numa_node(local_rank)  # before any pinned allocation
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from naturaldiffusion_b200 import _lib as nilib
    nilib.set_option("variant", args.variant)
    for kv in args.opt:
        name, val = kv.split("=")
        nilib.set_option(name, int(val))
    from naturaldiffusion_b200 import coeffs, generators
    from naturaldiffusion_b200.ops import philox_normal
    fname, batch, shape, m, cout, dts = CONFIGS[args.config]
    batch = args.batch or batch
    dtype = {"f32": torch.float32, "f16": torch.float16}[dts]
    markov = {"auto": "auto", "0": False, "1": True}[args.markov]
    if args.config in ("c2", "c3"):
        triple = ni.CoeffTriple.from_npz(os.path.join(WEIGHTS, fname))
        io = ni.io_score_vp(triple.node)
    elif args.config == "c4":
        triple = generators.ddpm_triple(250)
        c1, c2, _ = coeffs.ddim_x0_coeffs(250)
        io = ni.io_eps_cfg(c1, c2, 4.0)
    else:
        sig = coeffs.flow_match_sigmas(28)
        triple = ni.CoeffTriple.from_sd3_csv(os.path.join(WEIGHTS, fname), sig)
        io = ni.io_velocity_cfg(sig, 7.0)
    K = triple.K
    sampler = NaturalInferenceSampler(triple, io, batch, shape, device=dev, dtype=dtype, seed=888,
                                      eps0=args.eps0, sample_offset=rank * batch, markov=markov)
    # null denoiser: m pre-generated N(0,1) model-output tensors [B, cout, H, W], re-read from HBM every step
    outs = tuple(philox_normal((batch, cout) + shape[1:], seed=888, tensor_id=1000 + i, elem_offset=rank * batch * cout * shape[1] * shape[2],
                               dtype=dtype, device=dev) for i in range(m))
    den = (lambda x, k: outs[0]) if m == 1 else (lambda x, k: outs)
    numel = sampler.numel
    esize = torch.empty(0, dtype=dtype).element_size()
    stored0 = args.eps0 == "stored"
    units = sampler.plan.total_units(m, eps0_stored=stored0)
    bytes_per_traj = units * numel * esize
    launches_per_traj = sampler.kernel_launches_per_trajectory

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm
    noise = philox_normal((batch,) + shape, seed=888, tensor_id=0, elem_offset=sampler.elem_offset, dtype=dtype, device=dev)
    c0 = ni.launch_count()
    sampler.sample(den, noise=noise)
    torch.cuda.synchronize()
    counted = ni.launch_count() - c0
    assert counted == launches_per_traj, (counted, launches_per_traj)
    flavours = sampler.load_flavours()  # out0/out1 are patched into the descriptors by the trajectory above
    if args.no_graph:
        run = lambda: sampler.sample(den, noise=noise)
    else:
        sampler.capture(den, noise=noise)
        run = sampler.replay
    for _ in range(max(3, args.warmup)):
        run()
    barrier()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        run()
    e1.record()
    barrier()
    t_wall1 = time.perf_counter()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop(t_wall0, t_wall1) if clocks else None
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    value = world * batch * args.steps / (ms * 1e-3)
    ms_per_step = ms / args.steps
    achieved = bytes_per_traj / (ms_per_step * 1e-3) / 1e9  # GB/s per GPU == per launch (only ni_step kernels run)
    peak, peak_src = measured_peak()

    # ---- cold-L2 check: every launch timed on its own with the L2 flushed right before it (a 512 MB READ sweep, which
    # leaves clean lines -- a memset would leave 126 MB of dirty lines to be written back during the timed kernel), so no
    # launch can find the previous step's stores in the 126 MB L2 (the null denoiser leaves nothing between steps)
    cold = None
    if world == 1 and not args.no_cold and (args.config in ("c2", "c3") or args.cold):
        flush = torch.zeros(128 << 20, dtype=torch.float32, device=dev)
        sampler._graph = None
        sampler.sample(den, noise=noise)
        evs = []
        n_cold = 5 if K <= 30 else 1
        st_ptr = torch.cuda.current_stream(dev).cuda_stream
        for _ in range(n_cold):
            for k in range(K):
                flush.max()
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record()
                sampler.step(k, outs[0] if m == 1 else outs, st_ptr)
                b_.record()
                evs.append((a_, b_))
        torch.cuda.synchronize()
        cold_ms = sum(a_.elapsed_time(b_) for a_, b_ in evs) / n_cold
        cold = {"ms_per_trajectory": cold_ms, "achieved": bytes_per_traj / (cold_ms * 1e-3) / 1e9,
                "frac": bytes_per_traj / (cold_ms * 1e-3) / 1e9 / measured_peak()[0],
                "how": "each launch bracketed by its own CUDA events after a 512 MB read sweep (L2 flush); includes ~2 us event overhead per launch, no graph / PDL overlap"}
        del flush

    # ---- end-to-end arm: host buffers through the public sampler API
    e2e = None
    if not args.no_e2e:
        pixels = args.config in ("c2", "c3")  # image-space configs end in the uint8 stage; latent configs return the latent
        noise_h = torch.empty((batch,) + shape, dtype=dtype).pin_memory()
        noise_h.copy_(noise)
        out_h = (torch.empty((batch, shape[1], shape[2], shape[0]), dtype=torch.uint8) if pixels else torch.empty((batch,) + shape, dtype=dtype)).pin_memory()
        sampler._graph = None
        e2e_sampler = sampler  # same state slab; launches go through the non-graph path with the H2D / D2H copies in stream order
        # two pinned buffers per direction, alternated: batch i+1's H2D and batch i-1's D2H overlap batch i's steps
        noise_hs = [noise_h, noise_h.clone().pin_memory()]
        out_hs = [out_h, out_h.clone().pin_memory()]
        e2e_sampler.sample_host_many(den, [noise_hs[i % 2] for i in range(4)], [out_hs[i % 2] for i in range(4)], pixels=pixels)
        barrier()
        n_e2e = max(4, args.steps // 4)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        e2e_sampler.sample_host_many(den, [noise_hs[i % 2] for i in range(n_e2e)], [out_hs[i % 2] for i in range(n_e2e)], pixels=pixels)
        s1.record()
        barrier()
        # informational: the reference's own data flow (noise drawn on the device from a seed, only images come back)
        e2e_sampler.sample_host_many(den, None, [out_hs[i % 2] for i in range(4)], pixels=pixels, first_sample=rank * batch)
        barrier()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        e2e_sampler.sample_host_many(den, None, [out_hs[i % 2] for i in range(n_e2e)], pixels=pixels, first_sample=rank * batch)
        q1.record()
        barrier()
        seeded_ms = q0.elapsed_time(q1)
        ems = s0.elapsed_time(s1)
        if world > 1:
            t = torch.tensor([ems], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = t.item()
        e2e = {"value": world * batch * n_e2e / (ems * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": noise_h.numel() * noise_h.element_size(),
               "d2h_bytes_per_step": out_h.numel() * out_h.element_size(), "steps": n_e2e, "ms_per_step": ems / n_e2e,
               "seeded_variant": {"value": batch * n_e2e / (seeded_ms * 1e-3), "ms_per_step": seeded_ms / n_e2e, "h2d_bytes_per_step": 0,
                                  "note": "this rank only; noise drawn on the device from (seed, global sample index) as the reference does "
                                          "(torch.randn on the GPU): informational, NOT the e2e value"},
               "api": f"NaturalInferenceSampler.sample_host_many(pixels={pixels}), double-buffered copy streams: pinned {dts} noise in, " + ("NHWC uint8 out" if pixels else f"{dts} latent out")}

    # ---- optional comparator: the reference's own loop structure (oracle restatement: fp64 history, one torch kernel
    # per op, per-step H2D scalars) on THIS GPU with the same null denoiser -- "eager torch on B200", SURVEY 2.3
    eager = None
    if args.eager_comparator and args.config in ("c2", "c3") and world == 1:
        from oracle import ni_oracle as O
        A_, B_, node_ = O.load_triple(os.path.join(WEIGHTS, fname))

        def eager_traj():
            seq, x = [], noise
            for kk in range(K):
                vec_t = node_[kk, 0] * torch.ones(batch, device=dev)
                score = -outs[0] / O.vp_marginal_std(vec_t)[:, None, None, None]
                x64, s64 = x.to(torch.float64), score.to(torch.float64)
                e_ = torch.tensor(node_[kk, 2], dtype=torch.float64, device=dev)
                a_ = torch.tensor(node_[kk, 1], dtype=torch.float64, device=dev)
                seq.append((s64 * e_ ** 2 + x64) / a_)
                acc = torch.zeros_like(seq[0])
                for ii, x0 in enumerate(seq):
                    acc += x0 * A_[kk][ii]
                x = acc.to(torch.float32) + B_[kk, 0] * noise
            return x

        for _ in range(2):
            eager_traj()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(5):
            xe = eager_traj()
        g1.record()
        torch.cuda.synchronize()
        gms = g0.elapsed_time(g1) / 5
        mine = sampler.sample(den, noise=noise)
        eager = {"ms_per_step": gms, "value": batch / (gms * 1e-3), "unit": "samples/s", "speedup_of_fused_step": gms / ms_per_step,
                 "max_abs_diff_over_norm": float((mine - xe).abs().max() / xe.norm()),
                 "what": "reference loop structure (fp64 history, eager torch ops) on the same B200, same null denoiser"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        os.sched_setaffinity(0, all_cpus)  # the CPU arm gets every host core back, not only the GPU-local ones
        cb = {"c2": batch, "c3": 4096, "c4": 32, "c5": 4, "c5s": 4}[args.config]
        reps = {"c2": 10, "c3": 3}.get(args.config, 2)  # about 10 s of CPU work
        ts = time_cpu(args.config, cb, reps, 1)
        cpu = {"value": cb / min(ts), "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{cb}-sample batch of the same workload (full per-GPU batch is {batch}), best of {reps} trajectories after 1 warm-up "
                         f"({sum(ts):.1f} s CPU work); oracle port of the reference loop in the reference's dtypes; torch {torch.__version__} CPU, {cpu_model()}"}

    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dts,
        "data": "synthetic",
        "config": {"workload": f"{args.config}: NI update, {fname}, batch {batch}/GPU x {list(shape)}, K={K} fused steps per trajectory, "
                               f"null denoiser ({m} pre-generated N(0,1) model output(s) of {cout} channels re-read from HBM each step)",
                   "markov_fast_path": bool(sampler.plan.markov),
                   "shape": [batch] + list(shape), "eps0": args.eps0, "numa_bound_cpus": numa_cpus, "cuda_graph": not args.no_graph, "variant": args.variant, "opts": args.opt,
                   "load_flavours_per_step": "".join(str(f) for f in flavours) if K <= 32 else f"{sum(flavours)} of {K} steps streaming",
                   "load_flavour": "1 = plain ld.global, 0 = L1::no_allocate; auto per launch: ld.global.L1::no_allocate when the bytes it writes fit in 0.6 of the L2 and are >= 1/16 of its traffic, else plain ld.global (override: --opt load_policy=1|2)",
                   "l2": f"inputs larger than L2: per-trajectory working set {(sampler.state_bytes() + sum(o.numel() for o in outs) * esize) / 1e6:.0f} MB vs 126 MB L2",
                   "state_bytes": sampler.state_bytes()},
        "clocks": clk,
        "e2e": e2e,
        "gpu_launches": launches_per_traj * args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(args.config, args.eps0), "peak_source": peak_src,
                     "kernel": "ni_step_kernel (direct-load)" if args.variant != 2 else "ni_step_tma_kernel", "algorithmic_bytes_per_launch": bytes_per_traj / launches_per_traj,
                     "tensor_transfers_per_trajectory": units, "us_per_launch": 1e3 * ms_per_step / launches_per_traj,
                     "frac_of_nominal_8TBs": achieved / 8000.0, "cold_l2": cold},
        "cpu_baseline": cpu,
    }
    if eager is not None:
        line["reference_eager_gpu"] = eager
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        port = 29500 + os.getpid() % 1000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
