#!/usr/bin/env python
"""bench.py -- Natural Inference update throughput on B200 (contract: see the task's bench section).

Workload (BASELINE.json configs[1], "C2"): CIFAR-10 32x32 Natural Inference with the reference's
step_10_weight_42 coefficient matrix, batch 4096 per GPU, fp32 state.  One trajectory = K=10 fused `ni_step`
launches, 65 tensor-sized HBM transfers (3.27 GB algorithmic).  The denoiser is the *null denoiser* of SURVEY 8d
(a pre-generated N(0,1) model-output tensor re-read from HBM every step), so the timed region contains only our
kernels: the denoiser forward stays torch and is not what this repo accelerates.

One bench STEP = a block of trajectories sized for >= 50 ms (128 C2 trajectories = 1280 launches, 524288 samples), so
`--steps 20` is > 1 s of sustained work and the clock / power record means something.

  value     samples/s, inputs resident in HBM, whole job over all ranks (weak scaling: 4096 samples/GPU/trajectory)
  e2e       same metric through NaturalInferenceSampler.sample_host_many(graph=True): pinned host noise -> H2D -> 10
            steps (one CUDA graph per batch) -> fused uint8 pixel stage -> D2H, copies inside the timed region, >= 64
            batches whatever --steps is; `e2e.device_noise` is the reference's own data flow (noise drawn on the device,
            src/CIFAR10NaturalInference.py:290: only the images cross PCIe), aggregated over ranks; `e2e.copy_ceiling` is
            the bare cudaMemcpyAsync H2D+D2H rate of the same bytes on the same streams at the same N;
            `e2e.with_score_network` is the same host-buffer call with a real (random-init) NCSN++ in the loop -- the
            pipeline whose 1 -> 8 GPU scaling the north-star asks for (informational: the torch forward dominates)
  roofline  algorithmic bytes per ni_step launch / mean launch duration (CUDA events over the timed region)
  per_config  the other BASELINE shapes (C3, C4 dense + first-order, C5 default/sharp at B 64 and 256), same measurements
  cpu_baseline / --impl reference: the oracle's torch-CPU restatement of the reference loop
            (src/CIFAR10NaturalInference.py:219-238,294-304) on this box's host cores -- the reference is
            pure Python/torch, there is nothing to compile into oracle/_ref.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WEIGHTS = os.path.join(ROOT, "naturaldiffusion_b200", "data", "weights")
CONFIGS = {
    # name: (matrix, per-GPU batch, sample shape, model outputs m, model-output channels, state dtype)
    "c2": ("step_10_weight_42.npz", 4096, (3, 32, 32), 1, 3, "f32"),           # BASELINE configs[1] -- the bench headline
    "c3": ("step_15_weight_173.npz", 16384, (3, 32, 32), 1, 3, "f32"),         # configs[2]
    "c4": ("ddpm_250 (generated)", 1024, (4, 32, 32), 2, 8, "f32"),            # configs[3]: DiT-XL/2 shapes, CFG, 8-channel outputs
    "c5": ("sd3_step_28_weight.csv", 64, (16, 128, 128), 2, 16, "f16"),        # configs[4]: SD3 shapes, fp16 state
    "c5s": ("sd3_step_28_weight_sharp.csv", 64, (16, 128, 128), 2, 16, "f16"),
}
METRIC = "NI-update samples/sec (HBM GB/s and % peak in `roofline`)"
# (config, batch override, markov, label) lines of `per_config`
PER_CONFIG = [("c3", 0, "auto", "c3"), ("c4", 0, "0", "c4_dense"), ("c4", 0, "1", "c4_first_order"),
              ("c5", 64, "0", "c5_default_b64_dense"), ("c5", 64, "1", "c5_default_b64_first_order"), ("c5s", 64, "0", "c5_sharp_b64"),
              ("c5", 256, "0", "c5_default_b256_dense"), ("c5s", 256, "0", "c5_sharp_b256")]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (one step = a >= 50 ms block of trajectories; default 20; 10 for --impl reference)")
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--block", type=int, default=0, help="trajectories per step (default: enough for >= 50 ms)")
    ap.add_argument("--eps0", default="stored", choices=["stored", "regen"])
    ap.add_argument("--variant", type=int, default=0, help="0 auto (specialised kernels), 1 generic direct-load kernel, 2 TMA-staged kernel")
    ap.add_argument("--opt", action="append", default=[], help="ni_set_option name=value (tuning)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-batches", type=int, default=64)
    ap.add_argument("--no-score-network", action="store_true", help="skip e2e.with_score_network (real NCSN++ in the loop, ~5 s)")
    ap.add_argument("--no-per-config", action="store_true", help="skip the per_config lines (the other BASELINE shapes)")
    ap.add_argument("--only", default="", help="comma list of per_config labels to run")
    ap.add_argument("--cold", action="store_true", help="run the cold-L2 per-launch check on c4/c5 too (default: c2/c3 only)")
    ap.add_argument("--no-cold", action="store_true", help="skip the cold-L2 per-launch check (keeps profiler launch lists to the timed region)")
    ap.add_argument("--relay", default=None, choices=["auto", "off", "force"], help="host-copy routing at N > 1 (default $NI_HOST_RELAY or auto): relay the "
                    "host copies of ranks that reach host memory across the socket link through an NVLink peer, if a start-up probe says it pays")
    ap.add_argument("--no-numa", action="store_true", help="do not pin the process to its own slice of the GPU's NUMA node")
    ap.add_argument("--no-check", action="store_true", help="skip the cross-rank bit check (N > 1)")
    ap.add_argument("--eager-comparator", action="store_true", help="also time the reference's eager torch loop on this GPU (c2/c3)")
    ap.add_argument("--markov", default="auto", choices=["auto", "0", "1"], help="first-order fast path (c4/c5): auto|0|1")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 10 if args.impl == "reference" else 20
    if args.warmup is None:
        args.warmup = 1 if args.impl == "reference" else 5
    return args


def workload_config(cfg, batch):
    """The `config` object, IDENTICAL in both arms (ours and --impl reference)."""
    fname, _, shape, m, cout, dts = CONFIGS[cfg]
    K = {"c2": 10, "c3": 15, "c4": 250, "c5": 28, "c5s": 28}[cfg]
    return {"workload": f"{cfg}: Natural Inference update, {fname}, batch {batch}/GPU x {list(shape)}, K={K} fused steps per trajectory, "
                        f"null denoiser ({m} pre-generated N(0,1) model output(s) of {cout} channels re-read from HBM each step)",
            "shape": [batch] + list(shape),
            "l2": "inputs larger than L2: per-trajectory working set > 350 MB vs 126 MB L2 (no flush between iterations)"}


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


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arms use every core this process may run on."""
    import torch
    n = len(os.sched_getaffinity(0))
    torch.set_num_threads(n)
    return torch.get_num_threads()


def run_reference(args, rank):
    """`--impl reference`: rank 0 alone times the reference's CPU path; other ranks exit.  One step = ONE trajectory over
    the full per-GPU batch (a bounded sample of our arm's step, which is a block of such trajectories)."""
    if rank != 0:
        return
    import torch
    cores = use_all_host_threads()
    fname, full_batch, shape, m, cout, _ = CONFIGS[args.config]
    full_batch = args.batch or full_batch
    batch = full_batch
    if args.config not in ("c2",):  # the bigger shapes: largest batch whose (steps+warmup) trajectories fit in about two minutes
        pb = min(full_batch, 64)
        probe = time_cpu(args.config, pb, 1, 1)[0]
        budget = 120.0 / max(1, args.steps + args.warmup)
        batch = int(max(8, min(full_batch, pb * budget / probe / 4)))
        if batch >= 64:
            batch -= batch % 64
    ts = time_cpu(args.config, batch, args.steps, args.warmup)
    total = sum(ts)
    val = batch * args.steps / total
    sample = (f"one trajectory of {batch} samples per step x {args.steps} steps (ours: a block of such trajectories per step); "
              f"torch {torch.__version__} CPU, {cores} threads; {cpu_model()}")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.config, full_batch),
        "cpu_baseline": {"value": val, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "run": {"reference_dtypes": "f64 history / f32 state", "batch_timed": batch},
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
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(index)],
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
        time.sleep(0.1)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows]
        sm, mx, reasons, pw = [], [], set(), []
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_mhz_min": min(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
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


def kernel_source_sha():
    """hash of the sources of the dominant kernel (the specialised step kernels and the shared device helpers)"""
    import re
    h = hashlib.sha256()
    d = os.path.join(ROOT, "naturaldiffusion_b200", "csrc")
    for n in sorted(os.listdir(d)):
        if n.startswith("ni_step_lean") or n == "ni_common.cuh":
            src = open(os.path.join(d, n), "r").read()
            src = re.sub(r"//[^\n]*", "", src)          # comments and layout do not change the kernels
            h.update("".join(src.split()).encode())
    return h.hexdigest()[:16]


def ncu_traffic(label):
    """DRAM bytes per ni_step launch from the committed ncu --set full capture -- only if that capture was taken with THIS
    kernel source (profiles/traffic.json records the source hash); otherwise null rather than a stale constant."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = t.get(label)
        if isinstance(e, dict) and e.get("kernel_source_sha") == kernel_source_sha():
            return e.get("dram_bytes_per_launch"), e.get("source")
    except Exception:
        pass
    return None, None


class Ctx:
    """device, process group and the max-over-ranks helper"""

    def __init__(self, args, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.args, self.rank, self.world, self.local_rank = args, rank, world, local_rank
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        from naturaldiffusion_b200.hostutil import bind_rank_cpus
        self.all_cpus = os.sched_getaffinity(0)
        self.bound_cpus = None if args.no_numa else bind_rank_cpus(local_rank, world)  # before any pinned allocation
        self.relay, self.relay_info = {"d2h": None, "bidir": None}, {"mode": "off"}
        if world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            if not args.no_e2e:
                from naturaldiffusion_b200.hostutil import choose_host_relay
                self.relay, self.relay_info = choose_host_relay(rank, world, self.dev, mode=args.relay)

    def copy_ceiling_ms(self, h2d, d2h, reps, with_h2d, peer=None):
        """bare cudaMemcpyAsync of the same bytes, all ranks at once, ms per batch (max over ranks); peer = relay route"""
        from naturaldiffusion_b200.hostutil import HostCopyRig
        rig = HostCopyRig(self.dev, None if peer is None else self.torch.device("cuda", peer), h2d, d2h)
        rig.run(4, with_h2d)
        self.barrier()
        ms = rig.run(reps, with_h2d)
        self.barrier()
        return self.max_over_ranks(ms) / reps

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.item()


def build_workload(ctx, cfg, batch, markov, eps0):
    import naturaldiffusion_b200 as ni
    from naturaldiffusion_b200 import coeffs, generators
    from naturaldiffusion_b200.ops import philox_normal
    from naturaldiffusion_b200.sampler import NaturalInferenceSampler
    torch = ctx.torch
    fname, dbatch, shape, m, cout, dts = CONFIGS[cfg]
    batch = batch or dbatch
    dtype = {"f32": torch.float32, "f16": torch.float16}[dts]
    mk = {"auto": "auto", "0": False, "1": True}[markov]
    if cfg in ("c2", "c3"):
        triple = ni.CoeffTriple.from_npz(os.path.join(WEIGHTS, fname))
        io = ni.io_score_vp(triple.node)
    elif cfg == "c4":
        triple = generators.ddpm_triple(250)
        c1, c2, _ = coeffs.ddim_x0_coeffs(250)
        io = ni.io_eps_cfg(c1, c2, 4.0)
    else:
        sig = coeffs.flow_match_sigmas(28)
        triple = ni.CoeffTriple.from_sd3_csv(os.path.join(WEIGHTS, fname), sig)
        io = ni.io_velocity_cfg(sig, 7.0)
    sampler = NaturalInferenceSampler(triple, io, batch, shape, device=ctx.dev, dtype=dtype, seed=888, eps0=eps0,
                                      sample_offset=ctx.rank * batch, advance=ctx.world * batch, markov=mk)
    # null denoiser: m pre-generated N(0,1) model-output tensors [B, cout, H, W], re-read from HBM every step
    outs = tuple(philox_normal((batch, cout) + shape[1:], seed=888, tensor_id=1000 + i, elem_offset=ctx.rank * batch * cout * shape[1] * shape[2],
                               dtype=dtype, device=ctx.dev) for i in range(m))
    den = (lambda x, k: outs[0]) if m == 1 else (lambda x, k: outs)
    noise = philox_normal((batch,) + shape, seed=888, tensor_id=0, elem_offset=sampler.elem_offset, dtype=dtype, device=ctx.dev)
    esize = torch.empty(0, dtype=dtype).element_size()
    units = sampler.plan.total_units(m, eps0_stored=eps0 == "stored")
    return dict(sampler=sampler, den=den, outs=outs, noise=noise, batch=batch, shape=shape, m=m, cout=cout, dts=dts, dtype=dtype, K=triple.K,
                esize=esize, units=units, bytes_per_traj=units * sampler.numel * esize, launches_per_traj=sampler.kernel_launches_per_trajectory,
                pixels=cfg in ("c2", "c3"))


def time_resident(ctx, w, steps, warmup, block, graph=True, clocks=False):
    """`steps` steps of `block` trajectories each, inputs resident; returns (ms total max over ranks, clock record, burst ms per trajectory)."""
    import naturaldiffusion_b200 as ni
    torch = ctx.torch
    s, den, noise = w["sampler"], w["den"], w["noise"]
    c0, l0 = ni.launch_count(), ni._lib.lean_launch_count()
    s.sample(den, noise=noise)
    torch.cuda.synchronize()
    w["counted_launches"], w["lean_launches"] = ni.launch_count() - c0, ni._lib.lean_launch_count() - l0
    w["flavours"] = s.load_flavours()
    if graph:
        s.capture(den, noise=noise)
        run = s.replay
    else:
        run = lambda: s.sample(den, noise=noise)
    # burst: a ~10 ms window from an idle (cool, unthrottled) GPU, before the sustained run heats it
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    time.sleep(0.3)
    nb = max(2, min(block, int(0.010 / max(1e-6, w["bytes_per_traj"] / 6.5e12))))
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record()
    for _ in range(nb):
        run()
    b1.record()
    torch.cuda.synchronize()
    burst_ms = b0.elapsed_time(b1) / nb
    for _ in range(max(3, warmup)):
        for _ in range(block):
            run()
    ctx.barrier()
    clk = ClockSampler(ctx.local_rank) if (clocks and ctx.rank == 0) else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        for _ in range(block):
            run()
    e1.record()
    ctx.barrier()
    t1 = time.perf_counter()
    ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    return ms, (clk.stop(t0, t1) if clk else None), burst_ms


def cold_l2(ctx, w):
    """Every launch timed on its own with the L2 flushed right before it (a 512 MB READ sweep, which leaves clean lines --
    a memset would leave 126 MB of dirty lines to be written back during the timed kernel), so no launch can find the
    previous step's stores in the 126 MB L2 (the null denoiser leaves nothing between steps)."""
    torch = ctx.torch
    s, den, noise, outs, m, K = w["sampler"], w["den"], w["noise"], w["outs"], w["m"], w["K"]
    flush = torch.zeros(128 << 20, dtype=torch.float32, device=ctx.dev)
    s.sample(den, noise=noise)
    evs = []
    n_cold = 5 if K <= 30 else 1
    st_ptr = torch.cuda.current_stream(ctx.dev).cuda_stream
    for _ in range(n_cold):
        for k in range(K):
            flush.max()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            s.step(k, outs[0] if m == 1 else outs, st_ptr)
            b_.record()
            evs.append((a_, b_))
    torch.cuda.synchronize()
    cold_ms = sum(a_.elapsed_time(b_) for a_, b_ in evs) / n_cold
    ach = w["bytes_per_traj"] / (cold_ms * 1e-3) / 1e9
    return {"ms_per_trajectory": cold_ms, "achieved": ach, "frac": ach / measured_peak()[0],
            "how": "each launch bracketed by its own CUDA events after a 512 MB read sweep (L2 flush); includes ~2 us event overhead per launch, no graph / PDL overlap"}


def e2e_arm(ctx, w, n_batches, graph=True, ceiling=True):
    """End to end through the public API with HOST buffers.  Returns the `e2e` object (value = host-noise path)."""
    torch = ctx.torch
    s, den, batch, shape, dts, dtype, pixels = w["sampler"], w["den"], w["batch"], w["shape"], w["dts"], w["dtype"], w["pixels"]
    noise_h = torch.empty((batch,) + shape, dtype=dtype).pin_memory()
    noise_h.copy_(w["noise"])
    out_h = (torch.empty((batch, shape[1], shape[2], shape[0]), dtype=torch.uint8) if pixels else torch.empty((batch,) + shape, dtype=dtype)).pin_memory()
    noise_hs = [noise_h, noise_h.clone().pin_memory()]
    out_hs = [out_h, out_h.clone().pin_memory()]
    h2d, d2h = noise_h.numel() * noise_h.element_size(), out_h.numel() * out_h.element_size()

    def timed(noise_list_fn, n):
        s.sample_host_many(den, noise_list_fn(4), [out_hs[i % 2] for i in range(4)], pixels=pixels, first_sample=ctx.rank * batch, graph=graph)
        ctx.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s.sample_host_many(den, noise_list_fn(n), [out_hs[i % 2] for i in range(n)], pixels=pixels, graph=graph)
        b.record()
        ctx.barrier()
        return ctx.max_over_ranks(a.elapsed_time(b))

    s.set_host_relay(ctx.relay["bidir"])
    host_ms = timed(lambda n: [noise_hs[i % 2] for i in range(n)], n_batches)
    s.set_host_relay(ctx.relay["d2h"])
    dev_ms = timed(lambda n: None, n_batches)
    s.set_host_relay(None)
    agg = ctx.world * batch * n_batches
    e2e = {"value": agg / (host_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "batches": n_batches, "ms_per_batch": host_ms / n_batches, "cuda_graph_per_batch": graph,
           "bytes_are_per": f"batch of {batch} samples per GPU (one trajectory; the e2e region is {n_batches} such batches, not bench steps)",
           "device_noise": {"value": agg / (dev_ms * 1e-3), "unit": "samples/s", "ms_per_batch": dev_ms / n_batches, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": d2h,
                            "note": "aggregated over all ranks (max-over-ranks time); noise drawn on the device from (seed, global sample index) as the "
                                    "reference does with torch.randn on the GPU (src/CIFAR10NaturalInference.py:290): only the result crosses PCIe"},
           "api": f"NaturalInferenceSampler.sample_host_many(pixels={pixels}, graph={graph}), double-buffered copy streams: pinned {dts} noise in, "
                  + ("NHWC uint8 out" if pixels else f"{dts} latent out")}
    e2e["host_route"] = {"host_noise": ctx.relay["bidir"], "device_noise": ctx.relay["d2h"], "probe": ctx.relay_info,
                         "what": "GPU index this rank's host copies are relayed through over NVLink (null = its own PCIe link); decided per box by "
                                 "hostutil.choose_host_relay from the bandwidths in `probe` (rank 0's route shown; the probe lists every pair)"}
    if ceiling:
        # bare copies of the same bytes, all ranks at once: what the host/PCIe side allows at this N -- over every GPU's own
        # link ("copy_ceiling") and, when a relay route is in use, over the route the pipeline took ("copy_ceiling_routed")
        n_c = max(16, n_batches // 2)
        c_ms = ctx.copy_ceiling_ms(h2d, d2h, n_c, True)
        ceil_sps = ctx.world * batch / (c_ms * 1e-3)
        e2e["copy_ceiling"] = {"samples_per_s": ceil_sps, "gbs_aggregate": ctx.world * (h2d + d2h) / (c_ms * 1e-3) / 1e9, "ms_per_batch": c_ms,
                               "how": f"bare cudaMemcpyAsync of the same {h2d} B H2D + {d2h} B D2H per batch on two copy streams, every GPU over its own PCIe link, "
                                      f"all {ctx.world} rank(s) at once, max over ranks"}
        e2e["frac_of_copy_ceiling"] = e2e["value"] / ceil_sps
        e2e["copy_ceiling_gbs"] = e2e["copy_ceiling"]["gbs_aggregate"]
        # the device-noise variant only returns results: its ceiling is the D2H direction alone
        d_ms = ctx.copy_ceiling_ms(h2d, d2h, n_c, False)
        d_sps = ctx.world * batch / (d_ms * 1e-3)
        e2e["device_noise"]["copy_ceiling"] = {"samples_per_s": d_sps, "gbs_aggregate": ctx.world * d2h / (d_ms * 1e-3) / 1e9, "ms_per_batch": d_ms,
                                               "how": f"bare cudaMemcpyAsync D2H of the same {d2h} B per batch, every GPU over its own PCIe link, all {ctx.world} rank(s) at once"}
        e2e["device_noise"]["frac_of_copy_ceiling"] = e2e["device_noise"]["value"] / d_sps
        if ctx.world > 1 and ctx.relay_info.get("pairs") and "relayed_gbs_per_rank" in ctx.relay_info:  # (all ranks agree: the info is built from gathered numbers)
            r_ms = ctx.copy_ceiling_ms(h2d, d2h, n_c, True, peer=ctx.relay["bidir"])
            rd_ms = ctx.copy_ceiling_ms(h2d, d2h, n_c, False, peer=ctx.relay["d2h"])
            e2e["copy_ceiling_routed"] = {"samples_per_s": ctx.world * batch / (r_ms * 1e-3), "gbs_aggregate": ctx.world * (h2d + d2h) / (r_ms * 1e-3) / 1e9, "ms_per_batch": r_ms}
            e2e["frac_of_copy_ceiling_routed"] = e2e["value"] / e2e["copy_ceiling_routed"]["samples_per_s"]
            e2e["device_noise"]["copy_ceiling_routed"] = {"samples_per_s": ctx.world * batch / (rd_ms * 1e-3), "gbs_aggregate": ctx.world * d2h / (rd_ms * 1e-3) / 1e9, "ms_per_batch": rd_ms}
            e2e["device_noise"]["frac_of_copy_ceiling_routed"] = e2e["device_noise"]["value"] / e2e["device_noise"]["copy_ceiling_routed"]["samples_per_s"]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # and with nothing crossing PCIe at all: noise drawn on the device, images consumed on the device (the flow of
        # examples/cifar10_pipeline.py: uint8 images -> features -> FID statistics on the same GPU) -- what the pipeline itself scales like
        if graph:
            ob0 = torch.empty(out_h.shape, dtype=out_h.dtype, device=ctx.dev)
            kw = dict(pixels_out=ob0) if pixels else dict(out=ob0)
            g = s.capture(den, **kw)
            for _ in range(4):
                s.replay(g)
            ctx.barrier()
            a.record()
            for _ in range(n_batches):
                s.replay(g)
            b.record()
            ctx.barrier()
            r_ms = ctx.max_over_ranks(a.elapsed_time(b))
            e2e["device_resident_results"] = {"value": agg / (r_ms * 1e-3), "unit": "samples/s", "ms_per_batch": r_ms / n_batches, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                                              "note": "informational, NOT an end-to-end number: device noise, results left on the GPU for on-device evaluation"}
    return e2e


def e2e_score_network_arm(ctx, batch=512, n_batches=2):
    """End to end with HOST buffers and a REAL score network in the loop: pinned fp32 noise H2D -> K x (NCSN++ forward in torch +
    one fused ni_step) -> fused uint8 stage -> D2H, through NaturalInferenceSampler.sample_host_many.  The network is the
    reference's CIFAR-10 architecture (deps/score_sde_pytorch configs/vp/cifar10_ddpmpp_continuous.py, 61.8 M parameters) with
    random weights (no checkpoint offline); its forward stays torch by the north-star and dominates the time, so this is
    not a kernel number -- it is the end-to-end pipeline whose 1 -> 8 GPU scaling the north-star asks for, where the null
    denoiser of `e2e.value` is a PCIe stress test instead."""
    import naturaldiffusion_b200 as ni
    from naturaldiffusion_b200.adapters import ncsnpp_denoiser
    from naturaldiffusion_b200.denoisers import NCSNppVP
    from naturaldiffusion_b200.sampler import NaturalInferenceSampler
    torch = ctx.torch
    triple = ni.CoeffTriple.from_npz(os.path.join(WEIGHTS, CONFIGS["c2"][0]))
    torch.manual_seed(0)
    model = NCSNppVP().reinit_output().to(ctx.dev).eval()
    den = ncsnpp_denoiser(model, triple.node)
    s = NaturalInferenceSampler(triple, ni.io_score_vp(triple.node), batch, (3, 32, 32), device=ctx.dev, seed=888,
                                sample_offset=ctx.rank * batch, advance=ctx.world * batch)
    noise_h = [torch.randn(s.full_shape()).pin_memory() for _ in range(2)]
    out_h = [torch.empty((batch, 32, 32, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
    run = lambda n: s.sample_host_many(den, [noise_h[i % 2] for i in range(n)], [out_h[i % 2] for i in range(n)], pixels=True)
    with torch.no_grad():
        run(1)
        ctx.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run(n_batches)
        b.record()
        ctx.barrier()
    ms = ctx.max_over_ranks(a.elapsed_time(b)) / n_batches
    n_par = sum(p.numel() for p in model.parameters())
    res = {"value": ctx.world * batch / (ms * 1e-3), "unit": "samples/s", "ms_per_batch": ms, "batch_per_gpu": batch, "batches": n_batches,
           "h2d_bytes_per_step": noise_h[0].numel() * 4, "d2h_bytes_per_step": out_h[0].numel(), "K": triple.K,
           "denoiser": f"NCSN++ VP (random init, {n_par / 1e6:.1f} M parameters, fp32, eager torch)",
           "note": "informational: the torch forward dominates (the fused step is < 1 % of the batch); reported because it is the end-to-end "
                   "pipeline that scales with GPUs, while the null-denoiser e2e above is bound by the host's copy bandwidth"}
    del s, model
    torch.cuda.empty_cache()
    return res


def fid_arm(ctx):
    """The evaluation step behind the sampler: fp64 sufficient statistics of 2048-d activations on the fp64
    tensor cores (csrc/ni_fid.cu) and, at N > 1, the ONE collective of the north-star -- the all-reduce of the 34 MB
    statistics buffer over NCCL."""
    from naturaldiffusion_b200.fid import FidAccumulator
    torch = ctx.torch
    m, d = 8192, 2048
    x = torch.randn(m, d, device=ctx.dev)
    acc = FidAccumulator(dim=d, device=ctx.dev)
    for _ in range(2):
        acc.update(x)
    ctx.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        acc.update(x)
    b.record()
    ctx.barrier()
    ms = ctx.max_over_ranks(a.elapsed_time(b)) / 5
    out = {"activations": [m, d], "ms_per_update": ms, "fp64_tflops_symmetric": m * d * (d + 1.0) / ms / 1e9,
           "fp64_tflops_full_gemm_equivalent": 2.0 * m * d * d / ms / 1e9, "images_per_s": ctx.world * m / (ms * 1e-3),
           "kernel": "ni_fid_syrk_kernel (DMMA m8n8k4, upper-triangular 64x64 tiles + mirror) + ni_fid_colsum_kernel"}
    if ctx.world > 1:
        acc.all_reduce()
        ctx.barrier()
        a.record()
        for _ in range(5):
            acc.all_reduce()
        b.record()
        ctx.barrier()
        ar = ctx.max_over_ranks(a.elapsed_time(b)) / 5
        nbytes = acc.buf.numel() * 8
        out["allreduce"] = {"bytes": nbytes, "ms": ar, "bus_gbs": 2.0 * (ctx.world - 1) / ctx.world * nbytes / (ar * 1e-3) / 1e9, "ranks": ctx.world,
                            "what": "torch.distributed.all_reduce(SUM) of [n | sum x | sum x x^T] fp64 over NCCL, once per evaluated set"}
    return out


def samplers_via_matrix(ctx, batch=4096):
    """"Original sampler vs Natural Inference" on ONE code path: 15-model-call samplers on the CIFAR shape, each as its
    coefficient matrix through the same fused step kernel with the null denoiser -- the reference's optimised banded
    matrix next to the (dense) matrices of the original solvers that head results/FID/*_15step.csv."""
    import naturaldiffusion_b200 as ni
    from naturaldiffusion_b200 import generators as G
    from naturaldiffusion_b200.ops import philox_normal
    from naturaldiffusion_b200.sampler import NaturalInferenceSampler
    torch = ctx.torch
    shape = (3, 32, 32)
    out_t = philox_normal((batch,) + shape, seed=888, tensor_id=1000, device=ctx.dev)
    noise = philox_normal((batch,) + shape, seed=888, tensor_id=0, device=ctx.dev)
    den = lambda x, k: out_t
    cases = [("NI optimised step_15_weight_173", ni.CoeffTriple.from_npz(os.path.join(WEIGHTS, "step_15_weight_173.npz"))),
             ("DPM-Solver++ multistep-2 (2M)", G.dpm_solver_triple(15, "dpmsolver++", "multistep", 2)),
             ("DPM-Solver++ multistep-3", G.dpm_solver_triple(15, "dpmsolver++", "multistep", 3)),
             ("DPM-Solver multistep-3", G.dpm_solver_triple(15, "dpmsolver", "multistep", 3)),
             ("DPM-Solver++ singlestep-3", G.dpm_solver_triple(15, "dpmsolver++", "singlestep", 3)),
             ("DEIS tAB3", G.deis_triple(15, "t_ab")), ("DEIS rhoAB3", G.deis_triple(15, "rho_ab")), ("DEIS iPNDM", G.deis_triple(15, "ipndm")),
             ("DEIS rhoRK-3kutta (5 steps x 3 calls)", G.deis_triple(5, "rho_rk")),
             ("DDIM-15 (first-order path)", G.ddim_triple(15))]
    res = []
    for name, t in cases:
        io = ni.io_score_vp(t.node) if t.node[0, 0] <= 1.0 else ni.io_eps_cfg(*ni.coeffs.ddim_x0_coeffs(15)[:2], None)
        s = NaturalInferenceSampler(t, io, batch, shape, device=ctx.dev, seed=888)
        s.capture(den, noise=noise)
        for _ in range(5):
            s.replay()
        torch.cuda.synchronize()
        n = 40
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            s.replay()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        units = s.plan.total_units(1)
        res.append({"sampler": name, "K": t.K, "ms_per_trajectory": ms, "tensor_transfers": units, "x0_ring_slots": s.plan.n_x0_slots,
                    "first_order_path": bool(s.plan.markov), "achieved_gbs": units * s.numel * 4 / ms / 1e6, "samples_per_s": batch / (ms * 1e-3)})
        del s
    return res


def cross_rank_check(ctx):
    """N > 1: prove bits, not just speed.  (i) every rank hashes the x_K of its shard of one global batch (deterministic C2
    matrix, in-kernel initial noise) and of a stochastic DDPM-20 run (in-kernel fresh noise every step); rank 0 recomputes
    the LAST rank's shard on its own GPU from `sample_offset` alone and compares hashes.  (ii) the one collective of the
    path: FID statistics all-reduced over NCCL vs numpy mean/cov of the gathered features."""
    import naturaldiffusion_b200 as ni
    from naturaldiffusion_b200 import coeffs, generators
    from naturaldiffusion_b200.fid import FidAccumulator
    from naturaldiffusion_b200.sampler import NaturalInferenceSampler
    import numpy as np
    torch, dist = ctx.torch, ctx.dist
    den = lambda x, k: torch.tanh(0.7 * x) * (1.0 + 0.01 * k) + 0.1 * x
    B = 512
    t_c2 = ni.CoeffTriple.from_npz(os.path.join(WEIGHTS, "step_10_weight_42.npz"))
    t_dd = generators.ddpm_triple(20)
    c1, c2, _ = coeffs.ddim_x0_coeffs(20)

    def run(rank):
        a = NaturalInferenceSampler(t_c2, ni.io_score_vp(t_c2.node), B, (3, 32, 32), device=ctx.dev, seed=888, sample_offset=rank * B)
        b = NaturalInferenceSampler(t_dd, ni.io_eps_cfg(c1, c2, None), B, (4, 16, 16), device=ctx.dev, seed=5, sample_offset=rank * B)
        return a.sample(den).clone(), b.sample(den).clone()

    def digest(x):
        v = x.contiguous().view(torch.int32).to(torch.int64).flatten()
        return int(((v * (torch.arange(v.numel(), device=v.device) % 1000003 + 1)).sum() & 0x7FFFFFFFFFFFFFFF).item())

    xa, xb = run(ctx.rank)
    mine = torch.tensor([digest(xa), digest(xb)], device=ctx.dev, dtype=torch.int64)
    allh = [torch.zeros_like(mine) for _ in range(ctx.world)]
    dist.all_gather(allh, mine)
    # FID statistics: features = 64 fixed random projections of the image
    g = torch.Generator(device=ctx.dev).manual_seed(0)
    P = torch.randn(3 * 32 * 32, 64, device=ctx.dev, generator=g) / 55.0
    feats = xa.flatten(1) @ P
    acc = FidAccumulator(dim=64, device=ctx.dev).update(feats).all_reduce()
    gathered = [torch.zeros_like(feats) for _ in range(ctx.world)]
    dist.all_gather(gathered, feats)
    res = None
    if ctx.rank == 0:
        fa, fb = run(ctx.world - 1)
        mu, sigma = acc.finalize()
        allf = torch.cat(gathered).double().cpu().numpy()
        mu_ref, sig_ref = allf.mean(0), np.cov(allf, rowvar=False)
        res = {"shard_hash_match": bool(digest(fa) == int(allh[-1][0]) and digest(fb) == int(allh[-1][1])),
               "what": f"rank 0 recomputed rank {ctx.world - 1}'s shard (C2 matrix + DDPM-20 with in-kernel fresh noise, batch {B}/rank) from sample_offset alone",
               "fid_allreduce_max_rel_err": float(max(np.abs(mu - mu_ref).max() / np.abs(mu_ref).max(), np.abs(sigma - sig_ref).max() / np.abs(sig_ref).max())),
               "fid_samples": int(acc.n), "nccl_ranks": ctx.world}
    ctx.barrier()
    return res


def run_ours(args, rank, world, local_rank):
    import naturaldiffusion_b200 as ni
    from naturaldiffusion_b200 import _lib as nilib
    ctx = Ctx(args, rank, world, local_rank)
    torch = ctx.torch
    nilib.set_option("variant", args.variant)
    for kv in args.opt:
        name, val = kv.split("=")
        nilib.set_option(name, int(val))
    peak, peak_src = measured_peak()

    def block_for(w):
        est_ms = w["bytes_per_traj"] / 6.5e12 * 1e3
        return max(1, int(-(-60.0 // est_ms)))

    # ---- headline workload
    w = build_workload(ctx, args.config, args.batch, args.markov, args.eps0)
    block = args.block or block_for(w)
    ms, clk, burst_ms = time_resident(ctx, w, args.steps, args.warmup, block, graph=not args.no_graph, clocks=True)
    assert w["counted_launches"] == w["launches_per_traj"], (w["counted_launches"], w["launches_per_traj"])
    n_traj = args.steps * block
    value = world * w["batch"] * n_traj / (ms * 1e-3)
    ms_per_traj = ms / n_traj
    achieved = w["bytes_per_traj"] / (ms_per_traj * 1e-3) / 1e9  # GB/s per GPU == per launch (only ni_step kernels run)
    cold = None
    if world == 1 and not args.no_cold and (args.config in ("c2", "c3") or args.cold):
        cold = cold_l2(ctx, w)
    e2e = None if args.no_e2e else e2e_arm(ctx, w, max(64, args.e2e_batches))

    eager = None
    if args.eager_comparator and args.config in ("c2", "c3") and world == 1:
        eager = eager_comparator(ctx, w, ms_per_traj)

    flav = w["flavours"]
    K, launches_per_traj = w["K"], w["launches_per_traj"]
    traffic, traffic_src = ncu_traffic(f"{args.config}:{args.eps0}")
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": w["dts"],
        "data": "synthetic",
        "config": workload_config(args.config, w["batch"]),
        "run": {"trajectories_per_step": block, "samples_per_step_per_gpu": block * w["batch"], "launches_per_step": block * launches_per_traj,
                "ms_per_trajectory": ms_per_traj, "markov_fast_path": bool(w["sampler"].plan.markov), "eps0": args.eps0,
                "cpus_bound_to_this_rank": ctx.bound_cpus, "cuda_graph": not args.no_graph, "variant": args.variant, "opts": args.opt,
                "specialised_kernel_launches_per_trajectory": w["lean_launches"],
                "load_flavours_per_step": "".join(str(f) for f in flav) if K <= 32 else f"{sum(flav)} of {K} steps streaming",
                "load_flavour": "1 = plain ld.global, 0 = L1::no_allocate; auto per launch (csrc/ni_kernels.cu launch_streams)",
                "working_set_mb": (w["sampler"].state_bytes() + sum(o.numel() for o in w["outs"]) * w["esize"]) / 1e6,
                "kernel_source_sha": kernel_source_sha()},
        "clocks": clk,
        "e2e": e2e,
        "gpu_launches": launches_per_traj * n_traj,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "kernel": {0: "ni_step_lean_kernel (row-shape-specialised direct loads)", 1: "ni_step_kernel (generic direct loads)", 2: "ni_step_tma_kernel"}[args.variant],
                     "algorithmic_bytes_per_launch": w["bytes_per_traj"] / launches_per_traj,
                     "tensor_transfers_per_trajectory": w["units"], "us_per_launch": 1e3 * ms_per_traj / launches_per_traj,
                     "frac_of_nominal_8TBs": achieved / 8000.0,
                     "burst": {"achieved": w["bytes_per_traj"] / (burst_ms * 1e-3) / 1e9, "ms_per_trajectory": burst_ms,
                               "how": "~10 ms window from an idle GPU before the sustained run (rank 0's own GPU)"},
                     "cold_l2": cold},
    }

    # ---- the other BASELINE shapes
    if not args.no_per_config and args.config == "c2":
        only = set(x for x in args.only.split(",") if x)
        per = {}
        for cfg, b, mk, label in PER_CONFIG:
            if only and label not in only:
                continue
            del w
            torch.cuda.empty_cache()
            w = build_workload(ctx, cfg, b, mk, "stored")
            blk = block_for(w)
            st = 6 if w["K"] <= 30 else 3
            pms, _, pburst = time_resident(ctx, w, st, 3, blk, graph=True)
            ntr = st * blk
            ach = w["bytes_per_traj"] / (pms / ntr * 1e-3) / 1e9
            ent = {"workload": workload_config(cfg, w["batch"])["workload"], "markov_fast_path": bool(w["sampler"].plan.markov),
                   "ms_per_trajectory": pms / ntr, "samples_per_s": world * w["batch"] * ntr / (pms * 1e-3), "achieved": ach, "frac": ach / peak,
                   "frac_of_nominal_8TBs": ach / 8000.0, "tensor_transfers_per_trajectory": w["units"], "launches_per_trajectory": w["launches_per_traj"],
                   "us_per_launch": 1e3 * pms / ntr / w["launches_per_traj"], "specialised_kernel_launches_per_trajectory": w["lean_launches"],
                   "burst_achieved": w["bytes_per_traj"] / (pburst * 1e-3) / 1e9, "timed_ms": pms}
            if world == 1 and not args.no_cold and cfg == "c3":
                ent["cold_l2"] = cold_l2(ctx, w)
            if not args.no_e2e:
                nb = 24 if w["K"] <= 30 else 6
                try:
                    ee = e2e_arm(ctx, w, nb, ceiling=True)
                    ent["e2e"] = {k: ee[k] for k in ("value", "h2d_bytes_per_step", "d2h_bytes_per_step", "batches", "ms_per_batch", "frac_of_copy_ceiling")}
                    ent["e2e"]["device_noise"] = ee["device_noise"]["value"]
                    ent["e2e"]["device_noise_frac_of_d2h_ceiling"] = ee["device_noise"]["frac_of_copy_ceiling"]
                    ent["e2e"]["copy_ceiling_gbs"] = ee["copy_ceiling"]["gbs_aggregate"]
                    ent["e2e"]["d2h_ceiling_gbs"] = ee["device_noise"]["copy_ceiling"]["gbs_aggregate"]
                    if "copy_ceiling_routed" in ee:  # a relay route is in use at this N (e2e.host_route)
                        ent["e2e"]["relayed_via_gpu"] = {"host_noise": ee["host_route"]["host_noise"], "device_noise": ee["host_route"]["device_noise"]}
                        ent["e2e"]["frac_of_copy_ceiling_routed"] = ee["frac_of_copy_ceiling_routed"]
                        ent["e2e"]["device_noise_frac_of_d2h_ceiling_routed"] = ee["device_noise"]["frac_of_copy_ceiling_routed"]
                        ent["e2e"]["copy_ceiling_routed_gbs"] = ee["copy_ceiling_routed"]["gbs_aggregate"]
                        ent["e2e"]["d2h_ceiling_routed_gbs"] = ee["device_noise"]["copy_ceiling_routed"]["gbs_aggregate"]
                except (KeyError, ValueError, TypeError, AttributeError) as ex:  # a secondary line must not cost the headline
                    ent["e2e"] = {"error": f"{type(ex).__name__}: {ex}"}
            per[label] = ent
        line["per_config"] = per
        # C3 is the L2-clean roofline shape (201 MB tensors): quoted beside the headline
        if "c3" in per:
            line["roofline"]["c3"] = {k: per["c3"][k] for k in ("achieved", "frac", "frac_of_nominal_8TBs", "us_per_launch")}

    if world > 1 and not args.no_check:
        line["multi_gpu_check"] = cross_rank_check(ctx)
    if not args.no_per_config and args.config == "c2":
        del w
        torch.cuda.empty_cache()
        line["fid"] = fid_arm(ctx)
        if world == 1:
            line["samplers_via_matrix"] = samplers_via_matrix(ctx)

    if e2e is not None and not args.no_score_network and args.config == "c2":
        torch.cuda.empty_cache()
        try:
            e2e["with_score_network"] = e2e_score_network_arm(ctx)
        except Exception as ex:  # informational arm: say so in the line instead of losing the line
            e2e["with_score_network"] = {"error": f"{type(ex).__name__}: {ex}"}

    if rank != 0:
        if world > 1:
            ctx.dist.destroy_process_group()
        return

    if not args.no_cpu_baseline and world == 1:
        os.sched_setaffinity(0, ctx.all_cpus)  # the CPU arm gets every host core back, not only this rank's slice
        cores = use_all_host_threads()
        cb = {"c2": CONFIGS["c2"][1], "c3": 4096, "c4": 32, "c5": 4, "c5s": 4}[args.config]
        reps = {"c2": 10, "c3": 3}.get(args.config, 2)  # about 10 s of CPU work
        ts = time_cpu(args.config, cb, reps, 1)
        line["cpu_baseline"] = {"value": cb / min(ts), "unit": "samples/s", "cores": cores, "kind": "port",
                                "sample": f"{cb}-sample batch of the same workload, best of {reps} trajectories after 1 warm-up "
                                          f"({sum(ts):.1f} s CPU work); oracle port of the reference loop in the reference's dtypes; torch {torch.__version__} CPU, {cpu_model()}"}
    else:
        line["cpu_baseline"] = None
    if eager is not None:
        line["reference_eager_gpu"] = eager
    print(json.dumps(line))
    if world > 1:
        ctx.dist.destroy_process_group()


def eager_comparator(ctx, w, ms_per_traj):
    """the reference's own loop structure (oracle restatement: fp64 history, one torch kernel per op, per-step H2D scalars)
    on THIS GPU with the same null denoiser -- "eager torch on B200", SURVEY 2.3"""
    from oracle import ni_oracle as O
    torch = ctx.torch
    fname = CONFIGS[ctx.args.config][0]
    A_, B_, node_ = O.load_triple(os.path.join(WEIGHTS, fname))
    noise, outs, batch, K, dev = w["noise"], w["outs"], w["batch"], w["K"], ctx.dev

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
    mine = w["sampler"].sample(w["den"], noise=noise)
    return {"ms_per_trajectory": gms, "value": batch / (gms * 1e-3), "unit": "samples/s", "speedup_of_fused_step": gms / ms_per_traj,
            "max_abs_diff_over_norm": float((mine - xe).abs().max() / xe.norm()),
            "what": "reference loop structure (fp64 history, eager torch ops) on the same B200, same null denoiser"}


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
