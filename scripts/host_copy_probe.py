#!/usr/bin/env python
"""What bounds the host side of the end-to-end path at N GPUs?  (at 8 GPUs the null-denoiser e2e sits on the
box's copy ceiling.)  Bare cudaMemcpyAsync experiments, one process per GPU under torchrun, per-rank numbers kept (not only
the max): every rank alone, socket halves, PCIe-switch pairs, all ranks at once; pinned memory from cudaHostAlloc vs
transparent-huge-page backed + cudaHostRegister; small LLC-sized rings; two streams per rank; bigger chunks; both directions.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/host_copy_probe.py
"""
import ctypes
import json
import mmap
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

D2H_CHUNK = 4096 * 32 * 32 * 3          # one C2 batch of uint8 images
H2D_CHUNK = 4096 * 3 * 32 * 32 * 4      # one C2 batch of fp32 noise


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as ex:  # noqa: BLE001
        return f"<{ex}>"


class ThpPinned:
    """2 MB aligned anonymous memory with MADV_HUGEPAGE, touched, then cudaHostRegister'ed"""

    def __init__(self, nbytes):
        self.n = (nbytes + (2 << 20) - 1) // (2 << 20) * (2 << 20)
        self.mm = mmap.mmap(-1, self.n + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        base = ctypes.addressof(ctypes.c_char.from_buffer(self.mm))
        self.off = (-base) % (2 << 20)
        self.ptr = base + self.off
        libc = ctypes.CDLL("libc.so.6", use_errno=True)
        libc.madvise.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        self.madvise_rc = libc.madvise(ctypes.c_void_p(self.ptr), self.n, 14)  # MADV_HUGEPAGE
        self.t = torch.frombuffer(self.mm, dtype=torch.uint8, count=self.n, offset=self.off)
        self.t.fill_(1)  # first touch (on this rank's CPUs)
        rt = ctypes.CDLL("libcudart.so.12")
        rt.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
        self.reg_rc = rt.cudaHostRegister(ctypes.c_void_p(self.ptr), self.n, 0)
        self.rt = rt

    def close(self):
        self.rt.cudaHostUnregister(ctypes.c_void_p(self.ptr))


def anon_huge_kb():
    try:
        for ln in open("/proc/self/smaps_rollup"):
            if ln.startswith("AnonHugePages"):
                return int(ln.split()[1])
    except Exception:  # noqa: BLE001
        pass
    return -1


def main():
    rank, world, lr = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    from naturaldiffusion_b200.hostutil import bind_rank_cpus
    bound = bind_rank_cpus(lr, world)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out_path = os.path.join(ROOT, "gpurun_out", f"host_copy_probe_n{world}.json")
    res = {"world": world, "cpus_bound_per_rank": bound, "experiments": {}}
    if rank == 0:
        os.makedirs(os.path.dirname(out_path), exist_ok=True)
        res["env"] = {"topo": sh("nvidia-smi topo -m"), "pcie": sh("nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv,noheader"),
                      "lscpu": sh("lscpu | egrep 'Model name|Socket|NUMA|^CPU\\(s\\)|Thread|L3|Hypervisor'"), "mem": sh("egrep 'MemTotal|Huge|AnonHuge' /proc/meminfo"),
                      "thp": sh("cat /sys/kernel/mm/transparent_hugepage/enabled"), "nodes": sh("ls /sys/devices/system/node | grep node")}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def gather(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world == 1:
            return [x]
        lst = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(lst, t)
        return [float(v) for v in lst]

    def run(name, enqueue, nbytes_per_rep, reps, active=None, streams=None):
        """enqueue(i) issues copy i on its stream(s); per-rank GB/s of the active ranks, aggregate in both conventions"""
        active = list(range(world)) if active is None else active
        try:
            me = rank in active
            streams_ = streams or [torch.cuda.current_stream(dev)]
            if me:
                for i in range(2):
                    enqueue(i)
            barrier()
            ms = 0.0
            if me:
                e0 = [torch.cuda.Event(enable_timing=True) for _ in streams_]
                e1 = [torch.cuda.Event(enable_timing=True) for _ in streams_]
                t0 = time.perf_counter()
                for s, e in zip(streams_, e0):
                    e.record(s)
                for i in range(reps):
                    enqueue(i)
                for s, e in zip(streams_, e1):
                    e.record(s)
                torch.cuda.synchronize()
                ms = (time.perf_counter() - t0) * 1e3 if len(streams_) > 1 else e0[0].elapsed_time(e1[0])
            barrier()
            all_ms = gather(ms)
            if rank == 0:
                per = {r: round(nbytes_per_rep * reps / (all_ms[r] * 1e-3) / 1e9, 2) for r in active}
                worst = max(all_ms[r] for r in active)
                res["experiments"][name] = {"per_rank_gbs": per, "sum_gbs": round(sum(per.values()), 1),
                                            "max_over_ranks_gbs": round(len(active) * nbytes_per_rep * reps / (worst * 1e-3) / 1e9, 1)}
                with open(out_path, "w") as f:
                    json.dump(res, f, indent=1)
        except Exception as ex:  # noqa: BLE001
            if rank == 0:
                res["experiments"][name] = {"error": f"{type(ex).__name__}: {ex}"}

    # ---- buffers
    d_img = [torch.empty(D2H_CHUNK, dtype=torch.uint8, device=dev) for _ in range(2)]
    h_img = [torch.empty(D2H_CHUNK, dtype=torch.uint8).pin_memory() for _ in range(2)]
    d_noise = [torch.empty(H2D_CHUNK, dtype=torch.uint8, device=dev) for _ in range(2)]
    h_noise = [torch.empty(H2D_CHUNK, dtype=torch.uint8).pin_memory() for _ in range(2)]
    for t in h_img + h_noise:
        t.fill_(1)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def d2h(i):
        h_img[i % 2].copy_(d_img[i % 2], non_blocking=True)

    def h2d(i):
        d_noise[i % 2].copy_(h_noise[i % 2], non_blocking=True)

    REPS = 48
    run("d2h_all", d2h, D2H_CHUNK, REPS)
    run("h2d_all", h2d, H2D_CHUNK, 16)
    for r in range(world):
        run(f"d2h_solo_rank{r}", d2h, D2H_CHUNK, 24, active=[r])
    if world >= 4:
        half = world // 2
        run("d2h_first_half", d2h, D2H_CHUNK, REPS, active=list(range(half)))
        run("d2h_second_half", d2h, D2H_CHUNK, REPS, active=list(range(half, world)))
        run("d2h_even", d2h, D2H_CHUNK, REPS, active=list(range(0, world, 2)))
        run("d2h_odd", d2h, D2H_CHUNK, REPS, active=list(range(1, world, 2)))
        run("h2d_first_half", h2d, H2D_CHUNK, 16, active=list(range(half)))
        run("h2d_second_half", h2d, H2D_CHUNK, 16, active=list(range(half, world)))
    if world >= 2:
        for a in range(0, world, 2):
            run(f"d2h_pair_{a}_{a + 1}", d2h, D2H_CHUNK, 24, active=[a, a + 1])
        if world >= 4:
            run("d2h_pair_0_2", d2h, D2H_CHUNK, 24, active=[0, 2])
            run(f"d2h_pair_0_{world - 1}", d2h, D2H_CHUNK, 24, active=[0, world - 1])

    # both directions at once on two streams (what bench.py's e2e.copy_ceiling does)
    def both(i):
        with torch.cuda.stream(s1):
            d_noise[i % 2].copy_(h_noise[i % 2], non_blocking=True)
        with torch.cuda.stream(s2):
            h_img[i % 2].copy_(d_img[i % 2], non_blocking=True)

    run("h2d_plus_d2h_all", both, H2D_CHUNK + D2H_CHUNK, 16, streams=[s1, s2])

    # D2H on two streams per rank
    def d2h_two(i):
        with torch.cuda.stream(s1 if i % 2 == 0 else s2):
            h_img[i % 2].copy_(d_img[i % 2], non_blocking=True)

    run("d2h_all_two_streams", d2h_two, D2H_CHUNK, REPS, streams=[s1, s2])

    # small ring: 8 x 1.5 MB host buffers (fits the last-level cache; DDIO-style)
    small = D2H_CHUNK // 8
    h_small = [torch.empty(small, dtype=torch.uint8).pin_memory() for _ in range(8)]

    def d2h_small(i):
        h_small[i % 8].copy_(d_img[0][:small], non_blocking=True)

    run("d2h_all_small_ring_1p5MB", d2h_small, small, REPS * 8)

    # big chunks: 100 MB
    big_n = 8 * D2H_CHUNK
    d_big = torch.empty(big_n, dtype=torch.uint8, device=dev)
    h_big = torch.empty(big_n, dtype=torch.uint8).pin_memory()
    h_big.fill_(1)
    run("d2h_all_100MB_chunks", lambda i: h_big.copy_(d_big, non_blocking=True), big_n, 6)
    run("h2d_all_100MB_chunks", lambda i: d_big.copy_(h_big, non_blocking=True), big_n, 6)

    # transparent-huge-page backed, cudaHostRegister'ed
    try:
        thp = [ThpPinned(D2H_CHUNK) for _ in range(2)]
        thp_n = [ThpPinned(H2D_CHUNK) for _ in range(2)]
        info = gather(float(anon_huge_kb()))
        regs = gather(float(max(t.reg_rc for t in thp + thp_n)))
        if rank == 0:
            res["thp"] = {"anon_huge_kb_per_rank": info, "cudaHostRegister_rc_max_per_rank": regs, "madvise_rc": thp[0].madvise_rc}
        if max(regs) == 0:
            run("d2h_all_thp_registered", lambda i: thp[i % 2].t[:D2H_CHUNK].copy_(d_img[i % 2], non_blocking=True), D2H_CHUNK, REPS)
            run("h2d_all_thp_registered", lambda i: d_noise[i % 2].copy_(thp_n[i % 2].t[:H2D_CHUNK], non_blocking=True), H2D_CHUNK, 16)
    except Exception as ex:  # noqa: BLE001
        if rank == 0:
            res["thp"] = {"error": f"{type(ex).__name__}: {ex}"}

    # repeat the headline at the end (drift check)
    run("d2h_all_again", d2h, D2H_CHUNK, REPS)
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(res, f, indent=1)
        print(json.dumps({k: (v.get("sum_gbs"), v.get("max_over_ranks_gbs")) if "error" not in v else v for k, v in res["experiments"].items()}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
