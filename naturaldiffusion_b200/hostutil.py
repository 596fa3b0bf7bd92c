"""Host-side placement helpers for the end-to-end (host-buffer) path.  Pure OS plumbing, no compute."""
from __future__ import annotations

import os
from typing import Optional


def gpu_local_cpus(device_index: int) -> Optional[set]:
    """CPUs on the NUMA node the GPU's PCIe root hangs off (sysfs `local_cpulist`), or None if unknown."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        return cpus or None
    except Exception:
        return None


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process to the CPUs local to `device_index` BEFORE allocating pinned host buffers, so that the staging
    buffers are first-touched on the GPU's NUMA node and H2D/D2H DMA does not cross the socket interconnect.  With one
    process per GPU on an 8-GPU box this is what lets the host-buffer path scale.  Returns the number of CPUs bound."""
    cpus = gpu_local_cpus(device_index)
    if not cpus:
        return None
    try:
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except OSError:
        return None


def bind_rank_cpus(local_rank: int, local_world: int) -> Optional[int]:
    """One process per GPU: give THIS rank its own disjoint slice of CPUs, taken from the NUMA node of its GPU and shared
    out among the ranks whose GPUs hang off the same node (round 1 pinned all 8 ranks of a box to the same 32 CPUs, where
    their launch threads, copy-engine callbacks and NCCL proxies competed).  Call before allocating pinned host buffers so
    they are first-touched on the right node.  Returns the number of CPUs bound, or None if the topology is unknown."""
    mine = gpu_local_cpus(local_rank)
    if not mine:
        return None
    peers = [r for r in range(local_world) if gpu_local_cpus(r) == mine] if local_world > 1 else [local_rank]
    cpus = sorted(mine)
    if len(peers) > 1 and len(cpus) >= len(peers):
        i, n = peers.index(local_rank), len(peers)
        per = len(cpus) // n
        cpus = cpus[i * per:(i + 1) * per]
    try:
        os.sched_setaffinity(0, set(cpus))
        return len(cpus)
    except OSError:
        return None


# ----------------------------------------------------------------------------------------------
# host-copy routing: which PCIe link should carry this rank's host traffic?
# ----------------------------------------------------------------------------------------------
class HostCopyRig:
    """Bare host copies of one rank, direct (this GPU's PCIe link) or relayed through an NVLink peer (the peer's link),
    enqueued exactly like NaturalInferenceSampler.sample_host_many does.  Used by `choose_host_relay` and by bench.py for its copy ceilings."""

    def __init__(self, dev, peer, h2d_bytes: int, d2h_bytes: int):
        import torch
        self.torch, self.dev, self.peer = torch, dev, peer
        u8 = torch.uint8
        self.d_in, self.d_out = torch.empty(h2d_bytes, dtype=u8, device=dev), torch.empty(d2h_bytes, dtype=u8, device=dev)
        self.h_in, self.h_out = torch.empty(h2d_bytes, dtype=u8).pin_memory(), torch.empty(d2h_bytes, dtype=u8).pin_memory()
        self.h_in.fill_(1)
        self.s_h2d, self.s_d2h, self.push = (torch.cuda.Stream(device=dev) for _ in range(3))
        if peer is not None:
            self.r_in, self.r_out = torch.empty(h2d_bytes, dtype=u8, device=peer), torch.empty(d2h_bytes, dtype=u8, device=peer)
            self.r_h2d, self.r_d2h = torch.cuda.Stream(device=peer), torch.cuda.Stream(device=peer)

    def run(self, reps: int, with_h2d: bool) -> float:
        """ms for `reps` result copies (and as many input copies when with_h2d), both directions concurrently"""
        torch = self.torch
        main = torch.cuda.current_stream(self.dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for s in (self.s_h2d, self.s_d2h, self.push):
            s.wait_event(e0)
        last = None
        for _ in range(reps):
            if self.peer is None:
                if with_h2d:
                    with torch.cuda.stream(self.s_h2d):
                        self.d_in.copy_(self.h_in, non_blocking=True)
                with torch.cuda.stream(self.s_d2h):
                    self.h_out.copy_(self.d_out, non_blocking=True)
            else:
                if with_h2d:
                    with torch.cuda.stream(self.s_h2d), torch.cuda.stream(self.r_h2d):
                        self.r_in.copy_(self.h_in, non_blocking=True)
                        self.d_in.copy_(self.r_in, non_blocking=True)
                with torch.cuda.stream(self.push), torch.cuda.stream(self.r_d2h):
                    self.r_out.copy_(self.d_out, non_blocking=True)
                    self.h_out.copy_(self.r_out, non_blocking=True)
                    last = torch.cuda.Event()
                    last.record(self.r_d2h)
        main.wait_stream(self.s_h2d)
        main.wait_stream(self.s_d2h)
        main.wait_stream(self.push)
        if last is not None:
            main.wait_event(last)
        e1.record(main)
        torch.cuda.synchronize(self.dev)
        return e0.elapsed_time(e1)


def relay_pairs(ms_per_rank, mode: str = "auto"):
    """{slow rank: fast rank} from the per-rank times of the all-at-once direct copies, or None when the ranks do not split
    into a clearly slower and a clearly faster half (every slow rank at most 0.85 of the bandwidth of every fast one).
    mode "force": every rank relays through the next one."""
    world = len(ms_per_rank)
    if mode == "force":
        return {r: (r + 1) % world for r in range(world)} if world > 1 else None
    if world < 2 or world % 2:
        return None
    order = sorted(range(world), key=lambda r: ms_per_rank[r])  # fastest (fewest ms) first
    fast, slow = sorted(order[: world // 2]), sorted(order[world // 2:])
    if max(ms_per_rank[r] for r in fast) > 0.85 * min(ms_per_rank[r] for r in slow):
        return None
    return dict(zip(slow, fast))


def choose_host_relay(rank: int, world: int, device, h2d_bytes: int = 4096 * 3072 * 4, d2h_bytes: int = 4096 * 3072,
                      reps: int = 12, mode: Optional[str] = None):
    """One process per GPU on one box: decide, by measurement, whether this rank's host copies should be relayed through an
    NVLink peer (NaturalInferenceSampler.set_host_relay).  Needs an initialised torch.distributed group whose ranks are the
    local GPU indices.  Every rank first copies directly, all at once; if the ranks split into a clearly slower and a
    clearly faster half (ranks whose GPU reaches host memory across the socket interconnect vs ranks next to it), each slow
    rank is paired with a fast one and the copies are re-timed with the slow ranks relayed: the relay is adopted, separately
    for result-only traffic ("d2h") and for input+result traffic ("bidir"), when the slowest rank gets >= 15 % faster.
    Returns ({"d2h": peer|None, "bidir": peer|None}, info).  mode (or $NI_HOST_RELAY): "auto" | "off" | "force" (every rank
    relays through the next one -- exercises the path on any multi-GPU box)."""
    import torch
    import torch.distributed as dist
    mode = mode or os.environ.get("NI_HOST_RELAY", "auto")
    none = {"d2h": None, "bidir": None}
    info = {"mode": mode}
    if world < 2 or mode == "off" or not (dist.is_available() and dist.is_initialized()):
        return none, info
    dev = torch.device(device)
    own = dev.index if dev.index is not None else torch.cuda.current_device()

    def gather(x: float):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        lst = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(lst, t)
        return [float(v) for v in lst]

    def timed(rig, with_h2d):
        """(ms per rank, ok): every rank takes part in the collectives whatever happened locally"""
        ms, ok = 0.0, 1.0
        try:
            rig.run(2, with_h2d)
        except Exception:  # noqa: BLE001
            ok = 0.0
        torch.cuda.synchronize(dev)
        dist.barrier()
        if ok:
            try:
                ms = rig.run(reps, with_h2d)
            except Exception:  # noqa: BLE001
                ok = 0.0
        return gather(ms), min(gather(ok)) > 0

    try:
        direct = HostCopyRig(dev, None, h2d_bytes, d2h_bytes)
        built = 1.0
    except Exception:  # noqa: BLE001
        direct, built = None, 0.0
    if min(gather(built)) == 0:
        info["error"] = "could not allocate the probe buffers"
        return none, info
    a_d2h, ok1 = timed(direct, False)
    a_bi, ok2 = timed(direct, True)
    if not (ok1 and ok2):
        info["error"] = "direct probe failed"
        return none, info
    gbs = lambda nbytes, ms: [round(nbytes * reps / (m * 1e-3) / 1e9, 2) for m in ms]
    info["direct_gbs_per_rank"] = {"d2h": gbs(d2h_bytes, a_d2h), "bidir": gbs(h2d_bytes + d2h_bytes, a_bi)}
    pairs = relay_pairs(a_d2h, mode)
    if pairs is None:
        info["decision"] = "direct: no clearly slower half"
        return none, info
    info["pairs"] = {str(k): v for k, v in pairs.items()}
    peer = pairs.get(rank)
    rig, ok = direct, 1.0
    if peer is not None:
        try:
            if not torch.cuda.can_device_access_peer(own, peer):
                raise RuntimeError("no peer access")
            rig = HostCopyRig(dev, torch.device("cuda", peer), h2d_bytes, d2h_bytes)
        except Exception:  # noqa: BLE001
            ok = 0.0
    if min(gather(ok)) == 0:
        info["decision"] = "direct: a relay could not be set up"
        return none, info
    b_d2h, ok1 = timed(rig, False)
    b_bi, ok2 = timed(rig, True)
    if not (ok1 and ok2):
        info["decision"] = "direct: the relay trial failed"
        return none, info
    info["relayed_gbs_per_rank"] = {"d2h": gbs(d2h_bytes, b_d2h), "bidir": gbs(h2d_bytes + d2h_bytes, b_bi)}
    use_d2h = mode == "force" or max(b_d2h) <= 0.87 * max(a_d2h)
    use_bi = mode == "force" or max(b_bi) <= 0.87 * max(a_bi)
    info["slowest_rank_ms"] = {"d2h": [max(a_d2h), max(b_d2h)], "bidir": [max(a_bi), max(b_bi)]}
    info["decision"] = {"d2h": "relay" if use_d2h else "direct", "bidir": "relay" if use_bi else "direct"}
    del direct, rig
    torch.cuda.empty_cache()
    return {"d2h": peer if use_d2h else None, "bidir": peer if use_bi else None}, info
