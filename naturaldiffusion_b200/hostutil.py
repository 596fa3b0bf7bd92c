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
