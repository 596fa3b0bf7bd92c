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
