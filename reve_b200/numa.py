"""Host placement for the one-thread-per-GPU shape (the shared resources of the multi-GPU path are host
cores, host DRAM and the PCIe root complexes, not the GPUs).  A worker thread that allocates its pinned ring and drives
its context from the CPUs of the NUMA node its GPU hangs off keeps the frame traffic off the inter-socket link.
Linux only; every function degrades to a no-op when sysfs or the PCI address is not available."""
from __future__ import annotations

import os
from typing import Optional, Set


def _parse_cpulist(text: str) -> Set[int]:
    cpus: Set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_pci_bus_id(device: int) -> Optional[str]:
    """'0000:1b:00.0' of CUDA device `device` (through torch, which is already loaded wherever this is used)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        return f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:
        return None


def cpus_local_to_gpu(device: int, sysfs: str = "/sys") -> Optional[Set[int]]:
    """CPUs of the NUMA node the GPU is attached to, or None if unknown (single-node machines report -1)."""
    bdf = gpu_pci_bus_id(device)
    if not bdf:
        return None
    try:
        node = int(open(os.path.join(sysfs, "bus/pci/devices", bdf, "numa_node")).read())
        if node < 0:
            return None
        cpus = _parse_cpulist(open(os.path.join(sysfs, "devices/system/node", f"node{node}", "cpulist")).read())
        return cpus or None
    except (OSError, ValueError):
        return None


def bind_thread_to_gpu_node(device: int) -> Optional[int]:
    """Restricts the CALLING thread to the CPUs local to `device`; returns how many CPUs that is (None = left alone)."""
    cpus = cpus_local_to_gpu(device)
    if not cpus:
        return None
    try:
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)      # pid 0 = the calling thread
        return len(allowed)
    except (AttributeError, OSError):
        return None
