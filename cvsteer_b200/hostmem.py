"""Host-side placement helpers for the host-buffer entry points (cvs_*_run_batch_host, cvs_g2_lines_u8_host).

The library never allocates caller memory, so WHERE the caller's pinned frames live is the caller's choice -- and on a
two-socket box it decides whether eight GPUs can stream concurrently: a pinned buffer that sits on the other socket's
memory crosses the inter-socket link on every DMA.  One process per GPU (the deployment bench.py measures) should run on,
and allocate from, the NUMA node its GPU hangs off.  `bind_to_gpu_numa` does that with the affinity mask: Linux's default
memory policy is local allocation, so buffers pinned AFTER the call land on that node."""
from __future__ import annotations

import os
from typing import List, Optional


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_pci_bus_id(device: int) -> Optional[str]:
    """'0000:1b:00.0'-style id of CUDA device `device` (NVML order follows CUDA_VISIBLE_DEVICES only through torch)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        if hasattr(p, "pci_bus_id"):
            return "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, getattr(p, "pci_device_id", 0))
    except Exception:
        pass
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        return bus[-12:].lower()          # NVML pads the domain to 8 hex digits
    except Exception:
        return None


def gpu_numa_node(device: int) -> Optional[int]:
    bus = gpu_pci_bus_id(device)
    if not bus:
        return None
    try:
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        return node if node >= 0 else None
    except Exception:
        return None


def numa_cpus(node: int) -> List[int]:
    try:
        return _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
    except Exception:
        return []


def bind_to_gpu_numa(device: int) -> dict:
    """Restrict the calling process to the CPUs of the GPU's NUMA node (intersected with what it may already use).
    Returns what was done: {"node": n | None, "cpus": k, "bound": bool}."""
    node = gpu_numa_node(device)
    info = {"node": node, "cpus": 0, "bound": False}
    if node is None:
        return info
    try:
        allowed = set(os.sched_getaffinity(0))
        want = sorted(allowed.intersection(numa_cpus(node)))
        info["cpus"] = len(want)
        if want and len(want) < len(allowed):
            os.sched_setaffinity(0, want)
            info["bound"] = True
    except Exception:
        pass
    return info
