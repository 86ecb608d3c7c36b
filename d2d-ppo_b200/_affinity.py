"""Bind a rank's host threads (and therefore its first-touch pinned buffers) to the NUMA node of its GPU.

One process per GPU: the host side of ``step_host`` is a stream of pinned-memory copies (50 MB of actions per step at
1M envs).  With 8 ranks on a two-socket box, buffers that live on the other socket cross the inter-socket link on
their way to the GPU and the per-rank copy rate drops (round 1: 55 GB/s at 1 rank, 21 GB/s at 8).  Plumbing only;
best effort: silently does nothing where sysfs does not say.
"""
from __future__ import annotations

import os


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device_index):
    """NUMA node of CUDA device ``device_index`` from sysfs, or None."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa(device_index):
    """Restrict this process to the CPUs of the GPU's NUMA node (intersected with the current affinity mask).
    Returns (node, n_cpus) or None.  Call before allocating pinned host buffers."""
    node = gpu_numa_node(device_index)
    if node is None or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node, len(cpus)
    except Exception:
        return None
