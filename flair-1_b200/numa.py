"""Best-effort NUMA placement of a rank's host buffers next to its GPU.

Every rank of a sharded zone stages its raster rows, truth rows and its part of the output map in page-locked host
memory and moves them over its own PCIe link. On a two-socket host the pages should live on the socket that GPU
hangs off; otherwise eight ranks funnel their copies through one socket's memory controllers and the inter-socket
link. The kernel places pages on the node of the CPU that first touches them, so pinning the process to the CPUs
of the GPU's node before the buffers are allocated is enough -- no libnuma needed. Everything here is advisory:
any failure (no sysfs, a cpuset that excludes that node, a single-node host) leaves the process as it was.
"""
from __future__ import annotations

import os
from pathlib import Path


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device_index: int) -> int | None:
    """NUMA node of CUDA device `device_index` from sysfs, or None."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(Path(f"/sys/bus/pci/devices/{bdf}/numa_node").read_text().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_node(device_index: int) -> dict:
    """Restrict this process to the CPUs of the GPU's NUMA node (intersected with its current affinity).
    Returns a small report for logs: {"node": n or None, "cpus": count or None, "bound": bool}."""
    report = {"node": None, "cpus": None, "bound": False}
    node = gpu_numa_node(device_index)
    report["node"] = node
    if node is None or not hasattr(os, "sched_setaffinity"):
        return report
    try:
        cpus = _parse_cpulist(Path(f"/sys/devices/system/node/node{node}/cpulist").read_text())
        allowed = cpus & os.sched_getaffinity(0)
        if allowed and allowed != os.sched_getaffinity(0):
            os.sched_setaffinity(0, allowed)
            report["bound"] = True
        report["cpus"] = len(allowed)
    except Exception:
        pass
    return report
