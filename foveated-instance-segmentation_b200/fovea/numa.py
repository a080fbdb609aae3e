"""Bind a rank's host threads -- and with them, by first touch, its pinned staging buffers -- to the NUMA node its GPU
hangs off.

The host-buffer pipeline (fovea.pipeline.ResamplePipeline) moves ~0.75 GB per 64-frame batch over PCIe in both
directions.  Under `torchrun` every rank starts on whatever cores the scheduler picks; on a two-socket box half of the
ranks then stage their frames in the other socket's memory and every byte crosses the inter-socket link as well as
PCIe.  `bind_to_gpu_node(i)` reads the GPU's PCI address from the CUDA runtime, its NUMA node from sysfs, and restricts
the calling process to that node's cores before any pinned allocation is made.  Nothing is changed when the platform
exposes no NUMA information (single-node VMs report numa_node = -1): the returned record says so.
"""
from __future__ import annotations

import os


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_cpulist(s):
    cpus = set()
    for part in (s or "").split(","):
        part = part.strip()
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(index):
    """(pci address, numa node or None) of CUDA device `index`."""
    import torch
    p = torch.cuda.get_device_properties(index)
    dom = getattr(p, "pci_domain_id", None)
    bus = getattr(p, "pci_bus_id", None)
    devid = getattr(p, "pci_device_id", None)
    if bus is None or devid is None:
        return None, None
    addr = f"{(dom or 0):04x}:{bus:02x}:{devid:02x}.0"
    node = _read(f"/sys/bus/pci/devices/{addr}/numa_node")
    try:
        node = int(node)
    except (TypeError, ValueError):
        node = None
    return addr, (node if node is not None and node >= 0 else None)


def bind_to_gpu_node(index):
    """Restrict this process to the cores of the NUMA node of CUDA device `index`.  Returns a record for the logs:
    {"pci", "numa_node", "cpus_before", "cpus_after", "bound"}."""
    addr, node = gpu_numa_node(index)
    before = len(os.sched_getaffinity(0))
    rec = {"pci": addr, "numa_node": node, "cpus_before": before, "cpus_after": before, "bound": False,
           "nodes_online": _read("/sys/devices/system/node/online")}
    if node is None:
        return rec
    cpus = _parse_cpulist(_read(f"/sys/devices/system/node/node{node}/cpulist")) & os.sched_getaffinity(0)
    if not cpus:
        return rec
    os.sched_setaffinity(0, cpus)
    rec.update(cpus_after=len(cpus), bound=True)
    return rec
