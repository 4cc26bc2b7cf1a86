"""NUMA placement for the host-buffer pipeline: page-locked buffers should live on the memory node the GPU's PCIe root
hangs off, or every D2H byte crosses the socket interconnect (8 ranks pinning into one node was the round-1 e2e limit).

`bind_to_device_node(i)` restricts the calling process to the CPUs of GPU i's NUMA node, so that the pinned allocations
that follow (first touch by this process) land there.  Best effort: returns the node, or None when the platform does
not say (no sysfs entry, a single node, containers without affinity rights)."""
from __future__ import annotations

import os
from pathlib import Path
from typing import Optional


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def device_numa_node(device_index: int) -> Optional[int]:
    try:
        import torch

        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = Path(f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node")
        node = int(path.read_text().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_device_node(device_index: int) -> Optional[int]:
    node = device_numa_node(device_index)
    if node is None:
        return None
    try:
        cpus = _parse_cpulist(Path(f"/sys/devices/system/node/node{node}/cpulist").read_text())
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if not target:
            return None
        os.sched_setaffinity(0, target)
        return node
    except Exception:
        return None
