"""Bind a rank's host side to the NUMA node of its GPU.

The host-buffer entry (`maximum_path_from_prior_host`) is PCIe-bound: 250 MB of pinned memory per
step.  On an 8-GPU box the GPUs hang off two sockets; a rank whose pinned pages (first touch) or
whose thread live on the other socket pays the inter-socket link on every copy, and with 4-8 ranks
started by torchrun that placement is luck (round-1 SCALE: e2e 5.0 ms/step at N=1-2, 9.1 / 11.3 ms at
N=4 / 8).  Call `bind_to_device_node(local_rank)` BEFORE allocating or pinning host buffers.
Plain sysfs + sched_setaffinity: no libnuma needed; a no-op (returns None) wherever the topology
cannot be read.
"""
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


def device_numa_node(index: int) -> Optional[int]:
    """NUMA node of CUDA device `index` from /sys/bus/pci/devices/<bus id>/numa_node, or None."""
    try:
        import torch
        props = torch.cuda.get_device_properties(index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def node_cpus(node: int) -> List[int]:
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            return _parse_cpulist(f.read())
    except Exception:
        return []


def bind_to_device_node(index: int) -> Optional[dict]:
    """Restrict this process to the CPUs of the GPU's NUMA node (threads created afterwards
    inherit it; pages pinned afterwards are first-touched there).  Returns what was done."""
    node = device_numa_node(index)
    if node is None:
        return None
    cpus = node_cpus(node)
    try:
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
    except Exception:
        return None
    return {"numa_node": node, "cpus": len(allowed)}
