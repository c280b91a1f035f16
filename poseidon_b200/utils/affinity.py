"""CPU / NUMA affinity for the per-GPU process (the role of Bösen's optional NumaMgr: pin comm/worker threads
near their device).  reference: ps/src/petuum_ps/thread/numa_mgr.hpp:28-203 (policies Even / Center)."""
from __future__ import annotations

import os
from typing import List, Optional


def _read(path: str) -> Optional[str]:
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_cpulist(s: str) -> List[int]:
    out = []
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-")
            out.extend(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out


def gpu_numa_node(pci_bus_id: str) -> Optional[int]:
    v = _read(f"/sys/bus/pci/devices/{pci_bus_id.lower()}/numa_node")
    return int(v) if v not in (None, "", "-1") else None


def pin_to_gpu_numa(device_index: int, policy: str = "even", local_rank: int = 0, local_world: int = 1) -> List[int]:
    """Restrict this process to CPUs of the GPU's NUMA node ("center") or an even slice of all CPUs ("even")."""
    cpus = sorted(os.sched_getaffinity(0))
    chosen = cpus
    try:
        import torch
        if policy == "center" and torch.cuda.is_available():
            bus = torch.cuda.get_device_properties(device_index).pci_bus_id
            dom = f"{getattr(torch.cuda.get_device_properties(device_index), 'pci_domain_id', 0):04x}:{bus:02x}:00.0"
            node = gpu_numa_node(dom)
            if node is not None:
                lst = _read(f"/sys/devices/system/node/node{node}/cpulist")
                if lst:
                    chosen = [c for c in _parse_cpulist(lst) if c in cpus] or cpus
    except Exception:
        chosen = cpus
    if policy == "even" and local_world > 1:
        per = max(1, len(cpus) // local_world)
        chosen = cpus[local_rank * per:(local_rank + 1) * per] or cpus
    try:
        os.sched_setaffinity(0, chosen)
    except OSError:
        pass
    return chosen
