"""Best-effort NUMA placement of a rank's host threads (and with them its pinned staging buffers: pinned pages are placed by
first touch) next to its GPU.  One process per GPU on an 8-GPU host otherwise allocates wherever the launcher happened to run,
and host->device copies of remote-node memory cross the inter-socket link (50.7 -> 22.4 GB/s per GPU at 8 ranks)."""
from __future__ import annotations

import os


def _parse_cpulist(s: str):
    out = set()
    for part in s.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        out.update(range(int(lo), int(hi or lo) + 1))
    return out


def gpu_pci_bus_id(ordinal: int) -> str | None:
    try:
        import torch
        p = torch.cuda.get_device_properties(ordinal)
        return f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:
        pass
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(ordinal)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        return bus.lower()[-12:]            # NVML prints an 8-digit domain
    except Exception:
        return None


def bind_to_gpu(ordinal: int, apply: bool = True) -> dict:
    """Restricts this process to the CPUs of the NUMA node the GPU hangs off (intersected with what the process may use).
    Returns what it found and did; never raises."""
    info = {"gpu": ordinal, "bus": None, "numa_node": None, "local_cpus": None, "allowed_before": None, "bound": False}
    try:
        allowed = os.sched_getaffinity(0)
        info["allowed_before"] = len(allowed)
        bus = gpu_pci_bus_id(ordinal)
        info["bus"] = bus
        if bus is None:
            return info
        base = f"/sys/bus/pci/devices/{bus}"
        try:
            info["numa_node"] = int(open(base + "/numa_node").read())
        except Exception:
            pass
        local = _parse_cpulist(open(base + "/local_cpulist").read())
        info["local_cpus"] = len(local)
        want = local & allowed
        info["local_allowed"] = len(want)
        if apply and want and want != allowed:
            os.sched_setaffinity(0, want)
            info["bound"] = True
    except Exception as e:      # noqa
        info["error"] = repr(e)
    return info
