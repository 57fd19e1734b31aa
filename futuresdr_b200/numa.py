"""Pin a rank to the CPU cores (and thereby the NUMA node) next to its GPU.

One process per GPU moves 8 B/sample each way between pinned host memory and the device; with eight ranks the
traffic is bounded by host DRAM and the inter-socket link unless every rank's pinned buffers live on the node its
GPU hangs off.  Linux places pages on the node of the thread that first touches them, so it is enough to restrict
the rank's threads to that node's cores BEFORE allocating pinned memory.  (No reference counterpart: FutureSDR is
single-device.)"""
from __future__ import annotations

import os


def _cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def gpu_numa_info(device_index: int) -> dict:
    """{'node': int | None, 'cpus': sorted list} for the GPU, from NVML (CPU affinity mask) or sysfs."""
    info = {"node": None, "cpus": [], "source": None}
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(device_index)
        try:
            bus = nv.nvmlDeviceGetPciInfo(h).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            path = f"/sys/bus/pci/devices/{bus.lower()[-12:]}/numa_node"
            if os.path.exists(path):
                node = int(open(path).read().strip())
                if node >= 0:
                    info["node"] = node
                    cl = f"/sys/devices/system/node/node{node}/cpulist"
                    if os.path.exists(cl):
                        info["cpus"] = sorted(_cpulist(open(cl).read()))
                        info["source"] = "sysfs numa_node"
        except Exception:
            pass
        if not info["cpus"]:
            ncpu = os.cpu_count() or 64
            words = (ncpu + 63) // 64
            mask = nv.nvmlDeviceGetCpuAffinity(h, words)
            cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
            info["cpus"] = cpus
            info["source"] = "nvml cpu affinity"
    except Exception as e:  # noqa: BLE001
        info["error"] = repr(e)
    return info


def bind_to_gpu(device_index: int) -> dict:
    """Restrict this process to the cores next to the GPU (intersected with what it may already use).
    Returns what was done, for the bench line."""
    info = gpu_numa_info(device_index)
    try:
        allowed = os.sched_getaffinity(0)
        want = set(info["cpus"]) & allowed
        if want and want != allowed:
            os.sched_setaffinity(0, want)
            info["bound"] = True
        else:
            info["bound"] = False
        info["ncpus"] = len(os.sched_getaffinity(0))
    except Exception as e:  # noqa: BLE001
        info["bound"] = False
        info["error"] = repr(e)
    info.pop("cpus", None)
    return info


class local_to_gpu:
    """Context manager: allocate (and first-touch) pinned host buffers inside it so their pages land on the GPU's
    NUMA node, then give the process its original CPU set back (the DMA engines do not care where the thread runs;
    a CPU baseline timed afterwards must still see every core)."""

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.prev = None
        self.info = {}

    def __enter__(self):
        try:
            self.prev = os.sched_getaffinity(0)
        except Exception:  # noqa: BLE001
            self.prev = None
        self.info = bind_to_gpu(self.device_index)
        return self.info

    def __exit__(self, *exc):
        if self.prev is not None:
            try:
                os.sched_setaffinity(0, self.prev)
            except Exception:  # noqa: BLE001
                pass
        return False
