"""CPU / memory affinity for one-process-per-GPU runs: pin the process to the CPUs NVML reports as local to its GPU, so
that the pinned host buffers it allocates afterwards (first touch) and its copy threads sit on the GPU's NUMA node.
Eight ranks streaming 3.8 GB results per timepoint over PCIe otherwise share one socket's memory controllers."""
from __future__ import annotations

import os


def device_cpus(device):
    """CPUs local to CUDA device `device` (ordinal in this process), or None if NVML / the mapping is unavailable."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(int(device)).uuid)
        if not uuid.startswith('GPU-'):
            uuid = 'GPU-' + uuid
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, 'encode') else uuid)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [w * 64 + b for w in range(words) for b in range(64) if (int(mask[w]) >> b) & 1]
        return cpus or None
    except Exception:
        return None


def bind_to_device(device):
    """Restrict this process to the CPUs local to `device`; returns the CPU list or None (nothing changed)."""
    if os.environ.get('OF3D_NUMA_BIND', '1') == '0' or not hasattr(os, 'sched_setaffinity'):
        return None
    cpus = device_cpus(device)
    if not cpus:
        return None
    try:
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except OSError:
        return None
