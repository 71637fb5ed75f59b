"""CPU / NUMA placement of a one-GPU-per-process worker.

Pinned result segments are allocated (and first touched) by the calling thread, so they land on the
NUMA node that thread runs on.  On an 8-GPU HGX board half of the GPUs hang off each socket; a rank
whose pinned memory sits on the far socket pushes its D2H stream across the inter-socket link and
eight ranks doing so at once collapse the aggregate (measured: 91 GB/s vs 222 GB/s for 8 x 100 MB).
``bind_to_device`` pins the calling process to the CPUs NVML reports as local to the GPU, the same
thing ``numactl --cpunodebind`` would do for a job-backed worker.
"""
import os


def device_cpus(device_index):
    """CPU ids local to CUDA device ``device_index`` (empty list if NVML cannot tell)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            ncpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        finally:
            pynvml.nvmlShutdown()
    except Exception:
        return []
    cpus = []
    for w, word in enumerate(words):
        for b in range(64):
            if (int(word) >> b) & 1:
                cpus.append(w * 64 + b)
    return cpus


def bind_to_device(device_index):
    """Restrict the calling process to the GPU-local CPUs.  Returns the CPU list (``[]`` = unchanged)."""
    cpus = device_cpus(device_index)
    if not cpus or not hasattr(os, "sched_setaffinity"):
        return []
    allowed = set(os.sched_getaffinity(0))
    want = sorted(allowed.intersection(cpus))
    if not want:
        return []
    os.sched_setaffinity(0, want)
    return want
