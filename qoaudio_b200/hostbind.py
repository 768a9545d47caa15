"""Keep a rank's host side next to its GPU.

One process drives one GPU.  Its pinned staging buffers are the other end of every host<->device copy of
the end-to-end path, and Linux places their pages on the NUMA node of the thread that first touches them.  On a
two-socket box a rank that runs on the far socket sends all of its PCIe traffic across the socket interconnect, which
eight ranks saturate long before they saturate their PCIe links.  `bind_to_gpu` restricts the calling process to the
CPUs NVML reports as local to the GPU, before any pinned allocation is made.
"""
import os


def gpu_local_cpus(index):
    """CPUs NVML lists as closest to GPU `index` (NVML order), intersected with what this process may use."""
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(index)
    words = (max(os.cpu_count() or 1, 1) + 63) // 64
    mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
    cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
    return cpus & os.sched_getaffinity(0)


def bind_to_gpu(index):
    """Restrict this process to the GPU-local CPUs.  Returns (previous affinity, new affinity); the new one is None when
    nothing was changed (NVML missing, no locality information, or the process is already inside the local set)."""
    before = os.sched_getaffinity(0)
    if os.environ.get("QOA_B200_NO_BIND"):
        return before, None
    try:
        target = gpu_local_cpus(index)
        if target and target != before:
            os.sched_setaffinity(0, target)
            return before, target
    except Exception:
        pass
    return before, None


def restore(affinity):
    try:
        os.sched_setaffinity(0, affinity)
    except Exception:
        pass
