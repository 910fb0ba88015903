"""Host-side placement for the host-buffer entry points (``sagnn_propagate_host`` / ``sagnn_host_forward`` /
``sagnn_host_backward``): their time is the PCIe link's, and the link only runs at speed when the pinned pages
live on the NUMA node the GPU hangs off.  ``near_gpu(device)`` pins the calling thread to that node's cores for the
duration of a ``with`` block -- allocate (and touch) the pinned buffers inside it and the kernel's first-touch
policy places them there; the previous affinity is restored on exit, so CPU-side work keeps every core.

Uses NVML (``nvmlDeviceGetCpuAffinity``, looked up by the device's PCI bus id so CUDA_VISIBLE_DEVICES does not
matter).  Where NVML or the affinity call is unavailable the block runs unchanged and ``info["bound"]`` says so.
Framework plumbing: nothing on the reference's side corresponds to it (its tensors never leave the host).
"""
from __future__ import annotations

import contextlib
import os


def gpu_cpu_set(device_index):
    """Logical CPUs NVML reports as local to the CUDA device, or None."""
    try:
        import pynvml
        import torch
        props = torch.cuda.get_device_properties(device_index)
        pynvml.nvmlInit()
        try:
            bus = "%08x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            n_cpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
            cpus = {w * 64 + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        finally:
            pynvml.nvmlShutdown()
        return cpus or None
    except Exception:
        return None


@contextlib.contextmanager
def near_gpu(device_index, info=None):
    """``with near_gpu(i): buf = torch.empty(..., pin_memory=True)`` -> pages on the GPU's NUMA node."""
    info = {} if info is None else info
    info["bound"] = False
    old = None
    try:
        cpus = gpu_cpu_set(device_index)
        if cpus and hasattr(os, "sched_setaffinity"):
            old = os.sched_getaffinity(0)
            want = (cpus & old) or cpus
            os.sched_setaffinity(0, want)
            info["bound"] = True
            info["cpus"] = len(want)
    except Exception:
        old = None
    try:
        yield info
    finally:
        if old is not None:
            try:
                os.sched_setaffinity(0, old)
            except Exception:
                pass
