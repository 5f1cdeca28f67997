"""Probe (run under gpurun): which host CPUs / NUMA node NVML reports as local to each visible GPU, what this process may run
on, and pinned host->device copy bandwidth before and after binding the process to the GPU's CPUs (first-touch then places
the pinned pages on that node)."""
import os
import time

import pynvml
import torch

pynvml.nvmlInit()
n = pynvml.nvmlDeviceGetCount()
print("cpus allowed:", sorted(os.sched_getaffinity(0)), "cpu_count", os.cpu_count())
try:
    print("numa nodes:", sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")))
except OSError as e:
    print("no /sys numa info:", e)
for i in range(n):
    h = pynvml.nvmlDeviceGetHandleByIndex(i)
    words = (os.cpu_count() + 63) // 64
    try:
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
    except pynvml.NVMLError as e:
        cpus = repr(e)
    try:
        node = pynvml.nvmlDeviceGetNumaNodeId(h)
    except Exception as e:
        node = repr(e)
    print("gpu", i, pynvml.nvmlDeviceGetPciInfo(h).busId, "numa", node, "local cpus", cpus if isinstance(cpus, str) else (cpus[:4], "...", cpus[-4:], len(cpus)))


def h2d_gbs(dev, nbytes=2 << 30, reps=5):
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        d.copy_(host, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, nbytes / 1e9 / (e0.elapsed_time(e1) / 1e3))
    return best


dev = torch.device("cuda:0")
print("H2D pinned, as launched: %.1f GB/s" % h2d_gbs(dev))
h = pynvml.nvmlDeviceGetHandleByIndex(0)
try:
    pynvml.nvmlDeviceSetCpuAffinity(h)
    print("bound to:", sorted(os.sched_getaffinity(0))[:8], "...")
    print("H2D pinned, after nvmlDeviceSetCpuAffinity: %.1f GB/s" % h2d_gbs(dev))
except pynvml.NVMLError as e:
    print("nvmlDeviceSetCpuAffinity failed:", e)
