#!/usr/bin/env python
"""Concurrent host->device copy ceiling: every rank copies a pinned buffer to its own GPU at the same time (barrier-fenced CUDA
events), for 1..N ranks, next to NVML's view of each GPU's local CPUs / NUMA node.  Answers what bench.py's e2e line can reach at
every N (its `e2e.h2d_ceiling_gbs` is the same measurement on the bench's own buffers).

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py
"""
import json
import os

import torch
import torch.distributed as dist


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = 2 << 30
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    info = {}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        try:
            info["numa_node"] = pynvml.nvmlDeviceGetNumaNodeId(h)
        except Exception as e:
            info["numa_node"] = repr(e)
        info["cpus_allowed"] = len(os.sched_getaffinity(0))
    except Exception as e:
        info["nvml"] = repr(e)

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    rows = []
    for active in sorted({1, 2, 4, world} & set(range(1, world + 1))):
        d.copy_(host, non_blocking=True)
        fence()
        best = 0.0
        for _ in range(4):
            fence()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if rank < active:
                d.copy_(host, non_blocking=True)
            e1.record()
            fence()
            if rank < active:
                best = max(best, nbytes / 1e9 / (e0.elapsed_time(e1) / 1e3))
        t = torch.tensor([best if rank < active else 1e9, best if rank < active else 0.0], device=dev)
        lo, hi = t[:1].clone(), t[1:].clone()
        if world > 1:
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        rows.append({"ranks_copying": active, "slowest_gbs": round(lo.item(), 2), "fastest_gbs": round(hi.item(), 2),
                     "aggregate_gbs_at_slowest": round(lo.item() * active, 1)})
    if rank == 0:
        print(json.dumps({"world": world, "bytes_per_copy": nbytes, "rank0": info, "rows": rows}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
