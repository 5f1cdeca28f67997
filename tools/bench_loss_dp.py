#!/usr/bin/env python
"""BASELINE.json config 4: uncertainty-weighted (exp(-KLD)) rectified CE forward+backward at batch 64, 480x256, 5 greenhouse
classes, data parallel.  One process per GPU (torchrun); every rank runs the fused kernel K4 on its share of the batch with
the GLOBAL pixel count as the divisor of the means, then the scalar losses are summed over ranks (the logit gradients stay
local: they feed the model's own backward / DDP).  Prints one JSON line per configuration on rank 0.

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_loss_dp.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mspl_b200 import ops  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    h, w, k, steps, warm = 256, 480, 5, 50, 5
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    cw = torch.tensor([1.0, 1.0, 1.0, 1.0, 0.0], device=dev)
    for global_batch in (64, 64 * world):
        b = global_batch // world
        g = torch.Generator(device=dev).manual_seed(11 + rank)
        main_l = 3 * torch.randn((b, k, h, w), device=dev, generator=g)
        aux_l = main_l + 1.5 * torch.randn((b, k, h, w), device=dev, generator=g)
        target = torch.randint(1, 5, (b, h, w), device=dev, generator=g)
        norm = float(global_batch * h * w)

        def step():
            out3, dm, da = ops.uw_ce_fwd_bwd(main_l, aux_l, target, cw, norm_pixels=norm)
            if world > 1:
                dist.all_reduce(out3)          # global loss = sum of the per-rank partial means
            return out3
        for _ in range(warm):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = []
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out3 = step()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        t = torch.tensor([ms[len(ms) // 2]], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            pix = global_batch * h * w
            print(json.dumps({"name": "uw_ce_loss_fwd_bwd_dp", "n_gpus": world, "global_batch": global_batch, "batch_per_gpu": b,
                              "ms_median_max_over_ranks": round(t.item(), 4), "Mpix/s": round(pix / 1e6 / (t.item() / 1e3), 1),
                              "GB/s_per_gpu": round(b * h * w * 88 / 1e9 / (t.item() / 1e3), 1), "loss": round(out3[0].item(), 6),
                              "note": "K4 + all-reduce of the 3-float loss vector; L2 flushed between iterations"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
