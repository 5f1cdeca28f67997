#!/usr/bin/env python
"""Small driver for ncu captures of the kernels bench.py's headline launch list does not reach: K1-lowres (200 images), the
per-class K1 policy, and the threshold stage (bracket select, classify pass, cluster tail) on a 400-image job.

  ncu --set full --clock-control none --kernel-name regex:"lowres|cand_resolve|bracket|fuse_sources_tma" ... python tools/profile_targets.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mspl_b200 import ops  # noqa: E402
from mspl_b200.data_loader.segmentation.greenhouse import SOURCE_TABLES  # noqa: E402
from mspl_b200.pipeline import LabelGenerator  # noqa: E402

dev = torch.device("cuda:0")
luts = [SOURCE_TABLES[s] for s, _ in bench.SOURCES]
h, w = 256, 480
mains, auxs = bench.make_logits_device(torch, 200, h, w, dev, seed=9, lowres=True)
for _ in range(3):
    ops.fuse_sources_lowres(mains, auxs, luts, (h, w), policy="all")
del mains, auxs
mains, auxs = bench.make_logits_device(torch, 400, h, w, dev, seed=3)
for policy in ("all", "half"):
    gen = LabelGenerator(luts, policy=policy, portion=0.2)
    for _ in range(3):
        gen.run(mains, auxs)
torch.cuda.synchronize()
print("done")
