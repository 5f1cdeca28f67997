"""Host-side (Python + launch) time of the label-generation calls, measured without synchronising inside the loop: how long the
CPU needs to ENQUEUE one fuse_sources / select_and_apply / LabelGenerator.run, against the GPU time of the same work."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mspl_b200 import ops  # noqa: E402
from mspl_b200.data_loader.segmentation.greenhouse import SOURCE_TABLES  # noqa: E402
from mspl_b200.pipeline import LabelGenerator  # noqa: E402

dev = torch.device("cuda:0")
srcs = (("camvid", 13), ("cityscapes", 20), ("forest", 5))
luts = [SOURCE_TABLES[s] for s, _ in srcs]
for n in (8, 64, 400):
    g = torch.Generator(device=dev).manual_seed(1)
    mains = [3 * torch.randn((n, c, 256, 480), device=dev, generator=g) for _, c in srcs]
    auxs = [m + 1.5 * torch.randn(m.shape, device=dev, generator=g) for m in mains]
    gen = LabelGenerator(luts, policy="all")
    r = ops.fuse_sources(mains, auxs, luts, policy="all")
    for name, fn in (("fuse_sources", lambda: ops.fuse_sources(mains, auxs, luts, policy="all")),
                     ("select_and_apply", lambda: ops.select_and_apply(r.label, r.conf, 0.2)),
                     ("LabelGenerator.run", lambda: gen.run(mains, auxs))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(reps):
            fn()
        t_issue = time.perf_counter() - t0
        e1.record()
        torch.cuda.synchronize()
        print("n=%4d %-20s host issue %.3f ms/call   gpu %.3f ms/call" % (n, name, 1e3 * t_issue / reps, e0.elapsed_time(e1) / reps), flush=True)
