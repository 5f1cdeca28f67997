#!/usr/bin/env python
"""Secondary measurements (BASELINE.json configs 4 and 5 and the non-headline policies), one JSON line each.
Not the driver's bench (that is /bench.py); run on a B200:  python tools/bench_extra.py [--quick]

  loss_b64 / loss_b8 : K4 fused uncertainty-weighted CE forward+backward, K=5, 480x256 (config 4: batch 64, and the
                       8-per-GPU share of it under 8-way data parallelism), 88 algorithmic B/pixel
  loss_modules_b64   : the same loss through the reference-named modules (PixelwiseKLD + UncertaintyWeighted... + autograd)
  stress_1024x512    : config 5, 3 sources x 20 classes at 1024x512 (495 B/pixel incl. thresholds), one GPU's share
  policy_half / prob : K1 with per-class probabilities (vote threshold below S / probability fusion)
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mspl_b200 import ops  # noqa: E402
from mspl_b200.data_loader.segmentation.greenhouse import SOURCE_TABLES  # noqa: E402
from mspl_b200.loss_fns.segmentation_loss import PixelwiseKLD, UncertaintyWeightedSegmentationLoss  # noqa: E402
from mspl_b200.pipeline import LabelGenerator  # noqa: E402

PEAK = 6455.6
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, iters, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()                      # > L2 bytes written between iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2], ms[0]


def report(name, pix, bytes_per_pix, med, best, **extra):
    line = {"name": name, "Mpix/s": round(pix / 1e6 / (med / 1e3), 1), "ms_median": round(med, 4), "ms_min": round(best, 4),
            "GB/s": round(pix * bytes_per_pix / 1e9 / (med / 1e3), 1), "frac_of_measured_hbm_peak": round(pix * bytes_per_pix / 1e9 / (med / 1e3) / PEAK, 4),
            "algorithmic_bytes_per_pixel": bytes_per_pix}
    line.update(extra)
    print(json.dumps(line), flush=True)


def logits(n, c, h, w, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    m = 3 * torch.randn((n, c, h, w), device=dev, generator=g) + 3 * torch.randn((n, c, 1, 1), device=dev, generator=g)
    a = m + 1.5 * torch.randn((n, c, h, w), device=dev, generator=g)
    return m, a


def label_io_section(dev):
    """uint8 label maps on the GPU -> PNG files, asynchronous writer vs the reference's synchronous PIL save."""
    import tempfile
    import time
    from PIL import Image
    from mspl_b200.label_io import LabelWriter
    n, h, w = 256, 256, 480
    labels = torch.randint(0, 5, (n, h, w), device=dev, dtype=torch.uint8)
    with tempfile.TemporaryDirectory() as tmp:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with LabelWriter(dev, workers=8) as wr:
            for lo in range(0, n, 32):
                wr.submit(labels[lo:lo + 32], ["%s/a_%04d.png" % (tmp, i) for i in range(lo, lo + 32)])
        t_async = time.perf_counter() - t0
        t0 = time.perf_counter()
        host = labels[:32].cpu().numpy()
        for i in range(32):
            Image.fromarray(host[i]).save("%s/b_%04d.png" % (tmp, i))
        t_pil = (time.perf_counter() - t0) * n / 32
    print(json.dumps({"name": "label_io_256img_480x256", "async_writer_ms_per_image": round(1e3 * t_async / n, 3),
                      "pil_sync_ms_per_image": round(1e3 * t_pil / n, 3), "speedup": round(t_pil / t_async, 1),
                      "note": "D2H + PNG encode + file write; PIL figure extrapolated from 32 images"}), flush=True)


def lowres_section(dev, iters):
    """K1 with ESPDNetUE's final bilinear upsample (x2 main, x4 aux) fused in, against what the reference
    pipeline does on the GPU today: F.interpolate both heads to full size (PyTorch kernels), then fuse."""
    import torch.nn.functional as F
    n, h, w = 200, 256, 480
    srcs = (("camvid", 13), ("cityscapes", 20), ("forest", 5))
    g = torch.Generator(device=dev).manual_seed(5)
    mains = [3 * torch.randn((n, c, h // 2, w // 2), device=dev, generator=g) for _, c in srcs]
    auxs = [3 * torch.randn((n, c, h // 4, w // 4), device=dev, generator=g) for _, c in srcs]
    luts = [SOURCE_TABLES[s] for s, _ in srcs]
    lr_bytes = 4 * sum(c for _, c in srcs) * (1 / 4 + 1 / 16) + 9

    med, best = timed(lambda: ops.fuse_sources_lowres(mains, auxs, luts, (h, w), policy="all"), iters)
    report("k1_lowres_fused_upsample_all_%dimg" % n, n * h * w, lr_bytes, med, best,
           note="reads pre-upsample logits (%.1f B/pixel); instruction-bound" % lr_bytes)

    def unfused():
        um = [F.interpolate(m, size=(h, w), mode="bilinear", align_corners=True) for m in mains]
        ua = [F.interpolate(a, size=(h, w), mode="bilinear", align_corners=True) for a in auxs]
        return ops.fuse_sources(um, ua, luts, policy="all")
    med2, best2 = timed(unfused, iters)
    report("k1_after_torch_upsample_all_%dimg" % n, n * h * w, lr_bytes, med2, best2,
           note="F.interpolate x6 (PyTorch) + K1; speed-up of the fused kernel: %.2fx" % (med2 / med))
    um = [F.interpolate(m, size=(h, w), mode="bilinear", align_corners=True) for m in mains]
    ua = [F.interpolate(a, size=(h, w), mode="bilinear", align_corners=True) for a in auxs]
    med3, best3 = timed(lambda: ops.fuse_sources(um, ua, luts, policy="all"), iters)
    report("k1_fullres_only_all_%dimg" % n, n * h * w, 313, med3, best3)


def loss_lowres_section(dev, iters, flush):
    """SURVEY.md 8f-1, loss side: K4 taking ESPDNetUE's PRE-upsample heads (x2 main, x4 aux), against upsampling with
    F.interpolate (autograd) and running the full-resolution fused K4."""
    import torch.nn.functional as F
    b, k, h, w = 64, 5, 256, 480
    g = torch.Generator(device=dev).manual_seed(13)
    main_lr = 3 * torch.randn((b, k, h // 2, w // 2), device=dev, generator=g)
    aux_lr = 3 * torch.randn((b, k, h // 4, w // 4), device=dev, generator=g)
    target = torch.randint(1, 5, (b, h, w), device=dev)
    cw = torch.tensor([1.0, 1.0, 1.0, 1.0, 0.0], device=dev)
    lr_bytes = 2 * 4 * k * (1 / 4 + 1 / 16) + 8        # read + written low-res tensors, int64 target
    med, best = timed(lambda: ops.uw_ce_lowres_fwd_bwd(main_lr, aux_lr, target, cw), iters, flush=flush)
    report("loss_b64_lowres_fused_upsample_fwd_bwd", b * h * w, lr_bytes, med, best, launches_per_step=1,
           note="pre-upsample heads in, gradients w.r.t. them out (%.1f B/pixel); compute-bound" % lr_bytes)
    medf, bestf = timed(lambda: ops.uw_ce_lowres_fwd_bwd(main_lr, aux_lr, target, cw, backward=False), iters, flush=flush)
    report("loss_b64_lowres_fused_upsample_fwd_only", b * h * w, lr_bytes / 2 + 4, medf, bestf, launches_per_step=1)

    def unfused():
        m, a = main_lr.detach().requires_grad_(True), aux_lr.detach().requires_grad_(True)
        mu = F.interpolate(m, size=(h, w), mode="bilinear", align_corners=True)
        au = F.interpolate(a, size=(h, w), mode="bilinear", align_corners=True)
        ops.uw_ce_loss(mu, au, target, cw).backward()
    med2, best2 = timed(unfused, iters, flush=flush)
    report("loss_b64_torch_upsample_then_fused_k4_fwd_bwd", b * h * w, lr_bytes, med2, best2,
           note="F.interpolate x2 + K4 + upsample_bilinear2d_backward x2 (PyTorch); speed-up of the fused kernel: %.2fx" % (med2 / med))


def nid_section(dev, iters):
    """NIDLoss forward + backward at the training batch (64 x 480x256, 16 intensity bins, 5 label bins)."""
    from mspl_b200.loss_fns.segmentation_loss import NIDLoss
    b, h, w = 64, 256, 480
    g = torch.Generator(device=dev).manual_seed(2)
    camera = torch.rand((b, 3, h, w), device=dev, generator=g)
    label = 3 * torch.randn((b, 5, h, w), device=dev, generator=g)
    crit = NIDLoss(image_bin=16, label_bin=5)

    def step():
        lab = label.detach().requires_grad_(True)
        crit(camera, lab).backward()
    med, best = timed(step, max(3, iters // 2))
    report("nid_loss_b64_fwd_bwd", b * h * w, 4 * (3 + 5) * 2 + 4 * 5, med, best,
           note="one fused pass per direction; compute-bound (2 x (32 + 10) sigmoids per pixel and image)")
    med, best = timed(lambda: crit(camera, label), max(3, iters // 2))
    report("nid_loss_b64_fwd_only", b * h * w, 4 * (3 + 5), med, best)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--section", default="all", choices=("all", "loss", "policies", "stress", "io", "lowres", "nid", "loss_lowres"))
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    iters = 5 if args.quick else 20
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)      # 256 MB > 126 MB L2
    h, w, k = 256, 480, 5

    # ---- config 4: loss ------------------------------------------------------------------------------------------
    for b in ((64, 8, 256) if args.section in ("all", "loss") else ()):
        main_l, aux_l = logits(b, k, h, w, dev, 11)
        target = torch.randint(1, 5, (b, h, w), device=dev)
        cw = torch.tensor([1.0, 1.0, 1.0, 1.0, 0.0], device=dev)
        med, best = timed(lambda: ops.uw_ce_fwd_bwd(main_l, aux_l, target, cw), iters, flush=flush)
        report("loss_b%d_fused_fwd_bwd" % b, b * h * w, 88, med, best, launches_per_step=1)
        med, best = timed(lambda: ops.uw_ce_fwd_bwd(main_l, aux_l, target, cw, backward=False), iters, flush=flush)
        report("loss_b%d_fused_fwd_only" % b, b * h * w, 48, med, best, launches_per_step=1)
        target8 = target.to(torch.uint8)
        med, best = timed(lambda: ops.uw_ce_fwd_bwd(main_l, aux_l, target8, cw), iters, flush=flush)
        report("loss_b%d_fused_fwd_bwd_u8_targets" % b, b * h * w, 81, med, best, launches_per_step=1,
               note="labels read as the uint8 maps the generator wrote (1 B/pixel instead of int64's 8)")
        med, best = timed(lambda: ops.uw_ce_fwd_bwd(main_l, aux_l, target8, cw, backward=False), iters, flush=flush)
        report("loss_b%d_fused_fwd_only_u8_targets" % b, b * h * w, 41, med, best, launches_per_step=1)
        # the training loop's pair of statements (uest_seg_multi_os.py:1020-1032): loss, then miou_class.get_iou(pred, labels)
        counts = torch.zeros((3, k), dtype=torch.int64, device=dev)
        med, best = timed(lambda: ops.uw_ce_fwd_bwd(main_l, aux_l, target, cw, iou_counts=counts), iters, flush=flush)
        report("loss_b%d_fused_fwd_bwd_with_iou_counts" % b, b * h * w, 88, med, best, launches_per_step=1,
               note="loss forward+backward and the MIOU.get_iou counts of the same tensors in one launch")
        med, best = timed(lambda: ops.uw_ce_fwd_bwd(main_l, aux_l, target8, cw, iou_counts=counts), iters, flush=flush)
        report("loss_b%d_fused_fwd_bwd_with_iou_counts_u8_targets" % b, b * h * w, 81, med, best, launches_per_step=1)

        med, best = timed(lambda: ops.miou_counts(main_l, target, k, counts=counts), iters, flush=flush)
        report("miou_counts_from_logits_b%d" % b, b * h * w, 4 * k + 8, med, best, launches_per_step=1,
               note="stand-alone MIOU.get_iou counting kernel (argmax of the logits + int64 labels)")

        def two_launches():
            ops.uw_ce_fwd_bwd(main_l, aux_l, target, cw)
            ops.miou_counts(main_l, target, k, counts=counts)
        med, best = timed(two_launches, iters, flush=flush)
        report("loss_b%d_fused_fwd_bwd_then_miou_kernel" % b, b * h * w, 88 + 4 * k + 8, med, best, launches_per_step=2,
               note="the same work as two launches: K4, then the stand-alone metric kernel re-reading the main logits and labels")
        if b == 64 and args.section == "all":
            crit = UncertaintyWeightedSegmentationLoss(k, class_weights=cw.clone(), ignore_idx=4, device=dev)
            kld_layer = PixelwiseKLD()

            def modules():
                p, q = main_l.clone().requires_grad_(True), aux_l.clone().requires_grad_(True)
                kld = kld_layer(p, q)
                loss = crit(p + 0.5 * q, target, kld) * 20 + kld.mean()
                loss.backward()
            med, best = timed(modules, iters, flush=flush)
            report("loss_b64_reference_named_modules_autograd", b * h * w, 88, med, best,
                   note="PixelwiseKLD + UncertaintyWeightedSegmentationLoss kernels + torch glue (clone, add, mul, mean, autograd)")

            def torch_eager():
                import torch.nn.functional as F
                p, q = main_l.clone().requires_grad_(True), aux_l.clone().requires_grad_(True)
                p1, lp1, lp2 = F.softmax(p, 1), F.log_softmax(p, 1), F.log_softmax(q, 1)
                kld = (p1 * lp1 - p1 * lp2).sum(1)
                lp = -F.log_softmax(p + 0.5 * q, 1) * cw.reshape(1, -1, 1, 1)
                loss = (lp.gather(1, target.view(b, 1, h, w)) * torch.exp(-kld.reshape(b, 1, h, w))).mean() * 20 + kld.mean()
                loss.backward()
            med, best = timed(torch_eager, max(3, iters // 4), flush=flush)
            report("loss_b64_torch_eager_cuda_for_context", b * h * w, 88, med, best,
                   note="the reference's op sequence run by PyTorch eager on this GPU (not a CPU baseline)")
        del main_l, aux_l, target, target8

    if args.section in ("all", "io"):
        label_io_section(dev)
    if args.section in ("all", "lowres"):
        lowres_section(dev, iters)
    if args.section in ("all", "loss_lowres"):
        loss_lowres_section(dev, iters, flush)
    if args.section in ("all", "nid"):
        nid_section(dev, iters)
    if args.section in ("loss", "io", "lowres", "nid", "loss_lowres"):
        return
    # ---- non-headline policies on the 13/20/5 configuration -----------------------------------------------------------
    n = 100 if args.quick else 400
    srcs = (("camvid", 13), ("cityscapes", 20), ("forest", 5))
    mains, auxs = zip(*[logits(n, c, h, w, dev, 3 + i) for i, (_, c) in enumerate(srcs)])
    luts = [SOURCE_TABLES[s] for s, _ in srcs]
    for policy in ("all", "half", "prob"):
        med, best = timed(lambda: ops.fuse_sources(list(mains), list(auxs), luts, policy=policy), iters)
        report("k1_policy_%s_%dimg" % (policy, n), n * h * w, 313, med, best)
    for policy in ("all", "half"):
        med, best = timed(lambda: ops.fuse_sources(list(mains), list(auxs), luts, policy=policy, want_conf=False, want_unc=False,
                                                   want_conf_hist=False, count_marginal=False), iters)
        report("k1_labels_only_%s_%dimg" % (policy, n), n * h * w, 305, med, best,
               note="reference-exact output only (label map + class counts): no softmax needed")
    gen = LabelGenerator(luts, policy="all")
    med, best = timed(lambda: gen.run(list(mains), list(auxs)), iters)
    report("labelgen_all_thresholds_%dimg" % n, n * h * w, 319, med, best)
    del mains, auxs

    # ---- config 5: stress ---------------------------------------------------------------------------------------------
    hs, ws = 512, 1024
    ns = 16 if args.quick else 64
    mains, auxs = zip(*[logits(ns, 20, hs, ws, dev, 30 + i) for i in range(3)])
    luts = [SOURCE_TABLES["cityscapes"]] * 3
    med, best = timed(lambda: ops.fuse_sources(list(mains), list(auxs), luts, policy="all"), iters)
    report("stress_1024x512_3x20cls_k1_%dimg" % ns, ns * hs * ws, 489, med, best)
    gen = LabelGenerator(luts, policy="all")
    med, best = timed(lambda: gen.run(list(mains), list(auxs)), iters)
    report("stress_1024x512_3x20cls_labelgen_%dimg" % ns, ns * hs * ws, 495, med, best)


if __name__ == "__main__":
    main()
