#!/usr/bin/env python
"""Timing sweep of the threshold stage (bracketed protocol) on label/conf maps shaped like bench.py's: most pixels carry
the ignore label with conf 0, the rest a confidence in [0.3, 1].  MSPL_CLASSIFY_VARIANT selects a build-time variant of the
classify kernel when the library was compiled with several (development knob)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mspl_b200 import ops  # noqa: E402


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2], t[0]


def main():
    dev = torch.device("cuda:0")
    n, h, w = int(os.environ.get("SWEEP_IMAGES", "2000")), 256, 480
    gen = torch.Generator(device=dev).manual_seed(5)
    label = torch.full((n, h, w), 4, dtype=torch.uint8, device=dev)
    conf = torch.zeros((n, h, w), dtype=torch.float32, device=dev)
    for lo in range(0, n, 100):
        sl = slice(lo, min(n, lo + 100))
        u = torch.rand(label[sl].shape, device=dev, generator=gen)
        lab = torch.where(u < 0.0015, 1, torch.where(u < 0.119, 2, torch.where(u < 0.1237, 3, 4))).to(torch.uint8)
        c = 0.3 + 0.7 * torch.rand(label[sl].shape, device=dev, generator=gen) ** 0.5
        label[sl] = lab
        conf[sl] = torch.where(lab == 4, torch.zeros_like(c), c)
    npix = label.numel()
    print("pixels %.1f M" % (npix / 1e6))
    variants = [int(v) for v in os.environ.get("SWEEP_VARIANTS", "0").split(",")]
    for v in variants:
        os.environ["MSPL_CLASSIFY_VARIANT"] = str(v)
        hist0 = torch.zeros((5, ops.RADIX_BINS), dtype=torch.int64, device=dev)

        def with_hist():
            ops.select_and_apply(label, conf, 0.2, 1, 5, 4)

        med, mn = timed(with_hist)
        print("variant %d  conf_hist + select_and_apply: median %.3f ms  min %.3f ms" % (v, med, mn))
        ops._lib.check(ops._lib.load().mspl_conf_hist(ops._ptr(label), ops._ptr(conf), npix, h * w, 5, ops._ptr(hist0), 1,
                                                      ops._stream(dev)), "conf_hist")

        def staged():
            ops.select_and_apply(label, conf, 0.2, 1, 5, 4, conf_hist=hist0.clone())

        med, mn = timed(staged)
        print("variant %d  select_and_apply (histogram given): median %.3f ms  min %.3f ms  -> %.1f GB/s at 6 B/pixel" %
              (v, med, mn, npix * 6 / med / 1e6))
    # the classify kernel alone (raw ABI), bracket taken from a real run's thresholds
    lib = ops._lib.load()
    th, _, _, _, _ = ops.select_and_apply(label, conf, 0.2, 1, 5, 4)
    lo = torch.floor(th * 2048) / 2048
    bracket = torch.stack([lo, lo + 1.0 / 2048], 1).contiguous()
    bracket[0] = 1.0
    final = torch.empty_like(label)
    fh = torch.zeros(5, dtype=torch.int64, device=dev)
    cand = torch.empty(npix, dtype=torch.int32, device=dev)
    cnt = torch.zeros((), dtype=torch.int64, device=dev)
    st = ops._stream(dev)
    for v in variants:
        os.environ["MSPL_CLASSIFY_VARIANT"] = str(v)

        def classify():
            cnt.zero_()
            ops._lib.check(lib.mspl_bracket_classify(ops._ptr(label), ops._ptr(conf), ops._ptr(bracket), npix, 5, 4, ops._ptr(final),
                                                     None, ops._ptr(fh), ops._ptr(cand), ops._ptr(cnt), st), "classify")

        med, mn = timed(classify)
        print("variant %d  classify alone: median %.3f ms  min %.3f ms -> %.1f GB/s at 6 B/pixel, %d candidates" %
              (v, med, mn, npix * 6 / med / 1e6, int(cnt)))
    med, mn = timed(lambda: ops.apply_thresholds(label, conf, th, 4, want_mask=False, final_hist=fh))
    print("stand-alone apply_thresholds kernel: median %.3f ms  min %.3f ms -> %.1f GB/s at 6 B/pixel" % (med, mn, npix * 6 / med / 1e6))
    hist0 = torch.zeros((5, ops.RADIX_BINS), dtype=torch.int64, device=dev)
    med, mn = timed(lambda: lib.mspl_conf_hist(ops._ptr(label), ops._ptr(conf), npix, h * w, 5, ops._ptr(hist0), 1, st))
    print("stand-alone conf_hist kernel: median %.3f ms  min %.3f ms -> %.1f GB/s at 5 B/pixel" % (med, mn, npix * 5 / med / 1e6))
    med, mn = timed(lambda: final.copy_(label))
    print("torch copy u8 (2 B/pixel): median %.3f ms -> %.1f GB/s" % (med, npix * 2 / med / 1e6))
    conf2 = torch.empty_like(conf)
    med, mn = timed(lambda: conf2.copy_(conf))
    print("torch copy f32 (8 B/pixel): median %.3f ms -> %.1f GB/s" % (med, npix * 8 / med / 1e6))
    del conf2
    th, kept, final, _, fh = ops.select_and_apply(label, conf, 0.2, 1, 5, 4)
    print("thresholds", th.tolist(), "final_hist", fh.tolist())

    def old():
        t, _ = ops.cb_thresholds_radix(label, conf, 0.2)
        ops.apply_thresholds(label, conf, t, 4, want_mask=False)

    med, mn = timed(old, iters=10)
    print("generic radix (3 full passes) + apply: median %.3f ms  min %.3f ms" % (med, mn))


if __name__ == "__main__":
    main()
