"""-m gpu: seeded differential sweep of the fusion + threshold kernels against the oracle over random configurations --
number of sources, source class counts, target class count, ignore class, label tables (including target classes no source
class maps to and tables that map to the ignore class), vote policies, image shapes (aligned, odd, single row), logit scales
and exact ties -- beyond the handful of fixed cases in test_gpu_parity.py.  The configurations are drawn from a fixed seed,
so every run sees the same 48 cases."""
import random

import numpy as np
import pytest
import torch

from oracle import mspl_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-5
KLD_ATOL = 2e-6


def _cases(count, seed):
    rng = random.Random(seed)
    cases = []
    for i in range(count):
        S = rng.choice([1, 1, 2, 3, 3, 3])
        K = rng.choice([2, 3, 5, 5, 5, 8])
        ignore = rng.choice([K - 1, K - 1, 0 if K > 2 else K - 1, rng.randrange(K)])
        cls = [rng.choice([2, 3, 5, 7, 13, 20, 24]) for _ in range(S)]
        luts = [[rng.randrange(1 if rng.random() < 0.7 else 0, K) for _ in range(c)] for c in cls]
        policy = rng.choice(["half", "all", "prob", "prob", 1, S])
        shape = rng.choice([(2, 16, 24), (1, 9, 7), (3, 1, 40), (2, 31, 5), (1, 48, 80), (2, 8, 480), (1, 6, 962)])
        sigma = rng.choice([0.05, 1.0, 3.0, 3.0, 12.0, 40.0])
        ties = rng.random() < 0.2
        portion = rng.choice([0.05, 0.2, 0.5, 1.0])
        ds_rate = rng.choice([1, 1, 1, 2, 5])
        cases.append((i, S, K, ignore, cls, luts, policy, shape, sigma, ties, portion, ds_rate))
    return cases


@pytest.mark.parametrize("case", _cases(48, seed=20261018), ids=lambda c: "case%02d" % c[0])
def test_random_configuration_matches_oracle(case):
    from mspl_b200 import ops
    i, S, K, ignore, cls, luts, policy, (n, h, w), sigma, ties, portion, ds_rate = case
    dev = torch.device("cuda:0")
    mains, auxs = [], []
    for s, c in enumerate(cls):
        m, a = O.synthetic_logits(n, c, h, w, seed=1000 * i + s, sigma=sigma)
        if ties:                       # quantised logits: many exact ties between classes and between heads
            m, a = torch.round(m), torch.round(a * 2) / 2
        mains.append(m), auxs.append(a)
    saved = O.IGNORE_LABEL          # the reference hard-codes ignore id 4 in merge_outputs (:716); the kernels take it as a parameter
    O.IGNORE_LABEL = ignore
    try:
        ref = O.fuse_sources(mains, auxs, luts, policy, K, ignore)
    finally:
        O.IGNORE_LABEL = saved
    r = ops.fuse_sources([m.to(dev) for m in mains], [a.to(dev) for a in auxs], luts, policy=policy, num_classes=K,
                         ignore_label=ignore, ds_rate=ds_rate, want_kld=True)
    lab = r.label.cpu()
    diff = lab != ref["label"]
    assert not bool((diff & ~ref["marginal"]).any()), "%d mismatches outside near-ties" % int((diff & ~ref["marginal"]).sum())
    ok = ~diff
    torch.testing.assert_close(r.conf.cpu()[ok], ref["conf"][ok], rtol=RTOL, atol=1e-7)
    # KLD is a difference of O(|logit|) terms in fp32 (in the reference too): the absolute floor scales with the logits
    kld_atol = KLD_ATOL * max(1.0, sigma)
    torch.testing.assert_close(r.unc.cpu(), ref["unc"], rtol=RTOL, atol=kld_atol)
    for got, want in zip(r.kld, ref["kld"]):
        torch.testing.assert_close(got.cpu(), want, rtol=RTOL, atol=kld_atol)
    assert torch.equal(r.class_hist.cpu(), torch.bincount(lab.reshape(-1).long(), minlength=K))
    assert float(r.conf.min()) >= 0.0 and float(r.conf.max()) <= 1.0 + 1e-6

    # thresholds + selection on the kernel's own (label, conf): exact order statistic, exact selection
    th_ref, kept_ref = O.cb_thresholds(lab, r.conf.cpu(), portion, ds_rate, K, ignore=ignore)
    f_ref, m_ref = O.apply_thresholds(lab, r.conf.cpu(), th_ref, ignore)
    th, kept, final, mask, fh = ops.select_and_apply(r.label, r.conf, portion, ds_rate, K, ignore, conf_hist=r.conf_hist,
                                                     want_mask=True)
    assert torch.equal(th.cpu(), th_ref) and torch.equal(kept.cpu(), kept_ref)
    assert torch.equal(final.cpu(), f_ref) and torch.equal(mask.cpu(), m_ref)
    assert torch.equal(fh.cpu(), torch.bincount(f_ref.reshape(-1).long(), minlength=K))
    # the generic radix protocol resolves the same thresholds
    th2, _ = ops.cb_thresholds_radix(r.label, r.conf, portion, ds_rate, K)
    sel = torch.arange(K) != ignore
    assert torch.equal(th2.cpu()[sel], th_ref[sel])
    # labels-only kernel (vote policies): the reference-exact output
    if policy != "prob":
        lo = ops.fuse_sources([m.to(dev) for m in mains], [a.to(dev) for a in auxs], luts, policy=policy, num_classes=K,
                              ignore_label=ignore, want_conf=False, want_unc=False, want_conf_hist=False)
        assert torch.equal(lo.label, r.label) and torch.equal(lo.class_hist, r.class_hist)


def _loss_cases(count, seed):
    rng = random.Random(seed)
    cases = []
    for i in range(count):
        k = rng.choice([1, 2, 3, 5, 5, 5, 7, 8])
        shape = rng.choice([(2, 16, 24), (1, 9, 7), (3, 1, 40), (2, 31, 5), (1, 48, 80), (4, 8, 480), (1, 2, 2)])
        sigma = rng.choice([0.05, 1.0, 3.0, 3.0, 12.0, 40.0])
        independent = rng.random() < 0.3        # aux head unrelated to the main head: large KLD, tiny exp(-KLD)
        zero_w = rng.random() < 0.5
        bad_targets = rng.random() < 0.4
        u8 = rng.random() < 0.5
        cases.append((i, k, shape, sigma, independent, zero_w, bad_targets, u8))
    return cases


@pytest.mark.parametrize("case", _loss_cases(40, seed=77), ids=lambda c: "loss%02d" % c[0])
def test_random_loss_configuration_matches_oracle(case):
    """K4 (fused forward+backward, int64 / uint8 targets, with and without the folded MIOU counts), the reference-named
    modules (PixelwiseKLD, UncertaintyWeightedSegmentationLoss) and K0 (get_output's maps) against the fp64 / fp32 oracle."""
    from mspl_b200 import ops
    from mspl_b200.loss_fns.segmentation_loss import PixelwiseKLD, UncertaintyWeightedSegmentationLoss
    i, k, (b, h, w), sigma, independent, zero_w, bad_targets, u8 = case
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(5000 + i)
    main, aux = O.synthetic_logits(b, k, h, w, seed=7000 + i, sigma=sigma)
    if independent:
        aux = sigma * torch.randn(b, k, h, w, generator=gen)
    target = torch.randint(0, k, (b, h, w), generator=gen)
    cw = torch.rand(k, generator=gen) * 3 + 0.1
    if zero_w:
        cw[rng_index(gen, k)] = 0.0
    m, a, c = main.to(dev), aux.to(dev), cw.to(dev)

    # ---- K4 ----
    t_or, cw_or, m_or, a_or = target, cw, main, aux
    if bad_targets:                    # labels outside [0,K) (255 = the PNG ignore value) behave like a zero-weight class
        target = target.clone()
        target.view(-1)[:: 7] = 255
        pad = torch.full((b, 1, h, w), -1e4)
        t_or = target.clone()
        t_or[target == 255] = k
        cw_or = torch.cat([cw, torch.zeros(1)])
        m_or, a_or = torch.cat([main, pad], 1), torch.cat([aux, pad], 1)
    l64, gm64, ga64 = O.training_loss_and_grads(m_or, a_or, t_or, cw_or, dtype=torch.float64)
    gm64, ga64 = gm64[:, :k], ga64[:, :k]
    t_dev = target.to(dev).to(torch.uint8 if u8 else torch.int64)
    counts = torch.zeros((3, k), dtype=torch.int64, device=dev) if k >= 2 else None
    out3, dm, da = ops.uw_ce_fwd_bwd(m, a, t_dev, c, iou_counts=counts)
    assert abs(out3[0].item() - l64.item()) <= RTOL * abs(l64.item()) + 1e-9
    # per-pixel gradients: relative to the largest gradient of the tensor (they are 1/N-scaled sums of O(1) terms)
    scale = float(max(gm64.abs().max(), ga64.abs().max()))
    torch.testing.assert_close(dm.cpu().double(), gm64, rtol=1e-4, atol=1e-5 * scale)
    torch.testing.assert_close(da.cpu().double(), ga64, rtol=1e-4, atol=1e-5 * scale)
    if counts is not None:
        inter, union = O.miou_get_iou(main, target, num_classes=k)
        np.testing.assert_array_equal(counts[0].cpu().numpy().astype(np.float32), inter)
        np.testing.assert_array_equal((counts[1] + counts[2] - counts[0]).cpu().numpy().astype(np.float32) + np.float32(1e-6), union)
        assert torch.equal(counts, ops.miou_counts(m, target.to(dev), k))

    # ---- the reference-named modules, values and autograd gradients (in-range targets only: torch.gather needs them) ----
    if not bad_targets:
        p, q = m.clone().requires_grad_(True), a.clone().requires_grad_(True)
        kld = PixelwiseKLD()(p, q)
        crit = UncertaintyWeightedSegmentationLoss(k, class_weights=c.clone(), ignore_idx=None, device=dev)
        loss = crit(p + 0.5 * q, target.to(dev), kld) * 20 + kld.mean()
        loss.backward()
        kld_ref = O.pixelwise_kld(main.double(), aux.double())
        torch.testing.assert_close(kld.detach().cpu().double(), kld_ref, rtol=RTOL, atol=KLD_ATOL * max(1.0, sigma))
        assert abs(loss.item() - l64.item()) <= 2e-5 * abs(l64.item()) + 1e-9
        torch.testing.assert_close(p.grad.cpu().double(), gm64, rtol=1e-4, atol=2e-5 * scale)
        torch.testing.assert_close(q.grad.cpu().double(), ga64, rtol=1e-4, atol=2e-5 * scale)

    # ---- K0: get_output's softmax + KLD maps ----
    prob, kmap = ops.softmax_kld(m, a)
    want_p = torch.softmax(main + 0.5 * aux, dim=1)
    torch.testing.assert_close(prob.cpu(), want_p, rtol=RTOL, atol=1e-7)
    torch.testing.assert_close(kmap.cpu(), O.pixelwise_kld(main, aux), rtol=RTOL, atol=KLD_ATOL * max(1.0, sigma))


def rng_index(gen, k):
    return int(torch.randint(0, k, (1,), generator=gen))


def _lowres_cases(count, seed):
    rng = random.Random(seed)
    cases = []
    for i in range(count):
        H = rng.choice([8, 12, 20, 33, 48, 64, 70])
        W = rng.choice([16, 48, 96, 240, 480, 496])
        # head sizes between 1/8 of the output and the output itself (upsampling only), independent per axis; the fused
        # kernels bulk-copy whole source rows, so row lengths are multiples of 4 floats (their documented contract)
        hm, wm = rng.randint(max(1, H // 8), H), 4 * rng.randint(max(1, W // 32), W // 4)
        ha, wa = rng.randint(max(1, H // 8), H), 4 * rng.randint(max(1, W // 32), W // 4)
        espdnet = rng.random() < 0.4
        if espdnet:                                 # the ESPDNetUE geometry: x2 and x4
            hm, wm, ha, wa = (H + 1) // 2, W // 2, (H + 3) // 4, W // 4
        S = rng.choice([1, 2, 3])
        cls = [rng.choice([2, 5, 13, 20]) for _ in range(S)]
        K = rng.choice([3, 5, 5, 8])
        luts = [[rng.randrange(1, K) for _ in range(c)] for c in cls]
        policy = rng.choice(["half", "all", "prob"])
        n = rng.choice([1, 2, 3])
        cases.append((i, n, H, W, (hm, wm), (ha, wa), S, cls, K, luts, policy, espdnet))
    return cases


@pytest.mark.parametrize("case", _lowres_cases(32, seed=4242), ids=lambda c: "lowres%02d" % c[0])
def test_random_lowres_geometry_matches_oracle(case):
    """The fused-upsample kernels (K1-lowres, K4-lowres) over random output sizes, head sizes (integer and non-integer scale
    factors, heads already at full size, partial tiles) and source sets, against upsample-then-fuse / upsample-then-loss with
    autograd on the CPU (their parity definitions and tolerances, tests/test_gpu_parity.py)."""
    from mspl_b200 import ops
    i, n, H, W, (hm, wm), (ha, wa), S, cls, K, luts, policy, espdnet = case
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(9000 + i)
    mains = [(3 * torch.randn(n, c, hm, wm, generator=gen) + 3 * torch.randn(n, c, 1, 1, generator=gen)).contiguous() for c in cls]
    auxs = [(3 * torch.randn(n, c, ha, wa, generator=gen)).contiguous() for c in cls]
    ignore = K - 1
    saved = O.NEAR_TIE_MARGIN, O.IGNORE_LABEL
    O.NEAR_TIE_MARGIN, O.IGNORE_LABEL = 1e-5, ignore
    try:
        ref = O.fuse_sources_lowres(mains, auxs, luts, (H, W), policy, K, ignore)
    finally:
        O.NEAR_TIE_MARGIN, O.IGNORE_LABEL = saved
    try:
        r = ops.fuse_sources_lowres([m.to(dev) for m in mains], [a.to(dev) for a in auxs], luts, (H, W), policy=policy,
                                    num_classes=K, ignore_label=ignore)
    except NotImplementedError:
        r = None                      # source rows too wide for the shared-memory ring (documented): callers upsample instead
        assert not espdnet, "the ESPDNetUE geometry must be served by the fused kernel"
    if r is not None:
        diff = r.label.cpu() != ref["label"]
        assert not bool((diff & ~ref["marginal"]).any()), "%d mismatches outside near-ties" % int((diff & ~ref["marginal"]).sum())
        ok = ~diff
        torch.testing.assert_close(r.conf.cpu()[ok], ref["conf"][ok], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(r.unc.cpu(), ref["unc"], rtol=1e-4, atol=1e-5)
        assert torch.equal(r.class_hist, torch.bincount(r.label.reshape(-1).long(), minlength=K))

    # ---- K4-lowres on the first source's heads (K <= 8 classes) ----
    k = min(cls[0], 8)
    ml, al = mains[0][:, :k].contiguous(), auxs[0][:, :k].contiguous()
    target = torch.randint(0, k, (n, H, W), generator=gen)
    cw = torch.rand(k, generator=gen) * 2 + 0.1
    l64, gm64, ga64 = O.training_loss_lowres_and_grads(ml, al, target, cw, dtype=torch.float64)
    try:
        out3, dm, da = ops.uw_ce_lowres_fwd_bwd(ml.to(dev), al.to(dev), target.to(dev).to(torch.uint8 if i % 2 else torch.int64), cw.to(dev))
    except NotImplementedError:
        return
    assert abs(out3[0].item() - l64.item()) <= RTOL * abs(l64.item())
    scale = float(max(gm64.abs().max(), ga64.abs().max()))
    torch.testing.assert_close(dm.cpu().double(), gm64, rtol=1e-4, atol=5e-5 * scale)
    torch.testing.assert_close(da.cpu().double(), ga64, rtol=1e-4, atol=5e-5 * scale)


@pytest.mark.parametrize("num_classes,c,shape", [(2, 2, (2, 8, 12)), (5, 5, (3, 31, 7)), (21, 21, (2, 16, 20)), (32, 40, (1, 24, 36)),
                                                 (33, 33, (2, 16, 20)), (40, 48, (1, 24, 36)), (200, 64, (1, 16, 16)),
                                                 (5, 5, (40, 64, 96))])
def test_miou_counts_any_class_count(num_classes, c, shape):
    """MIOU.get_iou counting for class counts on both sides of the per-thread-counter / warp-aggregated split (32), more
    logit channels than metric classes (predictions outside histc's range are not counted), 255 and other out-of-range labels,
    logits / uint8 / int64 predictions, and enough pixels per thread to force mid-loop counter flushes."""
    from mspl_b200 import ops
    dev = torch.device("cuda:0")
    b, h, w = shape
    gen = torch.Generator().manual_seed(num_classes * 1000 + c)
    logits = torch.randn(b, c, h, w, generator=gen)
    logits[0, :, 0, :3] = 0.5                                   # ties -> first index
    target = torch.randint(0, max(c, num_classes) + 3, (b, h, w), generator=gen)
    target[0, 0, 3:6] = 255
    inter, union = O.miou_get_iou(logits, target, num_classes=num_classes)
    got = ops.miou_counts(logits.to(dev), target.to(dev), num_classes)
    np.testing.assert_array_equal(got[0].cpu().numpy().astype(np.float32), inter)
    np.testing.assert_array_equal((got[1] + got[2] - got[0]).cpu().numpy().astype(np.float32) + np.float32(1e-6), union)
    pred = logits.argmax(1)
    for p in (pred.to(dev), pred.to(torch.uint8).to(dev)):
        assert torch.equal(ops.miou_counts(p, target.to(dev), num_classes), got)
    # accumulation into a caller-provided tensor
    acc = got.clone()
    ops.miou_counts(logits.to(dev), target.to(dev), num_classes, counts=acc)
    assert torch.equal(acc, 2 * got)


def test_packed_counters_flush_at_scale():
    """The per-thread packed counters hold < 1,024 pixels per field and are folded into the CTA totals before that: run both
    counting kernels on enough pixels (> 1,016 per thread of a full persistent grid) that every thread flushes mid-loop, and
    compare with counts formed by torch ops on the device (MIOU.get_iou's definition, utilities/metrics/segmentation_miou.py:30-40)."""
    from mspl_b200 import ops
    dev = torch.device("cuda:0")
    k = 5
    gen = torch.Generator(device=dev).manual_seed(12)

    def expected(pred, target):
        ts = (target + 1) & 255
        ps = (pred.long() + 1) * (ts > 0)
        inter = ps * (ps == ts)
        return torch.stack([torch.bincount(x.reshape(-1), minlength=k + 3)[1:k + 1] for x in (inter, ps, ts)])

    # stand-alone kernel, label input: 1 pixel per thread and iteration, 148 x 6 CTAs x 256 threads
    npx = 148 * 6 * 256 * 1100
    pred = torch.randint(0, k, (npx,), device=dev, generator=gen, dtype=torch.uint8)
    target = torch.randint(0, k + 2, (npx,), device=dev, generator=gen)
    target[::97] = 255
    assert torch.equal(ops.miou_counts(pred, target, k), expected(pred, target))
    del pred, target

    # K4 forward with the counts folded in: 4 pixels per thread and iteration, 148 x 3 CTAs x 256 threads
    h, w = 10000, 13000
    main = torch.randn((1, k, h, w), device=dev, generator=gen)
    aux = main + 0.5
    target = torch.randint(0, k + 1, (1, h, w), device=dev, generator=gen).to(torch.uint8)
    counts = torch.zeros((3, k), dtype=torch.int64, device=dev)
    out3, _, _ = ops.uw_ce_fwd_bwd(main, aux, target, torch.ones(k, device=dev), backward=False, iou_counts=counts)
    assert torch.equal(counts, expected(main.argmax(1), target.long()))
    assert torch.isfinite(out3).all()


def _threshold_cases(count, seed):
    rng = random.Random(seed)
    cases = []
    for i in range(count):
        K = rng.choice([2, 5, 5, 8])
        ignore = rng.choice([K - 1, K - 1, 0, None])
        shape = rng.choice([(1, 7, 9), (2, 32, 48), (3, 50, 70), (1, 256, 480), (5, 64, 100), (1, 1, 4)])
        dist = rng.choice(["uniform", "few_values", "all_equal", "bin_edges", "near_one", "tiny", "mixed"])
        portion = rng.choice([0.0, 1e-4, 0.05, 0.2, 0.5, 0.999, 1.0])
        ds_rate = rng.choice([1, 1, 2, 3, 7])
        cases.append((i, K, ignore, shape, dist, portion, ds_rate))
    return cases


@pytest.mark.parametrize("case", _threshold_cases(36, seed=99), ids=lambda c: "thr%02d" % c[0])
def test_random_threshold_inputs_match_sort_definition(case):
    """Class-balanced thresholds + selection on adversarial confidence distributions -- heavy duplicates (every pixel inside
    the bracket bin), all-equal values, values exactly on histogram bin edges, values within an ulp of 1, denormal-sized
    values -- exactly equal to the sort-based definition, for the bracketed protocol and the generic radix one."""
    from mspl_b200 import ops
    i, K, ignore, (n, h, w), dist, portion, ds_rate = case
    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(300 + i)
    label = torch.randint(0, K, (n, h, w), generator=gen).to(torch.uint8)
    u = torch.rand(n, h, w, generator=gen)
    if dist == "few_values":
        conf = torch.tensor([0.25, 0.5, 0.5000001, 0.75, 0.9])[torch.randint(0, 5, (n, h, w), generator=gen)]
    elif dist == "all_equal":
        conf = torch.full((n, h, w), 0.3333333)
    elif dist == "bin_edges":
        conf = torch.randint(0, 2049, (n, h, w), generator=gen).float() / 2048
    elif dist == "near_one":
        conf = 1.0 - torch.randint(0, 4, (n, h, w), generator=gen).float() * 5.9604645e-08
    elif dist == "tiny":
        conf = u * 1e-38
    elif dist == "mixed":
        conf = torch.where(u < 0.3, torch.zeros(()), torch.where(u < 0.6, torch.ones(()), u))
    else:
        conf = u
    conf = conf.float().contiguous()
    if ignore is not None:                       # as the fusion kernel writes it: the ignore class carries conf 0
        conf = torch.where(label == ignore, torch.zeros(()), conf)
    th_ref, kept_ref = O.cb_thresholds(label, conf, portion, ds_rate, K, ignore=ignore)
    lab_d, conf_d = label.to(dev), conf.to(dev)
    th, kept, final, mask, fh = ops.select_and_apply(lab_d, conf_d, portion, ds_rate, K, ignore, want_final=ignore is not None,
                                                     want_mask=ignore is not None)
    assert torch.equal(th.cpu(), th_ref), (th.cpu(), th_ref)
    assert torch.equal(kept.cpu(), kept_ref)
    if ignore is not None:
        f_ref, m_ref = O.apply_thresholds(label, conf, th_ref, ignore)
        assert torch.equal(final.cpu(), f_ref) and torch.equal(mask.cpu(), m_ref)
        assert torch.equal(fh.cpu(), torch.bincount(f_ref.reshape(-1).long(), minlength=K))
        f2, m2, fh2 = ops.apply_thresholds(lab_d, conf_d, th, ignore_label=ignore)        # the stand-alone K3 kernel
        assert torch.equal(f2, final) and torch.equal(m2, mask) and torch.equal(fh2, fh)
    th2, kept2 = ops.cb_thresholds_radix(lab_d, conf_d, portion, ds_rate, K)
    sel = torch.arange(K) != (ignore if ignore is not None else -1)
    assert torch.equal(th2.cpu()[sel], th_ref[sel]) and torch.equal(kept2.cpu(), kept_ref)


def test_threshold_stage_beyond_2_31_pixels():
    """A 20,000-image target set thresholded on ONE GPU is 2,457,600,000 pixels -- more than 2^31 (BASELINE configs[2] kept
    as label + conf, 5 B/pixel): pixel counts and offsets are 64-bit, candidate indices unsigned 32-bit.  Checked by counting:
    thresh[k] is the j-th largest conf of class k iff #(conf >= t) >= j > #(conf > t); every kept pixel reaches its class
    threshold, every dropped one does not, histograms are exact -- including the pixels whose index exceeds 2^31."""
    from mspl_b200 import ops
    dev = torch.device("cuda:0")
    if torch.cuda.get_device_properties(dev).total_memory < 80 * 2 ** 30:
        pytest.skip("needs ~45 GB of device memory")
    n, h, w, K, ignore, portion = 20000, 256, 480, 5, 4, 0.2
    npix = n * h * w
    assert 2 ** 31 < npix < 2 ** 32
    gen = torch.Generator(device=dev).manual_seed(31)
    label = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
    conf = torch.empty((n, h, w), dtype=torch.float32, device=dev)
    step = 2000
    for lo in range(0, n, step):                       # filled in slices: no dataset-sized temporaries
        label[lo:lo + step] = torch.randint(0, K, (step, h, w), device=dev, generator=gen, dtype=torch.uint8)
        conf[lo:lo + step].uniform_(0.0, 1.0, generator=gen)
        conf[lo:lo + step] *= (label[lo:lo + step] != ignore)          # as the fusion kernel writes it
    thresh, kept, final, mask, final_hist = ops.select_and_apply(label, conf, portion, 1, K, ignore, want_mask=True)
    th = thresh.cpu()
    ge, gt, cnt, fh = (torch.zeros(K, dtype=torch.int64, device=dev) for _ in range(4))
    ok = torch.ones((), dtype=torch.bool, device=dev)
    for lo in range(0, n, step):
        lab, cf, fin, msk = label[lo:lo + step].long(), conf[lo:lo + step], final[lo:lo + step].long(), mask[lo:lo + step]
        t = thresh[lab]
        cnt += torch.bincount(lab.reshape(-1), minlength=K)
        fh += torch.bincount(fin.reshape(-1), minlength=K)
        ge += torch.bincount(lab[cf >= t], minlength=K)
        gt += torch.bincount(lab[cf > t], minlength=K)
        keep = (lab != ignore) & (cf >= t)
        ok &= ((fin == torch.where(keep, lab, torch.full_like(lab, ignore))).all()) & ((msk == (fin == ignore)).all())
    assert bool(ok), "a final label / mask differs from label-if-conf>=thresh-else-ignore"
    assert torch.equal(kept, cnt) and torch.equal(final_hist, fh)
    for k in range(K):
        if k == ignore:
            assert th[k] == float("inf")
            continue
        j = int(int(cnt[k]) * portion)
        assert int(ge[k]) >= j > int(gt[k]), (k, j, int(ge[k]), int(gt[k]))
        assert int(fh[k]) == int(ge[k])
    # the tail of the array (pixel indices above 2^31) against the CPU definition
    tail = slice(n - 4, n)
    assert (n - 4) * h * w > 2 ** 31
    f_ref, m_ref = O.apply_thresholds(label[tail].cpu(), conf[tail].cpu(), th, ignore)
    assert torch.equal(final[tail].cpu(), f_ref) and torch.equal(mask[tail].cpu(), m_ref)
