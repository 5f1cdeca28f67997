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
