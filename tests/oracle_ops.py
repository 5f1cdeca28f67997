"""CPU stand-in for mspl_b200.ops built on the oracle -- lets the host-side sharding / all-reduce logic of
mspl_b200.pipeline.LabelGenerator run under gloo without a GPU.  Test infrastructure only."""
import numpy as np
import torch

from oracle import mspl_oracle as O
from mspl_b200.ops import FuseResult, vote_threshold  # noqa: F401  (pure-Python helpers, no CUDA needed)

RADIX_BINS = 2048


def _keys(conf):
    b = conf.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    neg = (b & 0x80000000) != 0
    return torch.where(neg, (~b) & 0xFFFFFFFF, b | 0x80000000)


def _digit(keys, p):
    return (keys >> 21) if p == 0 else ((keys >> 10) & 0x7FF) if p == 1 else (keys & 0x3FF)


def _prefix(keys, p):
    return torch.zeros_like(keys) if p == 0 else (keys >> 21) if p == 1 else (keys >> 10)


def _conf_bin(conf):
    """The linear confidence bin of the bracketed protocol (include/mspl_b200.h): clamp(floor(conf * 2048), 0, 2047)."""
    return torch.clamp(torch.floor(conf.double() * RADIX_BINS), 0, RADIX_BINS - 1).long()


def fuse_sources(mains, auxs, luts, policy='half', num_classes=5, ignore_label=4, ds_rate=1, want_conf=True, want_unc=True,
                 want_kld=False, want_conf_hist=True, count_marginal=True, class_hist=None, conf_hist=None, marginal=None,
                 label_out=None, conf_out=None, unc_out=None):
    r = O.fuse_sources(mains, auxs, luts, policy, num_classes, ignore_label)
    h, w = r["label"].shape[-2:]
    ch = r["class_hist"].clone() if class_hist is None else class_hist.add_(r["class_hist"])
    hist = None
    if want_conf_hist:
        keep = (torch.arange(h * w) % ds_rate == 0).reshape(h, w).expand(r["label"].shape)
        idx = r["label"].long()[keep] * RADIX_BINS + _conf_bin(r["conf"])[keep]
        hist = torch.bincount(idx, minlength=num_classes * RADIX_BINS).reshape(num_classes, RADIX_BINS)
        if conf_hist is not None:
            hist = conf_hist.add_(hist)
    marg = r["marginal"].sum()
    if marginal is not None:
        marg = marginal.add_(marg)
    label, cf, unc = r["label"], r["conf"], r["unc"]
    if label_out is not None:
        label = label_out.copy_(label)
    if conf_out is not None:
        cf = conf_out.copy_(cf)
    if unc_out is not None:
        unc = unc_out.copy_(unc)
    return FuseResult(label, cf, unc, r["kld"] if want_kld else None, ch, hist, marg)


SELECT_AND_APPLY_LAUNCHES = 3
SELECT_AND_APPLY_LAUNCHES_SHARDED = 9


def new_label_stats(num_classes, device):
    K = num_classes
    buf = torch.zeros(K * RADIX_BINS + K + 1, dtype=torch.int64, device=device)
    return buf, buf[:K * RADIX_BINS].view(K, RADIX_BINS), buf[K * RADIX_BINS:K * RADIX_BINS + K], buf[K * RADIX_BINS + K:].view(())


def select_and_apply(label, conf, portion=0.2, ds_rate=1, num_classes=5, ignore_label=4, conf_hist=None, all_reduce=None,
                     want_final=True, want_mask=False, final_hist=None, hist_reduced=False):
    """The bracketed protocol written with torch ops (independent of the CUDA implementation): linear histogram ->
    [all-reduce] -> bracket -> candidates -> 3 radix passes over the candidates ([all-reduce] each) -> thresholds.
    Like the CUDA op, the returned final histogram is the GLOBAL one when an all_reduce is given."""
    h, w = label.shape[-2:]
    K = num_classes
    keep = (torch.arange(h * w) % ds_rate == 0).reshape(h, w).expand(label.shape)
    lab_all, bins_all = label.long(), _conf_bin(conf)
    if conf_hist is None:
        conf_hist = torch.bincount(lab_all[keep] * RADIX_BINS + bins_all[keep], minlength=K * RADIX_BINS).reshape(K, RADIX_BINS)
    hist = conf_hist
    if all_reduce is not None and not hist_reduced:
        all_reduce(hist)
    kept = hist.sum(1)
    rank = torch.zeros(K, dtype=torch.int64)
    sel_bin = torch.full((K,), -1, dtype=torch.int64)
    done = torch.zeros(K, dtype=torch.bool)
    thresh = torch.ones(K, dtype=torch.float32)
    for k in range(K):
        j = int(int(kept[k]) * float(portion))
        if j == 0 or k == ignore_label:
            done[k] = True
            if k == ignore_label:
                thresh[k] = float('inf')
            continue
        above = torch.flip(torch.cumsum(torch.flip(hist[k], [0]), 0), [0])      # above[d] = sum_{i >= d}
        b = int((above >= j).nonzero().max())
        sel_bin[k] = b
        rank[k] = j - (int(above[b]) - int(hist[k][b]))
    hist.zero_()
    cand = (bins_all == sel_bin[lab_all]) & keep            # the participating pixels inside their class's bracket
    keys, lab = _keys(conf)[cand], lab_all[cand]
    prefix = torch.zeros(K, dtype=torch.int64)
    for p in range(3):
        sel = (_prefix(keys, p) == prefix[lab]) & ~done[lab]
        hist = torch.bincount(lab[sel] * RADIX_BINS + _digit(keys, p)[sel], minlength=K * RADIX_BINS).reshape(K, RADIX_BINS)
        if all_reduce is not None:
            all_reduce(hist)
        for k in range(K):
            if done[k]:
                continue
            above = torch.flip(torch.cumsum(torch.flip(hist[k], [0]), 0), [0])
            d = int((above >= rank[k]).nonzero().max())
            rank[k] -= int(above[d]) - int(hist[k][d])
            prefix[k] = (prefix[k] << (10 if p == 2 else 11)) | d
            if p == 2:
                key = int(prefix[k])
                bits = (key & 0x7FFFFFFF) if key & 0x80000000 else (~key) & 0xFFFFFFFF
                thresh[k] = torch.tensor([bits if bits < 2 ** 31 else bits - 2 ** 32], dtype=torch.int64).to(torch.int32) \
                    .view(torch.float32)[0]
    final = mask = hist_out = None
    if want_final or want_mask or final_hist is not None:
        final, mask = O.apply_thresholds(label, conf, thresh, ignore_label)
        hist_out = torch.bincount(final.reshape(-1).long(), minlength=K)
        if all_reduce is not None:
            all_reduce(hist_out)
        if final_hist is not None:
            hist_out = final_hist.add_(hist_out)
    return thresh, kept, (final if want_final else None), (mask if want_mask else None), hist_out


def cb_thresholds(label, conf, portion=0.2, ds_rate=1, num_classes=5, conf_hist=None, all_reduce=None, ignore_label=None):
    thresh, kept, _, _, _ = select_and_apply(label, conf, portion, ds_rate, num_classes, ignore_label, conf_hist, all_reduce,
                                             want_final=False)
    return thresh, kept


def apply_thresholds(label, conf, thresh, ignore_label=4, want_mask=True, final_hist=None):
    final, mask = O.apply_thresholds(label, conf, thresh, ignore_label)
    hist = torch.bincount(final.reshape(-1).long(), minlength=thresh.numel())
    if final_hist is not None:
        hist = final_hist.add_(hist)
    return final, (mask if want_mask else None), hist
