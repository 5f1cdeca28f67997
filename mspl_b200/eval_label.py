"""Drop-in for the evaluation loop of the reference's eval_label.py (main, :156-211): how good are the pseudo-labels the source
models agree on?  The reference runs, per validation image, get_output for every source (two blocking device->host copies
each, :49-74), np.argmax on the host, the source->greenhouse tables, merge_outputs(thresh='all') (:76-100) and MIOU.get_iou
(torch.histc x3 on the CPU).  Here a batch of images goes through the sources, ONE fusion launch (the labels-only kernel: the
vote needs no softmax) and ONE counting launch; the per-class intersection / union areas accumulate on the device and are read
once at the end.

    iou, miou = evaluate_pseudo_labels(os_model_list, os_data_name_list, val_loader, device)

`iou` is the reference's ``inter_meter.sum / (union_meter.sum + 1e-10) * 100`` (float64 array of K entries) and `miou` its
``iou[[1, 2, 3]].mean()`` (:205-206)."""
import numpy as np
import torch

from . import ops
from .data_loader.segmentation.greenhouse import IGNORE_LABEL, SOURCE_TABLES
from .uest_seg_multi_os import _split_heads


def evaluate_pseudo_labels(model_list, os_data_list, val_loader, device='cuda', seg_classes=5, thresh='all', batch_images=8,
                           use_depth=False, return_counts=False):
    """model_list / os_data_list as in eval_label.main (:120-140); val_loader yields the reference's batches
    (image, label, name, ...) with any batch size.  Returns (iou %, miou %) [, int64 (3, K) area counts]."""
    dev = torch.device(device)
    luts = [SOURCE_TABLES[name] for name in os_data_list]
    for m in model_list:
        m.eval()
        m.to(dev)      # as the reference's get_output does (:57)
    counts = torch.zeros((3, seg_classes), dtype=torch.int64, device=dev)
    n_calls = 0        # every get_iou call of the reference adds its epsilon to the union once: once per LOADER batch (:201)
    pend_x, pend_t = [], []

    def flush():
        x = torch.cat(pend_x).to(dev, non_blocking=True)
        target = torch.cat(pend_t).to(dev, non_blocking=True)
        mains, auxs = [], []
        for m in model_list:
            pm, pa = _split_heads(m(x))
            mains.append(pm.float().contiguous()), auxs.append(pa.float().contiguous())
        r = ops.fuse_sources(mains, auxs, luts, policy=thresh, num_classes=seg_classes, ignore_label=IGNORE_LABEL,
                             want_conf=False, want_unc=False, want_conf_hist=False, count_marginal=False)
        ops.miou_counts(r.label, target.reshape(r.label.shape), seg_classes, counts=counts)

    with torch.no_grad():
        for batch in val_loader:
            image, target = batch[0], batch[1]
            if use_depth:
                raise NotImplementedError("depth inputs are not wired into the batched evaluation")
            pend_x.append(image), pend_t.append(target)
            n_calls += 1
            if sum(t.shape[0] for t in pend_x) >= batch_images:
                flush()
                pend_x, pend_t = [], []
        if pend_x:
            flush()
    c = counts.cpu().numpy().astype(np.float64)
    inter = c[0]
    # MIOU.get_iou: union = pred + mask - inter + 1e-6 (float32) per call; the meters sum the per-call values
    union = c[1] + c[2] - c[0] + n_calls * np.float64(np.float32(1e-6))
    iou = inter / (union + 1e-10) * 100
    miou = iou[[1, 2, 3]].mean()
    return (iou, miou, counts) if return_counts else (iou, miou)
