"""Whole-job label generation on logits that already exist (device tensors or pinned host buffers), single- or
multi-GPU.  This is the post-network body of generate_pseudo_label_multi_model (uest_seg_multi_os.py:891-950)
plus the [NEW] class-balanced thresholding stage, as one object whose steps enqueue without host syncs:

    K1 fuse_sources (+ linear confidence histogram + class histogram)  ->  [all-reduce histograms]  ->  bracket select
    ->  ONE pass over (label, conf): final label map, final class histogram, candidate list  ->  radix select on the
    candidates ([all-reduce] x3)  ->  candidate patch  ->  class weights

Multi-GPU: one process per GPU; target images are sharded by contiguous index range; the only collective is an
all-reduce(SUM) of small int64 histograms, so N-GPU results are bit-identical to 1-GPU results.
"""
from collections import namedtuple

import numpy as np
import torch

from . import ops as _cuda_ops
from .data_loader.segmentation.greenhouse import IGNORE_LABEL

LabelJob = namedtuple("LabelJob", "label final mask conf unc thresh kept class_hist final_hist marginal")


def shard_range(num_items, rank, world_size):
    """Contiguous block partition: rank r owns [r*N/R, (r+1)*N/R)."""
    lo = (num_items * rank) // world_size
    hi = (num_items * (rank + 1)) // world_size
    return lo, hi


def class_weights_from_histogram(class_hist, weighting='normal'):
    """uest_seg_multi_os.py:942-950 on a (K,) histogram (tensor or array) -> float32 CPU tensor."""
    class_array = np.asarray(class_hist.detach().cpu() if isinstance(class_hist, torch.Tensor) else class_hist, dtype=np.float64)
    if weighting == 'normal':
        class_array = class_array / class_array.sum()
        w = 1 / (class_array + 1e-10)
        w[0] = 0.0
    else:
        w = np.ones(len(class_array))
    return torch.from_numpy(w).float()


class LabelGenerator:
    """Fused multi-source pseudo-label generation with optional class-balanced thresholds.

    ops: the compute backend; defaults to the CUDA ops (mspl_b200.ops).  The hook exists so that the host-side
    sharding / all-reduce logic can be exercised on CPU by the test-suite with a stand-in; the product never
    passes anything but the default.
    group: torch.distributed process group (None = default group if initialised, else single process).
    """

    def __init__(self, luts, policy='all', num_classes=5, ignore_label=IGNORE_LABEL, portion=0.2, ds_rate=1,
                 thresholds=True, ops=None, group=None):
        self.luts = list(luts)
        self.policy = policy
        self.num_classes = num_classes
        self.ignore_label = ignore_label
        self.portion = portion
        self.ds_rate = ds_rate
        self.thresholds = thresholds
        self.ops = ops if ops is not None else _cuda_ops
        self.group = group
        self.launches = 0          # kernels of libmspl_b200.so launched so far (bench.py reports this)
        self.k1_events = None      # set to a list to collect (start, end) CUDA events around every K1 launch of run()

    # -- distributed plumbing -----------------------------------------------------------------------------------
    def _world(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.group)
        return 1

    def _all_reduce(self, t):
        if self._world() > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    # -- device-resident job ------------------------------------------------------------------------------------
    def run(self, mains, auxs, want_unc=True, want_mask=False):
        """mains/auxs: this rank's shard, lists of (N_local, C_s, H, W) device tensors.  Returns a LabelJob whose
        histograms / thresholds are GLOBAL (all-reduced) and whose maps are local."""
        ops = self.ops
        if self.k1_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        r = ops.fuse_sources(mains, auxs, self.luts, policy=self.policy, num_classes=self.num_classes,
                             ignore_label=self.ignore_label, ds_rate=self.ds_rate, want_conf=self.thresholds,
                             want_unc=want_unc, want_conf_hist=self.thresholds)
        if self.k1_events is not None:
            e1.record()
            self.k1_events.append((e0, e1))
        self.launches += 1
        class_hist = self._all_reduce(r.class_hist)
        marginal = self._all_reduce(r.marginal) if r.marginal is not None else None
        if not self.thresholds:
            return LabelJob(r.label, r.label, None, r.conf, r.unc, None, None, class_hist, class_hist, marginal)
        thresh, kept, final, mask, final_hist = ops.select_and_apply(
            r.label, r.conf, self.portion, self.ds_rate, self.num_classes, self.ignore_label, conf_hist=r.conf_hist,
            all_reduce=self._all_reduce if self._world() > 1 else None, want_final=True, want_mask=want_mask)
        self.launches += ops.SELECT_AND_APPLY_LAUNCHES
        final_hist = self._all_reduce(final_hist)
        return LabelJob(r.label, final, mask, r.conf, r.unc, thresh, kept, class_hist, final_hist, marginal)

    # -- host-resident job (end-to-end: H2D of every logit, D2H of the label maps) --------------------------------
    def run_from_host(self, mains_host, auxs_host, device, chunk_images=16, out_host=None):
        """mains_host/auxs_host: lists of (N_local, C_s, H, W) fp32 CPU tensors (pinned for full copy speed).
        Streams the logits to the device chunk by chunk on a copy stream while the previous chunk is fused, keeps only
        label (1 B/pix) and conf (4 B/pix) on the device, then thresholds and copies the final uint8 maps back.
        Returns (final label maps as a CPU uint8 tensor, LabelJob with device-side statistics)."""
        ops = self.ops
        dev = torch.device(device)
        S = len(mains_host)
        n, _, h, w = mains_host[0].shape
        K = self.num_classes
        label = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
        conf = torch.empty((n, h, w), dtype=torch.float32, device=dev)
        class_hist = torch.zeros(K, dtype=torch.int64, device=dev)
        conf_hist = torch.zeros((K, ops.RADIX_BINS), dtype=torch.int64, device=dev) if self.thresholds else None
        marginal = torch.zeros((), dtype=torch.int64, device=dev)
        compute = torch.cuda.current_stream(dev)
        copy = torch.cuda.Stream(dev)
        bufs = [[[torch.empty((chunk_images,) + tuple(t.shape[1:]), dtype=torch.float32, device=dev) for t in mains_host]
                 for _ in range(2)] for _ in range(2)]            # [slot][main|aux][source]
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        starts = list(range(0, n, chunk_images))

        def stage(i, slot):
            lo, hi = starts[i], min(starts[i] + chunk_images, n)
            with torch.cuda.stream(copy):
                if i >= 2:
                    copy.wait_event(freed[slot])
                for s in range(S):
                    bufs[slot][0][s][:hi - lo].copy_(mains_host[s][lo:hi], non_blocking=True)
                    bufs[slot][1][s][:hi - lo].copy_(auxs_host[s][lo:hi], non_blocking=True)
                ready[slot].record(copy)

        copy.wait_stream(compute)
        if starts:
            stage(0, 0)
        for i, lo in enumerate(starts):
            slot = i & 1
            hi = min(lo + chunk_images, n)
            if i + 1 < len(starts):
                stage(i + 1, slot ^ 1)
            compute.wait_event(ready[slot])
            ops.fuse_sources([b[:hi - lo] for b in bufs[slot][0]], [b[:hi - lo] for b in bufs[slot][1]], self.luts,
                                 policy=self.policy, num_classes=K, ignore_label=self.ignore_label, ds_rate=self.ds_rate,
                                 want_unc=False, want_conf_hist=self.thresholds, class_hist=class_hist,
                                 conf_hist=conf_hist, marginal=marginal, label_out=label[lo:hi], conf_out=conf[lo:hi])
            self.launches += 1
            freed[slot].record(compute)
        class_hist = self._all_reduce(class_hist)
        if self.thresholds:
            thresh, kept, final, _, final_hist = ops.select_and_apply(
                label, conf, self.portion, self.ds_rate, K, self.ignore_label, conf_hist=conf_hist,
                all_reduce=self._all_reduce if self._world() > 1 else None, want_final=True, want_mask=False)
            self.launches += ops.SELECT_AND_APPLY_LAUNCHES
            final_hist = self._all_reduce(final_hist)
        else:
            thresh = kept = None
            final, final_hist = label, class_hist
        if out_host is None:
            out_host = torch.empty((n, h, w), dtype=torch.uint8, pin_memory=True)
        out_host.copy_(final, non_blocking=True)
        compute.synchronize()          # a host-buffer API: the maps are in `out_host` when the call returns
        return out_host, LabelJob(label, final, None, conf, None, thresh, kept, class_hist, final_hist, marginal)
