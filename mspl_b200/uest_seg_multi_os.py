"""Drop-in replacements for the label-generation functions of the reference's uest_seg_multi_os.py (same names,
argument meaning and return types), routed through the CUDA kernels of libmspl_b200.so.

  get_output                          uest_seg_multi_os.py:669-693   (dups: eval_label.py:49-74, trav_mask_train.py:556-579)
  merge_outputs                       uest_seg_multi_os.py:695-718   (dup: eval_label.py:76-100)
  update_image_list                   uest_seg_multi_os.py:720-728
  generate_pseudo_label               uest_seg_multi_os.py:730-829
  generate_pseudo_label_multi_model   uest_seg_multi_os.py:832-956
  transfer_id_to_greenhouse           uest_seg_multi_os.py:1330-1332
  transfer_output_to_greenhouse       uest_seg_multi_os.py:1334-1350

The source networks stay ordinary PyTorch modules; this module starts where their (main, aux) logits leave the
network.  Differences from the reference, all deliberate:
  * no module-global ``args``: ``merge_outputs`` honours its ``seg_classes`` parameter (as eval_label.py:90 does);
  * the generators batch images through the networks and fuse all sources in ONE kernel pass on the device --
    no per-source softmax/KLD maps ever travel to the host; only the final uint8 label map does;
  * optional class-balanced confidence thresholds ([NEW], ``args.cb_thresholds``; the reference keeps only the
    CBST/CRST flags ``--init-tgt-port`` / ``--ds-rate`` at :216-219);
  * optional fused upsample ([NEW], ``args.fuse_upsample``): the sources' closing bilinear ``F.interpolate`` calls are
    intercepted (mspl_b200/lowres.py) and performed inside the fusion kernel;
  * multi-GPU ([NEW]): when ``torch.distributed`` is initialised (one process per GPU, torchrun) the generators shard the
    target images over the ranks round-robin by loader batch; only the integer histograms (class counts, confidence bins)
    are all-reduced, every rank returns the same class weights, and rank 0 writes ``tgt_train.lst`` in the loader's order.
    ``args.shard_target_images = False`` switches this off for callers that already hand each rank its own loader.
"""
import os
import os.path as osp
import time
from collections import OrderedDict

import numpy as np
import torch

from . import ops
from .label_io import LabelWriter
from .lowres import forward_lowres
from .data_loader.segmentation.greenhouse import (IGNORE_LABEL, SOURCE_TABLES, id_camvid_to_greenhouse,
                                                  id_cityscapes_to_greenhouse, id_forest_to_greenhouse)

__all__ = ["get_output", "merge_outputs", "update_image_list", "generate_pseudo_label",
           "generate_pseudo_label_multi_model", "transfer_id_to_greenhouse", "transfer_output_to_greenhouse",
           "id_camvid_to_greenhouse", "id_cityscapes_to_greenhouse", "id_forest_to_greenhouse"]


def _split_heads(out, model_name='espdnetue'):
    """(main, aux) from a model output: OrderedDict{'out','aux'} (torchvision deeplab) or a tuple (ESPDNetUE)."""
    if isinstance(out, (OrderedDict, dict)):
        return out['out'], out['aux']
    if model_name == 'espdnetue' or isinstance(out, (tuple, list)):
        return out[0], out[1]
    raise ValueError("model output of type %s carries no auxiliary head" % type(out).__name__)


def _cuda_device(device):
    dev = torch.device(device)
    if dev.type != 'cuda':
        raise ValueError("mspl_b200 runs on CUDA devices only (got %r); there is no CPU path" % (device,))
    return dev


def get_output(model, image, model_name='espdnetue', device='cuda'):
    """Forward ``image`` and return ``(softmax(main + 0.5*aux)[0] as ndarray (C,H,W) f32, kld[0] as ndarray (H,W) f32)``
    -- batch element 0 only, like the reference."""
    dev = _cuda_device(device)
    out = model(image.to(dev))
    pred, pred_aux = _split_heads(out, model_name)
    pred = pred.detach()[:1].float().contiguous()
    pred_aux = pred_aux.detach()[:1].float().contiguous()
    prob, kld = ops.softmax_kld(pred, pred_aux)
    return prob[0].cpu().numpy(), kld[0].cpu().numpy()


def merge_outputs(amax_outputs, seg_classes, thresh=None):
    """Per-pixel majority vote over (S,H,W) hard labels; pixels whose winning count is below the vote threshold
    (None/'half' -> S//2+1, 'all' -> S, int <= S -> itself) become class 4.  Returns int64 (H,W) like the reference;
    a CUDA tensor input gives a CUDA int64 tensor back."""
    is_tensor = isinstance(amax_outputs, torch.Tensor)
    lab = amax_outputs if is_tensor else torch.from_numpy(np.ascontiguousarray(amax_outputs))
    if not lab.is_cuda:
        lab = lab.cuda()
    # labels outside [0, seg_classes) match no class in the reference's count loop; 255 stays "no class" as uint8
    lab8 = torch.where((lab >= 0) & (lab < seg_classes), lab, torch.full_like(lab, 255)).to(torch.uint8).contiguous()
    merged = ops.vote_labels(lab8, seg_classes, thresh, IGNORE_LABEL).to(torch.int64)
    return merged if is_tensor else merged.cpu().numpy()


def update_image_list(tgt_train_lst, image_path_list, label_path_list, depth_path_list=None):
    """Write the CSV list "img,label[,depth]" the target dataset reads back (greenhouse.py:172-197)."""
    with open(tgt_train_lst, 'w') as f:
        for idx in range(len(image_path_list)):
            if depth_path_list:
                f.write("%s,%s,%s\n" % (image_path_list[idx], label_path_list[idx], depth_path_list[idx]))
            else:
                f.write("%s,%s\n" % (image_path_list[idx], label_path_list[idx]))
    return


def transfer_id_to_greenhouse(id_to_greenhouse, output_amax_np):
    """Table lookup source class id -> greenhouse class id (works on ndarrays and on tensors of any device)."""
    if isinstance(output_amax_np, torch.Tensor):
        table = torch.as_tensor(np.asarray(id_to_greenhouse), device=output_amax_np.device)
        return table[output_amax_np.long()]
    return np.asarray(id_to_greenhouse)[output_amax_np]


def transfer_output_to_greenhouse(id_to_greenhouse, output_np, seg_classes=5):
    """Probability-level conversion: G[0] = 0 and G[k] = max over source classes mapped to k of P[c] (0 if none);
    (C,H,W) -> (seg_classes,H,W) float64, like the reference."""
    table = np.asarray(id_to_greenhouse)
    is_tensor = isinstance(output_np, torch.Tensor)
    prob = output_np if is_tensor else torch.from_numpy(np.ascontiguousarray(output_np))
    planes = [torch.zeros_like(prob[0], dtype=torch.float64)]
    for k in range(1, seg_classes):
        sel = torch.from_numpy(table == k).to(prob.device)
        planes.append(prob[sel].max(dim=0).values.double() if bool(sel.any()) else torch.zeros_like(planes[0]))
    out = torch.stack(planes)
    return out if is_tensor else out.cpu().numpy()


def _class_weights_from_histogram(class_array, weighting, device):
    """uest_seg_multi_os.py:942-950."""
    class_array = np.asarray(class_array, dtype=np.float64)
    if weighting == 'normal':
        class_array = class_array / class_array.sum()
        class_weights = 1 / (class_array + 1e-10)
        class_weights[0] = 0.0
    else:
        class_weights = np.ones(len(class_array))
    print("class_weights : {}".format(class_weights))
    return torch.from_numpy(class_weights).float().to(device)


def _default_testloader(args):
    """The loader the reference builds inline (uest_seg_multi_os.py:845-849); the dataset class itself stays the
    reference's (file I/O + PIL transforms are outside this package's scope)."""
    if getattr(args, 'dataset', 'greenhouse') != 'greenhouse':
        raise ValueError("only the 'greenhouse' target dataset is wired, as in the reference")
    try:
        from data_loader.segmentation.greenhouse import GreenhouseRGBDSegmentation
    except ImportError as e:
        raise ImportError("pass testloader=... or put the reference's data_loader package on sys.path") from e
    from torch.utils import data
    ds = GreenhouseRGBDSegmentation(list_name=args.data_tgt_train_list, train=False,
                                    use_traversable=getattr(args, 'use_traversable', False),
                                    use_depth=getattr(args, 'use_depth', False))
    return data.DataLoader(ds, batch_size=1, shuffle=False, pin_memory=getattr(args, 'pin_memory', False))


def _image_and_names(batch, use_depth):
    if use_depth:
        image, _label, _depth, name = batch[0], batch[1], batch[2], batch[3]
    else:
        image, _label, name = batch[0], batch[1], batch[2]
    return image, ([name] if isinstance(name, str) else list(name))


def _dist_info(args):
    """(world_size, rank) of the label-generation job: one process per GPU when torch.distributed is initialised (and
    ``args.shard_target_images`` is not switched off), else a single process."""
    import torch.distributed as dist
    if getattr(args, 'shard_target_images', True) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def _all_reduce_sum(t):
    """In-place SUM over the ranks of a small integer tensor (histograms, counters).  NCCL reduces the device tensor on the
    compute stream; any other backend (gloo in the tests) goes through the host."""
    import torch.distributed as dist
    if dist.get_backend() == 'nccl':
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    else:
        c = t.cpu()
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        t.copy_(c)
    return t


def _generate(model_list, luts, device, save_path, round_idx, args, logger, testloader, batch_images, policy):
    dev = _cuda_device(device)
    world, rank = _dist_info(args)
    num_classes = args.classes
    use_depth = getattr(args, 'use_depth', False)
    if policy not in ('half', 'all', 'prob') and not isinstance(policy, int):
        policy = None           # un-typed CLI strings fall through to 'half', as at uest_seg_multi_os.py:702-705
    use_cb = bool(getattr(args, 'cb_thresholds', False))
    portion = float(getattr(args, 'init_tgt_port', 0.2))
    ds_rate = int(getattr(args, 'ds_rate', 1)) if use_cb else 1

    save_pred_path = osp.join(save_path, 'pred')
    os.makedirs(save_pred_path, exist_ok=True)
    tgt_train_lst = osp.join(save_path, 'tgt_train.lst')
    if testloader is None:
        testloader = _default_testloader(args)
    for m in model_list:
        m.train() if getattr(args, 'eval_training', False) else m.eval()
        m.to(dev)
    # The reference forwards ONE image at a time (batch_size=1, :888-897).  With modules in training mode (--eval-training)
    # batch norm normalises over the batch and updates its running statistics per forward, so batching the network would
    # change the labels: forward image by image then and batch only the fusion kernel.
    per_image_forward = any(mod.training for m in model_list for mod in m.modules())

    if logger is not None:
        logger.info('###### Start evaluating target domain train set in round {}! ######'.format(round_idx))
    start_eval = time.time()
    image_path_list, label_path_list, depth_path_list, names = [], [], [], []
    entry_order = []        # global position (loader batch index, index inside the batch) of every entry of the lists above
    class_hist = torch.zeros(num_classes, dtype=torch.int64, device=dev)
    conf_hist = torch.zeros((num_classes, ops.RADIX_BINS), dtype=torch.int64, device=dev) if use_cb else None
    count_ties = bool(getattr(args, 'count_near_ties', False))     # needs the softmax: off keeps the labels-only kernel
    marginal = torch.zeros((), dtype=torch.int64, device=dev)
    kept_labels, kept_confs = [], []

    writer = LabelWriter(dev, workers=int(getattr(args, 'label_writer_threads', 8)))

    def save_maps(label_u8, batch_names, batch_order):
        # PNG encode + file write happen on the writer's threads while the GPU works on the next batch
        out_paths = []
        entry_order.extend(batch_order)
        for path_name in batch_names:
            image_name = path_name.split('/')[-1].rsplit('.', 1)[0]
            out_paths.append('%s/%s.png' % (save_pred_path, image_name))
            image_path_list.append(path_name)
            label_path_list.append(out_paths[-1])
            if use_depth:
                depth_path_list.append(path_name.replace('color', 'depth'))
        writer.submit(label_u8, out_paths)

    # (a forward_lowres miss runs the network a second time, which in training mode would update the statistics twice)
    fuse_upsample = bool(getattr(args, 'fuse_upsample', False)) and not per_image_forward

    def forward_heads(m, x):
        if not per_image_forward or x.shape[0] == 1:
            return _split_heads(m(x))
        outs = [_split_heads(m(x[i:i + 1])) for i in range(x.shape[0])]
        return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])

    def table_for(s, num_src_classes):
        if luts[s] is not None:
            return luts[s]
        # a source without a table: the reference leaves its ids unconverted (:907-912) and merge_outputs never counts an id
        # >= args.classes (:708-711); an identity table covers the ids that can vote
        if num_src_classes > num_classes:
            raise ValueError("source %d has no label table (os_data name unknown) and %d > %d classes: the reference leaves such "
                             "ids unconverted, where those >= %d never win a vote; pass a table for it"
                             % (s, num_src_classes, num_classes, num_classes))
        return np.arange(num_src_classes)

    def flush(images, batch_names, batch_order):
        x = torch.cat(images).to(dev, non_blocking=True)      # (one image size per call, like the reference's fixed crop)
        kw = dict(policy=policy, num_classes=num_classes, ignore_label=IGNORE_LABEL, ds_rate=ds_rate, want_conf=use_cb,
                  want_unc=False, want_conf_hist=use_cb, class_hist=None if use_cb else class_hist, conf_hist=conf_hist,
                  count_marginal=count_ties, marginal=marginal if count_ties else None)
        r = None
        if fuse_upsample:
            # [NEW] take every source's logits before its closing bilinear upsample and interpolate inside the kernel
            heads = [forward_lowres(m, x) for m in model_list]
            if all(h is not None for h in heads):
                try:
                    r = ops.fuse_sources_lowres([h[0].float().contiguous() for h in heads],
                                                [h[1].float().contiguous() for h in heads],
                                                [table_for(s, h[0].shape[1]) for s, h in enumerate(heads)], x.shape[-2:], **kw)
                except NotImplementedError:
                    r = None        # geometry the fused kernel does not cover: fall through to the full-resolution path
        if r is None:
            mains, auxs = [], []
            for m in model_list:
                pred, pred_aux = forward_heads(m, x)
                mains.append(pred.float().contiguous())
                auxs.append(pred_aux.float().contiguous())
            r = ops.fuse_sources(mains, auxs, [table_for(s, t.shape[1]) for s, t in enumerate(mains)], **kw)
        if use_cb:      # labels wait on the device until the dataset-wide thresholds are known
            kept_labels.append(r.label), kept_confs.append(r.conf), names.append((batch_names, batch_order))
        else:
            save_maps(r.label, batch_names, batch_order)

    with torch.no_grad():
        pending, pending_names, pending_order = [], [], []
        for batch_idx, batch in enumerate(testloader):
            if batch_idx % world != rank:       # target images shard over the ranks, round-robin by loader batch
                continue
            image, batch_names = _image_and_names(batch, use_depth)
            pending.append(image), pending_names.extend(batch_names)
            pending_order.extend((batch_idx, j) for j in range(len(batch_names)))
            if len(pending_names) >= batch_images:
                flush(pending, pending_names, pending_order)
                pending, pending_names, pending_order = [], [], []
        if pending:
            flush(pending, pending_names, pending_order)
        if use_cb and (kept_labels or world > 1):
            # conf_hist already holds the linear confidence histogram of every batch; one more pass over the kept maps
            # settles every pixel outside its class's bracket and the few inside it are resolved from a candidate list.
            # With several ranks only the (K, 2048) histograms travel; a rank without images still takes part in them.
            if kept_labels:
                label_all, conf_all = torch.cat(kept_labels), torch.cat(kept_confs)
            else:
                label_all = torch.empty((0, 1, 1), dtype=torch.uint8, device=dev)
                conf_all = torch.empty((0, 1, 1), dtype=torch.float32, device=dev)
            _, _, final, _, _ = ops.select_and_apply(label_all, conf_all, portion, ds_rate, num_classes, IGNORE_LABEL,
                                                     conf_hist=conf_hist, all_reduce=_all_reduce_sum if world > 1 else None,
                                                     want_final=True, final_hist=class_hist)
            pos = 0
            for batch_names, batch_order in names:
                save_maps(final[pos:pos + len(batch_names)], batch_names, batch_order)
                pos += len(batch_names)
        if world > 1:       # the dataset-wide class histogram (-> class weights) and the near-tie count
            if not use_cb:      # (select_and_apply already returned the GLOBAL final histogram)
                _all_reduce_sum(class_hist)
            if count_ties:
                _all_reduce_sum(marginal)

    writer.close()       # every label map is on disk before the list that points at it is written
    if world > 1:
        # every rank hands over its share of the list; rank 0 writes it in the loader's order and all ranks wait for the file
        import torch.distributed as dist
        shares = [None] * world
        dist.all_gather_object(shares, (entry_order, image_path_list, label_path_list, depth_path_list))
        if rank == 0:
            rows = sorted((o, i, sh[1][i], sh[2][i], (sh[3][i] if sh[3] else None)) for sh in shares for i, o in enumerate(sh[0]))
            update_image_list(tgt_train_lst, [r[2] for r in rows], [r[3] for r in rows], [r[4] for r in rows] if use_depth else [])
        dist.barrier()
    else:
        update_image_list(tgt_train_lst, image_path_list, label_path_list, depth_path_list)
    class_weights = _class_weights_from_histogram(class_hist.cpu().numpy(), getattr(args, 'class_weighting', 'normal'), dev)
    if logger is not None:
        ties = ' ({} near-tie pixels)'.format(int(marginal.item())) if count_ties else ''
        logger.info('###### Finish evaluating target domain train set in round {}! Time cost: {:.2f} seconds.{} ######'.format(
            round_idx, time.time() - start_eval, ties))
    return tgt_train_lst, class_weights


def generate_pseudo_label(model, device, save_path, round_idx, tgt_num, label_2_id, valid_labels, args, logger,
                          class_encoding, writer, testloader=None, batch_images=32):
    """Single-model regeneration (--label-update rounds): argmax of the target model's own 5-class output, no table,
    no vote.  Same signature and return value as the reference: (path of tgt_train.lst, class_weights f32 tensor)."""
    identity = [np.arange(args.classes)]
    return _generate([model], identity, device, save_path, round_idx, args, logger, testloader, batch_images, policy=1)


def generate_pseudo_label_multi_model(model_list, os_data_list, device, save_path, round_idx, tgt_num, label_2_id,
                                      valid_labels, args, logger, class_encoding, writer, testloader=None, batch_images=32):
    """Multi-source pseudo-label generation: every source model labels every target image, labels are converted to
    greenhouse classes and merged by vote (``args.merge_label_policy``).  Writes ``<save_path>/pred/<name>.png`` and
    ``<save_path>/tgt_train.lst``; returns (list path, class_weights) like the reference."""
    luts = []
    for os_data in os_data_list:
        if os_data not in SOURCE_TABLES:     # the reference leaves such a source's ids unconverted (:907-912): see table_for
            luts.append(None)
        else:
            luts.append(SOURCE_TABLES[os_data])
    return _generate(list(model_list), luts, device, save_path, round_idx, args, logger, testloader, batch_images,
                     policy=getattr(args, 'merge_label_policy', None))
