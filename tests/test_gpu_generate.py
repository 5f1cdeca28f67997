"""-m gpu: the reference-named generators (generate_pseudo_label[_multi_model]) end to end with stand-in source models
and an in-memory loader: PNG label maps, tgt_train.lst and class weights against the oracle's restatement of the
reference loop (uest_seg_multi_os.py:888-950)."""
import os
import types

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import mspl_oracle as O

pytestmark = pytest.mark.gpu

SOURCES = (("camvid", 13), ("cityscapes", 20), ("forest", 5))
N, H, W = 5, 32, 48


class TinySource(torch.nn.Module):
    """A deterministic 'segmentation network' with a main and an auxiliary head (tuple output, like ESPDNetUE)."""

    def __init__(self, classes, seed, as_dict=False):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.main = torch.nn.Conv2d(3, classes, 3, padding=1)
        self.aux = torch.nn.Conv2d(3, classes, 3, padding=1)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * 2.0)
        self.as_dict = as_dict

    def forward(self, x):
        m, a = self.main(x), self.aux(x)
        return {"out": m, "aux": a} if self.as_dict else (m, a)


def _loader():
    g = torch.Generator().manual_seed(21)
    for i in range(N):
        image = torch.randn(1, 3, H, W, generator=g)
        yield image, torch.zeros(1, H, W, dtype=torch.long), ["/tmp/dataset/greenhouse/color/img_%03d.png" % i], torch.zeros(1)


def _args(**kw):
    a = types.SimpleNamespace(classes=5, use_depth=False, eval_training=False, class_weighting='normal',
                              merge_label_policy='half', dataset='greenhouse')
    a.__dict__.update(kw)
    return a


def _oracle_run(models, names, policy):
    mains, auxs = [], []
    images = torch.cat([b[0] for b in _loader()])
    with torch.no_grad():
        for m in models:
            out = m.cpu()(images)
            pm, pa = (out["out"], out["aux"]) if isinstance(out, dict) else out
            mains.append(pm.contiguous()), auxs.append(pa.contiguous())
    luts = [O.LUTS[n] for n in names]
    labels, class_array = O.multi_source_labels(mains, auxs, luts, policy)
    marg = O.fuse_sources(mains, auxs, luts, policy)["marginal"].numpy()
    return labels, class_array, marg


@pytest.mark.parametrize("policy", ["half", "all", 2])
def test_generate_pseudo_label_multi_model(tmp_path, policy):
    from mspl_b200 import uest_seg_multi_os as U
    models = [TinySource(c, 5 + i, as_dict=(i == 1)) for i, (_, c) in enumerate(SOURCES)]
    names = [n for n, _ in SOURCES]
    want, class_array, marg = _oracle_run(models, names, policy)
    lst, cw = U.generate_pseudo_label_multi_model(models, names, 'cuda:0', str(tmp_path), 0, N, None, None,
                                                  _args(merge_label_policy=policy), None, None, None, testloader=_loader(),
                                                  batch_images=2)
    assert lst == os.path.join(str(tmp_path), 'tgt_train.lst')
    rows = [ln.strip().split(',') for ln in open(lst)]
    assert len(rows) == N
    n_diff = 0
    for i, (img_path, lab_path) in enumerate(rows):
        assert img_path == "/tmp/dataset/greenhouse/color/img_%03d.png" % i
        assert lab_path == os.path.join(str(tmp_path), 'pred', 'img_%03d.png' % i)
        got = np.array(Image.open(lab_path))
        assert got.dtype == np.uint8 and got.shape == (H, W)
        diff = got != want[i]
        assert not (diff & ~marg[i]).any()
        n_diff += int(diff.sum())
    assert cw.is_cuda and cw.dtype == torch.float32 and cw.shape == (5,)
    if n_diff == 0:
        torch.testing.assert_close(cw.cpu(), O.class_weights_from_histogram(class_array, 'normal'))
    assert cw[0].item() == 0.0


def test_generate_pseudo_label_single_model(tmp_path):
    """--label-update rounds: the target model's own 5-class argmax, no table, no vote (uest_seg_multi_os.py:730-829)."""
    from mspl_b200 import uest_seg_multi_os as U
    model = TinySource(5, 77)
    images = torch.cat([b[0] for b in _loader()])
    with torch.no_grad():
        pm, pa = model(images)
    lst, cw = U.generate_pseudo_label(model, 'cuda:0', str(tmp_path), 1, N, None, None, _args(class_weighting='flat'), None,
                                      None, None, testloader=_loader(), batch_images=3)
    P = torch.softmax(pm + 0.5 * pa, dim=1)
    want = P.argmax(dim=1).numpy().astype(np.uint8)
    top2 = torch.topk(P, 2, dim=1).values
    marg = ((top2[:, 0] - top2[:, 1]) < 1e-6).numpy()
    for i, ln in enumerate(open(lst)):
        got = np.array(Image.open(ln.strip().split(',')[1]))
        assert not ((got != want[i]) & ~marg[i]).any()
    assert torch.equal(cw.cpu(), torch.ones(5))


def test_generate_with_class_balanced_thresholds(tmp_path):
    """[NEW] stage switched on: labels below their class threshold become the ignore class; kept fraction ~ portion."""
    from mspl_b200 import uest_seg_multi_os as U
    models = [TinySource(c, 5 + i) for i, (_, c) in enumerate(SOURCES)]
    names = [n for n, _ in SOURCES]
    base, _, _ = _oracle_run(models, names, 'half')
    lst, _ = U.generate_pseudo_label_multi_model(models, names, 'cuda:0', str(tmp_path), 0, N, None, None,
                                                 _args(cb_thresholds=True, init_tgt_port=0.3, ds_rate=1), None, None, None,
                                                 testloader=_loader(), batch_images=2)
    got = np.stack([np.array(Image.open(ln.strip().split(',')[1])) for ln in open(lst)])
    changed = got != base
    assert (got[changed] == 4).all()
    for k in (1, 2, 3):
        nk = int((base == k).sum())
        if nk >= 50:
            assert abs(int((got == k).sum()) - int(nk * 0.3)) <= max(3, nk // 50)


class UpsamplingSource(torch.nn.Module):
    """Main head at 1/2, aux head at 1/4 resolution, closed by the two bilinear align_corners=True upsamples of ESPDNetUE
    (model/segmentation/espdnet_ue.py:301-302), plus an unrelated internal interpolate that must NOT be intercepted."""

    def __init__(self, classes, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.main = torch.nn.Conv2d(3, classes, 3, padding=1)
        self.aux = torch.nn.Conv2d(3, classes, 3, padding=1)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * 2.0)

    def forward(self, x):
        import torch.nn.functional as F
        size = x.shape[-2:]
        half = F.interpolate(x, scale_factor=0.5, mode='bilinear', align_corners=True)
        quarter = F.interpolate(half, scale_factor=0.5, mode='bilinear', align_corners=True)
        return (F.interpolate(self.main(half), size=size, mode='bilinear', align_corners=True),
                F.interpolate(self.aux(quarter), size=size, mode='bilinear', align_corners=True))


def test_generate_with_fused_upsample(tmp_path):
    from mspl_b200 import uest_seg_multi_os as U
    from mspl_b200.lowres import forward_lowres
    models = [UpsamplingSource(c, 9 + i) for i, (_, c) in enumerate(SOURCES)]
    names = [n for n, _ in SOURCES]
    x = torch.randn(2, 3, H, W)
    heads = forward_lowres(models[0], x)
    assert heads is not None and heads[0].shape[-2:] == (H // 2, W // 2) and heads[1].shape[-2:] == (H // 4, W // 4)
    assert forward_lowres(TinySource(5, 1), x) is None                      # no closing upsample -> ordinary path
    import torch.nn.functional as F
    assert F.interpolate.__module__ == "torch.nn.functional"               # the wrapper is gone after the call
    plain, fused = tmp_path / "plain", tmp_path / "fused"
    lst_a, cw_a = U.generate_pseudo_label_multi_model(models, names, 'cuda:0', str(plain), 0, N, None, None, _args(), None, None,
                                                      None, testloader=_loader(), batch_images=2)
    lst_b, cw_b = U.generate_pseudo_label_multi_model(models, names, 'cuda:0', str(fused), 0, N, None, None,
                                                      _args(fuse_upsample=True), None, None, None, testloader=_loader(), batch_images=2)
    want, _, _ = _oracle_run(models, names, 'half')
    saved = O.NEAR_TIE_MARGIN
    O.NEAR_TIE_MARGIN = 1e-5
    try:
        _, _, marg = _oracle_run(models, names, 'half')
    finally:
        O.NEAR_TIE_MARGIN = saved
    for i, (la, lb) in enumerate(zip(open(lst_a), open(lst_b))):
        a = np.array(Image.open(la.strip().split(',')[1]))
        b = np.array(Image.open(lb.strip().split(',')[1]))
        assert not ((a != want[i]) & ~marg[i]).any()
        assert not ((b != want[i]) & ~marg[i]).any()


@pytest.mark.parametrize("batch_images", [1, 2, 8])
def test_evaluate_pseudo_labels_matches_reference_loop(batch_images):
    """eval_label.main's loop (:156-211): per-image get_output -> argmax -> table -> merge_outputs('all') -> MIOU.get_iou,
    restated with the oracle, against the batched GPU evaluation (one fusion + one counting launch per batch)."""
    from mspl_b200.eval_label import evaluate_pseudo_labels
    models = [TinySource(c, 5 + i, as_dict=(i == 1)) for i, (_, c) in enumerate(SOURCES)]
    names = [n for n, _ in SOURCES]
    g = torch.Generator().manual_seed(8)
    targets = torch.randint(0, 5, (N, 1, H, W), generator=g)
    targets[targets == 0] = 255                                  # unlabelled pixels of the greenhouse validation set

    def loader():
        for i, b in enumerate(_loader()):
            yield b[0], targets[i], b[2]

    labels, _, marg = _oracle_run(models, names, 'all')
    inter_sum, union_sum = np.zeros(5, np.float32), np.zeros(5, np.float32)
    for i in range(N):                                            # AverageMeter.sum of the per-image float32 arrays (:195-197)
        inter, union = O.miou_get_iou(torch.from_numpy(labels[i].astype(np.int64)), targets[i, 0], num_classes=5)
        inter_sum, union_sum = inter_sum + inter, union_sum + union
    want = inter_sum / (union_sum + 1e-10) * 100
    iou, miou, counts = evaluate_pseudo_labels([m.to('cuda:0') for m in models], names, loader(), 'cuda:0', batch_images=batch_images,
                                               return_counts=True)
    assert int(marg.sum()) == 0, "fixture has near-tie pixels; pick another seed"
    np.testing.assert_allclose(iou, want, rtol=1e-5, atol=1e-6)
    assert abs(miou - want[[1, 2, 3]].mean()) <= 1e-5 * max(1.0, abs(miou))
    assert counts.dtype == torch.int64 and counts.shape == (3, 5)


def _dist_worker(rank, world, port, save_path, use_cb, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mspl_b200 import uest_seg_multi_os as U
        models = [TinySource(c, 5 + i, as_dict=(i == 1)) for i, (_, c) in enumerate(SOURCES)]
        names = [n for n, _ in SOURCES]
        lst, cw = U.generate_pseudo_label_multi_model(models, names, 'cuda:0', save_path, 0, N, None, None,
                                                      _args(merge_label_policy='half', cb_thresholds=use_cb, init_tgt_port=0.3),
                                                      None, None, None, testloader=_loader(), batch_images=2)
        out[rank] = (lst, cw.cpu())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("use_cb", [False, True])
def test_generators_shard_over_ranks(tmp_path, use_cb):
    """Two processes (gloo rendezvous, both on cuda:0) share the target images round-robin: label maps, tgt_train.lst and class
    weights equal the single-process run; only integer histograms cross ranks."""
    import socket
    import torch.multiprocessing as mp
    from mspl_b200 import uest_seg_multi_os as U
    models = [TinySource(c, 5 + i, as_dict=(i == 1)) for i, (_, c) in enumerate(SOURCES)]
    names = [n for n, _ in SOURCES]
    single_dir, multi_dir = str(tmp_path / "single"), str(tmp_path / "multi")
    lst1, cw1 = U.generate_pseudo_label_multi_model(models, names, 'cuda:0', single_dir, 0, N, None, None,
                                                    _args(merge_label_policy='half', cb_thresholds=use_cb, init_tgt_port=0.3),
                                                    None, None, None, testloader=_loader(), batch_images=2)
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(_dist_worker, args=(2, port, multi_dir, use_cb, out), nprocs=2, join=True)
    rows1 = [ln.strip().split(',') for ln in open(lst1)]
    rows2 = [ln.strip().split(',') for ln in open(out[0][0])]
    assert out[0][0] == out[1][0] == os.path.join(multi_dir, 'tgt_train.lst')
    assert [r[0] for r in rows1] == [r[0] for r in rows2] and len(rows2) == N
    for (_, p1), (_, p2) in zip(rows1, rows2):
        assert os.path.basename(p1) == os.path.basename(p2)
        assert np.array_equal(np.array(Image.open(p1)), np.array(Image.open(p2)))
    for r in (0, 1):
        assert torch.equal(out[r][1], cw1.cpu())


class BatchNormSource(torch.nn.Module):
    """Like TinySource with a BatchNorm in front of the heads: in training mode its output depends on what else is in the batch."""

    def __init__(self, classes, seed):
        super().__init__()
        self.bn = torch.nn.BatchNorm2d(3)
        self.heads = TinySource(classes, seed)

    def forward(self, x):
        return self.heads(self.bn(x))


def test_eval_training_forwards_image_by_image(tmp_path):
    """--eval-training leaves the sources in train() mode (uest_seg_multi_os.py:873-878) and the reference forwards one image
    at a time: batch norm then normalises each image by its own statistics.  The generator must not batch the network there
    (it still batches the fusion kernel) -- labels equal the per-image train-mode forward, not the batched one."""
    from mspl_b200 import uest_seg_multi_os as U
    names = [n for n, _ in SOURCES]

    def fresh():
        return [BatchNormSource(c, 5 + i) for i, (_, c) in enumerate(SOURCES)]

    images = [b[0] for b in _loader()]
    mains, auxs = [], []
    for m in fresh():
        m.train()
        with torch.no_grad():
            outs = [m(x) for x in images]                         # one image per forward, as the reference
        mains.append(torch.cat([o[0] for o in outs]).contiguous()), auxs.append(torch.cat([o[1] for o in outs]).contiguous())
    luts = [O.LUTS[n] for n in names]
    want, _ = O.multi_source_labels(mains, auxs, luts, 'half')
    marg = O.fuse_sources(mains, auxs, luts, 'half')["marginal"].numpy()
    lst, _ = U.generate_pseudo_label_multi_model(fresh(), names, 'cuda:0', str(tmp_path), 0, N, None, None,
                                                 _args(merge_label_policy='half', eval_training=True), None, None, None,
                                                 testloader=_loader(), batch_images=4)
    n_batched_differs = 0
    for i, ln in enumerate(open(lst)):
        got = np.array(Image.open(ln.strip().split(',')[1]))
        assert not ((got != want[i]) & ~marg[i]).any()
    # the fixture is meaningful: a batched train-mode forward would have given different labels somewhere
    bm, ba = [], []
    for m in fresh():
        m.train()
        with torch.no_grad():
            o = m(torch.cat(images))
        bm.append(o[0].contiguous()), ba.append(o[1].contiguous())
    batched, _ = O.multi_source_labels(bm, ba, luts, 'half')
    n_batched_differs = int((np.asarray(batched) != np.asarray(want)).sum())
    assert n_batched_differs > 0


def test_source_without_table_votes_with_its_own_ids(tmp_path):
    """An os_data name with no table: the reference leaves that source's ids unconverted (:907-912).  With no more classes than
    the target it votes with its own ids; with more the generator refuses with an error that names the source."""
    from mspl_b200 import uest_seg_multi_os as U
    models = [TinySource(13, 5), TinySource(5, 6)]
    images = torch.cat([b[0] for b in _loader()])
    with torch.no_grad():
        outs = [m(images) for m in models]
    luts = [O.LUTS['camvid'], np.arange(5)]
    want, _ = O.multi_source_labels([o[0] for o in outs], [o[1] for o in outs], luts, 'all')
    marg = O.fuse_sources([o[0] for o in outs], [o[1] for o in outs], luts, 'all')["marginal"].numpy()
    lst, _ = U.generate_pseudo_label_multi_model(models, ['camvid', 'somewhere_else'], 'cuda:0', str(tmp_path / "a"), 0, N, None, None,
                                                 _args(merge_label_policy='all'), None, None, None, testloader=_loader(), batch_images=2)
    for i, ln in enumerate(open(lst)):
        got = np.array(Image.open(ln.strip().split(',')[1]))
        assert not ((got != want[i]) & ~marg[i]).any()
    with pytest.raises(ValueError, match="no label table"):
        U.generate_pseudo_label_multi_model([TinySource(13, 5), TinySource(20, 6)], ['camvid', 'somewhere_else'], 'cuda:0',
                                            str(tmp_path / "b"), 0, N, None, None, _args(merge_label_policy='all'), None, None, None,
                                            testloader=_loader(), batch_images=2)
