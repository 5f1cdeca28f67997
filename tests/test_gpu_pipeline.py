"""-m gpu: mspl_b200.pipeline.LabelGenerator -- the public whole-job API that bench.py measures (`run` on resident logits,
`run_from_host` on host buffers) -- against the oracle pipeline (the same class driven by the oracle-backed stand-in ops,
itself checked against the sort-based definition in tests/test_sharded_gloo.py)."""
import os
import sys

import pytest
import torch

from oracle import mspl_oracle as O

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_ops  # noqa: E402

pytestmark = pytest.mark.gpu

SOURCES = (("camvid", 13), ("cityscapes", 20), ("forest", 5))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _inputs(n, h, w, seed):
    mains, auxs = [], []
    for i, (_, c) in enumerate(SOURCES):
        m, a = O.synthetic_logits(n, c, h, w, seed=seed + i)
        mains.append(m), auxs.append(a)
    return mains, auxs, [O.LUTS[s] for s, _ in SOURCES]


@pytest.mark.parametrize("policy,portion,ds_rate", [("all", 0.2, 1), ("half", 0.3, 1), ("prob", 0.5, 2)])
def test_label_generator_matches_oracle_pipeline(dev, policy, portion, ds_rate):
    from mspl_b200.pipeline import LabelGenerator
    n, h, w = 7, 48, 64
    mains, auxs, luts = _inputs(n, h, w, seed=400)
    ref = LabelGenerator(luts, policy=policy, portion=portion, ds_rate=ds_rate, ops=oracle_ops).run(mains, auxs)
    fused = O.fuse_sources(mains, auxs, luts, policy)
    gen = LabelGenerator(luts, policy=policy, portion=portion, ds_rate=ds_rate)
    job = gen.run([m.to(dev) for m in mains], [a.to(dev) for a in auxs], want_mask=True)
    marginal = fused["marginal"]
    label_diff = job.label.cpu() != ref.label
    assert not bool((label_diff & ~marginal).any())
    k = torch.arange(5) != 4
    torch.testing.assert_close(job.thresh.cpu()[k], ref.thresh[k], rtol=1e-5, atol=0)
    assert job.thresh[4].item() == float("inf")
    if not bool(label_diff.any()):
        assert torch.equal(job.class_hist.cpu(), ref.class_hist) and torch.equal(job.kept.cpu(), ref.kept)
    # final maps: equal wherever the label agrees and conf is not within rounding of its class threshold
    th = ref.thresh[ref.label.long()]
    near = (ref.conf - th).abs() <= 2e-5 * th.clamp(max=1.0)
    assert not bool(((job.final.cpu() != ref.final) & ~label_diff & ~near).any())
    assert torch.equal(job.mask.cpu(), (job.final == 4).to(torch.uint8).cpu())
    assert torch.equal(job.final_hist.cpu(), torch.bincount(job.final.reshape(-1).long().cpu(), minlength=5))
    assert gen.launches == 1 + gen.ops.SELECT_AND_APPLY_LAUNCHES

    # host-buffer route: chunks that do not divide the image count, pinned and pageable inputs, caller-provided output
    want = job.final.cpu()
    for pin in (True, False):
        hm = [m.pin_memory() if pin else m for m in mains]
        ha = [a.pin_memory() if pin else a for a in auxs]
        out = torch.full((n, h, w), 77, dtype=torch.uint8).pin_memory()
        got, hjob = LabelGenerator(luts, policy=policy, portion=portion, ds_rate=ds_rate).run_from_host(
            hm, ha, dev, chunk_images=3, out_host=out)
        assert got is out and torch.equal(out, want)          # no synchronize needed: the maps are there on return
        assert torch.equal(hjob.class_hist, job.class_hist) and torch.equal(hjob.final_hist, job.final_hist)
        assert torch.equal(hjob.thresh, job.thresh) and int(hjob.marginal) == int(job.marginal)


def test_label_generator_without_thresholds_is_the_reference_loop(dev):
    """thresholds=False: exactly what the reference's loop keeps -- the voted label map and class_array."""
    from mspl_b200.pipeline import LabelGenerator, class_weights_from_histogram
    mains, auxs, luts = _inputs(5, 40, 36, seed=500)
    labels, class_array = O.multi_source_labels(mains, auxs, luts, "all")
    job = LabelGenerator(luts, policy="all", thresholds=False).run([m.to(dev) for m in mains], [a.to(dev) for a in auxs])
    fused = O.fuse_sources(mains, auxs, luts, "all")
    got = job.final.cpu()
    want = torch.as_tensor(labels).to(torch.uint8).reshape(got.shape)
    assert not bool(((got != want) & ~fused["marginal"]).any())
    if bool((got == want).all()):
        assert job.class_hist.cpu().tolist() == [int(x) for x in class_array]
        torch.testing.assert_close(class_weights_from_histogram(job.class_hist),
                                   O.class_weights_from_histogram(class_array, 'normal'))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_ops_follow_their_tensors_to_a_non_current_device():
    """Tensors on cuda:1 while cuda:0 is the current device (and a non-default stream is current on cuda:1): every launch must
    go to the tensors' device and stream -- results equal those computed on cuda:0."""
    from mspl_b200 import ops
    from mspl_b200.pipeline import LabelGenerator
    d0, d1 = torch.device("cuda:0"), torch.device("cuda:1")
    torch.cuda.set_device(d0)
    mains, auxs, luts = _inputs(3, 32, 48, seed=600)
    want = LabelGenerator(luts, policy="half").run([m.to(d0) for m in mains], [a.to(d0) for a in auxs])
    side = torch.cuda.Stream(d1)
    with torch.cuda.stream(side):
        assert torch.cuda.current_device() == 1
    assert torch.cuda.current_device() == 0
    m1, a1 = [m.to(d1) for m in mains], [a.to(d1) for a in auxs]
    torch.cuda.synchronize(d1)
    with torch.cuda.stream(side):           # current stream on cuda:1 is `side`; leave the block -> current device is cuda:0 again
        pass
    got = LabelGenerator(luts, policy="half").run(m1, a1)
    assert got.final.device == d1
    assert torch.equal(got.final.cpu(), want.final.cpu()) and torch.equal(got.thresh.cpu(), want.thresh.cpu())
    assert torch.equal(got.class_hist.cpu(), want.class_hist.cpu())
    k = 5
    main, aux = O.synthetic_logits(2, k, 24, 36, seed=601)
    target = torch.randint(0, k, (2, 24, 36), generator=torch.Generator().manual_seed(3))
    cw = torch.ones(k)
    o0 = ops.uw_ce_fwd_bwd(main.to(d0), aux.to(d0), target.to(d0), cw.to(d0))
    c1 = torch.zeros((3, k), dtype=torch.int64, device=d1)
    o1 = ops.uw_ce_fwd_bwd(main.to(d1), aux.to(d1), target.to(d1), cw.to(d1), iou_counts=c1)
    for x, y in zip(o0, o1):
        assert y.device == d1 and torch.equal(x.cpu(), y.cpu())
    assert torch.equal(c1.cpu(), ops.miou_counts(main.to(d0), target.to(d0), k).cpu())
    assert torch.cuda.current_device() == 0
