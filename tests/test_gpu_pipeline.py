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
