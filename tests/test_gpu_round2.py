"""-m gpu: round-2 additions.

* a whole BASELINE configs[0]-sized batch (8 images of 480x256, 3 sources) against the CPU oracle -- labels, confidence,
  uncertainty, class histogram, class-balanced thresholds and the final maps -- so that every ring wrap-around of the persistent
  CTAs (tiles 2..N of each CTA, partial waves) is checked against the oracle directly, not only through kernel-vs-kernel tests;
* the grouped class order: label tables with target 0, with absent targets, single-class sources, exact ties between targets;
* the sharded threshold protocol with the GLOBAL final histogram (one all-reduce per phase) through the raw C ABI;
* N-rank NCCL == 1-rank, bit for bit, when at least two GPUs are visible.
"""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch

from oracle import mspl_oracle as O

pytestmark = pytest.mark.gpu

SOURCES = (("camvid", 13), ("cityscapes", 20), ("forest", 5))
RTOL = 1e-5
KLD_ATOL = 2e-6


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def ops():
    from mspl_b200 import ops as _ops
    return _ops


def _inputs(n, h, w, seed):
    mains, auxs = [], []
    for i, (nm, c) in enumerate(SOURCES):
        m, a = O.synthetic_logits(n, c, h, w, seed=seed + i)
        mains.append(m), auxs.append(a)
    return mains, auxs, [O.LUTS[nm] for nm, _ in SOURCES]


@pytest.mark.parametrize("policy", ["all", "half", "prob"])
def test_full_size_batch_against_the_oracle(ops, dev, policy):
    """configs[0]'s size: 8 x 480x256 x (13 + 20 + 5) classes.  1,024 tiles over 148 persistent CTAs: every CTA wraps its ring
    several times and the last wave is partial."""
    from mspl_b200.pipeline import LabelGenerator
    n, h, w = 8, 256, 480
    mains, auxs, luts = _inputs(n, h, w, seed=300)
    ref = O.fuse_sources(mains, auxs, luts, policy)
    gen = LabelGenerator(luts, policy=policy, portion=0.2)
    job = gen.run([m.to(dev) for m in mains], [a.to(dev) for a in auxs])
    lab = job.label.cpu()
    diff = lab != ref["label"]
    assert not bool((diff & ~ref["marginal"]).any()), "%d label mismatches outside near-ties" % int((diff & ~ref["marginal"]).sum())
    ok = ~diff
    torch.testing.assert_close(job.conf.cpu()[ok], ref["conf"][ok], rtol=RTOL, atol=1e-7)
    torch.testing.assert_close(job.unc.cpu(), ref["unc"], rtol=RTOL, atol=KLD_ATOL)
    if not bool(diff.any()):
        assert torch.equal(job.class_hist.cpu(), ref["class_hist"])
    assert abs(int(job.marginal) - int(ref["marginal"].sum())) <= max(2, int(ref["marginal"].sum()) // 10)
    # thresholds: exact order statistics of the kernel's own confidences, and within 1e-5 of the oracle's
    th_own, kept_own = O.cb_thresholds(lab, job.conf.cpu(), 0.2, ignore=4)
    assert torch.equal(job.thresh.cpu(), th_own) and torch.equal(job.kept.cpu(), kept_own)
    th_ref, _ = O.cb_thresholds(ref["label"], ref["conf"], 0.2, ignore=4)
    torch.testing.assert_close(job.thresh.cpu()[:4], th_ref[:4], rtol=RTOL, atol=0)
    f_own, _ = O.apply_thresholds(lab, job.conf.cpu(), th_own)
    assert torch.equal(job.final.cpu(), f_own)
    assert torch.equal(job.final_hist.cpu(), torch.bincount(f_own.reshape(-1).long(), minlength=5))
    # the same shard labelled in two batches plus a pool cycle: identical maps, doubled statistics
    md, ad = [m.to(dev) for m in mains], [a.to(dev) for a in auxs]
    shard = gen.begin(n, h, w, dev)
    gen.fuse_batch(shard, 0, [m[:3] for m in md], [a[:3] for a in ad])
    gen.fuse_batch(shard, 3, [m[3:] for m in md], [a[3:] for a in ad])
    job2 = gen.finish(shard)
    assert torch.equal(job2.label, job.label) and torch.equal(job2.conf, job.conf) and torch.equal(job2.final, job.final)
    assert torch.equal(job2.thresh, job.thresh) and torch.equal(job2.final_hist, job.final_hist)
    job3 = gen.run(md, ad, cycles=2)
    assert torch.equal(job3.label[:n], job.label) and torch.equal(job3.label[n:], job.label)
    assert torch.equal(job3.class_hist, 2 * job.class_hist) and torch.equal(job3.final[n:], job3.final[:n])
    assert gen.launches == (1 + 3) + (2 + 3) + (2 + 3)


@pytest.mark.parametrize("policy", ["all", "half", "prob", 1])
def test_grouped_class_order_handles_any_table(ops, dev, policy):
    """Tables that map classes to target 0, leave targets without any class, interleave targets, or belong to a single-class
    source: the kernels visit classes grouped by target, the oracle in the original order."""
    n, h, w = 2, 24, 40
    g = torch.Generator().manual_seed(71)
    tables = [np.array([3, 0, 1, 0, 3, 1, 2, 2, 0, 1, 3]),        # target 0 present, interleaved
              np.array([2]),                                     # one class
              np.array([1, 1, 1, 1, 1, 1, 1]),                   # one target only
              np.array([3, 3, 2, 2, 1, 1, 0, 0, 3, 2, 1, 0, 3, 2, 1, 0, 2])]    # descending, 17 classes: 4 chunks
    mains, auxs = [], []
    for i, t in enumerate(tables):
        m, a = O.synthetic_logits(n, len(t), h, w, seed=500 + i)
        mains.append(m), auxs.append(a)
    ref = O.fuse_sources(mains, auxs, tables, policy)
    r = ops.fuse_sources([m.to(dev) for m in mains], [a.to(dev) for a in auxs], tables, policy=policy, want_kld=True)
    diff = r.label.cpu() != ref["label"]
    assert not bool((diff & ~ref["marginal"]).any())
    ok = ~diff
    torch.testing.assert_close(r.conf.cpu()[ok], ref["conf"][ok], rtol=RTOL, atol=1e-7)
    torch.testing.assert_close(r.unc.cpu(), ref["unc"], rtol=RTOL, atol=KLD_ATOL)
    for got, want in zip(r.kld, ref["kld"]):
        torch.testing.assert_close(got.cpu(), want, rtol=RTOL, atol=KLD_ATOL)
    # the scalar fallback kernel (4-byte-offset views) runs the same arithmetic
    off = lambda t: torch.cat([torch.zeros(1), t.reshape(-1)]).to(dev)[1:].view(t.shape)
    r2 = ops.fuse_sources([off(m) for m in mains], [off(a) for a in auxs], tables, policy=policy)
    assert torch.equal(r2.label, r.label) and torch.equal(r2.conf, r.conf) and torch.equal(r2.unc, r.unc)


def test_exact_ties_between_targets_follow_the_original_class_order(ops, dev):
    """Two classes of DIFFERENT targets with bit-identical fused logits: np.argmax takes the first class in the original order
    (uest_seg_multi_os.py:904), wherever the grouped visiting order puts it."""
    h, w = 16, 32
    table = np.array([3, 1, 2, 1, 3, 2, 1])
    C = len(table)
    g = torch.Generator().manual_seed(5)
    m = torch.randn(1, C, h, w, generator=g)
    a = torch.randn(1, C, h, w, generator=g)
    # make class pairs tie exactly for the maximum on parts of the image: (0, 1): targets 3 vs 1 -> original order says 3
    top = 9.0
    m[:, 0, :, :8], a[:, 0, :, :8] = top, 2.0
    m[:, 1, :, :8], a[:, 1, :, :8] = top, 2.0
    m[:, 5, :, 8:16], a[:, 5, :, 8:16] = top + 1, 0.0           # (5, 6): targets 2 vs 1 -> 2 (class 5 comes first)
    m[:, 6, :, 8:16], a[:, 6, :, 8:16] = top, 2.0
    m[:, 3, :, 16:24], a[:, 3, :, 16:24] = top, 4.0             # (3, 4): targets 1 vs 3 -> 1
    m[:, 4, :, 16:24], a[:, 4, :, 16:24] = top + 2, 0.0
    ref = O.fuse_sources([m], [a], [table], 'all')
    for tensors in (([m.to(dev)], [a.to(dev)]),):
        r = ops.fuse_sources(tensors[0], tensors[1], [table], policy='all')
        lab = r.label.cpu()
        assert torch.equal(lab, ref["label"])
        assert bool((lab[0, :, :8] == 3).all()) and bool((lab[0, :, 8:16] == 2).all()) and bool((lab[0, :, 16:24] == 1).all())
    # labels-only kernel (original class order, no softmax) agrees as well
    lean = ops.fuse_sources([m.to(dev)], [a.to(dev)], [table], policy='all', want_conf=False, want_unc=False, want_conf_hist=False,
                            count_marginal=False)
    assert torch.equal(lean.label.cpu(), ref["label"])


def test_sharded_protocol_global_final_hist_raw_abi(dev):
    """Two emulated ranks through the raw C ABI with the round-2 protocol: the linear histogram is all-reduced, bracket_select
    reads the GLOBAL settled counts off it, the last candidate select adds the GLOBAL patch, and cand_apply touches no counter:
    every rank ends with the same, global, final histogram without an all-reduce of its own."""
    from mspl_b200 import _lib
    lib = _lib.load()
    K, h, w = 5, 48, 64
    gen = torch.Generator().manual_seed(31)
    label = torch.randint(0, K, (6, h, w), generator=gen).to(torch.uint8).to(dev)
    conf = torch.rand((6, h, w), generator=gen).to(dev)
    conf[0, :4] = 1.0
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None

    class Rank:
        def __init__(self, lab, cf):
            self.lab, self.cf = lab.contiguous(), cf.contiguous()
            self.state = torch.zeros(lib.mspl_radix_state_bytes(K), dtype=torch.uint8, device=dev)
            self.thresh = torch.empty(K, dtype=torch.float32, device=dev)
            self.bracket = torch.empty((K, 2), dtype=torch.float32, device=dev)
            self.kept = torch.zeros(K, dtype=torch.int64, device=dev)
            self.hist = torch.zeros((K, 2048), dtype=torch.int64, device=dev)
            self.fh = torch.zeros(K, dtype=torch.int64, device=dev)
            self.final = torch.empty_like(self.lab)
            self.cand = torch.empty(self.lab.numel(), dtype=torch.int32, device=dev)
            self.count = torch.zeros((), dtype=torch.int64, device=dev)

    def all_reduce(ranks):
        total = sum(r.hist for r in ranks)
        for r in ranks:
            r.hist.copy_(total)

    for portion in (0.25, 1e-5):                               # 1e-5: every class has j == 0 -> threshold 1.0 (count-only passes)
        ranks = [Rank(label[:2], conf[:2]), Rank(label[2:], conf[2:])]
        for r in ranks:
            _lib.check(lib.mspl_conf_hist(p(r.lab), p(r.cf), r.lab.numel(), h * w, K, p(r.hist), 1, st), "conf_hist")
        all_reduce(ranks)
        for r in ranks:
            _lib.check(lib.mspl_bracket_select(p(r.hist), K, portion, 4, p(r.state), p(r.bracket), p(r.thresh), p(r.kept), None,
                                               p(r.fh), st), "bracket_select")
            _lib.check(lib.mspl_bracket_classify(p(r.lab), p(r.cf), p(r.bracket), r.lab.numel(), K, 4, p(r.final), None, None,
                                                 p(r.cand), p(r.count), st), "classify")
        for ps in range(3):
            for r in ranks:
                _lib.check(lib.mspl_cand_hist_pass(p(r.lab), p(r.cf), p(r.cand), p(r.count), h * w, K, ps, p(r.state), p(r.hist), 1, st),
                           "cand_hist")
            all_reduce(ranks)
            for r in ranks:
                _lib.check(lib.mspl_cand_select(p(r.hist), K, ps, p(r.state), p(r.thresh), p(r.fh), 4, st), "cand_select")
        for r in ranks:
            _lib.check(lib.mspl_cand_apply(p(r.lab), p(r.cf), p(r.thresh), p(r.cand), p(r.count), K, 4, p(r.final), None, None, st),
                       "cand_apply")
        th_ref, kept_ref = O.cb_thresholds(label.cpu(), conf.cpu(), portion, ignore=4)
        f_ref, _ = O.apply_thresholds(label.cpu(), conf.cpu(), th_ref)
        want_hist = torch.bincount(f_ref.reshape(-1).long(), minlength=K)
        for r in ranks:
            assert torch.equal(r.thresh.cpu(), th_ref) and torch.equal(r.kept.cpu(), kept_ref)
            assert torch.equal(r.fh.cpu(), want_hist)                        # GLOBAL on every rank
        assert torch.equal(torch.cat([r.final for r in ranks]).cpu(), f_ref)
        # ... and the single-launch tail of one rank over everything gives the same
        one = Rank(label, conf)
        _lib.check(lib.mspl_conf_hist(p(one.lab), p(one.cf), one.lab.numel(), h * w, K, p(one.hist), 1, st), "conf_hist")
        _lib.check(lib.mspl_bracket_select(p(one.hist), K, portion, 4, p(one.state), p(one.bracket), p(one.thresh), p(one.kept), None,
                                           p(one.fh), st), "bracket_select")
        _lib.check(lib.mspl_bracket_classify(p(one.lab), p(one.cf), p(one.bracket), one.lab.numel(), K, 4, p(one.final), None, None,
                                             p(one.cand), p(one.count), st), "classify")
        _lib.check(lib.mspl_cand_resolve(p(one.lab), p(one.cf), p(one.cand), p(one.count), h * w, K, 4, 1, p(one.state), p(one.thresh),
                                         p(one.final), None, p(one.fh), st), "cand_resolve")
        assert torch.equal(one.thresh.cpu(), th_ref) and torch.equal(one.final.cpu(), f_ref) and torch.equal(one.fh.cpu(), want_hist)


# ---- N ranks over NCCL == 1 rank ------------------------------------------------------------------------------------------
def _nccl_worker(rank, world, port, policy, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from mspl_b200.pipeline import LabelGenerator, shard_range
        mains, auxs, luts = _inputs(12, 64, 96, seed=900)
        lo, hi = shard_range(12, rank, world)
        gen = LabelGenerator(luts, policy=policy, portion=0.2)
        job = gen.run([m[lo:hi].to(dev) for m in mains], [a[lo:hi].to(dev) for a in auxs])
        torch.save(dict(lo=lo, hi=hi, final=job.final.cpu(), label=job.label.cpu(), thresh=job.thresh.cpu(), kept=job.kept.cpu(),
                        class_hist=job.class_hist.cpu(), final_hist=job.final_hist.cpu(), marginal=int(job.marginal),
                        collectives=gen.collectives), os.path.join(out_dir, "rank%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("policy", ["all", "half"])
def test_nccl_ranks_equal_one_rank(tmp_path, policy):
    """SURVEY.md section 4, tier T4 on hardware: 2 (or 4) ranks over NCCL give labels, thresholds and histograms bit-identical
    to the single-rank run on the same global image set, with 4 collectives per job."""
    import torch.multiprocessing as mp
    n_gpus = torch.cuda.device_count()
    if n_gpus < 2:
        pytest.skip("needs at least two GPUs (runs under `gpurun --gpus 2`)")
    world = 4 if n_gpus >= 4 else 2
    from mspl_b200.pipeline import LabelGenerator
    mains, auxs, luts = _inputs(12, 64, 96, seed=900)
    d0 = torch.device("cuda:0")
    single = LabelGenerator(luts, policy=policy, portion=0.2).run([m.to(d0) for m in mains], [a.to(d0) for a in auxs])
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_nccl_worker, args=(world, port, policy, str(tmp_path)), nprocs=world, join=True)
    finals, labels = [], []
    for r in range(world):
        o = torch.load(os.path.join(str(tmp_path), "rank%d.pt" % r))
        assert torch.equal(o["thresh"], single.thresh.cpu()) and torch.equal(o["kept"], single.kept.cpu())
        assert torch.equal(o["class_hist"], single.class_hist.cpu()) and torch.equal(o["final_hist"], single.final_hist.cpu())
        assert o["marginal"] == int(single.marginal) and o["collectives"] == 4
        finals.append(o["final"]), labels.append(o["label"])
    assert torch.equal(torch.cat(finals), single.final.cpu()) and torch.equal(torch.cat(labels), single.label.cpu())
