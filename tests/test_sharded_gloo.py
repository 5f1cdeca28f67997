"""world_size-2 gloo test (CPU) of the multi-GPU host logic: image sharding + histogram all-reduce give results
bit-identical to the single-process run."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import mspl_oracle as O

SOURCES = (("camvid", 13), ("cityscapes", 20), ("forest", 5))
N, H, W = 6, 24, 36


def _inputs():
    mains, auxs = [], []
    for i, (nm, c) in enumerate(SOURCES):
        m, a = O.synthetic_logits(N, c, H, W, seed=100 + i)
        mains.append(m), auxs.append(a)
    return mains, auxs, [O.LUTS[nm] for nm, _ in SOURCES]


def _worker(rank, world, port, policy, ds_rate, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mspl_b200.pipeline import LabelGenerator, shard_range
        import oracle_ops
        mains, auxs, luts = _inputs()
        lo, hi = shard_range(N, rank, world)
        gen = LabelGenerator(luts, policy=policy, portion=0.2, ds_rate=ds_rate, ops=oracle_ops)
        job = gen.run([m[lo:hi] for m in mains], [a[lo:hi] for a in auxs])
        out[rank] = dict(lo=lo, hi=hi, final=job.final, thresh=job.thresh, kept=job.kept, class_hist=job.class_hist,
                         final_hist=job.final_hist, marginal=job.marginal)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("policy,ds_rate", [("all", 1), ("half", 1), ("prob", 3)])
def test_two_ranks_equal_one(policy, ds_rate):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import oracle_ops
    from mspl_b200.pipeline import LabelGenerator, shard_range
    mains, auxs, luts = _inputs()
    single = LabelGenerator(luts, policy=policy, portion=0.2, ds_rate=ds_rate, ops=oracle_ops).run(mains, auxs)
    # the torch-op bracketed select used by the stand-in equals the sort-based definition
    ref = O.fuse_sources(mains, auxs, luts, policy)
    th_ref, kept_ref = O.cb_thresholds(ref["label"], ref["conf"], 0.2, ds_rate, ignore=4)
    assert torch.equal(single.thresh, th_ref) and torch.equal(single.kept, kept_ref)
    f_ref, _ = O.apply_thresholds(ref["label"], ref["conf"], th_ref)
    assert torch.equal(single.final, f_ref)
    th_all, _ = oracle_ops.cb_thresholds(ref["label"], ref["conf"], 0.2, ds_rate)          # every class resolved
    assert torch.equal(th_all, O.cb_thresholds(ref["label"], ref["conf"], 0.2, ds_rate)[0])

    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), policy, ds_rate, out), nprocs=world, join=True)
    finals = []
    for r in range(world):
        o = out[r]
        assert (o["lo"], o["hi"]) == shard_range(N, r, world)
        assert torch.equal(o["thresh"], single.thresh) and torch.equal(o["kept"], single.kept)
        assert torch.equal(o["class_hist"], single.class_hist) and torch.equal(o["final_hist"], single.final_hist)
        assert int(o["marginal"]) == int(single.marginal)
        finals.append(o["final"])
    assert torch.equal(torch.cat(finals), single.final)


def test_shard_range_covers_everything():
    from mspl_b200.pipeline import shard_range
    for n in (0, 1, 7, 2000, 20000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_class_weights_from_histogram():
    from mspl_b200.pipeline import class_weights_from_histogram
    hist = torch.tensor([0, 100, 300, 50, 550])
    assert torch.equal(class_weights_from_histogram(hist, 'normal'), O.class_weights_from_histogram(hist.numpy(), 'normal'))
    assert torch.equal(class_weights_from_histogram(hist, 'flat'), torch.ones(5))
