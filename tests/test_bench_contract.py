"""CPU tests of bench.py's host logic: the reference arm prints exactly one strict-JSON line with the contract's keys, the
image-parallel arrangement of the CPU port computes the same labels and class counts as the sequential loop, and non-finite
floats never reach the output."""
import json
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import mspl_oracle as O  # noqa: E402


def test_strict_json_has_no_nan_or_infinity():
    line = {"a": [1.0, float("inf"), float("-inf"), float("nan")], "b": {"c": (2, 3.5, float("inf"))}, "d": "inf", "e": None}
    out = json.dumps(bench._strict(line), allow_nan=False)
    assert json.loads(out) == {"a": [1.0, None, None, None], "b": {"c": [2, 3.5, None]}, "d": "inf", "e": None}


def test_image_parallel_arrangement_equals_sequential_loop():
    """Both arrangements of the CPU arm (the reference's sequential loop; images dealt to worker threads) give the same class
    weights, with the live reference's functions (when /root/reference or oracle/_ref is present) and with the oracle port."""
    mains, auxs = bench.make_logits_host(torch, 5, 24, 32, seed=3, pin=False)
    cpu = bench.CpuPath()
    for policy in ("all", "half"):
        w_seq = cpu.step(torch, np, mains, auxs, policy)
        with ThreadPoolExecutor(3) as pool:
            w_par = cpu.step(torch, np, mains, auxs, policy, pool, 3)
        assert torch.equal(w_seq, w_par)
        _, class_array = O.multi_source_labels(mains, auxs, [O.LUTS[s] for s, _ in bench.SOURCES], policy)
        assert torch.equal(w_seq, O.class_weights_from_histogram(class_array, 'normal'))     # reference arm == oracle port


def test_shard_plan_covers_the_named_configs():
    import argparse
    a = argparse.Namespace(images_total=0)
    assert bench.shard_plan(a, 1) == (2000, 2000, 1)              # configs[1]
    assert bench.shard_plan(a, 2) == (10000, 2500, 4)             # configs[2]: 20,000 images over the ranks
    assert bench.shard_plan(a, 4) == (5000, 2500, 2)
    assert bench.shard_plan(a, 8) == (2500, 2500, 1)
    a.images_total = 20000
    assert bench.shard_plan(a, 1) == (20000, 2500, 8)


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                        "--warmup", "1", "--ref-images", "2", "--height", "32", "--width", "48"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0], parse_constant=lambda c: (_ for _ in ()).throw(ValueError("non-strict JSON constant " + c)))
    assert d["impl"] == "reference" and d["unit"] == "Mpix/s" and d["higher_is_better"] is True
    assert d["metric"] == "pseudo-labelled Mpix/s (3-source fusion)" and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["gpu_launches"] == 0
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1", PYTHONDONTWRITEBYTECODE="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600)
    assert p.returncode == 0 and p.stdout.strip() == ""
