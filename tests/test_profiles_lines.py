"""The recorded bench lines under profiles/ keep the driver contract's keys and are mutually consistent: same metric and unit at
every rank count, configs[1] at N = 1 and configs[2] at N > 1, one rank-count-independent digest for the 20,000-image set, a
roofline block whose fraction is achieved / peak, an end-to-end block with the copies declared."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROFILES = os.path.join(ROOT, "profiles")


def _line(name):
    path = os.path.join(PROFILES, name)
    if not os.path.exists(path):
        pytest.skip(name + " not recorded")
    text = [ln for ln in open(path).read().splitlines() if ln.strip()]
    return json.loads(text[-1])


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_recorded_line_has_the_contract_keys(n):
    d = _line("r02_bench_n%d.json" % n)
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "roofline", "e2e", "clocks", "gpu_launches"):
        assert key in d, key
    assert d["n_gpus"] == n and d["unit"] == "Mpix/s" and d["higher_is_better"] is True and d["dtype"] == "f32"
    assert d["metric"] == "pseudo-labelled Mpix/s (3-source fusion)" and d["vs_baseline"] is None
    assert "workload" in d["config"] and d["gpu_launches"] > 0 and d["warmup"] >= 3
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 2e-3
    assert r["traffic"] is None or 0.9 < r["traffic"] / r["algorithmic_bytes_per_launch"] < 1.1
    e = d["e2e"]
    assert e["unit"] == "Mpix/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    assert "hw_slowdown" not in d["clocks"]["reasons"] and "hw_thermal_slowdown" not in d["clocks"]["reasons"]
    if n == 1:
        assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
        assert "2,000" in d["config"]["workload"] or "2000" in d["config"]["workload"]
    else:
        assert "20,000" in d["config"]["workload"] or "20000" in d["config"]["workload"]
        assert d["scaling"] == "strong" and d["collectives_per_step"] <= 4


def test_one_digest_for_the_20000_image_set_at_every_rank_count():
    digests = {_line("r02_bench_n1.json")["secondary"]["configs2_one_gpu"]["digest"]}
    for n in (2, 4, 8):
        digests.add(_line("r02_bench_n%d.json" % n)["results"]["digest"])
    assert len(digests) == 1, digests


def test_reference_arm_line():
    d = _line("r02_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "reference" and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
