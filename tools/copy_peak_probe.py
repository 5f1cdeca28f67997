#!/usr/bin/env python
"""How does the plain device-to-device copy bandwidth (the figure MEASURED_PEAKS.json's hbm_gbs is) depend on the footprint and
the duration of the copy?  K1 reads 75-93 GB per launch for 12-17 ms; the peak was measured on a 2 GiB -> 2 GiB copy (~0.7 ms).
Prints GB/s (read + written bytes) for growing sizes, best and mean of several repetitions."""
import json
import torch

dev = torch.device("cuda:0")
rows = []
for gib in (2, 8, 16, 32, 40):
    n = gib * (1 << 30) // 4
    a = torch.empty(n, dtype=torch.float32, device=dev).normal_()
    b = torch.empty_like(a)
    b.copy_(a)
    torch.cuda.synchronize()
    times = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        b.copy_(a)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    nbytes = 2 * n * 4
    # read-only sweep of the same buffer (sum): what a pure-read kernel can pull
    s = a.sum()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s = a.sum()
    e1.record()
    torch.cuda.synchronize()
    rows.append({"copy_gib_each_way": gib, "ms_best": round(min(times), 3), "copy_gbs_best": round(nbytes / 1e6 / min(times), 1),
                 "copy_gbs_mean": round(nbytes / 1e6 / (sum(times) / len(times)), 1),
                 "read_only_sum_gbs": round(n * 4 / 1e6 / e0.elapsed_time(e1), 1)})
    del a, b
    torch.cuda.empty_cache()
print(json.dumps(rows))


def sustained(seconds=4.0, gib=8):
    """The same copy back to back for `seconds`: GB/s of every copy in launch order next to nvidia-smi's SM clock / power /
    throttle reasons -- what the HBM system delivers once the board's power limiter has settled (K1 runs inside 130-ms steps that
    follow seconds of warm-up and data generation, not in a cold burst)."""
    import subprocess
    import time
    n = gib * (1 << 30) // 4
    a = torch.empty(n, dtype=torch.float32, device=dev).normal_()
    b = torch.empty_like(a)
    b.copy_(a)
    torch.cuda.synchronize()
    smi = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap",
                            "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
    events = []
    t0 = time.time()
    while time.time() - t0 < seconds:
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            b.copy_(a)
            e1.record()
            events.append((e0, e1))
        torch.cuda.synchronize()
    smi.terminate()
    rates = [2 * n * 4 / 1e6 / x.elapsed_time(y) for x, y in events]
    k = max(1, len(rates) // 8)
    buckets = [round(sum(rates[i:i + k]) / len(rates[i:i + k]), 1) for i in range(0, len(rates), k)]
    lines = [ln.strip() for ln in smi.stdout.read().splitlines() if ln.strip()]
    print(json.dumps({"sustained_copy_gib_each_way": gib, "seconds": seconds, "copies": len(rates),
                      "gbs_first": round(rates[0], 1), "gbs_by_eighth_of_run": buckets, "gbs_last_quarter": round(sum(rates[-len(rates) // 4:]) / (len(rates) // 4), 1),
                      "nvidia_smi_sm_mhz,mem_mhz,watts,sw_power_cap": lines[::max(1, len(lines) // 10)]}))


sustained()
