#!/usr/bin/env python
"""Benchmark of the hot path: 3-source pseudo-label generation (Cityscapes/CamVid/Forest logits -> greenhouse classes)
with class-balanced thresholds, BASELINE.json configs[1] per GPU.

  python bench.py --gpus N --steps K --warmup W              (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W    (CPU reference arm: the oracle port, rank 0 only)

A "step" is one whole pass of the path over this rank's resident synthetic batch: K1 fuse_sources (+ radix pass 0) ->
[all-reduce of the histograms] -> radix passes 1-2 + selects -> K3 apply_thresholds.  `value` = pixels labelled by all
ranks / max-over-ranks device time, inputs resident in HBM.  `e2e` = the same metric through the public host-buffer API
(LabelGenerator.run_from_host): pinned host logits -> H2D -> kernels -> D2H of the uint8 label maps, all inside the
timed region.  `roofline` is the fused kernel K1 alone (CUDA events around each launch, live) against the measured HBM
peak in MEASURED_PEAKS.json.  `cpu_baseline` is the oracle port of the reference's CPU path on a bounded sample, run both
as the reference's sequential per-image loop and with the images spread over all host cores (the better one is reported).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SOURCES = (("camvid", 13), ("cityscapes", 20), ("forest", 5))
SIGMA = 3.0
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback, used only if MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("native", "reference"), default="native")
    ap.add_argument("--images-per-gpu", type=int, default=2000, help="configs[1]: 2,000 synthetic 480x256 images on 1 B200")
    ap.add_argument("--height", type=int, default=256)
    ap.add_argument("--width", type=int, default=480)
    ap.add_argument("--policy", default="all", help="espdnet_greenhouse_uest_multi_os.sh: --merge-label-policy all")
    ap.add_argument("--portion", type=float, default=0.2, help="INIT_TGT_PORT (uest_seg_multi_os.py:90)")
    ap.add_argument("--e2e-images", type=int, default=128, help="images per end-to-end step (host buffers)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline budget (bounded sample)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--ref-images", type=int, default=32, help="--impl reference: images per step")
    return ap.parse_args()


# ---- synthetic inputs: main = sigma*randn + per-image class bias; aux = main + 0.5*sigma*randn ----------
def make_logits_device(torch, n, h, w, device, seed, chunk=50):
    gen = torch.Generator(device=device).manual_seed(seed)
    mains, auxs = [], []
    for _, c in SOURCES:
        m = torch.empty((n, c, h, w), dtype=torch.float32, device=device)
        a = torch.empty((n, c, h, w), dtype=torch.float32, device=device)
        for lo in range(0, n, chunk):
            hi = min(n, lo + chunk)
            m[lo:hi].normal_(0.0, SIGMA, generator=gen)
            m[lo:hi] += SIGMA * torch.randn((hi - lo, c, 1, 1), device=device, generator=gen)
            a[lo:hi].normal_(0.0, 0.5 * SIGMA, generator=gen)
            a[lo:hi] += m[lo:hi]
        mains.append(m), auxs.append(a)
    return mains, auxs


def make_logits_host(torch, n, h, w, seed, pin):
    gen = torch.Generator().manual_seed(seed)
    mains, auxs = [], []
    for _, c in SOURCES:
        m = SIGMA * torch.randn((n, c, h, w), generator=gen) + SIGMA * torch.randn((n, c, 1, 1), generator=gen)
        a = m + 0.5 * SIGMA * torch.randn((n, c, h, w), generator=gen)
        mains.append(m.pin_memory() if pin else m), auxs.append(a.pin_memory() if pin else a)
    return mains, auxs


# ---- clocks during the timed region (B200_PROFILING.md) ---------------------------------------------------------------
class ClockSampler:
    """nvidia-smi polled every 20 ms from process start; stop(t0, t1) keeps the samples taken inside the timed region."""
    QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self, t0, t1):
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        rows = []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), float(f[3]), [nm for nm, v in zip(names, f[5:9]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        inside = [r for r in rows if t0 <= r[0] <= t1]
        window = "timed region"
        if len(inside) < 3:      # very short timed region: widen to the neighbouring samples and say so
            inside = [r for r in rows if t0 - 0.25 <= r[0] <= t1 + 0.1]
            window = "timed region +-0.25 s (region shorter than 3 samples)"
        sm = sorted(r[1] for r in inside)
        reasons = sorted({nm for r in inside for nm in r[4]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((r[2] for r in inside), default=None),
                "power_w_max": max((r[3] for r in inside), default=None), "samples": len(inside), "window": window,
                "reasons": reasons}


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one K1 launch from the committed ncu capture, if any."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))
    except Exception:
        return None


# ---- CPU reference arm / baseline: the oracle port of uest_seg_multi_os.py:897-921 -----------------------------------------
def cpu_reference_step(O, mains, auxs, luts, policy):
    labels, class_array = O.multi_source_labels(mains, auxs, luts, policy)
    return O.class_weights_from_histogram(class_array, 'normal'), labels


def cpu_reference_step_image_parallel(O, mains, auxs, luts, policy, pool, workers):
    """The same per-image loop with the images dealt to `workers` host threads (torch's CPU ops release the GIL): the
    reference's loop is sequential and only gets torch's intra-op threads, which the small per-image tensors do not fill,
    so this is the arrangement that uses every host core; class_array is summed over the workers."""
    n = mains[0].shape[0]
    parts = list(pool.map(lambda i: O.multi_source_labels([m[i:i + 1] for m in mains], [a[i:i + 1] for a in auxs], luts, policy),
                          range(n)))
    class_array = sum(p[1] for p in parts)
    return O.class_weights_from_histogram(class_array, 'normal'), [p[0] for p in parts]


def cpu_rates(torch, O, mains, auxs, luts, policy, budget_s, max_images):
    """Mpix/s of the port, (a) as the reference runs it -- one image after the other, torch intra-op threads -- and (b) with
    the images spread over all host cores, each on `budget_s`/2 seconds of a bounded sample.  Returns the better one as
    the baseline, with both in the description."""
    from concurrent.futures import ThreadPoolExecutor
    n, h, w = mains[0].shape[0], mains[0].shape[2], mains[0].shape[3]
    threads = torch.get_num_threads()
    cores = os.cpu_count() or threads

    def timed(fn):
        done, t0 = 0, time.perf_counter()
        while True:
            fn()
            done += n
            el = time.perf_counter() - t0
            if el >= budget_s / 2 or done >= max_images:
                return done, el

    cpu_reference_step(O, [m[:1] for m in mains], [a[:1] for a in auxs], luts, policy)     # warm-up
    d_seq, t_seq = timed(lambda: cpu_reference_step(O, mains, auxs, luts, policy))
    torch.set_num_threads(1)
    try:
        with ThreadPoolExecutor(cores) as pool:
            cpu_reference_step_image_parallel(O, [m[:cores] for m in mains], [a[:cores] for a in auxs], luts, policy, pool, cores)
            d_par, t_par = timed(lambda: cpu_reference_step_image_parallel(O, mains, auxs, luts, policy, pool, cores))
    finally:
        torch.set_num_threads(threads)
    mpix = h * w / 1e6
    seq, par = d_seq * mpix / t_seq, d_par * mpix / t_par
    best, used = (par, cores) if par >= seq else (seq, threads)
    sample = ("synthetic %dx%d images x 3 sources, oracle port of get_output->argmax->LUT->merge_outputs('%s')->class_array "
              "(uest_seg_multi_os.py:897-921): %d images in %.1f s as the reference's sequential loop with %d torch threads "
              "(%.2f Mpix/s), %d images in %.1f s with the images spread over %d host threads (%.2f Mpix/s); value = the better"
              % (w, h, policy, d_seq, t_seq, threads, seq, d_par, t_par, cores, par))
    return round(best, 3), used, sample, (d_seq + d_par, t_seq + t_par)


def cpu_baseline(torch, args, budget_s):
    from oracle import mspl_oracle as O
    mains, auxs = make_logits_host(torch, 16, args.height, args.width, seed=3, pin=False)
    luts = [O.LUTS[s] for s, _ in SOURCES]
    value, used, sample, _ = cpu_rates(torch, O, mains, auxs, luts, args.policy, budget_s, 512)
    return {"value": value, "unit": "Mpix/s", "cores": used, "kind": "port", "host_cpus": os.cpu_count(), "sample": sample}


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import mspl_oracle as O
    from concurrent.futures import ThreadPoolExecutor
    n = args.ref_images
    mains, auxs = make_logits_host(torch, n, args.height, args.width, seed=3, pin=False)
    luts = [O.LUTS[s] for s, _ in SOURCES]
    threads, cores = torch.get_num_threads(), os.cpu_count() or 1
    # which arrangement of the port uses this box's cores best (short probe), then W warm-up + K timed steps of that one
    _, used, probe, _ = cpu_rates(torch, O, [m[:min(n, 8)] for m in mains], [a[:min(n, 8)] for a in auxs], luts, args.policy, 4.0, 64)
    parallel = used == cores and cores > 1
    pool = ThreadPoolExecutor(cores) if parallel else None
    if parallel:
        torch.set_num_threads(1)

    def step():
        if parallel:
            return cpu_reference_step_image_parallel(O, mains, auxs, luts, args.policy, pool, cores)
        return cpu_reference_step(O, mains, auxs, luts, args.policy)

    try:
        for _ in range(max(1, args.warmup)):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        el = time.perf_counter() - t0
    finally:
        if pool is not None:
            pool.shutdown()
        torch.set_num_threads(threads)
    mpix = args.steps * n * args.height * args.width / 1e6
    val = round(mpix / el, 3)
    sample = ("%d synthetic %dx%d images x 3 sources per step (bounded sample of configs[1]), oracle port of the reference CPU "
              "path (%s); /root/reference is pure Python and cannot travel to the GPU box.  Probe: %s"
              % (n, args.width, args.height,
                 "images spread over %d host threads" % cores if parallel else "sequential loop, %d torch threads" % threads, probe))
    emit({
        "impl": "reference", "metric": "pseudo-labelled Mpix/s (3-source fusion)", "value": val, "unit": "Mpix/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * el / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.images_per_gpu),
        "cpu_baseline": {"value": val, "unit": "Mpix/s", "cores": cores if parallel else threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})


def workload_config(args, images_per_gpu):
    return {"workload": "configs[1]: 3-source fusion (camvid 13 / cityscapes 20 / forest 5 logits -> 5 greenhouse classes), "
                        "class-balanced thresholds, %d synthetic %dx%d images per GPU" % (images_per_gpu, args.width, args.height),
            "images_per_gpu": images_per_gpu, "height": args.height, "width": args.width, "policy": args.policy,
            "portion": args.portion, "sharding": "images, contiguous blocks; histogram all-reduce only",
            "l2": "inputs (%.1f GB/GPU) far larger than the 126 MB L2; no flush needed" % (images_per_gpu * args.height * args.width * 304 / 1e9)}


_REAL_STDOUT = None


def claim_stdout():
    """Route everything libraries print on fd 1 (NCCL's version banner, torchrun notices) to stderr, so that stdout
    carries exactly ONE line: the JSON result written by emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr


def _strict(x):
    """Strict JSON has no NaN/Infinity tokens: non-finite floats become null."""
    if isinstance(x, float):
        return x if x == x and abs(x) != float("inf") else None
    if isinstance(x, dict):
        return {k: _strict(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_strict(v) for v in x]
    return x


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(_strict(line), allow_nan=False) + "\n")
    out.flush()


def full_size_checks(torch, job, pix_local, world, dist, portion):
    """Properties that must hold for any input, verified on the very outputs of the timed run (all ranks): histograms are
    exact pixel counts, thresholding only ever moves a label to the ignore class and keeps conf >= thresh, about `portion`
    of each class survives, and the global histograms equal the sum of the per-rank bincounts."""
    ok = {}
    K = job.class_hist.numel()
    local_hist = torch.bincount(job.label.reshape(-1).long(), minlength=K)[:K]
    local_final = torch.bincount(job.final.reshape(-1).long(), minlength=K)[:K]
    if dist is not None:
        dist.all_reduce(local_hist)
        dist.all_reduce(local_final)
    ok["class_hist_is_bincount_of_labels"] = bool(torch.equal(local_hist, job.class_hist))
    ok["final_hist_is_bincount_of_final"] = bool(torch.equal(local_final, job.final_hist))
    ok["histograms_count_every_pixel"] = int(job.class_hist.sum()) == pix_local * world == int(job.final_hist.sum())
    ok["kept_equals_class_hist"] = bool(torch.equal(job.kept, job.class_hist))
    changed = job.final != job.label
    ok["threshold_only_moves_to_ignore"] = bool((job.final[changed] == 4).all())
    keep = job.final != 4
    ok["kept_pixels_reach_their_threshold"] = bool((job.conf[keep] >= job.thresh[job.final[keep].long()]).all())
    ok["dropped_pixels_below_threshold"] = bool((job.conf[changed] < job.thresh[job.label[changed].long()]).all())
    frac = [(int(job.final_hist[k]) / max(1, int(job.class_hist[k]))) for k in range(1, 4)]
    ok["kept_fraction_close_to_portion"] = all(abs(f - portion) < 0.01 for f in frac)
    ok["conf_in_unit_interval"] = bool(((job.conf >= 0) & (job.conf <= 1.0000001)).all())
    ok["uncertainty_finite_nonnegative"] = bool(torch.isfinite(job.unc).all()) and float(job.unc.min()) > -1e-5
    ok["all_passed"] = all(v for v in ok.values())
    return ok


def main():
    args = parse_args()
    claim_stdout()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist
    from mspl_b200 import _lib
    from mspl_b200.data_loader.segmentation.greenhouse import SOURCE_TABLES
    from mspl_b200.pipeline import LabelGenerator

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n, h, w = args.images_per_gpu, args.height, args.width
    pix_local = n * h * w
    luts = [SOURCE_TABLES[s] for s, _ in SOURCES]
    mains, auxs = make_logits_device(torch, n, h, w, dev, seed=3 + 1000 * rank)
    gen = LabelGenerator(luts, policy=args.policy, portion=args.portion)
    torch.cuda.synchronize()

    gen.k1_events = []      # CUDA events around every K1 launch (roofline of the dominant kernel, measured live)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    job = None
    for _ in range(args.warmup):
        job = gen.run(mains, auxs)
    barrier()
    gen.k1_events.clear()
    launches0 = gen.launches
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    t0.record()
    for _ in range(args.steps):
        job = gen.run(mains, auxs)
    t1.record()
    barrier()
    wall1 = time.time()
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
    k1_events = list(gen.k1_events)
    k1_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in k1_events) / max(1, len(k1_events))], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(k1_ms, op=dist.ReduceOp.MAX)
    ms_per_step = ms.item() / args.steps
    launches = gen.launches - launches0
    value = world * pix_local / 1e6 / (ms_per_step / 1e3)

    # ---- roofline of the dominant kernel (K1): algorithmic bytes = 8*sum(C_s) + 9 per pixel ----
    peak, peak_src = hbm_peak()
    k1_bytes = pix_local * (8 * sum(c for _, c in SOURCES) + 9)
    achieved = k1_bytes / 1e9 / (k1_ms.item() / 1e3)
    traffic = ncu_traffic()
    roofline = {"bound": "hbm", "kernel": "fuse_sources_kernel (K1)", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "peak_source": peak_src, "ms_per_launch": round(k1_ms.item(), 4),
                "algorithmic_bytes_per_launch": k1_bytes, "traffic": traffic["bytes_per_launch"] if traffic else None,
                "traffic_note": traffic.get("note") if traffic else "no ncu capture committed yet",
                "whole_step_frac_319B_per_pix": round(world * pix_local * 319 / 1e9 / (ms_per_step / 1e3) / (peak * world), 4),
                "variant": lib.mspl_fuse_variant().decode()}

    # ---- end to end through the public host-buffer API -----------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        ne = min(args.e2e_images, n)
        hm, ha = make_logits_host(torch, ne, h, w, seed=7 + rank, pin=True)
        out_host = torch.empty((ne, h, w), dtype=torch.uint8, pin_memory=True)
        gen.run_from_host(hm, ha, dev, out_host=out_host)           # warm-up (allocations, first-touch)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.e2e_steps):
            gen.run_from_host(hm, ha, dev, out_host=out_host)
        e1.record()
        barrier()
        ems = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        ems_step = ems.item() / args.e2e_steps
        e2e = {"value": round(world * ne * h * w / 1e6 / (ems_step / 1e3), 1), "unit": "Mpix/s",
               "h2d_bytes_per_step": ne * h * w * 8 * sum(c for _, c in SOURCES), "d2h_bytes_per_step": ne * h * w,
               "images_per_step_per_gpu": ne, "ms_per_step": round(ems_step, 3),
               "api": "mspl_b200.pipeline.LabelGenerator.run_from_host (pinned host logits -> uint8 label maps on host)"}
        del hm, ha

    # ---- size-independent properties at the full benchmark size (the CPU oracle cannot run 2,000 images) -------------
    checks = full_size_checks(torch, job, pix_local, world, dist if world > 1 else None, args.portion)

    if rank == 0:
        cpu = None if args.no_cpu else cpu_baseline(torch, args, args.cpu_seconds)
        kept = job.kept.tolist() if job.kept is not None else None
        line = {
            "metric": "pseudo-labelled Mpix/s (3-source fusion)", "value": round(value, 1), "unit": "Mpix/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, n),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "checks": checks,
            "results": {"near_tie_pixels": int(job.marginal.item()) if job.marginal is not None else None,
                        "class_hist": job.class_hist.tolist(), "final_hist": job.final_hist.tolist(),
                        # the never-selected ignore class has threshold +inf: emit() writes it as null (strict JSON has no Infinity)
                        "thresholds": [round(x, 6) for x in job.thresh.tolist()] if job.thresh is not None else None, "kept": kept},
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
