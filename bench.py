#!/usr/bin/env python
"""bench.py -- headline benchmark of the grayskull hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5|ops|match] [--impl reference]

Default workload = BASELINE.json configs[1] ("c2"): gs_blur r=5 + gs_sobel on 4096x4096 synthetic
uint8 frames, batch 256 per GPU (weak scaling: every rank processes its own 256 frames; frames are
independent, no data-path collective).  A step = one pass of both ops over the batch.

One JSON line on stdout (rank 0):
  value      Mpixels/s, whole job, inputs resident in HBM, CUDA-event time, max over ranks
  e2e        the same metric through the C ABI with HOST (pinned) buffers: H2D + kernels + D2H in
             the timed region, chunked over two streams
  roofline   dominant kernel's algorithmic HBM bytes / its own CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline  the reference's own C code (oracle/_ref, built from /root/reference) timed on this
             box's host cores on a bounded sample of the same workload
With --impl reference the whole line is the reference CPU arm (rank 0 only).
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

W2, H2, B2, R2 = 4096, 4096, 256, 5            # c2
W3, H3, B3, NK3, T3 = 1920, 1080, 1024, 1250, 20   # c3
W4, H4, B4 = 3840, 2160, 256                  # c4 (sf 1.1, scales 1..4, step 2, max_rects 65536)
FALLBACK_HBM_GBS = 6650.0


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


# ------------------------------------------------------------------ reference CPU arm
def _cpu_worker(args):
    """one process: run the real reference (or the oracle port) on `nframes` frames"""
    kind, workload, seed, nframes = args
    import numpy as np
    import _libs as L
    units = 0
    busy = 0.0
    if kind == "reference":
        R = L.ref()
    else:
        O = L.oracle()
    for i in range(nframes):
        rng = np.random.default_rng(seed * 1000 + i)
        if workload == "c2":
            a = rng.integers(0, 256, (H2, W2), dtype=np.uint8)
            b = np.empty_like(a); s = np.zeros_like(a)
            t0 = time.perf_counter()
            if kind == "reference":
                R.gs_blur(L.img(b), L.img(a), R2); R.gs_sobel(L.img(s), L.img(b))
            else:
                O.gso_blur(L.ptr(b), L.ptr(a), W2, H2, R2); O.gso_sobel(L.ptr(s), L.ptr(b), W2, H2)
            units += W2 * H2
        elif workload == "c3":
            a = L.natural_like(W3, H3, seed * 1000 + i)
            k = np.zeros(NK3, L.KP_DTYPE); sm = np.zeros_like(a)
            t0 = time.perf_counter()
            if kind == "reference":
                R.gs_orb_extract(L.img(a), L.ptr(k), NK3, T3, L.ptr(sm))
            else:
                O.gso_orb_extract(L.ptr(a), W3, H3, L.ptr(k), NK3, T3, L.ptr(sm))
            units += W3 * H3
        else:  # c4
            a = L.natural_like(W4, H4, seed * 1000 + i)
            ii = np.empty(a.shape, np.uint32); r = np.zeros(65536, L.RECT_DTYPE)
            cas = None if kind == "reference" else L.HostCascade()
            t0 = time.perf_counter()
            if kind == "reference":
                R.gs_integral(L.img(a), L.ptr(ii))
                R.gs_lbp_detect(R.ref_frontalface(), L.ptr(ii), W4, H4, L.ptr(r), 65536, 1.1, 1.0, 4.0, 2)
            else:
                O.gso_integral(L.ptr(a), W4, H4, L.ptr(ii))
                O.gso_lbp_detect(cas.ptr, L.ptr(ii), W4, H4, L.ptr(r), 65536, 1.1, 1.0, 4.0, 2)
            units += 30016520
        busy += time.perf_counter() - t0
    return units, busy


def cpu_reference(workload, steps=1, warmup=0, frames_per_core=1, max_cores=None, budget_s=150.0):
    """Frame-parallel over the host cores (one process per core: gs_orb_extract's static buffer is
    not thread-safe, reference grayskull.h:655).  Returns (value per second, dict)."""
    import _libs as L
    kind = "reference" if L.have_ref() else "port"
    if kind == "port":
        L.oracle()
    cores = len(os.sched_getaffinity(0))
    if max_cores:
        cores = min(cores, max_cores)
    ctx = mp.get_context("fork")
    per_step = []
    t_begin = time.perf_counter()
    with ctx.Pool(cores) as pool:
        for s in range(warmup + steps):
            res = pool.map(_cpu_worker, [(kind, workload, s * 64 + c, frames_per_core) for c in range(cores)])
            if s >= warmup:   # all cores run concurrently: the step takes as long as the slowest one
                per_step.append((sum(u for u, _ in res), max(t for _, t in res)))
            # keep the whole arm within a few minutes: stop early once another step would not fit
            spent = time.perf_counter() - t_begin
            if per_step and spent + spent / (s + 1) > budget_s:
                break
    units = sum(u for u, _ in per_step)
    secs = sum(t for _, t in per_step)
    unit_name = "Mpixels/s" if workload in ("c2", "c3") else "windows/s"
    scale = 1e-6 if workload in ("c2", "c3") else 1.0
    sample = {"c2": "%d frames of 4096x4096 per step, gs_blur r=5 + gs_sobel" % (cores * frames_per_core),
              "c3": "%d frames of 1920x1080 per step, gs_orb_extract nkps=1250 t=20" % (cores * frames_per_core),
              "c4": "%d frames of 3840x2160 per step, gs_integral + gs_lbp_detect" % (cores * frames_per_core)}[workload]
    return units / secs * scale, {"value": units / secs * scale, "unit": unit_name, "cores": cores, "kind": kind,
                                 "sample": sample + " (gcc -std=c99 -O2, one process per core)",
                                 "ms_per_step": 1e3 * secs / max(len(per_step), 1), "steps_run": len(per_step)}


# ------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None
        t0 = time.time()
        while self.p is not None and time.time() - t0 < 3.0:     # first sample lands before timing starts
            if os.path.getsize(self.f.name) > 0:
                break
            time.sleep(0.01)

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, pw = [], 0, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                clk, cmax, power, util = float(parts[0]), float(parts[1]), float(parts[2]), float(parts[3])
            except ValueError:
                continue
            mx = max(mx, cmax)
            sm.append(clk); pw.append(power)
            for nme, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        os.unlink(self.f.name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "power_w": statistics.median(pw) if pw else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------ GPU arm
def gpu_main(args):
    import torch
    import torch.distributed as dist
    import grayskull_b200 as g
    from grayskull_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    lib = g.lib()
    g._lib.check(lib.gs_b200_set_device(local), "set_device")
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL may print a version banner on stdout (NCCL_DEBUG=VERSION): keep stdout for the ONE json line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()            # forces communicator creation now
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def time_steps(fn, steps, warmup, sample_clocks=False):
        """W warm-ups, then K steps between barrier+sync, CUDA events on the launching stream"""
        for _ in range(warmup):
            fn()
        barrier()
        sampler = ClockSampler(local) if (sample_clocks and rank == 0) else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        if sampler:
            clock_box.append(sampler.stop())
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    hbm, peak_kind = peaks()
    clock_box = []
    wl = args.workload
    torch.manual_seed(1234 + rank)
    extra = {}

    if wl == "c2":
        n, h, w = args.batch or B2, H2, W2
        src = torch.randint(0, 256, (n, h, w), dtype=torch.uint8, device=dev)
        blur = torch.empty_like(src)
        sob = torch.zeros_like(src)

        def step():
            api.blur_batch(src, R2, out=blur)
            api.sobel_batch(blur, out=sob)

        units_per_step = n * h * w
        unit, scale, metric = "Mpixels/s", 1e-6, "Mpixels/s, gs_blur(r=5) + gs_sobel, 4096x4096 uint8"
        cfg = {"workload": "c2: gs_blur r=5 + gs_sobel, 4096x4096 synthetic uint8, batch %d per GPU" % n,
               "frames_per_gpu": n, "l2": "inputs (%.1f GiB per GPU) exceed the 126 MB L2" % (n * h * w / 2**30)}
        kernels = {"gs_blur_r5": (lambda: api.blur_batch(src, R2, out=blur), 2.0 * n * h * w),
                   "gs_sobel": (lambda: api.sobel_batch(blur, out=sob), 1.0 * n * h * w + 1.0 * n * (h - 2) * (w - 2))}
        launches_per_step = 2
    elif wl == "c3":
        n, h, w = args.batch or B3, H3, W3
        noise = torch.randint(0, 256, (n, h, w), dtype=torch.uint8, device=dev)
        src = api.blur_batch(noise, 3)      # "blurred noise": natural-image-like autocorrelation
        del noise
        sm = torch.zeros_like(src)

        def step():
            api.orb_extract_batch(src, NK3, T3, scoremap=sm)

        units_per_step = n * h * w
        unit, scale, metric = "Mpixels/s", 1e-6, "Mpixels/s, gs_orb_extract (FAST-9 t=20 + BRIEF-256, nkps=1250), 1920x1080 uint8"
        cfg = {"workload": "c3: gs_orb_extract nkps=1250 t=20, 1920x1080 blurred-noise uint8, batch %d per GPU" % n,
               "frames_per_gpu": n, "l2": "inputs (%.1f GiB per GPU) exceed the 126 MB L2" % (n * h * w / 2**30)}
        kernels = {"gs_orb_extract": (step, 2.0 * n * h * w)}
        launches_per_step = 8
    elif wl == "c4":
        n, h, w = args.batch or B4, H4, W4
        cas = g.load_cascade()
        noise = torch.randint(0, 256, (n, h, w), dtype=torch.uint8, device=dev)
        src = api.blur_batch(noise, 3)
        del noise
        ii = torch.empty((n, h, w), dtype=torch.int32, device=dev)
        nwin = api.lbp_window_count(cas, w, h, 1.1, 1.0, 4.0, 2)

        def step():
            api.integral_batch(src, out=ii)
            api.lbp_detect_batch(cas, ii, 65536, 1.1, 1.0, 4.0, 2)

        units_per_step = n * nwin
        unit, scale, metric = "windows/s", 1.0, "LBP cascade windows/s, gs_integral + gs_lbp_detect frontalface, 3840x2160"
        cfg = {"workload": "c4: gs_integral + gs_lbp_detect frontalface sf=1.1 scales 1..4 step=2, 3840x2160, batch %d per GPU" % n,
               "frames_per_gpu": n, "windows_per_frame": nwin,
               "l2": "integral tables (%.1f GiB per GPU) exceed the 126 MB L2" % (n * h * w * 4 / 2**30)}
        kernels = {"gs_integral": (lambda: api.integral_batch(src, out=ii), 5.0 * n * h * w),
                   "gs_lbp_detect": (lambda: api.lbp_detect_batch(cas, ii, 65536, 1.1, 1.0, 4.0, 2), 4.0 * n * h * w)}
        launches_per_step = 5
    elif wl == "c5":
        # BASELINE configs[4]: blur -> sobel -> FAST/ORB -> integral + LBP on 1920x1080 frames, sharded by frame
        # (each rank owns its own frames; 8192 frames over 8 GPUs = 1024 per GPU).  FAST and LBP both run on
        # the sobel output.
        n, h, w = args.batch or 256, H3, W3
        cas = g.load_cascade()
        src = torch.randint(0, 256, (n, h, w), dtype=torch.uint8, device=dev)
        blur = torch.empty_like(src)
        sob = torch.zeros_like(src)
        sm = torch.zeros_like(src)
        ii = torch.empty((n, h, w), dtype=torch.int32, device=dev)

        def step():
            api.blur_batch(src, R2, out=blur)
            api.sobel_batch(blur, out=sob)
            api.orb_extract_batch(sob, NK3, T3, scoremap=sm)
            api.integral_batch(sob, out=ii)
            api.lbp_detect_batch(cas, ii, 65536, 1.1, 1.0, 4.0, 2)

        units_per_step = n * h * w
        unit, scale, metric = "Mpixels/s", 1e-6, "Mpixels/s, pipeline blur(r=5) -> sobel -> gs_orb_extract -> gs_integral + gs_lbp_detect, 1920x1080 uint8"
        cfg = {"workload": "c5: blur r=5 -> sobel -> orb_extract(nkps=1250,t=20) -> integral + lbp_detect(sf 1.1, scales 1..4, step 2), 1920x1080, %d frames per GPU" % n,
               "frames_per_gpu": n, "l2": "inputs (%.1f GiB per GPU) exceed the 126 MB L2" % (n * h * w / 2**30)}
        kernels = {"gs_blur_r5": (lambda: api.blur_batch(src, R2, out=blur), 2.0 * n * h * w),
                   "gs_sobel": (lambda: api.sobel_batch(blur, out=sob), 2.0 * n * h * w),
                   "gs_orb_extract": (lambda: api.orb_extract_batch(sob, NK3, T3, scoremap=sm), 2.0 * n * h * w),
                   "gs_integral": (lambda: api.integral_batch(sob, out=ii), 5.0 * n * h * w),
                   "gs_lbp_detect": (lambda: api.lbp_detect_batch(cas, ii, 65536, 1.1, 1.0, 4.0, 2), 4.0 * n * h * w)}
        launches_per_step = 14
    elif wl == "match":
        # gs_match_orb over frame pairs (1250 x 1250 descriptors each, the c3 keypoint budget).
        # Not HBM-bound: 8 POPC per descriptor comparison on the 16-lane/clk/SM POPC path bounds it.
        npairs, nd = args.batch or 2048, NK3
        gen = torch.Generator(device=dev); gen.manual_seed(77 + rank)
        k2 = torch.randint(-2**31, 2**31 - 1, (npairs, nd, 12), dtype=torch.int32, device=dev, generator=gen)
        k1 = k2[:, torch.randperm(nd, device=dev)].clone()
        flip = (torch.rand((npairs, nd, 8), device=dev, generator=gen) < 0.6).to(torch.int32) << torch.randint(0, 31, (npairs, nd, 8), device=dev, generator=gen).to(torch.int32)
        k1[:, :, 4:] ^= flip                     # near-duplicates: ~5 flipped bits per descriptor
        k1[:, nd // 2:, 4:] = torch.randint(-2**31, 2**31 - 1, (npairs, nd - nd // 2, 8), dtype=torch.int32, device=dev, generator=gen)
        del flip
        cnt = torch.full((npairs,), nd, dtype=torch.int32, device=dev)
        src = k1
        n, h, w = npairs, nd, nd

        def step():
            api.match_orb_batch(k1, cnt, k2, cnt, nd, 60.0)

        units_per_step = npairs * nd * nd
        unit, scale, metric = "Gcomparisons/s", 1e-9, "256-bit descriptor comparisons/s, gs_match_orb 1250 x 1250 per frame pair"
        cfg = {"workload": "match: gs_match_orb max_distance=60, 1250 x 1250 descriptors per pair, %d pairs per GPU" % npairs,
               "pairs_per_gpu": npairs, "l2": "descriptor sets (%.0f MB per GPU) exceed the 126 MB L2" % (2 * npairs * nd * 48 / 1e6)}
        kernels = {"gs_match_orb": (step, 2.0 * npairs * nd * 48 + 12.0 * npairs * nd)}
        launches_per_step = 2
    elif wl == "tmatch":
        # gs_match_template, one 32x32 template against 1920x1080 frames (compute-bound:
        # 11 instructions per 16 squared differences)
        n, h, w = args.batch or 16, H3, W3
        tw = th = 32
        noise = torch.randint(0, 256, (n, h, w), dtype=torch.uint8, device=dev)
        src = api.blur_batch(noise, 2)
        del noise
        tmpl = src[0, 500:500 + th, 900:900 + tw].contiguous()
        res = torch.empty((n, h - th + 1, w - tw + 1), dtype=torch.uint8, device=dev)

        def step():
            api.match_template_batch(src, tmpl, out=res)
            api.find_best_match_batch(res)

        units_per_step = n * (h - th + 1) * (w - tw + 1) * tw * th
        unit, scale, metric = "Gtaps/s", 1e-9, "squared differences/s, gs_match_template 32x32 template, 1920x1080 uint8"
        cfg = {"workload": "tmatch: gs_match_template + gs_find_best_match, 32x32 template, 1920x1080, batch %d per GPU" % n,
               "frames_per_gpu": n, "l2": "compute-bound; frames (%.0f MB per GPU) stream through L2" % (n * h * w / 1e6)}
        kernels = {"gs_match_template": (lambda: api.match_template_batch(src, tmpl, out=res), 2.0 * n * h * w)}
        launches_per_step = 3
    elif wl == "ops":
        # per-op table (every stencil / resampling op of the path at 4096x4096), not a driver line
        n, h, w = args.batch or 64, H2, W2
        src = torch.randint(0, 256, (n, h, w), dtype=torch.uint8, device=dev)
        out = torch.zeros_like(src)
        half = torch.empty((n, h // 2, w // 2), dtype=torch.uint8, device=dev)
        ii = torch.empty((n, h, w), dtype=torch.int32, device=dev)
        hist = torch.empty((n, 256), dtype=torch.int32, device=dev)
        import numpy as np
        K_SHARPEN = np.array([[0, -1, 0], [-1, 5, -1], [0, -1, 0]], np.int8)
        K_GAUSS = np.array([[1, 2, 1], [2, 4, 2], [1, 2, 1]], np.int8)
        oth = torch.empty((n,), dtype=torch.uint8, device=dev)
        px = float(n * h * w)
        kernels = {
            "gs_sobel": (lambda: api.sobel_batch(src, out=out), 2.0 * px),
            "gs_erode": (lambda: api.erode_batch(src, out=out), 2.0 * px),
            "gs_dilate": (lambda: api.dilate_batch(src, out=out), 2.0 * px),
            "gs_blur_r1": (lambda: api.blur_batch(src, 1, out=out), 2.0 * px),
            "gs_blur_r3": (lambda: api.blur_batch(src, 3, out=out), 2.0 * px),
            "gs_blur_r5": (lambda: api.blur_batch(src, 5, out=out), 2.0 * px),
            "gs_blur_r7": (lambda: api.blur_batch(src, 7, out=out), 2.0 * px),
            "gs_adaptive_threshold_r5": (lambda: api.adaptive_threshold_batch(src, 5, 2, out=out), 2.0 * px),
            "gs_downsample": (lambda: api.downsample_batch(src, out=half), 1.25 * px),
            "gs_resize_to_half": (lambda: api.resize_batch(src, w // 2, h // 2, out=half), 1.25 * px),
            "gs_integral": (lambda: api.integral_batch(src, out=ii), 5.0 * px),
            "gs_filter_sharpen": (lambda: api.filter_batch(src, K_SHARPEN, 1, out=out), 2.0 * px),
            "gs_filter_gaussian": (lambda: api.filter_batch(src, K_GAUSS, 16, out=out), 2.0 * px),
            "gs_histogram": (lambda: api.histogram_batch(src, out=hist), 1.0 * px),
            "gs_otsu_threshold": (lambda: api.otsu_threshold_batch(src, hist=hist, out=oth), 1.0 * px),
            "gs_threshold": (lambda: api.threshold_batch(out, 128), 2.0 * px),
        }

        def step():
            for fn, _ in kernels.values():
                fn()

        units_per_step = n * h * w * len(kernels)
        unit, scale, metric = "Mpixels/s", 1e-6, "Mpixels/s summed over the per-op table, 4096x4096 uint8"
        cfg = {"workload": "ops: every stencil/resampling op once, 4096x4096 synthetic uint8, batch %d per GPU" % n,
               "frames_per_gpu": n, "l2": "inputs (%.1f GiB per GPU) exceed the 126 MB L2" % (n * h * w / 2**30)}
        launches_per_step = len(kernels) + 1
    else:
        raise SystemExit("unknown workload " + wl)

    # ---- headline: device-resident steps ----
    l0 = lib.gs_b200_launch_count()
    ms = time_steps(step, args.steps, args.warmup, sample_clocks=True)
    launches = (lib.gs_b200_launch_count() - l0) * args.steps // (args.steps + args.warmup)
    clocks = clock_box[0] if clock_box else None
    value = units_per_step * world * args.steps / (ms * 1e-3) * scale

    # ---- per-kernel CUDA-event timing (roofline) ----
    kres = {}
    for name, (fn, algo_bytes) in kernels.items():
        ksteps = max(min(args.steps, 100) // 2, 5)
        kms = time_steps(fn, ksteps, 3) / ksteps
        kres[name] = {"ms": kms, "algorithmic_bytes": algo_bytes, "achieved_gbs": algo_bytes / (kms * 1e-3) / 1e9,
                      "frac": algo_bytes / (kms * 1e-3) / 1e9 / hbm}
    dom = max(kres, key=lambda k: kres[k]["ms"])
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f).get(wl, {}).get(dom)
            # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture, per frame,
            # scaled to this launch's frame count (profiles/r01_ncu_c2_summary.txt)
            traffic = t["dram_bytes_per_frame"] * n if t else None
    except Exception:
        pass
    roofline = {"kernel": dom, "bound": "hbm", "achieved": kres[dom]["achieved_gbs"], "peak": hbm,
                "peak_source": peak_kind + (" (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else " (B200_PROFILING.md)"),
                "unit": "GB/s", "frac": kres[dom]["frac"], "traffic": traffic,
                "algorithmic_bytes_per_launch": kres[dom]["algorithmic_bytes"], "ms_per_launch": kres[dom]["ms"]}

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    e2e = None
    if wl == "c2" and not args.no_e2e:
        ne = min(n, args.e2e_frames)
        chunk, nstreams = args.e2e_chunk, args.e2e_streams
        hin = torch.empty((ne, h, w), dtype=torch.uint8).pin_memory()
        hout = torch.empty((ne, h, w), dtype=torch.uint8).pin_memory()
        hin.copy_(src[:ne].cpu())
        streams = [torch.cuda.Stream(device=dev) for _ in range(nstreams)]
        dbuf = [(torch.empty((chunk, h, w), dtype=torch.uint8, device=dev), torch.empty((chunk, h, w), dtype=torch.uint8, device=dev),
                 torch.zeros((chunk, h, w), dtype=torch.uint8, device=dev)) for _ in range(nstreams)]
        fb = chunk * h * w

        def e2e_step():
            for ci, lo in enumerate(range(0, ne, chunk)):
                st = streams[ci % nstreams]
                a, b, c = dbuf[ci % nstreams]
                sp = C.c_void_p(st.cuda_stream)
                k = min(chunk, ne - lo)
                g._lib.check(lib.gs_b200_memcpy_h2d(C.c_void_p(a.data_ptr()), C.c_void_p(hin[lo].data_ptr()), k * h * w, sp))
                g._lib.check(lib.gs_b200_blur_batch(C.c_void_p(b.data_ptr()), C.c_void_p(a.data_ptr()), w, h, k, R2, sp))
                g._lib.check(lib.gs_b200_sobel_batch(C.c_void_p(c.data_ptr()), C.c_void_p(b.data_ptr()), w, h, k, sp))
                g._lib.check(lib.gs_b200_memcpy_d2h(C.c_void_p(hout[lo].data_ptr()), C.c_void_p(c.data_ptr()), k * h * w, sp))
            for st in streams:
                st.synchronize()

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        ksteps = max(3, min(args.steps, 100) // 10)
        for _ in range(ksteps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": ne * h * w * world * ksteps / dt * scale, "unit": unit, "h2d_bytes_per_step": ne * h * w,
               "d2h_bytes_per_step": ne * h * w, "frames_per_step": ne, "steps": ksteps,
               "path": "pinned host -> gs_b200_memcpy_h2d -> gs_b200_blur_batch -> gs_b200_sobel_batch -> gs_b200_memcpy_d2h, %d-frame chunks on %d streams" % (chunk, nstreams)}
        del hin, hout, dbuf

    # ---- CPU baseline (rank 0, N == 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and wl in ("c2", "c3", "c4"):
        _, cpu = cpu_reference(wl, steps=1, warmup=0, max_cores=args.cpu_cores or None)
    if rank == 0 and world == 1 and not args.no_cpu and wl == "match":
        import numpy as np
        import _libs as L
        kind = "reference" if L.have_ref() else "port"
        fn = L.ref().gs_match_orb if kind == "reference" else L.oracle().gso_match_orb
        a = api.kps_to_numpy(k1[:1], cnt[:1])[0]; b = api.kps_to_numpy(k2[:1], cnt[:1])[0]
        m = np.zeros(nd, L.MATCH_DTYPE)
        t0 = time.perf_counter(); reps = 0
        while time.perf_counter() - t0 < 5.0:
            fn(L.ptr(a), nd, L.ptr(b), nd, L.ptr(m), nd, 60.0); reps += 1
        dt = time.perf_counter() - t0
        cpu = {"value": reps * nd * nd / dt * scale, "unit": unit, "cores": 1, "kind": kind,
               "sample": "%d calls of gs_match_orb on one 1250 x 1250 pair (gcc -std=c99 -O2, single thread)" % reps}

    if rank == 0:
        out = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": {"c4": "u32", "match": "u32 (xor + popc)"}.get(wl, "u8"), "data": "synthetic (uniform iid u8, torch.randint; c3/c4: gs_blur r=3 of it)", "config": cfg, "clocks": clocks,
               "e2e": e2e, "gpu_launches": int(launches), "launches_per_step": int(launches) // max(args.steps, 1),
               "roofline": roofline, "kernels": kres, "cpu_baseline": cpu,
               "tma_path": bool(lib.gs_b200_uses_tma(w, h, src.data_ptr()))}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    value, cpu = cpu_reference(wl, steps=max(1, min(args.steps, 3)), warmup=0, budget_s=150.0)
    unit = cpu["unit"]
    metric = {"c2": "Mpixels/s, gs_blur(r=5) + gs_sobel, 4096x4096 uint8",
              "c3": "Mpixels/s, gs_orb_extract (FAST-9 t=20 + BRIEF-256, nkps=1250), 1920x1080 uint8",
              "c4": "LBP cascade windows/s, gs_integral + gs_lbp_detect frontalface, 3840x2160"}[wl]
    out = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
           "steps": cpu["steps_run"], "warmup": 0, "ms_per_step": cpu["ms_per_step"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8" if wl != "c4" else "u32",
           "data": "synthetic", "config": {"workload": cpu["sample"]}, "cpu_baseline": cpu,
           "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5", "ops", "match", "tmatch"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="frames per GPU (default: the BASELINE config's batch)")
    ap.add_argument("--e2e-frames", type=int, default=64)
    ap.add_argument("--e2e-chunk", type=int, default=16)
    ap.add_argument("--e2e-streams", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-cores", type=int, default=0, help="cap the cores of the cpu_baseline sample (default: all)")
    a = ap.parse_args()
    if a.impl == "reference":
        reference_main(a)
    else:
        gpu_main(a)
