#!/usr/bin/env python
"""Benchmark of the hot path: 1080p images/s through the CUDA SIFT pipeline (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step is one pass of the whole path (pyramid -> extrema -> elimination -> order replay -> orientation
-> descriptors) over one batch of synthetic 1920x1080 frames (5 octaves, 3 DoGs/octave, sigma 1.6,
k sqrt(2), no upsample: BASELINE.json configs[2]/[4]) per GPU.  `value` is measured with the frames
already resident in HBM; `e2e` goes through the same C-ABI call with pinned HOST buffers, so every
step pays the host->device copy of its frames and the device->host copy of keypoints + descriptors.
One process per GPU (torchrun for N > 1); images are independent, so ranks share nothing on the data
path (weak scaling) and torch.distributed only provides the barrier and the max-over-ranks reduction.

--impl reference times the reference's OWN sift.cpp + algorithms.cpp (oracle/_ref/libref_fast.so: compiled unmodified
against Vigra stand-in headers whose arrays are copy-on-write and whose blur is memoised, so that the reference's
O(pixels) copies per candidate and full-image blur per keypoint do not make a 1080p frame take the better part of an
hour) on all host cores, same metric and config; the restatement (oracle/) stands in only if that library is missing.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, OCTAVES, DPE = 1920, 1080, 5, 3
WORKLOAD = "1920x1080 synthetic frames (blobs+corners+-2 noise, seed=frame index), 5 octaves, 3 DoGs/octave, sigma 1.6, k sqrt(2), subpixel 0"


def pyramid_bytes(w, h, octaves, dpe, subpixel):
    """Algorithmic bytes of the pyramid + DoG stage, SURVEY.md §8(d): every image read/written once where
    semantically required, fp32."""
    in_px = w * h
    if subpixel:
        w, h = 2 * w, 2 * h
    px = []
    for _ in range(octaves):
        px.append(w * h)
        w, h = (w + 1) // 2, (h + 1) // 2
    a = 4 * (in_px + in_px) + 4 * (in_px + px[0]) if subpixel else 4 * (px[0] + px[0])
    b = sum(4 * p * (3 * dpe - 1) for p in px)
    c = sum(4 * (px[o] + px[o + 1]) for o in range(octaves - 1))
    return a + b + c


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index, self.t_begin = [], None, index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        """Start of the timed region: nvidia-smi is started before the warm-up (its first sample takes ~1 s), samples
        taken during the warm-up are only used if the timed region turns out shorter than one sampling period."""
        self.t_begin = time.perf_counter()

    def stop(self):
        if self.proc:
            self.proc.terminate()
        t_end = time.perf_counter()
        inside = [r for t, r in self.rows if self.t_begin is None or self.t_begin <= t <= t_end + 0.15]
        window = "timed region"
        if not inside:
            inside, window = [r for _, r in self.rows[-5:]], "warm-up just before the timed region (region shorter than one sample)"
        sm, mx, reasons = [], [], set()
        for r in inside:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def _cpu_impl():
    """(ctypes library, kind, description) of the CPU implementation of the path: the reference's own sources when oracle/_ref
    was built (it travels to the GPU box prebuilt), else the restatement."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol

    if ol.ref_available(fast=True):
        return ol, ol.ref_lib(fast=True), "reference", ("reference sift.cpp + algorithms.cpp compiled unmodified against the Vigra stand-in headers "
                                                         "(copy-on-write arrays, memoised blur: identical results, linear-time copies)")
    return ol, None, "port", "oracle restatement, 'hoisted' flavour (reference results without its O(pixels^2) copies)"


def run_reference(args):
    """CPU arm: the reference's own implementation of the path (see _cpu_impl) on all host cores; one step = one 1080p frame per
    host thread."""
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np

    from sift_b200.synth import synth_frame

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ol, L, kind, what = _cpu_impl()
    cores = len(os.sched_getaffinity(0)) or 1
    frames = [synth_frame(W, H, i) for i in range(min(cores, 16))]
    k = float(np.float32(np.sqrt(2.0)))
    oracles = [ol.Oracle(DPE, OCTAVES, 1.6, k, False, L=L) for _ in range(cores)]

    def step():
        def job(t):
            oracles[t].calculate(frames[t % len(frames)])
        with ThreadPoolExecutor(cores) as ex:
            list(ex.map(job, range(cores)))
        return cores

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    n = 0
    for _ in range(args.steps):
        n += step()
    dt = time.perf_counter() - t0
    v = n / dt
    line = {"impl": "reference", "metric": "1080p images/sec end-to-end SIFT", "value": v, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": cores},
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": kind,
                             "sample": f"{cores} 1080p frames per step, one per host thread; {what}"},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(seconds_budget=20.0):
    """Single-thread CPU baseline next to the GPU number: the reference's own sources (kind 'reference', see _cpu_impl) on
    1080p frames, bounded to about `seconds_budget` of CPU work."""
    import numpy as np

    from sift_b200.synth import synth_frame

    ol, L, kind, what = _cpu_impl()
    o = ol.Oracle(DPE, OCTAVES, 1.6, float(np.float32(np.sqrt(2.0))), False, L=L)
    n, t = 0, 0.0
    while t < seconds_budget and n < 16:
        dt, _ = o.time_calculate(synth_frame(W, H, n))
        t += dt
        n += 1
    return {"value": n / t, "unit": "images/s", "cores": 1, "kind": kind, "sample": f"{n} synthetic 1080p frames, 1 thread; {what}"}


def measure_configs(device, flags, literal_cpu=True):
    """BASELINE.json configs 1, 2 and 4 (config 3/5 are the main line): one image per call on one GPU — latency from a host frame to
    keypoints + descriptors in host memory, the pyramid stage's time and its share of the HBM roofline — and, for the sizes the
    reference's CPU path can finish (north_star: parrot, 600x600; 1500x1500 takes 9-11 minutes and is run with --literal-1500),
    the LITERAL reference (oracle/_ref/libref.so: every copy deep, no memoised blur) on one host thread, next to the
    README's own numbers (reference README.md:68-71: ~300x300 0.7 s, ~600x600 15 s, ~1500x1500 11 min)."""
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    from sift_b200 import capi
    from sift_b200.synth import synth_frame

    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    k = capi.SQRT2_F32
    parrot = np.load(os.path.join(ROOT, "tests", "golden", "parrot_r.npy")).astype(np.float32)
    cases = [("config1_parrot_488x600", parrot, 4, False, DPE, 1.6, k),
             ("config2_600x600_subpixel", synth_frame(600, 600, 0), 4, True, DPE, 1.6, k),
             ("config4_3840x2160_subpixel_6oct", synth_frame(3840, 2160, 0), 6, True, DPE, 1.6, k),
             # a schedule off the defaults (4 DoGs per octave, sigma 2.0, k 1.3: radii 6, 8, 10, 13, 17, ... instead of 5, 7, 10, 14, 19):
             # every level still runs on the sliding kernel (next instantiated radius up, zero taps added)
             ("nondefault_1080p_dpe4_sigma2.0_k1.3", synth_frame(1920, 1080, 0), 4, False, 4, 2.0, 1.3)]
    out = {}
    for name, img, octaves, sub, dpe, sigma, kk in cases:
        h, w = img.shape
        g = capi.SiftGpu(dpe, octaves, sigma, kk, sub, max_width=w, max_height=h, max_batch=1, device=device, flags=flags | capi.FLAG_SERIAL)
        lat, pyr, n = [], [], 0
        for i in range(7):
            t0 = time.perf_counter()
            r = g.run([img], raise_on_error=False)[0]
            lat.append(1e3 * (time.perf_counter() - t0))
            pyr.append(g.timings()["pyramid_ms"])
            n = int(r["kps"].size)
        g.close()
        lat, pyr = sorted(lat[2:]), sorted(pyr[2:])
        pb = pyramid_bytes(w, h, octaves, dpe, sub)
        out[name] = {"gpu_latency_ms": lat[len(lat) // 2], "pyramid_ms": pyr[len(pyr) // 2], "pyramid_algorithmic_bytes": pb,
                     "pyramid_gbs": pb / (pyr[len(pyr) // 2] * 1e-3) / 1e9, "pyramid_frac_of_hbm_peak": pb / (pyr[len(pyr) // 2] * 1e-3) / 1e9 / peak,
                     "keypoints": n, "octaves": octaves, "subpixel": sub, "dogs_per_octave": dpe, "sigma": sigma, "k": float(kk)}
    if literal_cpu and ol.ref_available(fast=False):
        L = ol.ref_lib(fast=False)
        for name, img, octaves, readme in (("literal_cpu_300x300", synth_frame(300, 300, 0), 4, "~0.7 s"),
                                            ("literal_cpu_parrot_488x600", parrot, 4, None),
                                            ("literal_cpu_600x600", synth_frame(600, 600, 0), 4, "~15 s")):
            o = ol.Oracle(DPE, octaves, 1.6, float(np.float32(np.sqrt(2.0))), False, L=L)
            dt, nk = o.time_calculate(img)
            out[name] = {"seconds": dt, "keypoints": nk, "threads": 1, "readme_says": readme,
                         "what": "reference sift.cpp + algorithms.cpp compiled unmodified (deep copies, no memoisation), CLI defaults"}
        out["literal_cpu_1500x1500"] = {"seconds": None, "readme_says": "~11 min",
                                        "note": "526 s measured in the build container (BASELINE.md); bench.py --literal-1500 repeats it"}
    return out


def bind_to_gpu_cpus(index):
    """Multi-GPU host: keep this rank's threads and its pinned buffers on the CPUs (NUMA node) next to its GPU, so the frame
    uploads do not cross the socket interconnect.  Best effort; returns the number of CPUs bound to, or None."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        near = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        mine = near & os.sched_getaffinity(0)
        if mine:
            os.sched_setaffinity(0, mine)
            return len(mine)
    except Exception:
        pass
    return None


def run_ours(args):
    import numpy as np
    import torch

    from sift_b200 import capi, shard
    from sift_b200.synth import synth_frame

    rank, local_rank, world = shard.world()
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_cpus(local_rank) if world > 1 else None
    dev = torch.device("cuda", local_rank)
    dist = shard.init_process_group("nccl") if world > 1 else None

    B, n_distinct = args.batch, args.frames
    # weak scaling: every rank owns its own range of the global frame sequence
    lo, _ = shard.shard_range(n_distinct * world, rank, world)
    host_frames = torch.empty((n_distinct, H, W), dtype=torch.float32).pin_memory()
    for i in range(n_distinct):
        host_frames[i] = torch.from_numpy(synth_frame(W, H, lo + i))
    dev_frames = host_frames.to(dev)  # 64 x 8.3 MB = 531 MB > L2 (126 MB): successive steps cannot be served from L2
    torch.cuda.synchronize()

    # 192-frame passes x 5 slots take ~145 GB of the B200's 180 GB; on a device with less free memory fall back to smaller passes
    # (same results, a few per cent slower) instead of failing the run
    while True:
        try:
            g = capi.SiftGpu(DPE, OCTAVES, 1.6, capi.SQRT2_F32, False, max_width=W, max_height=H, max_batch=args.device_batch,
                             device=local_rank, flags=args.flags)
            break
        except capi.SiftGpuError as e:
            if args.device_batch <= 16:
                raise
            print(f"note: context with {args.device_batch}-frame passes does not fit ({e}); trying {args.device_batch // 2}", file=sys.stderr)
            args.device_batch //= 2
            torch.cuda.empty_cache()

    def descs(base_ptr, memory, step, dtype=capi.DTYPE_F32):
        arr = (capi.Image * B)()
        esz = 4 if dtype == capi.DTYPE_F32 else 1
        for j in range(B):
            f = (step * B + j) % n_distinct
            arr[j] = capi.Image(base_ptr + f * W * H * esz, W, H, 0, dtype, memory, None)
        return arr

    def do_steps(base_ptr, memory, n_steps, first, dtype=capi.DTYPE_F32):
        acc = {"pyramid_ms": 0.0, "span_ms": 0.0, "launches": 0, "kps": 0, "cands": 0, "surv": 0, "device_total_ms": 0.0,
               "host_order_ms": 0.0, "stages": {}, "h2d_bytes": 0, "packed": 0}
        for s in range(n_steps):
            rc, res = g.run_raw(descs(base_ptr, memory, first + s, dtype), B)
            if rc != 0:
                raise g._err(rc)
            t = g.timings()
            acc["pyramid_ms"] += t["pyramid_ms"]; acc["span_ms"] += t["span_ms"]; acc["launches"] += int(t["kernel_launches"])
            acc["device_total_ms"] += t["device_total_ms"]; acc["host_order_ms"] += t["host_order_ms"]
            acc["h2d_bytes"] += int(t["h2d_bytes"]); acc["packed"] += int(t["packed_images"])
            for k2, v in t.items():
                if k2.endswith("_ms"):
                    acc["stages"][k2] = acc["stages"].get(k2, 0.0) + v
            for r in res:
                acc["kps"] += r.n; acc["cands"] += r.n_candidates; acc["surv"] += r.n_survivors
        return acc

    def timed(base_ptr, memory, dtype=capi.DTYPE_F32):
        sampler = ClockSampler(local_rank) if rank == 0 else None
        if sampler:
            sampler.start()
        do_steps(base_ptr, memory, args.warmup, 0, dtype)
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.mark_begin()
        t0 = time.perf_counter()
        acc = do_steps(base_ptr, memory, args.steps, args.warmup, dtype)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if dist:
            dist.barrier()
        clocks = sampler.stop() if sampler else None
        return wall, acc, clocks

    if args.quick_e2e:
        # development: only the end-to-end leg, with and without the packed upload (SIFT_GPU_HOST_PACK), one line
        out = {}
        for mode in ("default", "0", "1", "2"):
            if mode == "default":
                os.environ.pop("SIFT_GPU_HOST_PACK", None)
            else:
                os.environ["SIFT_GPU_HOST_PACK"] = mode
            w, a, _ = timed(host_frames.data_ptr(), capi.MEM_HOST)
            out["pack_" + mode] = {"images_per_s": args.steps * B / w, "h2d_bytes_per_step": a["h2d_bytes"] // args.steps,
                                   "packed_images_per_step": a["packed"] // args.steps, "host_order_ms_per_image": a["host_order_ms"] / (args.steps * B)}
        os.environ.pop("SIFT_GPU_HOST_PACK", None)
        w, a, _ = timed(dev_frames.data_ptr(), capi.MEM_DEVICE)
        out["value"] = {"images_per_s": args.steps * B / w, "host_order_ms_per_image": a["host_order_ms"] / (args.steps * B)}
        out["host_cpus"] = len(os.sched_getaffinity(0))
        print(json.dumps({"quick_e2e": out}), flush=True)
        g.close()
        return

    wall_d, acc_d, clocks = timed(dev_frames.data_ptr(), capi.MEM_DEVICE)
    # end to end: pinned host f32 frames through the C-ABI call.  The library packs 8-bit-valued f32 frames to bytes on its host
    # threads inside the call (lossless, include/sift_gpu.h) where its policy allows; the bytes that travelled are its own count.
    wall_h, acc_h, _ = timed(host_frames.data_ptr(), capi.MEM_HOST)
    wall_f32, acc_f32 = wall_h, acc_h
    (any_packed,), _ = shard.reduce_max_sum(dist, dev, [float(acc_h["packed"] > 0)], [0.0])  # every rank takes the same branch (barriers inside)
    if any_packed:
        # supplementary: the same call with the packing switched off (every frame travels as 4 bytes per pixel)
        os.environ["SIFT_GPU_HOST_PACK"] = "0"
        wall_f32, acc_f32, _ = timed(host_frames.data_ptr(), capi.MEM_HOST)
        os.environ.pop("SIFT_GPU_HOST_PACK", None)
    # supplementary: the same frames as 8-bit pixels (what a decoder produces before Vigra's importImage widens them,
    # main.cpp:52-54); the library widens on the device, results are identical, the upload is 4x smaller
    host_u8 = host_frames.to(torch.uint8).pin_memory()
    wall_u8, acc_u8, _ = timed(host_u8.data_ptr(), capi.MEM_HOST, capi.DTYPE_U8)

    # single-image latency, one 1080p frame from pinned host memory to keypoints + descriptors in host
    # memory (no batching, nothing to overlap with)
    lat = []
    one = (capi.Image * 1)()
    for i in range(12):
        one[0] = capi.Image(host_frames.data_ptr() + (i % n_distinct) * W * H * 4, W, H, 0, capi.DTYPE_F32, capi.MEM_HOST, None)
        t0 = time.perf_counter()
        g.run_raw(one, 1)
        lat.append(1e3 * (time.perf_counter() - t0))
    latency_ms = sorted(lat[2:])[len(lat[2:]) // 2]

    # Roofline of the pyramid + DoG stage: in the pipelined runs above three device passes overlap, so a stage's
    # CUDA-event duration includes other passes' kernels.  Time the stage on a serial context (one pass at a time,
    # same kernels, same frames resident in HBM, CUDA events on the library's stream) for the roofline figure.
    g.close()
    g = capi.SiftGpu(DPE, OCTAVES, 1.6, capi.SQRT2_F32, False, max_width=W, max_height=H, max_batch=args.device_batch,
                     device=local_rank, flags=args.flags | capi.FLAG_SERIAL)
    do_steps(dev_frames.data_ptr(), capi.MEM_DEVICE, 1, 0)
    torch.cuda.synchronize()
    n_roof = max(1, min(args.steps, 4))
    acc_r = do_steps(dev_frames.data_ptr(), capi.MEM_DEVICE, n_roof, 1)
    # the same stage in the other blur arithmetic (exact <-> fused multiply-add), measured the same way, for the record
    other_flags = args.flags ^ capi.FLAG_FMA_BLUR
    g.close()
    g = capi.SiftGpu(DPE, OCTAVES, 1.6, capi.SQRT2_F32, False, max_width=W, max_height=H, max_batch=args.device_batch,
                     device=local_rank, flags=other_flags | capi.FLAG_SERIAL)
    do_steps(dev_frames.data_ptr(), capi.MEM_DEVICE, 1, 0)
    torch.cuda.synchronize()
    acc_o = do_steps(dev_frames.data_ptr(), capi.MEM_DEVICE, n_roof, 1)

    # run() is synchronous (it returns after its last stream sync), so the host clock around the K steps equals the
    # device-side span; span_ms (CUDA events on the library's stream) is reported beside it.  Max over ranks.
    (mx, sm) = shard.reduce_max_sum(dist, dev, [wall_d, wall_h, acc_d["span_ms"], acc_r["pyramid_ms"], wall_u8, acc_o["pyramid_ms"], wall_f32],
                                    [args.steps * B, acc_d["launches"], acc_d["kps"], acc_d["cands"]])
    if rank != 0:
        g.close()
        dist.barrier()
        dist.destroy_process_group()
        return
    wall_d, wall_h, span_d, pyr_ms, wall_u8, pyr_other_ms, wall_f32 = mx
    images, launches, kps, cands = sm
    value = images / wall_d
    e2e = images / wall_h

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    pyr_bytes = pyramid_bytes(W, H, OCTAVES, DPE, False)
    achieved = pyr_bytes * n_roof * B / (pyr_ms * 1e-3) / 1e9  # per rank: bytes of this rank's frames / its pyramid-stage time
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "pyramid_traffic.json"))).get("dram_bytes_per_image")
    except Exception:
        pass
    per_img_h2d = W * H * 4
    d2h = (acc_h["kps"] * (20 + 512) + acc_h["surv"] * 12) / max(1, args.steps) + 8 * B
    line = {
        "metric": "1080p images/sec end-to-end SIFT", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall_d / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": B, "distinct_frames_per_gpu": n_distinct, "device_batch": args.device_batch,
                   "l2": "inputs larger than L2 (distinct frames cycled: %d MB per GPU)" % (n_distinct * per_img_h2d // 1000000),
                   "order": "canonical" if args.flags & capi.FLAG_ORDER_CANONICAL else "reference std::sort replay",
                   "blur": "fma" if args.flags & capi.FLAG_FMA_BLUR else "exact mul+add (bit-identical to the oracle)",
                   "keypoints_per_image": kps / images, "candidates_per_image": cands / images,
                   "host_cpus": len(os.sched_getaffinity(0)), "bound_to_gpu_numa_cpus": numa},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": int(acc_h["h2d_bytes"] // max(1, args.steps)), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * wall_h / args.steps, "host_buffers": "pinned f32 frames, %d bytes each" % per_img_h2d,
                "upload": ("%d of %d frames per step were 8-bit valued and travelled as bytes: packed by the library's host threads inside the timed call, "
                           "widened on the device, results bit-identical (include/sift_gpu.h, tests/test_gpu_host_pack.py)" % (acc_h["packed"] // max(1, args.steps), B))
                          if acc_h["packed"] else "f32 frames as they are (packing off: fewer than 8 host threads for this context, or SIFT_GPU_HOST_PACK=0)"},
        "e2e_f32_upload": {"value": images / wall_f32, "unit": "images/s", "h2d_bytes_per_step": int(acc_f32["h2d_bytes"] // max(1, args.steps)),
                           "ms_per_step": 1e3 * wall_f32 / args.steps,
                           "note": "supplementary: the same call with SIFT_GPU_HOST_PACK=0, every frame uploaded as 4 bytes per pixel (PCIe-bound)"},
        "e2e_u8_input": {"value": images / wall_u8, "unit": "images/s", "h2d_bytes_per_step": B * W * H, "ms_per_step": 1e3 * wall_u8 / args.steps,
                         "note": "supplementary: same call with SIFT_GPU_DTYPE_U8 host frames (identical results)"},
        "gpu_launches": int(launches),
        "single_image_latency_ms": latency_ms,
        "roofline": {"bound": "hbm", "kernel": "pyramid+DoG stage (blur/DoG/decimation launches of one device pass)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                     "algorithmic_bytes_per_image": pyr_bytes, "pyramid_ms_per_image": pyr_ms / (n_roof * B),
                     "measured": "serial context (SIFT_GPU_FLAG_SERIAL), %d steps, CUDA events around the stage" % n_roof, "traffic": traffic,
                     "serial_stage_ms_per_image": {k2: v / (n_roof * B) for k2, v in acc_r["stages"].items()}},
        "roofline_other_blur": {"blur": "exact mul+add" if other_flags & capi.FLAG_FMA_BLUR == 0 else "fma (SIFT_GPU_FLAG_FMA_BLUR: DoG within 1e-6 relative, same keypoint set to 99.9 %, "
                                        "descriptors comparable only under a pinned order; tests/test_gpu_fma_mode.py)",
                                "achieved": pyr_bytes * n_roof * B / (pyr_other_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                "frac": pyr_bytes * n_roof * B / (pyr_other_ms * 1e-3) / 1e9 / peak, "pyramid_ms_per_image": pyr_other_ms / (n_roof * B)},
        "device_span_ms_per_step": span_d / args.steps,
        "timing": "host clock around K synchronous library calls, bracketed by cuda synchronize + barrier, max over ranks; device_span_ms_per_step is the CUDA-event span (first to last stream event of each call) of the same steps",
        "stage_ms_per_image": {k2: v / (args.steps * B) for k2, v in acc_d["stages"].items()},
    }
    g.close()
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline_sample()
    else:
        line["cpu_baseline"] = None
    if world == 1 and not args.no_configs:
        line["configs"] = measure_configs(local_rank, args.flags, literal_cpu=not args.no_cpu_baseline)
    print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="frames per step per GPU (BASELINE config 5: a batch of 4096 frames per shard)")
    ap.add_argument("--frames", type=int, default=64, help="distinct synthetic frames per GPU (cycled)")
    ap.add_argument("--device-batch", type=int, default=192, help="frames per device pass (ctx max_batch); several passes are in flight")
    ap.add_argument("--flags", type=int, default=0, help="SIFT_GPU_FLAG_* bits (1 canonical order, 4 FMA blur)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick-e2e", action="store_true", help="development: only the end-to-end leg with the packed upload off / on")
    ap.add_argument("--no-configs", action="store_true", help="skip the single-image measurements of BASELINE configs 1, 2, 4")
    ap.add_argument("--literal-1500", action="store_true", help="only time the literal reference CPU path on a 1500x1500 frame (minutes)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.literal_1500:
        import numpy as np

        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as ol
        from sift_b200.synth import synth_frame

        o = ol.Oracle(DPE, 4, 1.6, float(np.float32(np.sqrt(2.0))), False, L=ol.ref_lib(fast=False))
        dt, nk = o.time_calculate(synth_frame(1500, 1500, 0))
        print(json.dumps({"literal_cpu_1500x1500": {"seconds": dt, "keypoints": nk, "threads": 1, "readme_says": "~11 min"}}))
        return
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
