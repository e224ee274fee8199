#!/usr/bin/env python
"""Benchmark of the à trous hot path on B200 (see DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload transform|wow]

Workload at N=1 (BASELINE.json configs[1]): B3spline 2-D à trous transform of one 4096x4096 fp32 frame over 10
scales.  One "step" = one full transform (10 per-scale launches) of a device-resident frame; `value` is
Mpixel*scales/s over all ranks (weak scaling: every rank transforms its own frame).  `e2e` is the same metric
through the public Python API with HOST buffers (AtrousTransform.stream): pinned-host -> device copy of the frame,
transform, device -> pinned-host copy of all 11 planes, every step, consecutive steps overlapped on three streams.  `roofline` is for the dominant kernel (atrous_rows_kernel):
algorithmic bytes 3*sizeof(T) per pixel per launch / measured launch time, against MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference` time the reference's own CPU algorithm (oracle port: the same cv2.filter2D
calls the reference makes) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SIDE = 4096
LEVELS = 10
SF_NAME = "b3spline"
METRIC = "atrous_transform_throughput"
UNIT = "Mpixel*scales/s"
WORKLOAD = "cfg2: B3spline 2-D a trous, 4096x4096 fp32, 10 scales (BASELINE.json configs[1])"


def measured_peaks():
    """HBM roofline denominator: the driver-written MEASURED_PEAKS.json (the SUSTAINED figure when the file tells burst
    and sustained apart -- K1 is timed inside a long step), else the fallback of B200_PROFILING.md."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as fh:
                d = json.load(fh)

            def num(v):
                if isinstance(v, (int, float)) and v > 0:
                    return float(v)
                if isinstance(v, dict):
                    for k in ("sustained", "sustained_gbs", "gbs", "value", "burst"):
                        if k in v and isinstance(v[k], (int, float)) and v[k] > 0:
                            return float(v[k])
                return None

            for key in ("hbm_gbs_sustained", "hbm_sustained_gbs", "hbm_gbs"):
                if key in d and num(d[key]):
                    return num(d[key]), f"MEASURED_PEAKS.json[{key}] (of measured)"
            for key, v in d.items():
                if "hbm" in key.lower() and "burst" not in key.lower() and num(v):
                    return num(v), f"MEASURED_PEAKS.json[{key}] (of measured)"
        except (OSError, ValueError):
            pass
    return 6650.0, "B200_PROFILING.md fallback (of fallback)"


def ncu_traffic():
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu capture (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh).get("dram_bytes_per_launch")
    return None


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, uuid=None, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:  # robust against CUDA_VISIBLE_DEVICES renumbering
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_reference_rate(side, levels, repeats=1):
    """Mpixel*scales/s of the reference's CPU algorithm (oracle port; cv2 backend = the reference's own calls)."""
    from oracle import atrous_oracle as orc
    backend = orc.default_backend()
    img = np.random.default_rng(0).standard_normal((side, side)).astype(np.float32)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        orc.atrous_transform(img, levels, SF_NAME, backend=backend)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    threads = 1
    if backend == "cv2":
        import cv2
        threads = cv2.getNumThreads()
    return side * side * levels / best / 1e6, best, backend, threads


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port) on the host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import atrous_oracle as orc
    backend = orc.default_backend()
    # Bounded sample: FULL-SIZE frames of the named workload, as many of the K requested steps as fit in ~100 s of CPU
    # work.  A crop would not be a sample of the same workload: the reference convolves with the dense dilated kernel
    # (2049 x 2049 taps at scale 9) through a DFT whose cost hardly depends on the image size (measured here: 1.7 / 4.5
    # / 8.4 Mpixel*scales/s on 256^2 / 512^2 / 1024^2 crops against 29 on the 4096^2 frame), so shrinking the frame to
    # fit K steps would understate the reference by an order of magnitude.
    side = N_SIDE
    img = np.random.default_rng(0).standard_normal((side, side)).astype(np.float32)
    t0 = time.perf_counter()
    orc.atrous_transform(img, LEVELS, SF_NAME, backend=backend)  # warm-up (one frame, whatever W: ~6 s each)
    t_frame = time.perf_counter() - t0
    steps = max(1, min(args.steps, int(100.0 / max(t_frame, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.atrous_transform(img, LEVELS, SF_NAME, backend=backend)
    dt = time.perf_counter() - t0
    value = side * side * LEVELS * steps / dt / 1e6
    threads = 1
    if backend == "cv2":
        import cv2
        threads = cv2.getNumThreads()
    sample = (f"{steps} full {side}x{side} fp32 frames x {LEVELS} scales timed ({args.steps} steps requested, capped to "
              f"~100 s of CPU work; 1 warm-up frame), oracle port backend={backend}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "steps_requested": args.steps, "warmup": 1, "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def solar_like_device(n, dtype, device, seed=2, flux=0.05):
    """Synthetic solar-like frame generated on the device (SURVEY.md 8(d) 'S': limb-darkened disk + exponential corona
    + 30 Gaussian active regions + background, photon noise, integer counts)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    ax = torch.arange(n, device=device, dtype=torch.float64)
    y, x = torch.meshgrid(ax, ax, indexing="ij")
    r = torch.hypot(x - n / 2, y - n / 2) / (0.4 * n)
    img = torch.where(r < 1, 2000 * (0.4 + 0.6 * torch.sqrt(torch.clamp(1 - r ** 2, min=0))),
                      800 * torch.exp(-(torch.clamp(r, min=1) - 1) / 0.15))
    for _ in range(30):
        cx, cy = ((torch.rand(2, generator=g, device=device) * 0.6 + 0.2) * n).tolist()
        sg = float(torch.rand(1, generator=g, device=device) * 27 + 3) * n / 1024
        amp = float(torch.rand(1, generator=g, device=device) * 7500 + 500)
        img += amp * torch.exp(-((x - cx) ** 2 + (y - cy) ** 2) / (2 * sg ** 2))
    img = (img + 20) * flux
    img = img + torch.sqrt(img) * torch.randn(img.shape, generator=g, device=device, dtype=torch.float64)
    return torch.round(torch.clamp(img, min=0)).to(dtype)


def time_wow(wb, img, reps, peak, **kw):
    """Device-resident wow(): frames/s and the algorithmic-byte roofline of SURVEY.md 8(d), (5L+3)*sizeof(T) B/pixel."""
    import torch
    for _ in range(3):
        _, co = wb.wow(img, **kw)
    levels = len(co) - 1
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        wb.wow(img, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    algo = (5 * levels + 3) * img.element_size() * img.numel()
    return {"frames_per_s": 1e3 / ms, "ms_per_frame": ms, "scales": levels, "algorithmic_bytes_per_frame": algo,
            "achieved_gbs": algo / ms / 1e6, "frac_of_hbm_peak": algo / ms / 1e6 / peak}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import wavelets_b200 as wb
    from wavelets_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: wavelets_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load(require_cuda=True)

    h = w = N_SIDE
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    img = torch.randn((h, w), generator=gen, device=dev, dtype=torch.float32)
    planes = torch.empty((LEVELS + 1, h, w), dtype=torch.float32, device=dev)
    scratch = torch.empty((2, h, w), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)
    assert lib.wb_atrous_scale_path(h, w, w, w, 0, _lib.WB_B3SPLINE, _lib.WB_F32, img.data_ptr(), scratch.data_ptr(),
                                    planes.data_ptr()) == 1, "expected the TMA row-pipeline kernel on this shape"

    def step():
        _lib.check(lib.wb_atrous_transform(img.data_ptr(), planes.data_ptr(), scratch.data_ptr(), 1, h, w, w, 0,
                                           LEVELS, _lib.WB_B3SPLINE, _lib.WB_F32, stream.cuda_stream))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank, uuid=getattr(torch.cuda.get_device_properties(dev), "uuid", None))
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # keep the GPU busy long enough for the clock sampler to see clocks under load, without changing K
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    if args.steps * 0.4 < 300:  # short run: extend the sampling window with untimed identical steps
        t_end = time.perf_counter() + 0.5
        while time.perf_counter() < t_end:
            for _ in range(50):
                step()
            torch.cuda.synchronize(dev)
    clocks = sampler.stop()

    # ---- end-to-end through the public API with host buffers ------------------------------------------------
    e2e_steps = max(1, min(args.steps, 20))
    host_in = torch.randn((h, w), dtype=torch.float32).pin_memory()
    host_out = torch.empty((LEVELS + 1, h, w), dtype=torch.float32).pin_memory()
    transform = wb.AtrousTransform(wb.B3spline)

    # One step = one frame through AtrousTransform.stream(): pinned-host -> device copy of the frame, the transform,
    # device -> pinned-host copy of all 11 planes; consecutive steps overlap on three CUDA streams (the upload of step
    # n+1 and the kernels of step n hide behind the download of step n-1).  Every step re-reads the same host frame and
    # rewrites the same host planes (stride-0 views), so the pinned footprint stays at one frame + one set of planes.
    frames_view = host_in.unsqueeze(0).expand(e2e_steps, h, w)
    out_view = host_out.unsqueeze(0).expand(e2e_steps, LEVELS + 1, h, w)

    def e2e_run(n):
        transform.stream(frames_view[:n], LEVELS, out=out_view[:n])

    e2e_run(2)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    e2e_run(e2e_steps)
    f1.record(stream)
    barrier()
    e2e_ms = f0.elapsed_time(f1)

    times = torch.tensor([elapsed_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms = times.tolist()

    units_per_step = h * w * LEVELS / 1e6  # Mpixel*scales
    value = units_per_step * args.steps * world / (elapsed_ms / 1e3)
    e2e_value = units_per_step * e2e_steps * world / (e2e_ms / 1e3)

    # ---- WOW frames/s (BASELINE.json configs[2]), device-resident, rank 0 reports its own GPU ---------------------
    wow_keys = {}
    if rank == 0 and not args.no_wow:
        peak_w, _ = measured_peaks()
        solar = solar_like_device(N_SIDE, torch.float32, dev)
        wow_keys["wow"] = time_wow(wb, solar, 30, peak_w)
        wow_keys["wow_bilateral"] = time_wow(wb, solar, 15, peak_w, bilateral=1, denoise_coefficients=[5, 2])
        wow_keys["wow"]["call"] = "wow(4096x4096 fp32 solar-like)"
        wow_keys["wow_bilateral"]["call"] = "wow(4096x4096 fp32 solar-like, bilateral=1, denoise_coefficients=[5, 2])"
        wow_keys["wow_bilateral"]["bound"] = "instruction issue and FMA/MUFU pipes (24 exp + ~160 fp32 ops per pixel per scale), not HBM"
        # BASELINE.json configs[3] flavour: a stack of frames, one launch per scale for the whole stack (wow_batch)
        nb = 8
        stack = solar.unsqueeze(0).repeat(nb, 1, 1)
        stack += torch.sqrt(stack.clamp(min=1)) * 0.1 * torch.randn(stack.shape, device=dev)
        for _ in range(2):
            _, pl, _ = wb.wow_batch(stack)
        levels_b = pl.shape[1] - 1
        del pl
        torch.cuda.synchronize(dev)
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(5):
            wb.wow_batch(stack)
        b1.record()
        torch.cuda.synchronize(dev)
        ms_b = b0.elapsed_time(b1) / 5 / nb
        algo_b = (5 * levels_b + 3) * 4 * h * w
        wow_keys["wow_batch"] = {"frames_per_s": 1e3 / ms_b, "ms_per_frame": ms_b, "scales": levels_b,
                                 "frames_per_launch": nb, "algorithmic_bytes_per_frame": algo_b,
                                 "achieved_gbs": algo_b / ms_b / 1e6, "frac_of_hbm_peak": algo_b / ms_b / 1e6 / peak_w,
                                 "call": "wow_batch(8 x 4096x4096 fp32 solar-like)"}
        del stack
    if world > 1:
        dist.barrier()

    if rank == 0:
        peak, peak_src = measured_peaks()
        launch_ms = elapsed_ms / (args.steps * LEVELS)
        algo_bytes = 3 * 4 * h * w  # read c_s, write c_{s+1}, write w_s
        achieved = algo_bytes / (launch_ms / 1e3) / 1e9
        # the CPU baseline is timed on rank 0 at N = 1 only (the contract); multi-GPU lines carry null
        cpu_base = None
        if world == 1:
            cpu_val, cpu_s, backend, threads = cpu_reference_rate(N_SIDE, LEVELS)
            cpu_base = {"value": cpu_val, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"one full {N_SIDE}x{N_SIDE} fp32 frame, {LEVELS} scales, {cpu_s:.2f} s, "
                                  f"oracle port backend={backend} ({os.cpu_count()} host cpus)"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_rank": 1, "sharding": "one frame per rank, no collective",
                       "l2": "working set 14 planes x 64 MiB = 896 MiB per step >> 126 MB L2 (no flush needed)",
                       "e2e_steps": e2e_steps, "kernel": "atrous_rows_lean_kernel<5> (TMA row pipeline, packed fp32x2, PDL)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": algo_bytes, "launch_ms": launch_ms},
            "cpu_baseline": cpu_base,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h * w * 4,
                    "d2h_bytes_per_step": (LEVELS + 1) * h * w * 4, "ms_per_step": e2e_ms / e2e_steps},
            "gpu_launches": args.steps * LEVELS,
            "clocks": clocks,
        }
        line.update(wow_keys)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-wow", action="store_true", help="skip the extra WOW frames/s measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
