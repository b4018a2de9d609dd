#!/usr/bin/env python
"""bench.py - throughput of the ABC-OCT B-scan reconstruction path (BASELINE.json metric: A-scans/s, B-scans/s,
HBM GB/s vs peak at 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5-2048] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic frames (fdoct_b200.synth, after the reference's
generator Matlab files/wangOCTimg2.m).  One JSON line is printed by rank 0:

  value        device-resident A-scans/s, whole job (frames already in HBM, larger than L2, CUDA-event timed,
               max over ranks), `ms_per_step` the matching time
  e2e          the same metric through the host-buffer C-ABI call abcoct_process_bscans (pinned host frames in,
               host display B-scans out, H2D + D2H inside the timed region)
  roofline     fused reconstruction kernel: algorithmic bytes per launch / mean launch duration (CUDA events on the
               launching stream, inside the timed region) against the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline the reference's own processing block compiled verbatim (oracle/_ref, kind "reference"; the oracle port when that
               module is absent, kind "port") timed on this box's host cores on a bounded sample

`--impl reference` times that CPU implementation only (all host cores, one process per core, disjoint B-scans).
oracle/ is used here ONLY as the thing timed in those two legs; the product path never touches it.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LMIN, LMAX = 840.5e-9, 859.5e-9

# name -> parameters; `frames` = frames per step and per GPU (weak scaling)
WORKLOADS = {
    "c1": dict(w=1280, h=960, N=1280, D=640, A=1, variant=0, frames=1024, seed=1001,
               desc="C1 1280x960 u16, N=1280, D=640, averages=1, FFT variant"),
    "c2": dict(w=1280, h=960, N=1280, D=640, A=8, variant=1, frames=1024, seed=1002,
               desc="C2 1280x960 u16, N=1280, D=640, averages=8, DARK variant"),
    "c3": dict(w=1920, h=1200, N=3840, D=1024, A=1, variant=0, frames=48, seed=1003, m=2, twelve_bit=True,
               desc="C3 1920x1200 u16 (12-bit), increasefftpointsmultiplier=2 -> N=3840, D=1024 (general pre-processing path)"),
    "c4": dict(w=1920, h=1200, N=1920, D=960, A=1, variant=0, frames=125, seed=1004, twelve_bit=True,
               desc="C4 1920x1200 u16 (12-bit), N=1920, D=960, 125 B-scans per GPU"),
    "c5-1024": dict(w=1024, h=1024, N=1024, D=512, A=1, variant=0, frames=2048, seed=1005,
                    desc="C5 1024-sample u16 spectra, 1024 A-scans/frame, D=512"),
    "c5-2048": dict(w=2048, h=1024, N=2048, D=1024, A=1, variant=0, frames=1024, seed=1005,
                    desc="C5 2048-sample u16 spectra, 1024 A-scans/frame, D=1024 (north_star target config)"),
    "c5-4096": dict(w=4096, h=1024, N=4096, D=2048, A=1, variant=0, frames=512, seed=1005,
                    desc="C5 4096-sample u16 spectra, 1024 A-scans/frame, D=2048"),
}
DEFAULT_WORKLOAD = "c5-2048"


def oracle_params(wl):
    from oracle.abcoct_oracle import Params

    return Params(w=wl["w"], h=wl["h"], numfftpoints=wl["N"], numdisplaypoints=wl["D"], averages=wl["A"],
                  variant=wl["variant"], fft_multiplier=wl.get("m", 1), lambdamin=LMIN, lambdamax=LMAX)


def abi_params(api, wl):
    return api.default_params(w=wl["w"], h=wl["h"], bpp=16, binx=1, biny=1, averages=wl["A"], numfftpoints=wl["N"],
                              numdisplaypoints=wl["D"], lambdamin=LMIN, lambdamax=LMAX, mediann=0, movavgn=0, fft_multiplier=wl.get("m", 1),
                              rowwisenormalize=0, donotnormalize=1, variant=wl["variant"], weight_mode=0)


def make_inputs(wl, nframes, n_unique=4):
    """Synthetic frames + calibration for one rank. Few unique interferograms, tiled to the batch size."""
    from fdoct_b200 import synth

    w, h, A, seed = wl["w"], wl["h"], wl["A"], wl["seed"]
    dark = wl["variant"] == 1
    fs = 4095 if wl.get("twelve_bit") else 65535
    nu = max(A, n_unique) if nframes >= max(A, n_unique) else nframes
    uniq = synth.make_frames(nu, w, h, seed=seed, dark=dark, full_scale=fs)
    frames = uniq[np.arange(nframes) % nu]
    yd = None
    if dark:
        yd = synth.make_dark_frames(2, w, h, seed=seed + 2).mean(axis=0)
        yr = synth.make_background_frames(2, w, h, seed=seed + 1, dark=True, full_scale=fs).mean(axis=0)
        ys = yd + 0.02 * (yr - yd)
        yb = (yr - yd) + (ys - yd)  # BscanDark.cpp:996
    else:
        yb = synth.make_background_frames(2, w, h, seed=seed + 1, full_scale=fs).mean(axis=0)
    return frames, uniq, yb, yd


# --------------------------------------------------------------------------------------------- CPU reference leg
_CPU_SHARED = {}  # inputs of the worker processes: set BEFORE the pool forks, so nothing is pickled per call


def cpu_reference_module(wl):
    """oracle/_ref: the reference's own processing block compiled verbatim (oracle/build_ref.py) - BscanFFT.cpp for variant 0,
    BscanDark.cpp for variant 1.  None when it was not built / shipped: the CPU legs then time the oracle port."""
    try:
        from oracle import build_ref

        return build_ref.load("abcoct_ref_dark" if wl["variant"] == 1 else "abcoct_ref")
    except Exception:
        return None


def cpu_kind(wl):
    import cv2

    if cpu_reference_module(wl) is not None:
        src = "BscanDark.cpp" if wl["variant"] == 1 else "BscanFFT.cpp"
        return "reference", f"the reference's processing block ({src}) compiled verbatim, OpenCV {cv2.__version__} kernels through cv2"
    return "port", f"oracle = Python restatement calling OpenCV {cv2.__version__} kernels"


def cpu_process(wl, frames, yb, yd):
    """`frames` through the reference's CPU implementation of the path: oracle/_ref when present, else the oracle port."""
    mod = cpu_reference_module(wl)
    if mod is not None:
        prm = dict(w=wl["w"], h=wl["h"], averages=wl["A"], binvalue=1, numfftpoints=wl["N"], numdisplaypoints=wl["D"], movavgn=0,
                   clampupper=False, lambdamin=LMIN, lambdamax=LMAX, mediann=0, fft_multiplier=wl.get("m", 1), bscanthreshold=-30.0,
                   rowwisenormalize=False, donotnormalize=True, bandpassfilter=False)
        ybc = np.ascontiguousarray(yb, dtype=np.float64)
        ydc = None if yd is None else np.ascontiguousarray(yd, dtype=np.float64)
        step = 8 * wl["A"]  # the module returns every B-scan it made: bounded chunks keep a worker's memory flat (the one-off
        for i in range(0, len(frames), step):  # table precompute it repeats per call is ~1e-3 of a chunk)
            mod.run_block(prm, frames[i:i + step], ybc, None, ydc)
        return
    from oracle.abcoct_oracle import Oracle

    o = Oracle(oracle_params(wl))
    o.set_background(yb)
    if yd is not None:
        o.set_dark(yd)
    o.process_bscans(frames)


def _cpu_worker(nbscans):
    """One worker process: the reference block over `nbscans` B-scans of the shared sample. Returns A-scans processed."""
    import cv2

    wl, frames, yb, yd = _CPU_SHARED["wl"], _CPU_SHARED["frames"], _CPU_SHARED["yb"], _CPU_SHARED["yd"]
    cv2.setNumThreads(1)
    n = nbscans * wl["A"]
    cpu_process(wl, frames[:n], yb, yd)
    return n * wl["h"]


def _cpu_sample(wl, uniq, bscans):
    A = wl["A"]
    return np.concatenate([uniq] * ((bscans * A + len(uniq) - 1) // len(uniq)))[: bscans * A]


def time_cpu_reference(wl, uniq, yb, yd, cores, bscans_per_core):
    """All host cores, one oracle process per core, `bscans_per_core` B-scans each (inputs inherited by fork); returns
    (A-scans/s, seconds, sample description)."""
    import multiprocessing as mp

    _CPU_SHARED.update(wl=wl, frames=_cpu_sample(wl, uniq, bscans_per_core), yb=yb, yd=yd)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_worker, [1] * cores)  # warm the workers (imports, cv2 init)
        t0 = time.perf_counter()
        n = sum(pool.map(_cpu_worker, [bscans_per_core] * cores, chunksize=1))
        dt = time.perf_counter() - t0
    return n / dt, dt, f"{cores} processes x {bscans_per_core} B-scans ({bscans_per_core * wl['A']} frames of {wl['h']} A-scans each)"


def time_cpu_single_process(wl, uniq, yb, yd, bscans):
    """The reference-like mode: ONE process, OpenCV's own internal threads (what the reference binary does)."""
    import cv2

    cv2.setNumThreads(-1)
    frames = _cpu_sample(wl, uniq, bscans)
    cpu_process(wl, frames[: wl["A"]], yb, yd)
    t0 = time.perf_counter()
    cpu_process(wl, frames, yb, yd)
    dt = time.perf_counter() - t0
    return {"value": frames.shape[0] * wl["h"] / dt, "unit": "A-scans/s", "bscans": bscans, "seconds": dt,
            "cv2_threads": int(cv2.getNumThreads())}


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def calibrate_cpu_sample(wl, uniq, yb, yd, target_s, lo=16, hi=64):
    """B-scans per core: about `target_s` seconds of CPU work, never fewer than `lo` (a short sample under-measures the CPU:
    worker start-up and the first-call costs of cv2 would dominate)."""
    _CPU_SHARED.update(wl=wl, frames=_cpu_sample(wl, uniq, 1), yb=yb, yd=yd)
    t0 = time.perf_counter()
    _cpu_worker(1)
    one = max(time.perf_counter() - t0, 1e-3)
    return int(max(lo, min(hi, round(target_s / one))))


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock, power and throttle reasons of one GPU through NVML every ~2 ms from a thread, so that even a
    timed region of a few tens of milliseconds holds samples (nvidia-smi -lms is too coarse for that)."""

    def __init__(self, gpu_index):
        self.rows, self.stop_flag, self.h, self.nv = [], False, None, None
        try:
            import pynvml as nv

            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_sm = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()
        except Exception as e:  # pragma: no cover
            self.err = repr(e)
            self.h = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.perf_counter(), sm, pw, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self, t0, t1):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")]}
        self.stop_flag = True
        self.th.join(timeout=1.0)
        nv = self.nv
        rows = [r for r in self.rows if t0 <= r[0] <= t1]
        in_region = len(rows)
        if not rows:
            rows = self.rows[-3:]
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted(n for n, bit in names.items() if any(r[3] & bit for r in rows))
        return {"sm_mhz": float(np.median([r[1] for r in rows])) if rows else None, "sm_max_mhz": float(self.max_sm),
                "power_w_max": max(r[2] for r in rows) if rows else None, "samples": in_region, "reasons": reasons}


TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_traffic.json")


def measured_traffic(workload, nframes, live=True):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the fused kernel.  Measured in this run by a one-launch
    ncu pass over a child bench process (counters only - no timing is taken from it); when ncu is not available here the
    number recorded by the last profiled run (profiles/r02_traffic.json) is reported instead.  Returns (bytes, source)."""
    if live:
        try:
            cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k",
                   "regex:wres_kernel|wrow_kernel|recon_kernel", "-s", "3", "-c", "1", "--csv", sys.executable, os.path.abspath(__file__),
                   "--workload", workload, "--frames", str(nframes), "--steps", "1", "--warmup", "3", "--no-cpu", "--no-traffic",
                   "--e2e-steps", "1"]
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=240).stdout
            vals = {}
            import csv as _csv

            for r in _csv.reader(out.splitlines()):
                if len(r) > 10 and r[0].isdigit():
                    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[-2], 1)
                    vals[r[-3]] = float(r[-1].replace(",", "")) * scale
            if len(vals) == 2:
                return int(sum(vals.values())), "ncu, one launch, this run"
        except Exception:  # noqa: BLE001
            pass
    try:
        with open(TRAFFIC_FILE) as f:
            t = json.load(f)[workload][str(nframes)]
        return int(t["dram_read"]) + int(t["dram_write"]), "profiles/r02_traffic.json (recorded ncu capture; ncu unavailable in this run)"
    except Exception:  # noqa: BLE001
        return None, "not measured"


def ncu_summary(workload):
    """Secondary-roof evidence of the committed ncu capture (issue slots, FMA pipe, shared-memory wavefronts), if any."""
    try:
        with open(TRAFFIC_FILE) as f:
            return json.load(f)[workload]["ncu"]
    except Exception:
        return None


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"



# --------------------------------------------------------------------------------------------- multi-GPU volume leg
def volume_leg(api, torch, dist, dev, rank, world, local_rank, per_gpu):
    """BASELINE config 4 in the shape `north_star` names: a triggered-capture volume of 1920x1200 frames, `per_gpu` B-scans per
    GPU (125 -> 1000 B-scans on 8 GPUs), host buffers in, host display B-scans out, FINAL HOST GATHER INSIDE the timed region.
    Three measurements (all ranks take part; rank 0 returns the dict):
      ranks_gather   one process per GPU through fdoct_b200.shard.process_sharded: every rank reconstructs its shard from pinned
                     host memory, the display images are gathered to rank 0 over a host (gloo) group
      one_context    ONE abcoct context spanning all GPUs (abcoct_create(ngpu = world), one feeder thread per GPU) on rank 0:
                     the whole volume in one pinned host buffer, all outputs into one host array - the gather is implicit
      copy_ceiling   the same bytes, H2D and D2H concurrently on every GPU, no kernel: what the box's host side can feed
    and a self-check: every rank's shard and the one-context output are identical (all shards hold the same frames)."""
    from fdoct_b200 import shard

    wl = WORKLOADS["c4"]
    w, h, D, A = wl["w"], wl["h"], wl["D"], wl["A"]
    frames, uniq, yb, yd = make_inputs(wl, per_gpu)
    gloo = dist.new_group(backend="gloo")
    ctx = api.Context(abi_params(api, wl), gpu_ids=[local_rank])
    ctx.set_background(yb)
    pin_in = api.PinnedArray(frames.shape, np.uint16)
    pin_in.array[...] = frames
    pin_out = api.PinnedArray((per_gpu, D, h), np.uint8)

    def fn(fr):
        ctx.process_bscans(fr, out8=pin_out.array)
        return pin_out.array

    src = shard.FrameSource(per_gpu * world, lambda lo, hi: pin_in.array)  # every rank's shard = its own pinned copy
    fn(pin_in.array)  # warm-up (allocates the ring)
    t = torch.zeros(3, dtype=torch.float64, device=dev)

    def sync_all():
        torch.cuda.synchronize()
        dist.barrier(group=gloo)

    sync_all()
    t0 = time.perf_counter()
    whole = shard.process_sharded(fn, src, A, (D, h), rank=rank, world=world, group=gloo)
    dist.barrier(group=gloo)
    t[0] = time.perf_counter() - t0

    # copy ceiling: the shard's input up, the shard's output down, concurrently on two streams, no kernel
    d_in = torch.empty(frames.nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(per_gpu * D * h, dtype=torch.uint8, device=dev)
    h_in = torch.from_numpy(pin_in.array.view(np.uint8).reshape(-1))
    h_out = torch.from_numpy(pin_out.array.reshape(-1))
    keep = pin_out.array.copy()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for timed in (False, True):
        sync_all()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        sync_all()
        if timed:
            t[1] = time.perf_counter() - t0
    same_shards = True
    ctx.close()

    # one context over all GPUs, on rank 0 (the other ranks wait)
    one_ok = None
    if rank == 0:
        big_in = api.PinnedArray((per_gpu * world,) + frames.shape[1:], np.uint16)
        for r in range(world):
            big_in.array[r * per_gpu:(r + 1) * per_gpu] = frames
        big_out = api.PinnedArray((per_gpu * world, D, h), np.uint8)
        with api.Context(abi_params(api, wl), gpu_ids=list(range(world))) as mctx:
            mctx.set_background(yb)
            mctx.process_bscans(big_in.array, out8=big_out.array)  # warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            mctx.process_bscans(big_in.array, out8=big_out.array)
            t[2] = time.perf_counter() - t0
        one_ok = all(np.array_equal(big_out.array[r * per_gpu:(r + 1) * per_gpu], keep) for r in range(world))
        same_shards = all(np.array_equal(whole[r * per_gpu:(r + 1) * per_gpu], keep) for r in range(world))
        big_in.free()
        big_out.free()
    dist.barrier(group=gloo)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    pin_in.free()
    pin_out.free()
    if rank != 0:
        return None
    t_rg, t_cp, t_one = [float(x) for x in t.cpu()]
    ascans = per_gpu * world * A * h
    in_b, out_b = int(frames.nbytes) * world, per_gpu * D * h * world
    ceiling = ascans / t_cp
    return {"workload": wl["desc"], "bscans_total": per_gpu * world, "bscans_per_gpu": per_gpu, "host_bytes_in": in_b, "host_bytes_out": out_b,
            "ranks_gather": {"value": ascans / t_rg, "unit": "A-scans/s", "seconds": t_rg,
                             "note": "torchrun ranks via shard.process_sharded, gloo gather of the display images to rank 0 inside the timed region"},
            "one_context": {"value": ascans / t_one, "unit": "A-scans/s", "seconds": t_one,
                            "note": "one abcoct context over all GPUs (one feeder thread per GPU), one pinned host buffer in, one out"},
            "copy_ceiling": {"value": ceiling, "unit": "A-scans/s", "seconds": t_cp, "h2d_gbs": in_b / t_cp / 1e9, "d2h_gbs": out_b / t_cp / 1e9,
                             "note": "same bytes, cudaMemcpyAsync H2D + D2H concurrently on every GPU, no kernel"},
            "one_context_frac_of_ceiling": (ascans / t_one) / ceiling, "ranks_gather_frac_of_ceiling": (ascans / t_rg) / ceiling,
            "outputs_identical": bool(one_ok and same_shards)}

# --------------------------------------------------------------------------------------------- main arms
def run_reference(args, wl, rank):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    _, uniq, yb, yd = make_inputs(wl, max(wl["A"], 4))
    per_core = calibrate_cpu_sample(wl, uniq, yb, yd, target_s=10 * args.cpu_seconds / max(1, args.steps + args.warmup))
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, sample = time_cpu_reference(wl, uniq, yb, yd, cores, per_core)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals]) * 1e3)
    import cv2

    line = {
        "impl": "reference", "metric": "A-scans/s", "value": value, "unit": "A-scans/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64/f32 (OpenCV)", "data": "synthetic",
        "config": {"workload": wl["desc"], "name": args.workload, "sample_per_step": sample},
        "bscans_per_s": value / (wl["h"] * wl["A"]),
        "cpu_baseline": {"value": value, "unit": "A-scans/s", "cores": cores, "kind": cpu_kind(wl)[0], "cpu": cpu_model(),
                         "sample": sample + "; " + cpu_kind(wl)[1],
                         "single_process": time_cpu_single_process(wl, uniq, yb, yd, 4)},
        "e2e": {"value": value, "unit": "A-scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_near_gpu(gpu_index):
    """Multi-rank runs: keep this rank's threads - and with them its pinned staging buffers (first touch) - on the CPUs next to
    its GPU (NVML's ideal affinity), so that 8 ranks do not pull their H2D traffic across the socket interconnect.  Returns a
    short description for the JSON line; never fatal."""
    try:
        import pynvml as nv

        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(gpu_index)
        ncpu = os.cpu_count() or 1
        words = nv.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        near = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = near & allowed
        if not use or use == allowed:
            return "none (NVML affinity covers every allowed CPU)" if use else "none (no allowed CPU near the GPU)"
        os.sched_setaffinity(0, use)
        return f"{len(use)} of {len(allowed)} CPUs"
    except Exception as e:  # noqa: BLE001
        return "none (" + type(e).__name__ + ")"


def run_ours(args, wl, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from fdoct_b200 import api

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    affinity = bind_near_gpu(local_rank) if world > 1 and os.environ.get("ABCOCT_BENCH_NO_AFFINITY") != "1" else "not set"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "NONE"  # keep stdout to the one JSON line (NCCL prints its version banner there)
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    w, h, N, D, A = wl["w"], wl["h"], wl["N"], wl["D"], wl["A"]
    nframes = args.frames or wl["frames"]
    nframes = max(A, nframes // A * A)
    nB = nframes // A
    frames, uniq, yb, yd = make_inputs(wl, nframes)
    ctx = api.Context(abi_params(api, wl), gpu_ids=[local_rank])
    ctx.set_background(yb)
    if yd is not None:
        ctx.set_dark(yd)

    # ---- device-resident leg
    d_in = torch.from_numpy(frames.view(np.int16)).to(dev)
    d_out = torch.empty((nB, D, h), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step_dev():
        ctx.process_bscans_device(d_in.data_ptr(), nframes, d_out.data_ptr(), None, stream=stream.cuda_stream)

    for _ in range(max(3, args.warmup)):
        step_dev()
    barrier()
    ctx.timing_reset()
    l0 = ctx.info().kernel_launches
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step_dev()
    e1.record(stream)
    barrier()
    t1 = time.perf_counter()
    dev_ms = e0.elapsed_time(e1)
    launches = ctx.info().kernel_launches - l0
    nchunks, recon_ms, norm_ms = ctx.timing_read(0)
    clocks = sampler.stop(t0, t1) if sampler else None

    # ---- end-to-end leg: pinned host frames -> abcoct_process_bscans -> host display B-scans
    pin_in = api.PinnedArray(frames.shape, np.uint16)
    pin_in.array[...] = frames
    pin_out = api.PinnedArray((nB, D, h), np.uint8)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    ctx.process_bscans(pin_in.array, out8=pin_out.array)  # warm-up (allocates the ring)
    barrier()
    t2 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.process_bscans(pin_in.array, out8=pin_out.array)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t2
    ok = bool(np.array_equal(pin_out.array, d_out.cpu().numpy()))

    # ---- max over ranks
    t = torch.tensor([dev_ms, e2e_s * 1e3, recon_ms / max(nchunks, 1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, recon_launch_ms = [float(x) for x in t.cpu()]
    info = ctx.info()

    volume = None
    if world > 1 and not args.no_volume:
        volume = volume_leg(api, torch, dist, dev, rank, world, local_rank, args.volume_bscans)

    if rank == 0:
        ascans_step = nframes * h * world
        value = ascans_step * args.steps / (dev_ms * 1e-3)
        e2e_value = ascans_step * e2e_steps / (e2e_ms * 1e-3)
        peak, peak_src = measured_hbm_peak()
        traffic, traffic_src = (None, "skipped") if args.no_traffic else measured_traffic(args.workload, nframes, live=world == 1)
        bytes_per_ascan = 2 * w + D / A  # u16 pixels in + u8 display pixels out, dB output off
        ascans_per_launch = nframes * h / max(nchunks // max(args.steps, 1), 1)
        achieved = bytes_per_ascan * ascans_per_launch / (recon_launch_ms * 1e-3) / 1e9 if recon_launch_ms > 0 else None
        line = {
            "metric": "A-scans/s", "value": value, "unit": "A-scans/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "name": args.workload, "frames_per_step_per_gpu": nframes,
                       "ascans_per_step": ascans_step, "input_bytes_per_step_per_gpu": int(frames.nbytes),
                       "l2": "inputs larger than L2 (no flush needed)" if frames.nbytes > 256e6 else "inputs SMALLER than 2x L2",
                       "parallelism": f"{world} GPU(s), B-scans sharded by rank, no collective on the data path"},
            "bscans_per_s": value / (h * A),
            "gpu_launches": int(launches),
            "kernel_ms_per_step": {"recon": recon_ms / args.steps, "normalise": norm_ms / args.steps},
            "e2e": {"value": e2e_value, "unit": "A-scans/s", "h2d_bytes_per_step": int(frames.nbytes) * world,
                    "d2h_bytes_per_step": int(nB * D * h) * world, "steps": e2e_steps, "matches_device_leg": ok},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": bytes_per_ascan * ascans_per_launch,
                         "secondary_roofs_ncu": ncu_summary(args.workload),
                         "kernel": ["recon_kernel (fused reconstruction + display normalisation, thread group per row pair, dB scratch in L2)",
                                    "wrow_kernel (fused reconstruction + display normalisation, one warp per A-scan, dB scratch in L2)",
                                    "wres_kernel (fused reconstruction + display normalisation, one warp per A-scan, dB rows resident in tensor memory)",
                                    "generic_recon_kernel (any N = 2^a 3^b 5^c / row width / D <= N, shared-memory Stockham, run-time radices)"][info.kernel_kind],
                         "bytes_per_ascan": bytes_per_ascan,
                         "ascans_per_launch": ascans_per_launch, "launch_ms": recon_launch_ms, "peak_source": peak_src,
                         "whole_step_frac": bytes_per_ascan * value / world / 1e9 / peak},
            "plan": {"fft_threads": info.fft_threads, "radix": list(info.fft_radix), "groups_per_cta": info.groups_per_cta,
                     "smem_bytes": info.smem_bytes, "regs_per_thread": info.regs_per_thread, "sm_count": info.sm_count,
                     "kernel_kind": info.kernel_kind, "slots_per_warp": info.slots_per_warp},
            "clocks": clocks,
        }
        line["config"]["cpu_affinity"] = affinity
        if volume is not None:
            line["volume"] = volume
        if not args.no_cpu and world == 1:  # the CPU baseline is a single-GPU-run figure (rank 0 at N = 1 only)
            cores = os.cpu_count() or 1
            per_core = calibrate_cpu_sample(wl, uniq, yb, yd, target_s=args.cpu_seconds)
            v, dt, sample = time_cpu_reference(wl, uniq, yb, yd, cores, per_core)
            line["cpu_baseline"] = {"value": v, "unit": "A-scans/s", "cores": cores, "kind": cpu_kind(wl)[0], "cpu": cpu_model(),
                                    "sample": sample + f"; {dt:.1f} s; " + cpu_kind(wl)[1],
                                    "single_process": time_cpu_single_process(wl, uniq, yb, yd, 4)}
        print(json.dumps(line), flush=True)
    pin_in.free()
    pin_out.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="frames per step per GPU (default: the workload's)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work per timed CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-traffic", action="store_true", help="skip the one-launch ncu pass that measures roofline.traffic")
    ap.add_argument("--no-volume", action="store_true", help="multi-GPU runs: skip the C4 volume leg (sharded volume + host gather, one context over all GPUs, copy ceiling)")
    ap.add_argument("--volume-bscans", type=int, default=125, help="B-scans per GPU in the volume leg (125 x 8 GPUs = the 1000-B-scan volume of BASELINE config 4)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world == 1 and args.gpus > 1:
        # convenience: relaunch under torchrun when started as plain `python bench.py --gpus N`
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl, rank)
    else:
        run_ours(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
