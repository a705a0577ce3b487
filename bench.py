#!/usr/bin/env python
"""bench.py -- 1080p30 frames/s through EVM + ROI + BPM (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c4|c2|c3|c1] [--roi rect|poly]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A STEP is one pass of the hot path over one synthetic clip per rank: fused pyrDown cascade ->
temporal ideal bandpass (x alpha) -> collapse + add-back with fused ROI means (float32 magnified
frames written to HBM) -> ROI finalize -> BPM (float32 detrend + FFT peak).  --roi rect (default):
the reference's clamped cheek rectangle (analysis/utils/roi.py:43-59); --roi poly: forehead + two
cheek landmark polygons (36 vertices each), rasterised per frame and fused into the same pass,
BPM from the ROI trace with the strongest peak (estimate_bpm.py:59-64).
Workload c4 (default): 1920x1080, 30 FPS, 60 s (T = 1800), 4 pyramid levels, 0.7-4 Hz, alpha 50;
clip i carries a pulse of 0.8 + i*2.4/63 Hz.  Clips are independent, so ranks take clips
round-robin (weak scaling, no data-path collective; one final gather of the BPMs).

`value`    device-timed: clips resident in HBM before the timed region (each clip is 11.2 GB,
           far larger than the 126 MB L2, so no L2 flush is needed between steps).
`e2e`      the same metric through the host-buffer call (Engine.evm_roi_host ->
           vhr_evm_roi_host): pinned host frames in, H2D + kernels + D2H of the ROI trace and
           the BPM inside the timed region.
`roofline` dominant kernel (collapse + add-back): algorithmic bytes per launch / its mean
           CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
`bpm_ok`   every timed clip's BPM and peak bin equal the CPU ORACLE's for that clip
           (tests/golden/configs.npz, made by tests/golden/make_config_golden.py).
`cpu_baseline` / `--impl reference`: the CPU oracle (cv2.pyrDown/pyrUp + np.fft EVM port and the
           reference's own ROI/BPM functions restated) on the host cores, bounded sample.
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
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (W, H, fps, T, description)
    "c4": (1920, 1080, 30.0, 1800, "c4: 1920x1080 30 FPS 60 s clips, 4-level pyramid, 0.7-4 Hz, alpha 50, fp32 out"),
    "c2": (1280, 720, 30.0, 1800, "c2: 1280x720 30 FPS 60 s clip, 4-level pyramid, 0.7-4 Hz, alpha 50, fp32 out"),
    "c3": (640, 480, 30.0, 1800, "c3: 640x480 30 FPS 60 s stream, 10 s windows / 1 s hop"),
    "c1": (256, 144, 5.0, 150, "c1: 256x144 5 FPS 30 s clip"),
}
LEVELS, F_LO, F_HI, ALPHA = 4, 0.7, 4.0, 50.0
METRIC = "1080p30 frames/sec (EVM+ROI+BPM)"


def pulse_hz(clip: int) -> float:
    return 0.8 + (clip % 64) * (2.4 / 63.0)


def recorded_traffic(kernel: str):
    """(DRAM bytes per launch of `kernel`, file) from the newest committed ncu capture record
    (profiles/r2_traffic.json, else profiles/r1_traffic.json); (None, None) without a record."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            rec = json.load(open(os.path.join(ROOT, "profiles", name)))[kernel]
            return int(rec["traffic_bytes"]), "profiles/" + name
        except Exception:
            continue
    return None, None


def oracle_goldens(workload: str, roi: str):
    """{clip: (bpm, bin)} from tests/golden/configs.npz: the CPU oracle's result for the clips this
    workload generates (a data file -- the oracle package itself is not imported on the GPU arm)."""
    key = {"c4": "c4", "c1": "bench_c1", "c2": "bench_c2", "c3": "bench_c3"}[workload] + ("poly" if roi == "poly" else "")
    try:
        g = np.load(os.path.join(ROOT, "tests", "golden", "configs.npz"))
        return {int(c): (float(b), int(k)) for c, b, k in zip(g[key + "_ids"], g[key + "_bpm"], g[key + "_bin"])}
    except Exception:
        return {}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
            time.sleep(0.25)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self):
        return time.perf_counter()

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [l for (ts, l) in self.lines if (t0 is None or ts >= t0) and (t1 is None or ts <= t1 + 0.15)]
        if not rows:
            rows = [l for (_, l) in self.lines]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in rows:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU baseline
def cpu_reference_run(W, H, fps, T_sample, steps, warmup, frames=None):
    """The oracle (EVM port on cv2/np.fft + reference ROI/BPM restatement) on the host cores.
    Returns (frames_per_s, seconds_per_step list, cores, bpm)."""
    import cv2
    from oracle import bpm as obpm, evm as oevm, fast as ofast, roi as oroi, synth as osynth
    cores = os.cpu_count() or 1
    cv2.setNumThreads(cores)
    p = osynth.SynthParams(T=T_sample, H=H, W=W, fps=fps, pulse_hz=1.2, seed=0, clip=0)
    if frames is None:
        frames = ofast.synth_frames(p)
    lm = p.landmarks()
    rect = oroi.cheek_roi_from_bbox(oroi.bbox_from_landmarks_clamped(lm[:, 0], lm[:, 1], W, H), W, H)
    rects = np.tile(np.array(rect, dtype=np.int32), (T_sample, 1, 1))
    times, bpm = [], None
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        _, _, _, means = oevm.evm_clip_cv2(frames, fps, LEVELS, F_LO, F_HI, ALPHA, rects=rects, keep_out=False)
        g = means[:, 0, 1].astype(np.float32)
        bpm = obpm.estimate_bpm_analysis(g - np.mean(g), fps)[0]
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    return T_sample / min(times), times, cores, bpm


def run_reference(args):
    """--impl reference: rank 0 alone measures; other ranks exit 0."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    W, H, fps, T, desc = WORKLOADS[args.workload]
    T_sample = 48 if W >= 1280 else min(T, 150)
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 2))
    t_start = time.perf_counter()
    fps_cpu, times, cores, bpm = cpu_reference_run(W, H, fps, T_sample, steps, warmup)
    total = float(np.sum(times))
    value = steps * T_sample / total
    sample = (f"{T_sample}-frame {W}x{H} sub-clip per step through the full EVM+ROI+BPM path "
              f"(cv2.pyrDown/pyrUp float32 with {cores} threads, np.fft float64 bandpass, rectangle ROI mean, FFT-peak BPM)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * total / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc + f" [CPU arm: each step is a {T_sample}-frame sub-clip of the {T}-frame clip; "
                                          "frames/s is per frame]", "levels": LEVELS, "band_hz": [F_LO, F_HI], "alpha": ALPHA,
                       "frames_per_step": T_sample, "frames_per_clip": T, "roi": "1 cheek rectangle"},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t_start}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import video_heart_rate_b200 as vhr
    from video_heart_rate_b200 import host, parallel
    from video_heart_rate_b200.pipeline import ANALYSIS_BAND

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    affinity = None
    if args.affinity == "on" or (args.affinity == "auto" and world > 1):
        # each rank on its own share of the host cores, before any pinned memory is allocated (first-touch locality
        # of the clip it uploads in the e2e leg; tools/h2d_probe.py measures what this is worth on the box)
        ncpu = os.cpu_count() or 1
        per = max(1, ncpu // world)
        try:
            os.sched_setaffinity(0, set(range(local_rank * per, min(ncpu, (local_rank + 1) * per))))
            affinity = f"cores {local_rank * per}-{min(ncpu, (local_rank + 1) * per) - 1} per rank"
        except OSError:
            affinity = None
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = vhr.Engine(local_rank)
    dev = eng.tdev
    W, H, fps, T, desc = WORKLOADS[args.workload]
    K, Wm = args.steps, max(3, args.warmup)
    dims = eng.pyr_dims(W, H, LEVELS)
    wl, hl = dims[-1]
    p = wl * hl
    u8out = args.out == "u8"                                   # saturated uint8 magnified frames instead of float32 ones
    bytes_per_frame = (9 if u8out else 18) * W * H + 48 * p    # BASELINE.md section 3 (fp32 output: the contract figure)
    collapse_bytes_per_frame = (6 if u8out else 15) * W * H + 12 * p   # S3: read L4 + u8 frame, write the output frame
    peak_gbs, peak_src = peaks()

    # ---- resident inputs -------------------------------------------------------------------
    n_res = max(1, min(args.resident, K + Wm))
    my_clips = [rank + world * j for j in range(n_res)]
    specs = [vhr.SynthSpec(T=T, H=H, W=W, fps=fps, pulse_hz=pulse_hz(c), seed=c, clip=c) for c in my_clips]
    clips = [eng.synth_clip(s) for s in specs]
    lm = specs[0].landmarks()
    rect = host.slice_rects(host.cheek_roi_clamped(host.bbox_clamped(lm[None], W, H), W, H), W, H)[0]
    rects = torch.as_tensor(np.tile(rect, (T, 1, 1)).astype(np.int32), device=dev)
    poly = args.roi == "poly"
    polys_np, nverts_np = host.face_polygons(np.broadcast_to(lm, (T,) + lm.shape), W, H, n_vertices=36)
    polys_d, nverts_d = torch.as_tensor(polys_np, device=dev), torch.as_tensor(nverts_np, device=dev)
    out = torch.empty((T, H, W, 3), dtype=torch.uint8 if u8out else torch.float32, device=dev)
    lvl = torch.empty((T, hl, wl, 3), dtype=torch.float32, device=dev)
    starts = torch.zeros(1, dtype=torch.int32, device=dev)      # BPM window list lives on the device
    lens = torch.full((1,), T, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()

    stream = torch.cuda.current_stream(dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    kernel_events = {"pyrdown": [], "bandpass": [], "collapse": [], "bpm": []}

    def step(i, timed):
        fr = clips[i % n_res]
        e = [ev() for _ in range(5)] if timed else None
        if timed: e[0].record(stream)
        eng.pyrdown(fr, LEVELS, out=lvl)
        if timed: e[1].record(stream)
        eng.bandpass(lvl, fps, F_LO, F_HI, ALPHA, out=lvl)
        if timed: e[2].record(stream)
        if poly:
            _, _, means = eng.collapse(lvl, fr, LEVELS, out_f32=False if u8out else out, out_u8=out if u8out else False,
                                       polys=polys_d, nverts=nverts_d)
        else:
            _, _, means = eng.collapse(lvl, fr, LEVELS, out_f32=False if u8out else out, out_u8=out if u8out else False,
                                       rects=rects)
        if timed: e[3].record(stream)
        # green column(s) of the (T,K,3) trace, read in place (row / column strides)
        bpm, kbin = eng.bpm_fft(means[:, :, 1] if poly else means[:, 0, 1], starts, lens, fps, ANALYSIS_BAND,
                                detrend=vhr.DETREND_F32, mode=vhr.FFT_ANALYSIS, max_len=T)
        if timed:
            e[4].record(stream)
            for name, a, b in (("pyrdown", 0, 1), ("bandpass", 1, 2), ("collapse", 2, 3), ("bpm", 3, 4)):
                kernel_events[name].append((e[a], e[b]))
        return bpm, kbin

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(Wm):
        step(i, False)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    t_mark0 = sampler.mark()
    e_start, e_end = ev(), ev()
    barrier()
    e_start.record(stream)
    results = []
    for i in range(K):
        results.append(step(Wm + i, True))
    e_end.record(stream)
    barrier()
    t_mark1 = sampler.mark()
    launches = eng.launch_count() - launches0
    clocks = sampler.stop(t_mark0, t_mark1) if rank == 0 else None
    elapsed_ms = e_start.elapsed_time(e_end)
    t_el = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_el, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t_el.item())
    value = world * K * T / (elapsed_ms / 1e3)

    kt = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in kernel_events.items()}
    col_ms = kt["collapse"]
    achieved = collapse_bytes_per_frame * T / (col_ms / 1e3) / 1e9
    path_ms = kt["pyrdown"] + kt["bandpass"] + kt["collapse"]
    path_achieved = bytes_per_frame * T / (path_ms / 1e3) / 1e9

    # ---- BPM check: every timed clip must give the CPU oracle's BPM and peak bin for that clip ------------
    gold = oracle_goldens(args.workload, args.roi)
    n_fft = T
    bpm_ok = True
    n_checked = 0
    bpm_local = {}
    for i, (bpm, kbin) in enumerate(results):
        c = my_clips[(Wm + i) % n_res]
        got, got_k = float(bpm[0].item()), int(kbin[0].item())
        if c in gold:
            bpm_ok &= (got == gold[c][0] and got_k == gold[c][1])
            n_checked += 1
        else:       # no oracle record for this clip (other workloads at N > 1): the injected pulse's bin
            k_true = int(round(pulse_hz(c) * n_fft / fps))
            bpm_ok &= abs(got - k_true * (1.0 / (n_fft * (1.0 / fps))) * 60.0) < 1e-9
        bpm_local[(Wm + i) % n_res] = [got]
    gathered = parallel.gather_results({my_clips[j]: v for j, v in bpm_local.items()}, world * n_res, 1, device=dev)

    # ---- e2e through the host-buffer call ----------------------------------------------------
    e2e = None
    if not args.no_e2e:
        # one clip per rank in pinned host memory (11.2 GB at 1080p); if the host cannot pin that much for
        # every rank, the same call runs on the first third of the clip (frames/s is per frame either way)
        T_e = T
        try:
            h_frames = torch.empty((T_e, H, W, 3), dtype=torch.uint8, pin_memory=True)
        except RuntimeError:
            torch.cuda.empty_cache()
            T_e = max(T // 3, 1)
            h_frames = torch.empty((T_e, H, W, 3), dtype=torch.uint8, pin_memory=True)
        if world > 1:                 # every rank runs the same number of frames
            t_min = torch.tensor([T_e], dtype=torch.int64, device=dev)
            dist.all_reduce(t_min, op=dist.ReduceOp.MIN)
            T_e = int(t_min.item())
            h_frames = h_frames[:T_e]
        h_frames.copy_(clips[0][:T_e])
        torch.cuda.synchronize()
        clips.clear()                 # free the device-timed buffers before the host path allocates its own
        out = lvl = None
        torch.cuda.empty_cache()
        fr_np = h_frames.numpy()
        rects_np = np.tile(rect, (T_e, 1, 1)).astype(np.int32)
        polys_e, nverts_e = np.ascontiguousarray(polys_np[:T_e]), np.ascontiguousarray(nverts_np[:T_e])
        lens_e = torch.full((1,), T_e, dtype=torch.int32, device=dev)
        ke = max(2, min(args.e2e_steps, K))

        def e2e_step():
            if poly:                                                                       # H2D + kernels + D2H
                means = eng.evm_roi_host(fr_np, fps, None, LEVELS, F_LO, F_HI, ALPHA, polys_np=polys_e, nverts_np=nverts_e)
                sig = means[:, :, 1]
            else:
                means = eng.evm_roi_host(fr_np, fps, rects_np, LEVELS, F_LO, F_HI, ALPHA)
                sig = means[:, 0, 1]
            bpm, _ = eng.bpm_fft(sig, starts, lens_e, fps, ANALYSIS_BAND, detrend=vhr.DETREND_F32, max_len=T_e)
            return float(bpm[0].item())                                                    # D2H of the result

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            b = e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t_e = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        dt = float(t_e.item())
        roi_bytes = int(polys_e.nbytes + nverts_e.nbytes) if poly else int(T_e * 16)
        e2e = {"value": world * ke * T_e / dt, "unit": "frames/s", "h2d_bytes_per_step": int(T_e * H * W * 3) + roi_bytes,
               "d2h_bytes_per_step": int(T_e * 24 + 8), "steps": ke, "ms_per_step": 1e3 * dt / ke, "bpm": b,
               "frames_per_step_per_gpu": T_e,
               "api": ("Engine.evm_roi_host (vhr_evm_poly_host)" if poly else "Engine.evm_roi_host (vhr_evm_roi_host)")
                      + " + Engine.bpm_fft, pinned host frames; ROI-only collapse (no magnified frame is materialised)"}

    # ---- CPU baseline (rank 0, N == 1 only) ---------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        T_s = 240 if W >= 1280 else min(T, 150)
        sp = vhr.SynthSpec(T=T_s, H=H, W=W, fps=fps, pulse_hz=1.2, seed=0, clip=0)
        fr_s = eng.synth_clip(sp).cpu().numpy()       # bit-identical to the oracle's generator (tests)
        eng.trim()                                    # give the host-path arena back before the CPU leg
        v, times, cores, bpm_cpu = cpu_reference_run(W, H, fps, T_s, steps=1, warmup=0, frames=fr_s)
        cpu = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": f"{T_s}-frame {W}x{H} sub-clip, one pass of the oracle EVM port (cv2 float32 pyrDown/pyrUp with "
                         f"{cores} threads + np.fft float64) + reference ROI mean + FFT-peak BPM; {times[0]:.1f} s",
               "bpm": bpm_cpu}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": desc.replace("fp32 out", "uint8 out (--out u8; not the contract configuration)") if u8out else desc,
                           "levels": LEVELS, "band_hz": [F_LO, F_HI], "alpha": ALPHA,
                           "frames_per_step_per_gpu": T, "resident_clips_per_gpu": n_res,
                           "roi": "forehead + 2 cheeks, 36-vertex polygons (fused row masks)" if poly else "1 cheek rectangle",
                           "bpm": "whole-clip window, float32 detrend + FFT peak",
                           "l2": "inputs (11.2 GB/clip at 1080p) larger than L2; no flush", "parallelism": f"clip-sharded x{world}",
                           "host_affinity": affinity},
                "roofline": {"bound": "hbm", "kernel": "collapse_sep_kernel", "achieved": achieved, "peak": peak_gbs,
                             "unit": "GB/s", "frac": achieved / peak_gbs,
                             "traffic": recorded_traffic("collapse_sep_kernel")[0] if args.workload == "c4" and not u8out else None,
                             "traffic_source": f"ncu --set full capture of the same command, {recorded_traffic('collapse_sep_kernel')[1]}",
                             "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": collapse_bytes_per_frame * T, "ms_per_launch": col_ms},
                "path_roofline": {"stages": "pyrdown+bandpass+collapse", "bytes_per_frame": bytes_per_frame,
                                  "achieved": path_achieved, "frac": path_achieved / peak_gbs, "unit": "GB/s"},
                "kernel_ms": kt, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches) * world,
                "clocks": clocks, "bpm_ok": bool(bpm_ok),
                "bpm_check": f"{n_checked} of {len(results)} timed clips on rank 0 against the CPU oracle's BPM and bin "
                             "(tests/golden/configs.npz); the rest against the injected pulse's bin",
                "bpm_gathered": None if gathered is None else [round(float(x), 4) for x in gathered[:, 0]]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--resident", type=int, default=3, help="distinct clips resident in HBM per GPU")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--affinity", default="off", choices=["off", "on", "auto"],
                    help="pin each rank to its share of the host cores (auto: when N > 1)")
    ap.add_argument("--out", default="f32", choices=["f32", "u8"], help="magnified frames: float32 (the BASELINE contract) or saturated uint8")
    ap.add_argument("--roi", default="rect", choices=["rect", "poly"], help="ROI stage: cheek rectangle, or forehead + cheek polygons")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
