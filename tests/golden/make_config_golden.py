"""Generate tests/golden/configs.npz: the CPU oracle's ROI traces / BPMs / peak bins for the
BASELINE.json configurations that are benchmarked (c2, c3, c4, c5), so that the GPU path and
bench.py's ``bpm_ok`` are held to the ORACLE on those shapes and not to the injected pulse.

    python tests/golden/make_config_golden.py [--procs 7] [--only c4,c4poly,c2,c3,c5,bench] [--c4-clips 64]

Per clip / window the oracle is: oracle.fast.synth_frames (bit-identical to the CUDA generator;
tests) [-> oracle.fast.add_noise for c5] -> oracle.evm.evm_roi_trace_streaming (cv2.pyrDown /
cv2.pyrUp float32 + np.fft float64, the arithmetic of evm_clip_cv2) on the clamped cheek rectangle
(oracle.roi, analysis/utils/roi.py:43-59) -> float32 detrend (analysis/measurement/green_avg.py:42-43)
-> oracle.bpm.estimate_bpm_analysis (analysis/utils/estimate_bpm.py:12-65).
Needs cv2 + numpy only (no /root/reference); ~20 minutes on 8 cores for everything.
"""
from __future__ import annotations

import argparse
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

LEVELS, F_LO, F_HI, ALPHA = 4, 0.7, 4.0, 50.0
RES = {144: 256, 240: 426, 360: 640, 480: 854, 720: 1280, 1080: 1920}
FPS = [5, 10, 15, 25, 30]
NOISE = [0, 5, 10, 20, 40]


def c4_pulse_hz(clip: int) -> float:
    return 0.8 + (clip % 64) * (2.4 / 63.0)


def c5_windows(n=512):
    pulses = [1.0, 1.33, 1.75, 2.2]
    out = []
    for f in pulses:
        for h in RES:
            for fps in FPS:
                for sg in NOISE:
                    out.append((h, fps, sg, f, len(out)))
    return out[:n]


def _host_module():
    """video-heart-rate_b200/host.py (NumPy-only ROI geometry) without importing the CUDA package."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("vhr_host", os.path.join(ROOT, "video-heart-rate_b200", "host.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def oracle_unit(job):
    """job = (tag, synth kwargs, noise_sigma, t0, t1) -> (tag, trace (n,) float64 [(n,K) for polygon jobs], bpm, bin)."""
    import cv2
    cv2.setNumThreads(1)
    from oracle import bpm as obpm, evm as oevm, fast, roi as oroi, synth as osynth
    tag, kw, sigma, t0, t1 = job
    p = osynth.SynthParams(**kw)
    n = t1 - t0
    lm = p.landmarks()
    rect = oroi.cheek_roi_from_bbox(oroi.bbox_from_landmarks_clamped(lm[:, 0], lm[:, 1], p.W, p.H), p.W, p.H)
    rects = np.tile(np.array(rect, dtype=np.int32), (n, 1, 1))

    def chunks():
        step = max(1, (64 << 20) // (p.H * p.W * 3))
        for a in range(t0, t1, step):
            b = min(t1, a + step)
            fr = fast.synth_frames(p, a, b)
            if sigma > 0:
                fr = fast.add_noise(fr, sigma, seed=p.seed, clip=p.clip, t0=a)
            yield fr

    if tag[0].endswith("poly"):
        # forehead + two cheeks: the product's own host geometry (pure NumPy) turns the landmarks into polygons;
        # rasterisation, masked means and the multi-column estimator are the oracle's
        polys, nv = _host_module().face_polygons(np.broadcast_to(lm, (n,) + lm.shape), p.W, p.H)
        _, _, means, _ = oevm.evm_roi_trace_streaming(chunks, n, p.H, p.W, p.fps, None, LEVELS, F_LO, F_HI, ALPHA, polys=polys, nverts=nv)
        g = means[:, :, 1]
        g32 = g.astype(np.float32)
        bpm, k, _ = obpm.estimate_bpm_analysis(g32 - np.mean(g32, axis=0), p.fps)
        return tag, g, (np.nan if bpm is None else bpm), k
    _, _, means = oevm.evm_roi_trace_streaming(chunks, n, p.H, p.W, p.fps, rects, LEVELS, F_LO, F_HI, ALPHA)
    g = means[:, 0, 1]
    g32 = g.astype(np.float32)
    bpm, k, _ = obpm.estimate_bpm_analysis(g32 - np.mean(g32), p.fps)
    return tag, g, (np.nan if bpm is None else bpm), k


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", type=int, default=max(1, (os.cpu_count() or 2) - 1))
    ap.add_argument("--only", default="c4,c4poly,c2,c3,c5,bench")
    ap.add_argument("--c4-clips", type=int, default=64)
    ap.add_argument("--out", default=os.path.join(HERE, "configs.npz"))
    args = ap.parse_args()
    only = set(args.only.split(","))
    jobs = []
    if "c4" in only:
        for c in range(args.c4_clips):
            jobs.append((("c4", c), dict(T=1800, H=1080, W=1920, fps=30.0, pulse_hz=c4_pulse_hz(c), seed=c, clip=c), 0, 0, 1800))
    if "c4poly" in only:
        for c in (0, 1, 2, 21, 42, 63):
            jobs.append((("c4poly", c), dict(T=1800, H=1080, W=1920, fps=30.0, pulse_hz=c4_pulse_hz(c), seed=c, clip=c), 0, 0, 1800))
    if "c2" in only:
        jobs.append((("c2", 0), dict(T=1800, H=720, W=1280, fps=30.0, pulse_hz=1.2, seed=2), 0, 0, 1800))
    if "c3" in only:
        for w in range(51):
            jobs.append((("c3", w), dict(T=1800, H=480, W=640, fps=30.0, pulse_hz=1.4, seed=3), 0, 30 * w, 30 * w + 300))
    if "c5" in only:
        for (h, fps, sg, f, i) in c5_windows():
            jobs.append((("c5", i), dict(T=int(fps * 10.0), H=h, W=RES[h], fps=float(fps), pulse_hz=f, seed=i, clip=i),
                         sg, 0, int(fps * 10.0)))
    if "bench" in only:
        # bench.py --workload c1|c2|c3 (its clips 0..2: seed = clip = c, pulse c4_pulse_hz(c)); c4 is covered above
        for name, (W, H, fps, T) in (("bench_c1", (256, 144, 5.0, 150)), ("bench_c2", (1280, 720, 30.0, 1800)),
                                     ("bench_c3", (640, 480, 30.0, 1800))):
            for c in range(3):
                jobs.append(((name, c), dict(T=T, H=H, W=W, fps=fps, pulse_hz=c4_pulse_hz(c), seed=c, clip=c), 0, 0, T))
    # heaviest first so the pool drains evenly
    jobs.sort(key=lambda j: -(j[4] - j[3]) * j[1]["H"] * j[1]["W"])
    t0 = time.time()
    res = {}
    with mp.Pool(args.procs) as pool:
        for i, (tag, g, bpm, k) in enumerate(pool.imap_unordered(oracle_unit, jobs)):
            res[tag] = (g, bpm, k)
            print(f"[{time.time() - t0:7.1f}s] {i + 1}/{len(jobs)} {tag} bpm {bpm} bin {k}", flush=True)
    out = {}
    if os.path.exists(args.out):
        out = dict(np.load(args.out))
    for cfg in ("c4", "c4poly", "c2", "c3", "c5", "bench_c1", "bench_c2", "bench_c3"):
        ids = sorted(i for (c, i) in res if c == cfg)
        if not ids:
            continue
        out[f"{cfg}_ids"] = np.asarray(ids, dtype=np.int32)
        out[f"{cfg}_bpm"] = np.asarray([res[(cfg, i)][1] for i in ids], dtype=np.float64)
        out[f"{cfg}_bin"] = np.asarray([res[(cfg, i)][2] for i in ids], dtype=np.int32)
        if cfg != "c5":
            out[f"{cfg}_trace"] = np.stack([res[(cfg, i)][0] for i in ids]).astype(np.float64)
    np.savez_compressed(args.out, **out)
    print("wrote", args.out, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
