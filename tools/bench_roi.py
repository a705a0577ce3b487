#!/usr/bin/env python
"""Micro-benchmark of the stand-alone ROI kernels on one 1080p, 1800-frame clip resident in HBM:
rectangle mean (1 cheek rectangle), polygon mean on uint8 frames and on float32 (magnified) frames
(forehead + two cheeks, 36-vertex polygons following a jittered landmark track).  CUDA events."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def polygons(T, W, H, V=36, seed=0):
    rng = np.random.default_rng(seed)
    cx = W / 2 + np.cumsum(rng.normal(0, 0.4, T))
    cy = H / 2 + np.cumsum(rng.normal(0, 0.4, T))
    parts = [(0.0, -0.27, 0.135, 0.085), (-0.10, 0.06, 0.06, 0.10), (0.10, 0.06, 0.06, 0.10)]   # forehead, cheeks
    ang = np.linspace(0, 2 * np.pi, V, endpoint=False)
    out = np.zeros((T, len(parts), V, 2), dtype=np.int32)
    for k, (ox, oy, rx, ry) in enumerate(parts):
        x = cx[:, None] + ox * W + rx * W * np.cos(ang)[None] * (1 + 0.05 * np.sin(3 * ang)[None])
        y = cy[:, None] + oy * H + ry * H * np.sin(ang)[None]
        out[:, k, :, 0] = np.floor(x).astype(np.int32)
        out[:, k, :, 1] = np.floor(y).astype(np.int32)
    return out, np.full((T, len(parts)), V, dtype=np.int32)


def main():
    import torch
    import video_heart_rate_b200 as vhr
    from video_heart_rate_b200 import host
    eng = vhr.Engine(0)
    T, H, W = (int(os.environ.get("ROI_T", 1800)), 1080, 1920)
    spec = vhr.SynthSpec(T=T, H=H, W=W, fps=30.0, pulse_hz=1.2, seed=0, clip=0)
    fr = eng.synth_clip(spec)
    polys, nv = polygons(T, W, H)
    lm = spec.landmarks()
    rect = host.slice_rects(host.cheek_roi_clamped(host.bbox_clamped(lm[None], W, H), W, H), W, H)[0]
    rects = np.tile(rect, (T, 1, 1)).astype(np.int32)
    pd, nd = torch.as_tensor(polys, device=eng.tdev), torch.as_tensor(nv, device=eng.tdev)
    rd = torch.as_tensor(rects, device=eng.tdev)

    def timeit(fn, reps=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    res = {"T": T, "polygons": int(polys.shape[1]), "vertices": int(polys.shape[2])}
    res["rect_mean_u8_ms"] = timeit(lambda: eng.roi_mean_rect(fr, rd))
    res["poly_mean_u8_ms"] = timeit(lambda: eng.roi_mean_poly(fr, pd, nd))
    means, counts = eng.roi_mean_poly(fr, pd, nd)
    res["pixels_per_polygon"] = [int(c) for c in counts[0].tolist()]
    f32 = fr[: min(T, 300)].float()
    res["poly_mean_f32_ms_per_1800"] = timeit(lambda: eng.roi_mean_poly(f32, pd[: f32.shape[0]], nd[: f32.shape[0]])) * T / f32.shape[0]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
