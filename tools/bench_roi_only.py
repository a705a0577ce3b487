#!/usr/bin/env python
"""What the measurement plugins' ROI-only mode is worth: one 1080p / 1800-frame clip resident in
HBM through Engine.evm with and without a frame output, rectangle and polygon ROIs; CUDA-event time per stage and the
identity of the ROI traces of both calls (bit for bit).  One JSON line."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import video_heart_rate_b200 as vhr
    from video_heart_rate_b200 import host
    eng = vhr.Engine(0)
    T, H, W = int(os.environ.get("ROI_T", 1800)), 1080, 1920
    spec = vhr.SynthSpec(T=T, H=H, W=W, fps=30.0, pulse_hz=1.2, seed=0, clip=0)
    fr = eng.synth_clip(spec)
    lm = np.broadcast_to(spec.landmarks(), (T, 4, 2))
    rects = torch.as_tensor(host.slice_rects(host.cheek_roi_clamped(host.bbox_clamped(lm, W, H), W, H), W, H)[:, None, :], device=eng.tdev)
    polys_np, nv_np = host.face_polygons(lm, W, H)
    polys, nv = torch.as_tensor(polys_np, device=eng.tdev), torch.as_tensor(nv_np, device=eng.tdev)
    out = torch.empty((T, H, W, 3), dtype=torch.float32, device=eng.tdev)
    lvl = eng.pyrdown(fr, 4)
    filt = eng.bandpass(lvl, 30.0, 0.7, 4.0, 50.0)

    def timeit(fn, reps=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            r = fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, r

    res = {"T": T, "frame": [W, H], "roi_px": {"rect": int((rects[0, 0, 2] - rects[0, 0, 0]) * (rects[0, 0, 3] - rects[0, 0, 1]))}}
    ms_full, a = timeit(lambda: eng.collapse(filt, fr, 4, out_f32=out, rects=rects))
    ms_only, b = timeit(lambda: eng.collapse(filt, fr, 4, out_f32=False, rects=rects))
    res["collapse_rect_ms"] = {"full_frame_f32": ms_full, "roi_only": ms_only, "traces_bit_identical": bool(torch.equal(a[2], b[2]))}
    ms_full, a = timeit(lambda: eng.collapse(filt, fr, 4, out_f32=out, polys=polys, nverts=nv, want_counts=True))
    ms_only, b = timeit(lambda: eng.collapse(filt, fr, 4, out_f32=False, polys=polys, nverts=nv, want_counts=True))
    res["collapse_poly_ms"] = {"full_frame_f32": ms_full, "roi_only": ms_only, "traces_bit_identical": bool(torch.equal(a[2], b[2]))}
    res["roi_px"]["poly"] = [int(c) for c in a[3][0].tolist()]
    ms_evm_full, _ = timeit(lambda: eng.evm(fr, 30.0, 4, 0.7, 4.0, 50.0, rects=rects, out_f32=out), reps=3)
    ms_evm_only, _ = timeit(lambda: eng.evm(fr, 30.0, 4, 0.7, 4.0, 50.0, polys=polys, nverts=nv, out_f32=False), reps=3)
    res["evm_ms"] = {"rect_full_frame": ms_evm_full, "poly_roi_only (evm_b200.measure's call)": ms_evm_only}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
