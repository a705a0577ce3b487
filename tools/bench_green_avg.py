#!/usr/bin/env python
"""The reference's own measurement (analysis/measurement/green_avg.py: cheek-rectangle mean green per
frame, 10 -> 30 s growing window, float32 detrend, FFT peak per frame) on one 1080p 60 s clip resident
in HBM: time of the rectangle means, of the 1501 windowed BPM estimates, and of the whole measure()."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import video_heart_rate_b200 as vhr
    from video_heart_rate_b200 import host
    from video_heart_rate_b200.pipeline import ANALYSIS_BAND, green_avg_measure
    eng = vhr.Engine(0)
    T, H, W, fps = 1800, 1080, 1920, 30.0
    spec = vhr.SynthSpec(T=T, H=H, W=W, fps=fps, pulse_hz=1.2, seed=0)
    fr = eng.synth_clip(spec)
    lm = spec.landmarks()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def timeit(fn, reps=5):
        fn(); torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    rect = host.slice_rects(host.cheek_roi_clamped(host.bbox_clamped(lm[None], W, H), W, H), W, H)[0]
    rects = torch.as_tensor(np.tile(rect, (T, 1, 1)).astype(np.int32), device=eng.tdev)
    res = {"rect": [int(v) for v in rect]}
    res["rect_mean_ms"] = timeit(lambda: eng.roi_mean_rect(fr, rects))
    green = eng.roi_mean_rect(fr, rects)[:, 0, 1].contiguous()
    fi, st, ln = host.green_avg_windows(T, fps)
    std, lnd = torch.as_tensor(st, device=eng.tdev), torch.as_tensor(ln, device=eng.tdev)
    res["windows"] = int(len(fi))
    res["bpm_fft_ms"] = timeit(lambda: eng.bpm_fft(green, std, lnd, fps, ANALYSIS_BAND, detrend=vhr.DETREND_F32,
                                                   mode=vhr.FFT_ANALYSIS, max_len=int(max(ln))))
    t0 = time.perf_counter()
    out = green_avg_measure(eng, fr, fps, lm)
    torch.cuda.synchronize()
    res["measure_wall_ms"] = 1e3 * (time.perf_counter() - t0)
    res["rows"] = int(out.shape[0])
    res["bpm_last"] = float(out[-1, 1])
    print(json.dumps(res))


if __name__ == "__main__":
    main()
