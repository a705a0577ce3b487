#!/usr/bin/env python
"""Every kernel of the library once, at small shapes, for `compute-sanitizer --tool memcheck python tools/sanitize_smoke.py`."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import video_heart_rate_b200 as vhr
    from video_heart_rate_b200.pipeline import ANALYSIS_BAND, VIDEO_BAND, design_filters
    eng = vhr.Engine(0)
    rng = np.random.default_rng(0)
    for (T, H, W, L) in [(6, 72, 128, 4), (5, 70, 192, 3), (4, 61, 67, 3), (3, 40, 64, 2), (3, 36, 48, 1), (2, 130, 320, 6)]:
        fr = torch.as_tensor(rng.integers(0, 256, (T, H, W, 3), dtype=np.uint8), device=eng.tdev)
        rects = np.tile(np.array([W // 4, H // 4, 3 * W // 4, 3 * H // 4], dtype=np.int32), (T, 1, 1))
        r = eng.evm(fr, 30.0, L, 0.7, 4.0, 50.0, rects=rects, out_f32=True, out_u8=True)
        assert torch.isfinite(r["out_f32"]).all()
    spec = vhr.SynthSpec(T=150, H=72, W=128, fps=5.0, pulse_hz=1.2, seed=0)
    clip = eng.synth_clip(spec)
    lvl = eng.pyrdown(clip, 3)
    eng.bandpass(lvl, 5.0, 0.7, 4.0, 50.0)
    eng.bandpass(eng.pyrdown(clip[:149], 3), 5.0, 0.7, 2.0, 1.0)          # T = 149: DFT fallback
    T, H, W = 150, 72, 128
    rect = np.tile(np.array([30, 20, 90, 50], dtype=np.int32), (T, 1, 1))
    m = eng.roi_mean_rect(clip, rect)
    poly = np.tile(np.array([[20, 10], [100, 12], [110, 60], [60, 70], [15, 50]], dtype=np.int32), (T, 1, 1, 1))
    nv = np.full((T, 1), 5, dtype=np.int32)
    eng.roi_mean_poly(clip, poly, nv)
    eng.roi_mean_poly(clip.float(), poly, nv)
    eng.poly_mask(4, H, W, poly[:4], nv[:4])
    g = m[:, 0, 1].contiguous()
    eng.bpm_fft(g, [0, 10], [150, 100], 5.0, ANALYSIS_BAND, detrend=vhr.DETREND_F32, mode=vhr.FFT_ANALYSIS)
    f = design_filters(30.0, VIDEO_BAND)
    x = np.sin(2 * np.pi * 1.2 * np.arange(300) / 30.0) + 0.1 * rng.standard_normal(300)
    eng.bpm_welch(x, [0], [300], 30.0, VIDEO_BAND, vhr.DETREND_F64, vhr.FILT_SOS, f["butter"], want_filtered=True)
    eng.bpm_welch(x, [0], [300], 30.0, VIDEO_BAND, vhr.DETREND_F64, vhr.FILT_FIR, f["fir"], want_filtered=True)
    fr_np = rng.integers(0, 256, (8, 72, 128, 3), dtype=np.uint8)
    eng.evm_roi_host(fr_np, 30.0, np.tile(np.array([30, 20, 90, 50], dtype=np.int32), (8, 1, 1)), 3)
    torch.cuda.synchronize()
    print("sanitize smoke OK, launches:", eng.launch_count())
    eng.close()


if __name__ == "__main__":
    main()
