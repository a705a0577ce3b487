"""Measurement plugin: Eulerian video magnification + ROI + BPM on the B200
(``measure(video_path) -> (N,2) [t_sec, bpm]``, contract of ``analysis/main.py:29-31``).

EVM parameters follow BASELINE.json (4-level Gaussian pyramid, 0.7-4 Hz ideal bandpass,
alpha = 50).  ROIs are the landmark polygons BASELINE.json's north_star names -- forehead and
two cheeks (``host.face_polygons``: through the FaceMesh landmarks when the side-car has the mesh
topology, otherwise ellipses in the reference's own ratio rectangles, rppg_VIDEO.py:102-103) --
or, with ``VHR_EVM_ROI=rect``, the reference's clamped cheek rectangle
(analysis/utils/roi.py:43-59).  The call is ROI-only: no magnified frame is written, and only the
image parts under a ROI are collapsed.  The green means of the magnified ROIs then go through the
same rolling window / float32 detrend / FFT-peak estimator as ``green_avg`` (green_avg.py:24-50);
with several ROIs the estimator keeps the trace with the strongest in-band peak, its own
multi-column rule (analysis/utils/estimate_bpm.py:59-64).
"""
from __future__ import annotations

import os

import numpy as np

from .green_avg_b200 import load_landmarks, read_video

LEVELS, BAND, ALPHA = 4, (0.7, 4.0), 50.0


def measure(video_path: str) -> np.ndarray:
    from video_heart_rate_b200 import DETREND_F32, FFT_ANALYSIS, default_engine, host
    from video_heart_rate_b200.pipeline import ANALYSIS_BAND
    import torch
    eng = default_engine()
    frames, fps = read_video(video_path)
    if not frames:
        return np.zeros((0, 2))
    lm, valid = load_landmarks(video_path, frames, fps)
    fr = torch.as_tensor(np.stack(frames), device=eng.tdev)
    T, H, W, _ = fr.shape
    usable = np.ones(T, dtype=bool)
    if valid is not None:
        lm, usable = host.hold_landmarks(lm, valid)
    lm = np.broadcast_to(lm, (T,) + lm.shape[-2:]) if lm.ndim == 2 else lm
    if os.environ.get("VHR_EVM_ROI", "poly") == "rect":
        rects = host.slice_rects(host.cheek_roi_clamped(host.bbox_clamped(lm, W, H), W, H), W, H)
        r = eng.evm(fr, fps, LEVELS, BAND[0], BAND[1], ALPHA, rects=rects[:, None, :], out_f32=False, out_u8=False)
    else:
        polys, nverts = host.face_polygons(lm, W, H)
        r = eng.evm(fr, fps, LEVELS, BAND[0], BAND[1], ALPHA, polys=polys, nverts=nverts, out_f32=False, out_u8=False)
    green = r["roi_mean"][:, :, 1]                       # (T,K) green means of the magnified ROIs
    idx = np.arange(T)[usable]
    green = green[torch.as_tensor(idx, device=eng.tdev)].contiguous()
    fi, st, ln = host.green_avg_windows(int(green.shape[0]), fps)
    if len(fi) == 0:
        return np.zeros((0, 2))
    bpm, _ = eng.bpm_fft(green, st, ln, fps, ANALYSIS_BAND, detrend=DETREND_F32, mode=FFT_ANALYSIS)
    bpm = bpm.cpu().numpy()
    ok = ~np.isnan(bpm)
    return np.column_stack([(idx[fi] * (1 / fps))[ok], bpm[ok]])
