"""Measurement plugin: the ``green_avg_psd_plot`` pipeline on the B200 -- cheek ROI green mean,
float32 z-score, Butterworth(2) sosfiltfilt, periodogram peak over a 10 s rolling window
(``analysis/measurement/green_avg_psd_plot.py:117-185``; the interactive PSD plots and the ROI-mean
cache of that module are out of scope).  ``measure(video_path) -> (N,2) [t_sec, bpm]`` with NaN
BPM during the acquisition period, exactly like the reference.
"""
from __future__ import annotations

import numpy as np

from .green_avg_b200 import load_landmarks, read_video


def measure(video_path: str) -> np.ndarray:
    from video_heart_rate_b200 import default_engine
    from video_heart_rate_b200.pipeline import green_avg_psd_series, green_avg_trace
    eng = default_engine()
    frames, fps = read_video(video_path)
    if not frames:
        return np.zeros((0, 2))
    lm, valid = load_landmarks(video_path, frames, fps)
    means, usable = green_avg_trace(eng, np.stack(frames), lm, valid)
    green = means[:, 1].contiguous()
    if not usable.all():
        import torch
        green = green[torch.as_tensor(np.nonzero(usable)[0], device=eng.tdev)].contiguous()
    return green_avg_psd_series(eng, green, fps)
