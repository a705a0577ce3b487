"""Measurement plugin: drop-in for ``analysis/measurement/ica.py`` on the B200
(``measure(video_path) -> (N,2) [t_sec, bpm]``, contract of ``analysis/main.py:29-31``).

Same windowing (10 s rolling, 5 s acquisition, ica.py:10-11,27-29), same float32 / ddof=1 standardisation
(:56-61), FastICA with the reference's arguments (:36-44) as a batched CUDA kernel (one warp per window),
windows that do not converge skipped (:64-69), ``estimate_bpm`` over the three sources (:72).
Tolerance contract, not bit parity: scikit-learn runs float32 LAPACK for float32 input, the kernel float64;
on every window where scikit-learn converges the spectral-peak bin is the same (tests), but the kernel also
converges on windows where scikit-learn stops at ``max_iter`` -- those rows are emitted here and skipped there.
"""
from __future__ import annotations

import numpy as np

from .green_avg_b200 import load_landmarks, read_video


def measure(video_path: str) -> np.ndarray:
    from video_heart_rate_b200 import default_engine
    from video_heart_rate_b200.pipeline import ica_measure
    frames, fps = read_video(video_path)
    if not frames:
        return np.zeros((0, 2))
    lm, valid = load_landmarks(video_path, frames, fps)
    return ica_measure(default_engine(), np.stack(frames), fps, lm, valid)
