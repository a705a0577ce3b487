"""Degradations and metric of the analysis harness -- restatement (oracle; test infrastructure).

``quantise_colour``  analysis/degradation/colour_quantisation.py:12-25 (verbatim arithmetic)
``add_noise``        analysis/degradation/colour_noise.py:11-24 with the reference's
                     ``np.random.normal`` replaced by a counter-based 12-term Irwin-Hall draw (sum of
                     twelve hash bytes: mean 0, std sigma, tails to +-5.98 sigma, excess kurtosis
                     -0.1; same clip + ``astype(uint8)`` truncation), pure integer arithmetic so the
                     CUDA kernel can be held to it bit for bit.
``align_truth`` / ``mae``  analysis/utils/video_io.py:80-106, analysis/metrics/mae.py:32-36
"""
from __future__ import annotations

import numpy as np

from .synth import mix32

NOISE_SUM_STD = float(np.sqrt(12.0 * (256.0**2 - 1.0) / 12.0))   # twelve uniform bytes: 255.998


def quantise_colour(frame: np.ndarray, bits: int) -> np.ndarray:
    levels = 2 ** bits
    scale = 256 // levels
    with np.errstate(divide="ignore"):
        return (frame // scale) * scale


def noise_gain(sigma: float) -> int:
    """Q16 gain per unit of the centred twelve-byte sum."""
    return int(round(sigma * 65536.0 / NOISE_SUM_STD))


def add_noise(frames: np.ndarray, sigma: float, seed: int = 0, clip: int = 0, t0: int = 0) -> np.ndarray:
    T, H, W, C = frames.shape
    gain = noise_gain(sigma)
    with np.errstate(over="ignore"):
        idx = np.arange(H * W * C, dtype=np.uint32) * np.uint32(0x27D4EB2F)
        t = np.arange(t0, t0 + T).astype(np.uint32)
        keys = mix32(np.uint32(seed & 0xFFFFFFFF) * np.uint32(0x9E3779B1) + np.uint32(clip & 0xFFFFFFFF) * np.uint32(0x85EBCA77)
                     + t * np.uint32(0xC2B2AE3D) + np.uint32(0x3C6EF372))
    out = np.empty_like(frames)
    flat = frames.reshape(T, -1).astype(np.int32)
    for i in range(T):
        s = np.full(idx.shape, -1530, dtype=np.int32)
        for w in range(3):
            with np.errstate(over="ignore"):
                r = mix32((keys[i] + np.uint32(w) * np.uint32(0x9E3779B9)) ^ idx)
            s += ((r & np.uint32(255)) + ((r >> np.uint32(8)) & np.uint32(255)) + ((r >> np.uint32(16)) & np.uint32(255))
                  + (r >> np.uint32(24))).astype(np.int32)
        v = (flat[i] * 65536 + s * gain) >> 16
        out[i] = np.clip(v, 0, 255).astype(np.uint8).reshape(H, W, C)
    return out


def align_truth(t_truth, hr_truth, measured) -> np.ndarray:
    """video_io.interpolate_hr_to_frames: last truth time <= t, clamped -> (N,2) [t, hr]."""
    t_truth = np.asarray(t_truth, dtype=float)
    hr_truth = np.asarray(hr_truth, dtype=float)
    measured = np.asarray(measured)
    t_meas = measured[:, 0].astype(float)
    idx = np.searchsorted(t_truth, t_meas, side='right') - 1
    idx = np.clip(idx, 0, len(t_truth) - 1)
    return np.column_stack([t_meas, hr_truth[idx]])


def mae(t_truth, hr_truth, measured) -> float:
    aligned = align_truth(t_truth, hr_truth, measured)
    return float(np.mean(np.abs(np.asarray(measured)[:, 1].astype(float) - aligned[:, 1].astype(float))))
