"""Synthetic clip specification (product side of the generator in ``csrc/synth.cu``).

A clip is a pure integer function of (seed, clip, t, y, x, c) -- see the kernel for the
formula -- so that clips far larger than host memory (config c4: 64 x 11.2 GB) can be
generated on the device and regenerated bit for bit by the CPU oracle for parity checks.
This module only derives the integer parameters; it evaluates no pixels.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

BYTE_SUM_STD = math.sqrt(4.0 * (256.0 ** 2 - 1.0) / 12.0)
BG_RGB = (128, 128, 128)
SKIN_RGB = (200, 150, 130)
PULSE_AMP_RGB = (0.75, 1.5, 0.5)


@dataclasses.dataclass(frozen=True)
class SynthSpec:
    T: int
    H: int
    W: int
    fps: float
    pulse_hz: float
    seed: int = 0
    clip: int = 0
    noise_sigma: float = 2.0
    face: tuple | None = None      # half-open [x0,x1) x [y0,y1); default central 50 % x 70 %

    def face_rect(self):
        if self.face is not None:
            return tuple(int(v) for v in self.face)
        return (self.W // 4, (self.H * 15) // 100, self.W - self.W // 4, self.H - (self.H * 15) // 100)

    def noise_gain(self) -> int:
        return int(round(self.noise_sigma * 256.0 / BYTE_SUM_STD))

    def pulse_table(self) -> np.ndarray:
        t = np.arange(self.T, dtype=np.float64)
        s = np.sin(2.0 * np.pi * self.pulse_hz * t / self.fps)
        return np.rint(256.0 * s[:, None] * np.asarray(PULSE_AMP_RGB, dtype=np.float64)[None, :]).astype(np.int32)

    def base_q8(self) -> np.ndarray:
        return (np.asarray([BG_RGB, SKIN_RGB], dtype=np.int32) * 256).astype(np.int32)

    def landmarks(self) -> np.ndarray:
        """(4,2) normalised (x,y) landmarks = face-rectangle corners at pixel centres."""
        x0, y0, x1, y1 = self.face_rect()
        xs = np.array([x0 + 0.5, x1 - 0.5]) / self.W
        ys = np.array([y0 + 0.5, y1 - 0.5]) / self.H
        return np.array([[xs[0], ys[0]], [xs[1], ys[0]], [xs[1], ys[1]], [xs[0], ys[1]]])

    def polygons(self):
        """Forehead / left-cheek / right-cheek int32 outlines inside the face rectangle
        (stand-ins for MediaPipe landmark polygons) -> (K, Vmax, 2) int32, (K,) int32."""
        x0, y0, x1, y1 = self.face_rect()
        fw, fh = x1 - x0, y1 - y0

        def P(*uv):
            return [[x0 + int(u * fw), y0 + int(v * fh)] for u, v in uv]

        polys = [P((0.25, 0.06), (0.40, 0.03), (0.60, 0.03), (0.75, 0.06), (0.78, 0.20), (0.60, 0.24), (0.50, 0.21),
                   (0.40, 0.24), (0.22, 0.20)),
                 P((0.14, 0.45), (0.30, 0.42), (0.40, 0.52), (0.36, 0.66), (0.24, 0.70), (0.15, 0.60)),
                 P((0.86, 0.45), (0.70, 0.42), (0.60, 0.52), (0.64, 0.66), (0.76, 0.70), (0.85, 0.60))]
        vmax = max(len(p) for p in polys)
        arr = np.zeros((len(polys), vmax, 2), dtype=np.int32)
        nv = np.zeros(len(polys), dtype=np.int32)
        for k, p in enumerate(polys):
            arr[k, :len(p)] = p
            nv[k] = len(p)
        return arr, nv
