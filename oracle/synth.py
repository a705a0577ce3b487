"""Synthetic clip generator -- NumPy side (oracle; test infrastructure).

A clip is a PURE INTEGER FUNCTION of (seed, clip, t, y, x, c) so that the CUDA
generator (``csrc/synth.cu``) reproduces the same bytes bit for bit; the reference has
no generator (its videos are git-ignored, ``.gitignore:1-5``), the spec below is ours
(SURVEY.md section 7.1 / 8d):

    h_t   = mix32(seed*0x9E3779B1 + clip*0x85EBCA77 + t*0xC2B2AE3D + 0x165667B1)   (u32)
    r     = mix32(h_t ^ (((y*W + x)*3 + c) * 0x27D4EB2F))                            (u32)
    s     = byte0(r) + byte1(r) + byte2(r) + byte3(r) - 510          in [-510, 510]
    v_q8  = base_q8[face][c] + face * pulse_q8[t][c] + s * noise_gain              (i32)
    pixel = clamp((v_q8 + 128) >> 8, 0, 255)                     (arithmetic shift)

``face`` is 1 inside the half-open face rectangle.  ``pulse_q8`` is a host-computed
int32 table (T,3) -- ``round(256 * A_c * sin(2 pi f t / fps))`` -- handed to both
generators, so no transcendental is evaluated on the device.  The byte sum has standard
deviation 147.8; ``noise_gain = round(sigma * 256 / 147.8)`` gives noise of ``sigma``
LSB, which also dithers the sub-LSB pulse through the u8 quantiser.
"""
from __future__ import annotations

import dataclasses
import numpy as np

BYTE_SUM_STD = float(np.sqrt(4.0 * (256.0**2 - 1.0) / 12.0))  # 147.80...

# (R, G, B) -- frames are uint8 RGB as BASELINE.json:north_star states.
BG_RGB = (128, 128, 128)
SKIN_RGB = (200, 150, 130)
PULSE_AMP_RGB = (0.75, 1.5, 0.5)  # LSB; strongest in G


@dataclasses.dataclass(frozen=True)
class SynthParams:
    T: int
    H: int
    W: int
    fps: float
    pulse_hz: float
    seed: int = 0
    clip: int = 0
    noise_sigma: float = 2.0
    # half-open face rectangle [x0,x1) x [y0,y1); default = central 50 % x 70 %
    face: tuple | None = None

    def face_rect(self):
        if self.face is not None:
            return tuple(int(v) for v in self.face)
        x0 = self.W // 4
        x1 = self.W - self.W // 4
        y0 = (self.H * 15) // 100
        y1 = self.H - (self.H * 15) // 100
        return x0, y0, x1, y1

    def noise_gain(self) -> int:
        return int(round(self.noise_sigma * 256.0 / BYTE_SUM_STD))

    def pulse_table(self) -> np.ndarray:
        """(T,3) int32, Q8 LSB.  Host float64 -> int; the same table feeds the GPU."""
        t = np.arange(self.T, dtype=np.float64)
        s = np.sin(2.0 * np.pi * self.pulse_hz * t / self.fps)
        amp = np.asarray(PULSE_AMP_RGB, dtype=np.float64)
        return np.rint(256.0 * s[:, None] * amp[None, :]).astype(np.int32)

    def base_q8(self) -> np.ndarray:
        """(2,3) int32: [background, skin] * 256."""
        return (np.asarray([BG_RGB, SKIN_RGB], dtype=np.int32) * 256).astype(np.int32)

    def landmarks(self) -> np.ndarray:
        """(4,2) float64 normalised (x,y) 'landmarks' = face-rect corners at pixel
        centres (the +0.5 keeps ``int(min(xs)*w)`` away from a float round-down)."""
        x0, y0, x1, y1 = self.face_rect()
        xs = np.array([x0 + 0.5, x1 - 0.5], dtype=np.float64) / self.W
        ys = np.array([y0 + 0.5, y1 - 0.5], dtype=np.float64) / self.H
        return np.array([[xs[0], ys[0]], [xs[1], ys[0]], [xs[1], ys[1]], [xs[0], ys[1]]])


def mix32(x: np.ndarray) -> np.ndarray:
    """lowbias32 integer finaliser on uint32 arrays (wrap-around arithmetic)."""
    x = x.astype(np.uint32, copy=True)
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7FEB352D)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846CA68B)
    x ^= x >> np.uint32(16)
    return x


def frame_key(seed: int, clip: int, t: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        k = (np.uint32(seed & 0xFFFFFFFF) * np.uint32(0x9E3779B1)
             + np.uint32(clip & 0xFFFFFFFF) * np.uint32(0x85EBCA77)
             + t.astype(np.uint32) * np.uint32(0xC2B2AE3D) + np.uint32(0x165667B1))
    return mix32(k)


def synth_frames(p: SynthParams, t0: int = 0, t1: int | None = None) -> np.ndarray:
    """uint8 (t1-t0, H, W, 3) RGB frames of the clip ``p`` (frames t0..t1-1)."""
    t1 = p.T if t1 is None else t1
    H, W = p.H, p.W
    x0, y0, x1, y1 = p.face_rect()
    pulse = p.pulse_table()
    base = p.base_q8()
    gain = np.int32(p.noise_gain())
    yy, xx = np.mgrid[0:H, 0:W]
    face = ((xx >= x0) & (xx < x1) & (yy >= y0) & (yy < y1))
    with np.errstate(over="ignore"):
        idx = ((yy * W + xx)[:, :, None] * 3 + np.arange(3)[None, None, :]).astype(np.uint32)
        idxm = idx * np.uint32(0x27D4EB2F)
    out = np.empty((t1 - t0, H, W, 3), dtype=np.uint8)
    keys = frame_key(p.seed, p.clip, np.arange(t0, t1))
    facei = face.astype(np.int32)[:, :, None]
    base_img = np.where(face[:, :, None], base[1][None, None, :], base[0][None, None, :]).astype(np.int32)
    for i, t in enumerate(range(t0, t1)):
        r = mix32(keys[i] ^ idxm)
        s = ((r & np.uint32(255)) + ((r >> np.uint32(8)) & np.uint32(255))
             + ((r >> np.uint32(16)) & np.uint32(255)) + (r >> np.uint32(24))).astype(np.int32) - 510
        v = base_img + facei * pulse[t][None, None, :] + s * gain
        out[i] = np.clip((v + 128) >> 8, 0, 255).astype(np.uint8)
    return out
