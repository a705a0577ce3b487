"""ctypes front of the oracle's C helpers (oracle/csrc/synth_ref.c) -- TEST INFRASTRUCTURE.

``synth_frames`` / ``add_noise`` return exactly what ``oracle.synth.synth_frames`` /
``oracle.degrade.add_noise`` return (tests/test_oracle.py), ~50x faster; they fall back to the
NumPy forms when the C library has not been built.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import degrade as _degrade, synth as _synth

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "liboracle_c.so")
        if not os.path.exists(path):
            try:
                from .build import build
                build()
            except Exception:
                _LIB = False
                return None
        lib = C.CDLL(path)
        lib.oracle_synth_frames.restype = None
        lib.oracle_synth_frames.argtypes = [C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        lib.oracle_add_noise.restype = None
        lib.oracle_add_noise.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int32, C.c_uint32, C.c_uint32,
                                         C.c_int]
        _LIB = lib
    return _LIB or None


def synth_frames(p: _synth.SynthParams, t0: int = 0, t1: int | None = None) -> np.ndarray:
    lib = _lib()
    if lib is None:
        return _synth.synth_frames(p, t0, t1)
    t1 = p.T if t1 is None else t1
    out = np.empty((t1 - t0, p.H, p.W, 3), dtype=np.uint8)
    face = np.asarray(p.face_rect(), dtype=np.int32)
    base = np.ascontiguousarray(p.base_q8(), dtype=np.int32)
    pulse = np.ascontiguousarray(p.pulse_table(), dtype=np.int32)
    lib.oracle_synth_frames(p.seed & 0xFFFFFFFF, p.clip & 0xFFFFFFFF, t0, t1 - t0, p.H, p.W, face.ctypes.data,
                            base.ctypes.data, p.noise_gain(), pulse.ctypes.data, out.ctypes.data)
    return out


def add_noise(frames: np.ndarray, sigma: float, seed: int = 0, clip: int = 0, t0: int = 0) -> np.ndarray:
    lib = _lib()
    if lib is None:
        return _degrade.add_noise(frames, sigma, seed, clip, t0)
    fr = np.ascontiguousarray(frames, dtype=np.uint8)
    T = fr.shape[0]
    out = np.empty_like(fr)
    lib.oracle_add_noise(fr.ctypes.data, out.ctypes.data, T, fr[0].size, _degrade.noise_gain(sigma), seed & 0xFFFFFFFF,
                         clip & 0xFFFFFFFF, t0)
    return out
