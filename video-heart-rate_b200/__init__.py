"""video-heart-rate_b200 -- B200-native (sm_100a) rPPG signal path.

uint8 frames + per-frame landmark ROIs in, heart-rate estimates out: fused pyrDown cascade,
temporal ideal bandpass, collapse + add-back with fused ROI means, ROI reductions and BPM
estimation, each a hand-written CUDA kernel behind the C ABI in ``include/vhr_b200.h``.
Importing this package never touches the GPU; constructing an ``Engine`` does, and fails
loudly when the library or a CUDA device is missing (there is no CPU fallback).
"""
from . import host
from ._lib import LIB_PATH, VhrError
from .engine import (DETREND_F32, DETREND_F64, DETREND_NONE, DETREND_ZSCORE_F32, FFT_ANALYSIS, FFT_VIDEO, FILT_FIR, FILT_NONE, FILT_SOS,
                     Engine, default_engine)
from .synth import SynthSpec

__all__ = ["Engine", "default_engine", "SynthSpec", "VhrError", "LIB_PATH", "host", "DETREND_NONE", "DETREND_F64",
           "DETREND_F32", "DETREND_ZSCORE_F32", "FFT_ANALYSIS", "FFT_VIDEO", "FILT_NONE", "FILT_SOS", "FILT_FIR"]
__version__ = "0.1.0"
