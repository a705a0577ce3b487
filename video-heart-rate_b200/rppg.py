"""Drop-in signal functions with the names, arguments and error behaviour of the reference
scripts (boundary b4), computed by the CUDA library.

    rppg_VIDEO.py:       get_roi_coords, get_avg, process_frame, estimate_bpm,
                         estimate_bpm_welch, bandpass_butterworth, bandpass_fir, bandpass_cheby2
    rppg_LIVESTREAM.py:  bandpass_butterworth_eqn, live_sos_init, live_sos_reset, live_sos_push
                         (+ the twins of the functions above; the only difference between the
                         two scripts is the band, selected with ``set_band``)

A maintainer swaps ``from video_heart_rate_b200.rppg import *`` in after the reference's own
definitions.  Plot arguments (``ax``) are accepted and ignored: plotting is
out of scope.  Drawing side effects on the frame (cv.rectangle) are kept when cv2 is present,
because the reference's ROI mean depends on them (the overdraw quirk, rppg_VIDEO.py:54,100-106).
"""
from __future__ import annotations

from collections import deque

import numpy as np

from . import host
from .engine import (DETREND_NONE, FFT_VIDEO, FILT_FIR, FILT_NONE, FILT_SOS, default_engine)

# signal storage (rppg_VIDEO.py:15-20)
green_signal_forehead = deque(maxlen=500)
green_signal_cheek = deque(maxlen=1000)
green_signal_filtered = deque(maxlen=500)      # rppg_LIVESTREAM.py:17

FREQ_LOW = 0.7     # rppg_VIDEO.py:33-34
FREQ_HIGH = 2


def set_band(freq_low: float, freq_high: float):
    """rppg_VIDEO.py uses 0.7-2 Hz (:33-34), rppg_LIVESTREAM.py 40-150 BPM (:34-35)."""
    global FREQ_LOW, FREQ_HIGH
    FREQ_LOW, FREQ_HIGH = freq_low, freq_high


def _draw(frame, p1, p2, colour):
    try:
        import cv2
        cv2.rectangle(frame, p1, p2, colour, 2)
    except ImportError:       # drawing is cosmetic once the mean is computed on the device
        pass


def get_roi_coords(bb_x1, bb_y1, bb_x2, bb_y2, horizontal_ratio, top_ratio, bottom_ratio, frame_bgr):
    """rppg_VIDEO.py:49-55."""
    r = host.roi_coords([[bb_x1, bb_y1, bb_x2, bb_y2]], horizontal_ratio, top_ratio, bottom_ratio)[0]
    roi_x1, roi_y1, roi_x2, roi_y2 = (int(v) for v in r)
    if frame_bgr is not None:
        _draw(frame_bgr, (roi_x1, roi_y1), (roi_x2, roi_y2), (255, 0, 0))
    return roi_x1, roi_y1, roi_x2, roi_y2


def get_avg(roi, color):
    """rppg_VIDEO.py:60-66: np.mean(roi[:, :, color]).  uint8 ROIs (what the scripts pass): exact integer sum on
    the device, one float64 division -- bit-identical to NumPy.  Any other dtype np.mean accepts goes through the
    float32 masked-mean kernel (float64 accumulation of the float32-cast values)."""
    import torch
    eng = default_engine()
    roi = np.asarray(roi)
    if roi.ndim != 3 or not (0 <= color < roi.shape[2] or -roi.shape[2] <= color < 0):
        raise IndexError("get_avg expects an (h, w, C) ROI and a channel index inside it")
    h, w = roi.shape[:2]
    if h == 0 or w == 0:
        return float("nan")
    if roi.dtype == np.uint8 and roi.shape[2] == 3:
        fr = torch.as_tensor(np.ascontiguousarray(roi)[None], device=eng.tdev)
        m = eng.roi_mean_rect(fr, np.array([[[0, 0, w, h]]], dtype=np.int32))
        return float(m[0, 0, color].item())
    plane = np.ascontiguousarray(roi[:, :, color], dtype=np.float32)
    fr = torch.as_tensor(np.repeat(plane[None, :, :, None], 3, axis=3), device=eng.tdev)
    poly = np.array([[[[0, 0], [w - 1, 0], [w - 1, h - 1], [0, h - 1]]]], dtype=np.int32)
    m, _ = eng.roi_mean_poly(fr, poly, np.array([[4]], dtype=np.int32))
    return float(m[0, 0, 0].item())


def process_frame(frame_bgr, landmarks):
    """rppg_VIDEO.py:91-110: bbox from landmarks (``.x``/``.y``), outlines drawn into the frame,
    mean green of the cheek slice appended to ``green_signal_cheek``."""
    import torch
    from .pipeline import video_trace
    eng = default_engine()
    lm = np.array([[p.x, p.y] for p in landmarks], dtype=np.float64)
    fr = torch.as_tensor(np.ascontiguousarray(frame_bgr)[None], device=eng.tdev)
    g = video_trace(eng, fr, lm[None], overdraw=True)
    h, w = frame_bgr.shape[:2]
    bb = host.bbox_video(lm[None], w, h)[0]
    _draw(frame_bgr, (int(bb[0]), int(bb[1])), (int(bb[2]), int(bb[3])), (0, 255, 0))
    get_roi_coords(*[int(v) for v in bb], 0.25, 0.00, 0.25, frame_bgr)
    get_roi_coords(*[int(v) for v in bb], 0.15, 0.4, 0.65, frame_bgr)
    green_signal_cheek.append(float(g[0].item()))


def _one_window(signal):
    x = np.asarray(signal, dtype=np.float64)
    if x.ndim != 1:
        raise ValueError("the estimators take one trace at a time (shape [T,])")
    return x, np.zeros(1, dtype=np.int32), np.array([x.shape[0]], dtype=np.int32)


def _windows_of(signal):
    """(T,) -> one window; (T,N) -> N windows, one per column (the reference filters along axis 0)."""
    x = np.asarray(signal, dtype=np.float64)
    if x.ndim not in (1, 2):
        raise ValueError("signal must be 1D (T,) or 2D (T, N) with time along axis 0")
    cols = x[:, None] if x.ndim == 1 else x
    T, N = cols.shape
    flat = np.ascontiguousarray(cols.T).reshape(-1)
    return x, flat, (np.arange(N, dtype=np.int32) * T).astype(np.int32), np.full(N, T, dtype=np.int32)


def _filtfilt(signal, kind, coef, padlen):
    x, flat, st, ln = _windows_of(signal)
    if x.shape[0] <= padlen:
        # scipy's message (signal/_signaltools.py:_validate_pad), raised by the reference at
        # rppg_VIDEO.py:404 for 5 FPS clips
        raise ValueError(f"The length of the input vector x must be greater than padlen, which is {padlen}.")
    eng = default_engine()
    _, _, filt = eng.bpm_welch(flat, st, ln, 1.0, (0.0, 1.0), detrend=DETREND_NONE, filt_kind=kind, coef=coef,
                               want_filtered=True)
    out = filt.cpu().numpy()
    return out[0] if x.ndim == 1 else np.ascontiguousarray(out.T)


def _sos_padlen(sos):
    ntaps = 2 * sos.shape[0] + 1
    ntaps -= min(int((sos[:, 2] == 0).sum()), int((sos[:, 5] == 0).sum()))
    return 3 * ntaps


def bandpass_butterworth(signal, fps, freq_lo, freq_high, order):
    """rppg_VIDEO.py:241-255: Butterworth band-pass as SOS, zero-phase (sosfiltfilt)."""
    import scipy.signal as sp
    nyquist = 0.5 * fps
    sos = sp.butter(order, [freq_lo / nyquist, freq_high / nyquist], btype='band', output='sos')
    return _filtfilt(signal, FILT_SOS, sos, _sos_padlen(sos))


def bandpass_fir(signal, fps, freq_lo, freq_high, numtaps=41):
    """rppg_VIDEO.py:259-271: Hamming FIR, zero-phase (filtfilt)."""
    import scipy.signal as sp
    nyquist = 0.5 * fps
    b = sp.firwin(numtaps, [freq_lo / nyquist, freq_high / nyquist], pass_zero=False, window='hamming')
    return _filtfilt(signal, FILT_FIR, b, 3 * numtaps)


def bandpass_cheby2(signal, fps, freq_lo, freq_high, order=4, stopband_atten=40):
    """rppg_VIDEO.py:274-289."""
    import scipy.signal as sp
    nyquist = 0.5 * fps
    sos = sp.cheby2(order, stopband_atten, [freq_lo / nyquist, freq_high / nyquist], btype='band', output='sos')
    return _filtfilt(signal, FILT_SOS, sos, _sos_padlen(sos))


def estimate_bpm_welch(signal, fps, ax=None, label='Welch PSD'):
    """rppg_VIDEO.py:172-235 (``ax`` ignored)."""
    x, st, ln = _one_window(signal)
    eng = default_engine()
    bpm, _, _ = eng.bpm_welch(x, st, ln, fps, (FREQ_LOW, FREQ_HIGH), detrend=DETREND_NONE, filt_kind=FILT_NONE)
    v = float(bpm[0].item())
    return None if np.isnan(v) else v


def estimate_bpm(signal, fps, ax=None, label='FFT'):
    """rppg_VIDEO.py:129-168 (``ax`` ignored)."""
    x, st, ln = _one_window(signal)
    eng = default_engine()
    bpm, _ = eng.bpm_fft(x, st, ln, fps, (FREQ_LOW, FREQ_HIGH), detrend=DETREND_NONE, mode=FFT_VIDEO)
    v = float(bpm[0].item())
    return None if np.isnan(v) else v


# ---------- live SOS filter (rppg_LIVESTREAM.py:200-251) ----------
_live_sos = None
_live_zi = None


def bandpass_butterworth_eqn(signal, fps, freq_lo, freq_high, order):
    """rppg_LIVESTREAM.py:207-220: returns the SOS coefficients."""
    import scipy.signal as sp
    nyquist = 0.5 * fps
    return sp.butter(order, [freq_lo / nyquist, freq_high / nyquist], btype='band', output='sos')


def live_sos_init(sos):
    global _live_sos, _live_zi
    import torch
    eng = default_engine()
    _live_sos = np.asarray(sos, dtype=np.float64)
    _live_zi = torch.zeros((_live_sos.shape[0], 2), dtype=torch.float64, device=eng.tdev)


def live_sos_reset():
    if _live_zi is not None:
        _live_zi.zero_()


def live_sos_push(x: float) -> float:
    if _live_sos is None or _live_zi is None:
        raise RuntimeError("live_sos_init(sos) must be called before live_sos_push().")
    eng = default_engine()
    y = eng.sos_causal(np.array([x], dtype=np.float64), _live_sos, _live_zi)
    return float(y[0].item())
