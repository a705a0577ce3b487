"""BPM-stage restatement (oracle; test infrastructure).

Each function restates one reference function and makes the same third-party calls
(numpy.fft / scipy.signal) at the same call sites; pinned by executing the reference's
own bodies verbatim (``oracle/ref_loader.py``) in ``tests/golden/make_golden.py``.
Functions return the in-band peak BIN as well as the BPM so the GPU path can be held
to "identical spectral-peak bin".
"""
from __future__ import annotations

import numpy as np
import scipy.signal as sp

# band edges per entry point (SURVEY.md section 5 "Config")
VIDEO_BAND = (0.7, 2.0)                   # rppg_VIDEO.py:33-34
LIVE_BAND = (40 / 60, 150 / 60)           # rppg_LIVESTREAM.py:34-35
ANALYSIS_BAND = (40 / 60, 200 / 60)       # analysis/utils/estimate_bpm.py:6-7


def estimate_bpm_analysis(signal, fs: float, band=ANALYSIS_BAND):
    """analysis/utils/estimate_bpm.py:12-65 -> (bpm | None, k | -1, col).
    k is the FFT bin index (1..N-1 over the positive half) of the chosen peak."""
    if signal is None:
        return None, -1, -1
    X = np.asarray(signal, dtype=np.float64)
    if X.ndim == 1:
        X = X[:, None]
    elif X.ndim != 2:
        raise ValueError("signal must be 1D (T,) or 2D (T, C) with time along axis 0.")
    N = X.shape[0]
    if N < 8:
        return None, -1, -1
    fft_vals = np.fft.fft(X, axis=0)
    freqs = np.fft.fftfreq(N, d=1 / fs)
    pos = freqs > 0
    if not np.any(pos):
        return None, -1, -1
    freqs_pos = freqs[pos]
    mags = np.abs(fft_vals[pos, ...])
    bandm = (freqs_pos >= band[0]) & (freqs_pos <= band[1])
    if not np.any(bandm):
        return None, -1, -1
    band_mags = mags[bandm, :]
    peak_idx_per_col = np.argmax(band_mags, axis=0)
    peak_mag_per_col = band_mags[peak_idx_per_col, np.arange(band_mags.shape[1])]
    best_col = int(np.argmax(peak_mag_per_col))
    kk = np.nonzero(pos)[0][bandm]
    k = int(kk[peak_idx_per_col[best_col]])
    dom_freq = float(freqs_pos[bandm][peak_idx_per_col[best_col]])
    return dom_freq * 60.0, k, best_col


def green_avg_series(green, fps: float, window_s: float = 30.0, acq_s: float = 10.0,
                     band=ANALYSIS_BAND):
    """analysis/measurement/green_avg.py:24-52 given the per-frame ROI means
    (``float(np.mean(roi[:,:,1]))`` of :34) -> ((M,2) [t_sec, bpm], bins (M,))."""
    from collections import deque
    window_len = int(window_s * fps)
    acquisition_len = int(acq_s * fps)
    dq = deque(maxlen=window_len)
    ts, bpms, bins = [], [], []
    for i, g in enumerate(green):
        dq.append(float(g))
        if len(dq) < acquisition_len:
            continue
        sig = np.asarray(dq, dtype=np.float32)
        sig = sig - np.mean(sig)
        bpm, k, _ = estimate_bpm_analysis(sig, fps, band)
        if bpm is not None:
            ts.append(i * (1 / fps))
            bpms.append(bpm)
            bins.append(k)
    return np.column_stack([ts, bpms]) if ts else np.zeros((0, 2)), np.asarray(bins, dtype=np.int64)


def estimate_bpm_video_fft(signal, fps: float, band=VIDEO_BAND):
    """rppg_VIDEO.py:129-147 (defined, never called by the script) -> (bpm | None, k)."""
    signal = np.asarray(signal)
    freqs = np.fft.fftfreq(len(signal), d=1 / fps)
    magnitudes = np.abs(np.fft.fft(signal))
    mask = (freqs >= band[0]) & (freqs <= band[1])
    if not np.any(mask):
        return None, -1
    kk = np.nonzero(mask)[0]
    k = int(np.argmax(magnitudes[mask]))
    return float(freqs[mask][k]) * 60.0, int(kk[k])


def estimate_bpm_welch(signal, fps: float, band=VIDEO_BAND):
    """rppg_VIDEO.py:172-203 (twin rppg_LIVESTREAM.py:133-164) -> (bpm | None, k, nperseg)."""
    x = np.asarray(signal, dtype=np.float32)
    x = x - np.nanmean(x)
    window_seconds = 9
    nperseg = int(min(len(x), fps * window_seconds))
    noverlap = nperseg // 2
    freqs, psd = sp.welch(x, fs=fps, window='hann', nperseg=nperseg, noverlap=noverlap,
                          detrend='constant', scaling='density', average='mean')
    band_mask = (freqs >= band[0]) & (freqs <= band[1])
    if not np.any(band_mask):
        return None, -1, nperseg
    kk = np.nonzero(band_mask)[0]
    k = int(np.argmax(psd[band_mask]))
    return float(freqs[band_mask][k] * 60.0), int(kk[k]), nperseg


def bandpass_butterworth(signal, fps, freq_lo, freq_high, order):
    """rppg_VIDEO.py:241-255."""
    nyquist = 0.5 * fps
    sos = sp.butter(order, [freq_lo / nyquist, freq_high / nyquist], btype='band', output='sos')
    return sp.sosfiltfilt(sos, signal, axis=0)


def bandpass_cheby2(signal, fps, freq_lo, freq_high, order=4, stopband_atten=40):
    """rppg_VIDEO.py:274-289."""
    nyquist = 0.5 * fps
    sos = sp.cheby2(order, stopband_atten, [freq_lo / nyquist, freq_high / nyquist],
                    btype='band', output='sos')
    return sp.sosfiltfilt(sos, signal, axis=0)


def bandpass_fir(signal, fps, freq_lo, freq_high, numtaps=41):
    """rppg_VIDEO.py:259-271 (raises ValueError for len(signal) <= 123, as the
    reference does)."""
    nyquist = 0.5 * fps
    b = sp.firwin(numtaps, [freq_lo / nyquist, freq_high / nyquist], pass_zero=False,
                  window='hamming')
    return sp.filtfilt(b, [1.0], signal, axis=0)


def video_window_bpm(green, fps: float, band=VIDEO_BAND, window_seconds: float = 10,
                     maxlen: int = 1000):
    """The per-frame sliding-window block of rppg_VIDEO.py:392-409 replayed over a
    whole trace: after frame i the deque (maxlen 1000, :16) holds the last samples; if
    ``len > window_len`` the last ``window_len`` samples are mean-detrended, filtered
    three ways and Welch-peaked.  -> list of (i, bpm_butter, bpm_cheby2, bpm_fir|None,
    (k_butter, k_cheby2, k_fir)).  FIR entry is None where the reference raises."""
    from collections import deque
    dq = deque(maxlen=maxlen)
    window_len = int(fps * window_seconds)
    out = []
    for i, g in enumerate(green):
        dq.append(g)
        if len(dq) > window_len:
            w = np.array(dq)[-window_len:]
            w = w - np.mean(w)
            fb = bandpass_butterworth(w, fps, band[0], band[1], order=2)
            fc = bandpass_cheby2(w, fps, band[0], band[1], order=4)
            b1, k1, _ = estimate_bpm_welch(fb, fps, band)
            b2, k2, _ = estimate_bpm_welch(fc, fps, band)
            try:
                ff = bandpass_fir(w, fps, band[0], band[1])
                b3, k3, _ = estimate_bpm_welch(ff, fps, band)
            except ValueError:
                b3, k3 = None, -1
            out.append((i, b1, b2, b3, (k1, k2, k3)))
    return out


class LiveSOS:
    """rppg_LIVESTREAM.py:207-251: causal SOS filter with carried (n_sections,2) state."""

    def __init__(self, sos):
        self.sos = np.asarray(sos, dtype=np.float64)
        self.zi = np.zeros((self.sos.shape[0], 2), dtype=np.float64)

    def reset(self):
        self.zi[...] = 0.0

    def push(self, x: float) -> float:
        y, self.zi = sp.sosfilt(self.sos, [x], zi=self.zi)
        return float(y[0])


def live_sos_design(fps: float, band=LIVE_BAND, order: int = 4):
    """rppg_LIVESTREAM.py:207-220 / 294-301."""
    nyq = 0.5 * fps
    return sp.butter(order, [band[0] / nyq, band[1] / nyq], btype='band', output='sos')


def psd_plot_estimate(green_window, fps: float, band=ANALYSIS_BAND, order: int = 2):
    """analysis/measurement/green_avg_psd_plot.py:173-183 + :34-63 for one window:
    float32 z-score -> Butterworth sosfiltfilt (clamped edges) -> periodogram peak.
    -> (bpm, k)."""
    signal = np.asarray(green_window, dtype=np.float32)
    signal = (signal - np.mean(signal)) / np.std(signal)
    nyq = 0.5 * fps
    low = max(1e-6, band[0] / nyq)
    high = min(0.999, band[1] / nyq)
    filtered = signal
    if high > low:
        sos = sp.butter(order, [low, high], btype="band", output="sos")
        filtered = sp.sosfiltfilt(sos, signal, axis=0)
    if filtered.size < 8:
        return float("nan"), -1
    fft_vals = np.fft.fft(filtered)
    freqs = np.fft.fftfreq(filtered.shape[0], d=1.0 / fps)
    pos = freqs > 0
    kk = np.nonzero(pos)[0]
    freqs = freqs[pos]
    mags = np.abs(fft_vals[pos]) ** 2 / (fps * len(filtered))
    bandm = (freqs >= band[0]) & (freqs <= band[1])
    if not np.any(bandm):
        return float("nan"), -1
    j = int(np.argmax(mags[bandm]))
    return float(freqs[bandm][j] * 60.0), int(kk[bandm][j])
