"""Clip-level pipelines: the reference's per-frame loops restated as batched device work.

``green_avg_measure``   analysis/measurement/green_avg.py:11-52 (the measure() plugin body)
``video_bpm_series``    rppg_VIDEO.py:354-416 signal part (process_frame + sliding window)
``evm_bpm``             EVM (pyramid -> ideal bandpass -> collapse) + ROI mean + BPM, the
                        BASELINE.json configs c2/c4 (no reference code for the EVM part)
``SlidingEvm``          rppg_LIVESTREAM.py-style sliding window (BASELINE.json config c3): frames
                        arrive hop by hop from HOST memory, one BPM per hop
"""
from __future__ import annotations

import numpy as np

from . import host
from .engine import (DETREND_F32, DETREND_F64, DETREND_NONE, DETREND_ZSCORE_F32, FFT_ANALYSIS, FILT_FIR, FILT_NONE, FILT_SOS,
                     Engine)

VIDEO_BAND = (0.7, 2.0)                 # rppg_VIDEO.py:33-34
LIVE_BAND = (40 / 60, 150 / 60)         # rppg_LIVESTREAM.py:34-35
ANALYSIS_BAND = (40 / 60, 200 / 60)     # analysis/utils/estimate_bpm.py:6-7
EVM_BAND = (0.7, 4.0)                   # BASELINE.json configs
# paint colours of rppg_VIDEO.py:100 (bbox, green) and :54 (forehead, cheek: blue) in BGR
VIDEO_PAINT_BGR = ((0, 255, 0), (255, 0, 0), (255, 0, 0))


def _frames_to_device(eng: Engine, frames):
    import torch
    if isinstance(frames, torch.Tensor):
        return frames.to(eng.tdev).contiguous()
    return torch.as_tensor(np.ascontiguousarray(frames), device=eng.tdev)


def green_avg_trace(eng: Engine, frames, landmarks, valid=None):
    """Per-frame cheek-ROI channel means (T,3) float64 (device) the analysis way: clamped
    bbox -> cheek rectangle -> mean (analysis/utils/roi.py:43-59,104-109; green_avg.py:34).
    Returns (means, usable)."""
    fr = _frames_to_device(eng, frames)
    T, H, W, _ = fr.shape
    lm = np.asarray(landmarks, dtype=np.float64)
    if lm.ndim == 2:
        lm = np.broadcast_to(lm, (T,) + lm.shape)
    usable = np.ones(T, dtype=bool)
    if valid is not None:
        lm, usable = host.hold_landmarks(lm, valid)
    rects = host.slice_rects(host.cheek_roi_clamped(host.bbox_clamped(lm, W, H), W, H), W, H)
    means = eng.roi_mean_rect(fr, rects[:, None, :])[:, 0, :]
    return means, usable


def green_avg_measure(eng: Engine, frames, fps: float, landmarks, valid=None, channel: int = 1,
                      window_s: float = 30.0, acq_s: float = 10.0, band=ANALYSIS_BAND) -> np.ndarray:
    """analysis/measurement/green_avg.py:measure on decoded frames + landmarks -> (M,2)
    float64 [t_sec, bpm]."""
    import torch
    means, usable = green_avg_trace(eng, frames, landmarks, valid)
    green = means[:, channel]
    idx_all = np.arange(green.shape[0])
    if not usable.all():
        keep = torch.as_tensor(np.nonzero(usable)[0], device=eng.tdev)
        green = green[keep].contiguous()
        idx_all = idx_all[usable]
    n = int(green.shape[0])
    fi, st, ln = host.green_avg_windows(n, fps, window_s, acq_s)
    if len(fi) == 0:
        return np.zeros((0, 2))
    bpm, _ = eng.bpm_fft(green, st, ln, fps, band, detrend=DETREND_F32, mode=FFT_ANALYSIS)
    bpm = bpm.cpu().numpy()
    ok = ~np.isnan(bpm)                                   # `if bpm is not None` (green_avg.py:47)
    ts = idx_all[fi] * (1 / fps)                          # ts = i * (1 / fps)  (green_avg.py:48)
    return np.column_stack([ts[ok], bpm[ok]])


def ica_measure(eng: Engine, frames, fps: float, landmarks, valid=None, window_s: float = 10.0, acq_s: float = 5.0,
                band=ANALYSIS_BAND, max_iter: int = 300, tol: float = 1e-6, return_details: bool = False):
    """analysis/measurement/ica.py:measure on decoded frames + landmarks -> (M,2) float64 [t_sec, bpm]:
    per-frame mean BGR of the cheek ROI (:48), growing 5 s -> 10 s window (:27-29, :52), float32 cast + ddof=1
    std normalisation (:56-61), FastICA (3 components, parallel, logcosh, tol 1e-6, 300 iterations, seed 0;
    :36-44) on the device, windows whose fixed point does not converge are skipped (:64-69), BPM = FFT peak
    of the source with the strongest in-band peak (:72, estimate_bpm.py:59-64)."""
    import torch
    means, usable = green_avg_trace(eng, frames, landmarks, valid)          # (T,3) in the frames' channel order (BGR from cv2)
    idx_all = np.arange(means.shape[0])
    if not usable.all():
        keep = torch.as_tensor(np.nonzero(usable)[0], device=eng.tdev)
        means = means[keep]
        idx_all = idx_all[usable]
    means = means.contiguous()
    n = int(means.shape[0])
    fi, st, ln = host.green_avg_windows(n, fps, window_s, acq_s)
    if len(fi) == 0:
        return (np.zeros((0, 2)), None) if return_details else np.zeros((0, 2))
    src, nit = eng.ica_fastica(means, st, ln, max_iter=max_iter, tol=tol)
    nw, ml, _ = src.shape
    starts2 = (np.arange(nw, dtype=np.int64) * ml).astype(np.int32)
    bpm, kbin = eng.bpm_fft(src.reshape(nw * ml, 3), starts2, ln, fps, band, detrend=DETREND_NONE, mode=FFT_ANALYSIS, max_len=ml)
    bpm, nit_h = bpm.cpu().numpy(), nit.cpu().numpy()
    ok = (nit_h > 0) & ~np.isnan(bpm)
    rows = np.column_stack([(idx_all[fi] * (1 / fps))[ok], bpm[ok]])
    if return_details:
        return rows, {"frame": idx_all[fi], "bpm": bpm, "bin": kbin.cpu().numpy(), "n_iter": nit_h}
    return rows


def green_avg_psd_series(eng: Engine, green, fps: float, band=ANALYSIS_BAND, window_s: float = 10.0, acq_s: float = 10.0,
                         order: int = 2) -> np.ndarray:
    """The BPM series of analysis/measurement/green_avg_psd_plot.py:160-185 given the per-frame green
    means: rolling ``int(round(window_s*fps))`` window, NaN rows until ``int(round(acq_s*fps))`` samples
    are in, then float32 z-score -> Butterworth(order) sosfiltfilt with clamped edges (:34-43) ->
    periodogram peak (:46-63).  -> (T,2) [t_sec, bpm]."""
    import scipy.signal as sp
    import torch
    g = green if isinstance(green, torch.Tensor) else torch.as_tensor(np.asarray(green, dtype=np.float64), device=eng.tdev)
    n = int(g.numel())
    window_len, acq = int(round(window_s * fps)), int(round(acq_s * fps))
    i = np.arange(n, dtype=np.int64)
    length = np.minimum(i + 1, window_len)
    keep = length >= acq
    ts = i * (1 / fps)
    out = np.column_stack([ts, np.full(n, np.nan)])
    if not keep.any():
        return out
    st = (i + 1 - length)[keep].astype(np.int32)
    ln = length[keep].astype(np.int32)
    nyq = 0.5 * fps
    low, high = max(1e-6, band[0] / nyq), min(0.999, band[1] / nyq)
    if high > low:
        sos = sp.butter(order, [low, high], btype="band", output="sos")
        _, _, filt = eng.bpm_welch(g, st, ln, fps, band, detrend=DETREND_ZSCORE_F32, filt_kind=FILT_SOS, coef=sos,
                                   want_filtered=True)
    else:
        _, _, filt = eng.bpm_welch(g, st, ln, fps, band, detrend=DETREND_ZSCORE_F32, filt_kind=FILT_NONE, want_filtered=True)
    nw, ml = filt.shape
    starts2 = (np.arange(nw, dtype=np.int64) * ml).astype(np.int32)
    bpm, _ = eng.bpm_fft(filt.reshape(-1), starts2, ln, fps, band, detrend=DETREND_NONE, mode=FFT_ANALYSIS, max_len=ml)
    out[keep, 1] = bpm.cpu().numpy()
    return out


def video_trace(eng: Engine, frames_bgr, landmarks, overdraw: bool = True, channel: int = 1):
    """The value process_frame appends per frame (rppg_VIDEO.py:91-110): mean of ``channel``
    over the unclamped cheek slice, by default with the outline-overdraw quirk.  (T,) device."""
    fr = _frames_to_device(eng, frames_bgr)
    T, H, W, _ = fr.shape
    lm = np.asarray(landmarks, dtype=np.float64)
    if lm.ndim == 2:
        lm = np.broadcast_to(lm, (T,) + lm.shape)
    bb = host.bbox_video(lm, W, H)
    fh = host.roi_coords(bb, *host.FOREHEAD)
    ck = host.roi_coords(bb, *host.CHEEK)
    rects = host.slice_rects(ck, W, H)[:, None, :]
    if overdraw:
        paint = np.stack([bb, fh, ck], 1).astype(np.int32)
        means = eng.roi_mean_rect(fr, rects, paint=paint, paint_rgb=np.asarray(VIDEO_PAINT_BGR, dtype=np.uint8))
    else:
        means = eng.roi_mean_rect(fr, rects)
    return means[:, 0, channel].contiguous()


def design_filters(fps: float, band=VIDEO_BAND):
    """Coefficient design exactly where the reference designs them (rppg_VIDEO.py:252,266,284):
    Butterworth order 2 and Chebyshev-II order 4 / 40 dB as SOS, 41-tap Hamming FIR."""
    import scipy.signal as sp
    nyq = 0.5 * fps
    lo, hi = band[0] / nyq, band[1] / nyq
    return {"butter": sp.butter(2, [lo, hi], btype="band", output="sos"),
            "cheby2": sp.cheby2(4, 40, [lo, hi], btype="band", output="sos"),
            "fir": sp.firwin(41, [lo, hi], pass_zero=False, window="hamming")}


def video_bpm_series(eng: Engine, green, fps: float, band=VIDEO_BAND, window_seconds: float = 10):
    """rppg_VIDEO.py:392-409 over a whole trace: every frame once the deque holds more than
    ``window_len`` samples -> dict(frame (M,), butter (M,), cheby2 (M,), fir (M,)) BPM arrays
    (NaN where the reference raises / returns None) and the chosen bins."""
    import torch
    g = green if isinstance(green, torch.Tensor) else torch.as_tensor(np.asarray(green, dtype=np.float64), device=eng.tdev)
    fi, st, ln = host.video_windows(int(g.numel()), fps, window_seconds)
    out = {"frame": fi}
    if len(fi) == 0:
        for k in ("butter", "cheby2", "fir"):
            out[k] = np.zeros(0)
            out[k + "_bin"] = np.zeros(0, dtype=np.int32)
        return out
    f = design_filters(fps, band)
    for name, kind in (("butter", FILT_SOS), ("cheby2", FILT_SOS), ("fir", FILT_FIR)):
        bpm, kbin, _ = eng.bpm_welch(g, st, ln, fps, band, detrend=DETREND_F64, filt_kind=kind, coef=f[name])
        out[name] = bpm.cpu().numpy()
        out[name + "_bin"] = kbin.cpu().numpy()
    return out


def evm_bpm(eng: Engine, frames, fps: float, rects=None, levels: int = 4, band=EVM_BAND, alpha: float = 50.0,
            channel: int = 1, window_len: int | None = None, hop: int | None = None, bpm_band=ANALYSIS_BAND,
            out_f32=True, out_u8=False, polys=None, nverts=None):
    """EVM + ROI + BPM on device-resident frames.  ROIs: ``rects`` (T,K,4) already sliced to the
    frame, or landmark polygons ``polys`` (T,K,V,2) + ``nverts`` (T,K) (forehead / cheeks,
    ``host.face_polygons``).  The BPM comes from the ``channel`` (green) means of the MAGNIFIED
    frames through the analysis estimator (float32 detrend + FFT peak, green_avg.py:42-44 /
    estimate_bpm.py) over sliding windows (default: one window = the whole clip): ROI 0 for
    rectangles; for polygons all K ROI traces go in as the columns of a (T,K) signal and the
    estimator keeps the column with the strongest in-band peak (estimate_bpm.py:59-64).
    -> dict(roi_mean (T,K,3) device, roi_count, bpm (n_win,), bin (n_win,), out_f32/out_u8)."""
    fr = _frames_to_device(eng, frames)
    T = fr.shape[0]
    r = eng.evm(fr, fps, levels, band[0], band[1], alpha, rects=rects, out_f32=out_f32, out_u8=out_u8, polys=polys,
                nverts=nverts)
    # strided views: vhr_bpm_fft reads them through its row / column strides
    green = r["roi_mean"][:, 0, channel] if polys is None else r["roi_mean"][:, :, channel]
    wl = T if window_len is None else window_len
    st, ln = host.sliding_windows(T, wl, wl if hop is None else hop)
    bpm, kbin = eng.bpm_fft(green, st, ln, fps, bpm_band, detrend=DETREND_F32, mode=FFT_ANALYSIS)
    r.update({"bpm": bpm, "bin": kbin, "win_start": st, "win_len": ln})
    return r


class SlidingEvm:
    """Sliding-window EVM + ROI + BPM for a live stream (BASELINE.json config c3: 10 s window,
    1 s hop; the reference's live loop is rppg_LIVESTREAM.py:311-349, which re-estimates on its
    whole deque every frame).  ``push(frames_host, rects)`` takes the hop's new uint8 frames from
    HOST memory (pinned for an asynchronous copy) and returns the window's BPM once the window is
    full.  Per hop only the NEW frames are uploaded and reduced by the pyrDown cascade (a frame's
    pyramid does not depend on its neighbours); the level-L window and the frame window slide on
    the device (two buffers, device-to-device copies); the temporal bandpass and the ROI-only
    collapse run on the whole window.  Every number equals ``evm_bpm`` on the same window, bit for
    bit (same kernels on the same per-frame data)."""

    def __init__(self, eng: Engine, H: int, W: int, fps: float, window_len: int, hop: int, K: int = 1, levels: int = 4,
                 band=EVM_BAND, alpha: float = 50.0, bpm_band=ANALYSIS_BAND, channel: int = 1):
        import torch
        self.eng, self.fps, self.n, self.hop, self.levels = eng, float(fps), int(window_len), int(hop), levels
        self.band, self.alpha, self.bpm_band, self.channel = band, alpha, bpm_band, channel
        wl, hl = eng.pyr_dims(W, H, levels)[-1]
        d = eng.tdev
        self.frames = [torch.empty((self.n, H, W, 3), dtype=torch.uint8, device=d) for _ in range(2)]
        self.level = [torch.empty((self.n, hl, wl, 3), dtype=torch.float32, device=d) for _ in range(2)]
        self.rects = [torch.zeros((self.n, K, 4), dtype=torch.int32, device=d) for _ in range(2)]
        self.filt = torch.empty((self.n, hl, wl, 3), dtype=torch.float32, device=d)
        self.cur, self.have = 0, 0
        self.st = torch.zeros(1, dtype=torch.int32, device=d)
        self.ln = torch.full((1,), self.n, dtype=torch.int32, device=d)

    def push(self, frames_host, rects):
        """frames_host uint8 (m,H,W,3) host tensor / array (m <= window), rects int32 (m,K,4).
        -> (bpm float, bin int) of the window ending with these frames, or None while filling."""
        import torch
        fh = frames_host if isinstance(frames_host, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(frames_host))
        rh = torch.as_tensor(np.ascontiguousarray(rects, dtype=np.int32)).reshape(fh.shape[0], -1, 4)
        m = int(fh.shape[0])
        src, dst = self.cur, 1 - self.cur
        keep = min(self.have, self.n - m)
        if keep > 0:            # slide: the newest `keep` frames of the old window move to the front of the other buffer
            a = self.have - keep
            self.frames[dst][:keep].copy_(self.frames[src][a:a + keep], non_blocking=True)
            self.level[dst][:keep].copy_(self.level[src][a:a + keep], non_blocking=True)
            self.rects[dst][:keep].copy_(self.rects[src][a:a + keep], non_blocking=True)
        self.frames[dst][keep:keep + m].copy_(fh, non_blocking=True)                      # H2D of the hop
        self.rects[dst][keep:keep + m].copy_(rh, non_blocking=True)
        self.eng.pyrdown(self.frames[dst][keep:keep + m], self.levels, out=self.level[dst][keep:keep + m])
        self.cur, self.have = dst, keep + m
        if self.have < self.n:
            return None
        e = self.eng
        e.bandpass(self.level[dst], self.fps, self.band[0], self.band[1], self.alpha, out=self.filt)
        _, _, means = e.collapse(self.filt, self.frames[dst], self.levels, out_f32=False, out_u8=False, rects=self.rects[dst])
        bpm, kbin = e.bpm_fft(means[:, 0, self.channel], self.st, self.ln, self.fps, self.bpm_band, detrend=DETREND_F32,
                              mode=FFT_ANALYSIS, max_len=self.n)
        self.last_means = means
        return float(bpm[0].item()), int(kbin[0].item())                                   # D2H of the result
