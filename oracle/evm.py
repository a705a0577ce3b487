"""Eulerian-video-magnification stages -- frozen NumPy restatement (oracle; test infra).

PARITY UNPINNED BY THE REFERENCE: ``/root/reference`` names EVM only in its README
(``README.md:36-39``) and carries no pyramid / temporal filter / collapse code and no
test for one.  This file freezes the published algorithm
(Wu et al. 2012 colour magnification, Gaussian pyramid variant) with the exact
arithmetic of the third-party calls a NumPy/OpenCV implementation would make, pinned by
probe against cv2 4.13.0 (``tests/test_oracle_evm.py`` repeats the probes):

* ``pyrdown``  == ``cv2.pyrDown(float32)``: separable [1 4 6 4 1]/16 per axis,
  BORDER_REFLECT_101, output ``(n+1)//2``, samples at even coordinates.
* ``pyrup``    == ``cv2.pyrUp(float32, dstsize=...)``: per axis even outputs
  ``(s[i-1]+6 s[i]+s[i+1])/8``, odd outputs ``(s[i]+s[i+1])/2``; border low side
  reflect-101 (``s[-1]=s[1]``), high side replicate (``s[n]=s[n-1]``).
* ``ideal_bandpass`` == ``irfft(mask * rfft(x))`` along time in float64, mask inclusive
  ``f_lo <= rfftfreq(T, 1/fps) <= f_hi`` (the reference's own mask style,
  ``rppg_VIDEO.py:140,196``; ``analysis/utils/estimate_bpm.py:51``), DC always zero.
* ``collapse_addback``: ``out = float32(frame) + pyrUp^L(alpha * filtered)``, all three
  channels equally, RGB space (the reference's YIQ helper is dead code,
  ``rppg_VIDEO.py:120-124``).

Everything here computes in float64 and rounds once at the end, so it is the
"true" value the fp32 GPU kernels are held to (tolerance stated in the tests).
"""
from __future__ import annotations

import numpy as np


def pyr_dims(W: int, H: int, levels: int):
    """[(w0,h0), (w1,h1), ... (wL,hL)] with n_{l+1} = (n_l + 1)//2."""
    dims = [(int(W), int(H))]
    for _ in range(levels):
        w, h = dims[-1]
        dims.append(((w + 1) // 2, (h + 1) // 2))
    return dims


def _reflect101(i: np.ndarray, n: int) -> np.ndarray:
    if n == 1:
        return np.zeros_like(i)
    period = 2 * (n - 1)
    i = np.mod(i, period)
    return np.where(i >= n, period - i, i)


def _pyrdown_axis(x: np.ndarray, axis: int) -> np.ndarray:
    n = x.shape[axis]
    no = (n + 1) // 2
    c = 2 * np.arange(no)

    def at(i):
        return np.take(x, _reflect101(i, n), axis=axis)

    return (at(c - 2) + 4.0 * at(c - 1) + 6.0 * at(c) + 4.0 * at(c + 1) + at(c + 2)) / 16.0


def pyrdown(img: np.ndarray) -> np.ndarray:
    """One Gaussian-pyramid reduction of (..., H, W, C) in float64."""
    x = np.asarray(img, dtype=np.float64)
    return _pyrdown_axis(_pyrdown_axis(x, -3), -2)


def pyrdown_cascade(frames: np.ndarray, levels: int) -> np.ndarray:
    """uint8/float (T,H,W,C) -> float64 (T,hL,wL,C) after ``levels`` reductions."""
    x = np.asarray(frames, dtype=np.float64)
    for _ in range(levels):
        x = pyrdown(x)
    return x


def _pyrup_axis(s: np.ndarray, axis: int, nout: int) -> np.ndarray:
    n = s.shape[axis]
    if nout not in (2 * n, 2 * n - 1):   # the only sizes a pyrDown chain produces
        raise ValueError(f"pyrup: dst size {nout} not reachable from {n}")
    s = np.moveaxis(s, axis, 0)
    i = np.arange(n)
    im1 = np.where(i - 1 < 0, min(1, n - 1), i - 1)   # reflect-101 low side
    ip1 = np.where(i + 1 >= n, n - 1, i + 1)          # replicate high side
    ev = (s[im1] + 6.0 * s[i] + s[ip1]) / 8.0
    od = (s[i] + s[ip1]) / 2.0
    out = np.empty((nout,) + s.shape[1:], dtype=s.dtype)
    out[0::2] = ev[: (nout + 1) // 2]
    out[1::2] = od[: nout // 2]
    return np.moveaxis(out, 0, axis)


def pyrup(img: np.ndarray, dst_wh) -> np.ndarray:
    """One expansion of (..., h, w, C) to (..., dst_h, dst_w, C) in float64."""
    x = np.asarray(img, dtype=np.float64)
    dw, dh = dst_wh
    return _pyrup_axis(_pyrup_axis(x, -3, dh), -2, dw)


def band_bins(T: int, fps: float, f_lo: float, f_hi: float) -> np.ndarray:
    """rfft bin indices kept by the ideal bandpass (inclusive edges, DC dropped)."""
    f = np.fft.rfftfreq(T, d=1.0 / fps)
    keep = (f >= f_lo) & (f <= f_hi)
    keep[0] = False
    return np.nonzero(keep)[0]


def ideal_bandpass(x: np.ndarray, fps: float, f_lo: float, f_hi: float) -> np.ndarray:
    """Ideal temporal bandpass along axis 0 of (T, ...) in float64."""
    x = np.asarray(x, dtype=np.float64)
    T = x.shape[0]
    X = np.fft.rfft(x, axis=0)
    keep = np.zeros(X.shape[0], dtype=bool)
    keep[band_bins(T, fps, f_lo, f_hi)] = True
    X[~keep] = 0.0
    return np.fft.irfft(X, n=T, axis=0)


def collapse(level: np.ndarray, W: int, H: int, levels: int) -> np.ndarray:
    """pyrUp^levels of (T,hL,wL,C) back to (T,H,W,C), walking the pyrDown size chain."""
    dims = pyr_dims(W, H, levels)
    x = np.asarray(level, dtype=np.float64)
    for l in range(levels, 0, -1):
        x = pyrup(x, dims[l - 1])
    return x


def collapse_addback(level: np.ndarray, frames: np.ndarray, levels: int) -> np.ndarray:
    T, H, W, _ = frames.shape
    return np.asarray(frames, dtype=np.float64) + collapse(level, W, H, levels)


def to_u8(out: np.ndarray) -> np.ndarray:
    """Declared u8 output: clip(0,255) then round-half-up."""
    return np.floor(np.clip(out, 0.0, 255.0) + 0.5).astype(np.uint8)


def evm_clip(frames: np.ndarray, fps: float, levels: int = 4, f_lo: float = 0.7,
             f_hi: float = 4.0, alpha: float = 50.0, chunk: int = 64):
    """Full EVM over a clip.  Returns (level_L, filtered_amplified_L, out) float64.
    ``out`` is produced in chunks of frames to bound memory."""
    T, H, W, C = frames.shape
    lv = np.concatenate([pyrdown_cascade(frames[i:i + chunk], levels) for i in range(0, T, chunk)])
    filt = alpha * ideal_bandpass(lv, fps, f_lo, f_hi)
    out = np.concatenate([collapse_addback(filt[i:i + chunk], frames[i:i + chunk], levels)
                          for i in range(0, T, chunk)])
    return lv, filt, out


# --------------------------------------------------------------------------- cv2 path
def evm_clip_cv2(frames: np.ndarray, fps: float, levels: int = 4, f_lo: float = 0.7,
                 f_hi: float = 4.0, alpha: float = 50.0, rects=None, keep_out: bool = False):
    """The same pipeline the way a NumPy/OpenCV user would write it (cv2.pyrDown /
    cv2.pyrUp in float32, np.fft in float64).  Used as the timed CPU baseline
    (``bench.py`` cpu_baseline / ``--impl reference``) and as a cross-check of the
    restatement above.  ``rects`` (T,K,4) [x1,y1,x2,y2) -> per-frame mean RGB (T,K,3)."""
    import cv2
    T, H, W, C = frames.shape
    dims = pyr_dims(W, H, levels)
    wl, hl = dims[-1]
    lv = np.empty((T, hl, wl, C), dtype=np.float32)
    for t in range(T):
        x = frames[t].astype(np.float32)
        for _ in range(levels):
            x = cv2.pyrDown(x)
        lv[t] = x
    filt = (alpha * ideal_bandpass(lv, fps, f_lo, f_hi)).astype(np.float32)
    means = None if rects is None else np.full((T, rects.shape[1], 3), np.nan)
    out = np.empty((T, H, W, C), dtype=np.float32) if keep_out else None
    for t in range(T):
        x = filt[t]
        for l in range(levels, 0, -1):
            x = cv2.pyrUp(x, dstsize=dims[l - 1])
        o = frames[t].astype(np.float32) + x
        if keep_out:
            out[t] = o
        if rects is not None:
            for k in range(rects.shape[1]):
                x1, y1, x2, y2 = (int(v) for v in rects[t, k])
                roi = o[y1:y2, x1:x2]
                if roi.size:
                    means[t, k] = roi.reshape(-1, 3).mean(axis=0, dtype=np.float64)
    return lv, filt, out, means


def evm_roi_trace_streaming(frame_chunks, T: int, H: int, W: int, fps: float, rects: np.ndarray | None = None,
                            levels: int = 4, f_lo: float = 0.7, f_hi: float = 4.0, alpha: float = 50.0,
                            polys: np.ndarray | None = None, nverts: np.ndarray | None = None):
    """``evm_clip_cv2(...)[3]`` (the (T,K,3) ROI means of the magnified frames) for clips too large
    to hold: ``frame_chunks()`` yields consecutive uint8 (n,H,W,3) chunks of the clip.  Same calls in
    the same order per frame (cv2.pyrDown / cv2.pyrUp float32, np.fft float64); the only difference
    is that the add-back ``frame.astype(float32) + up`` is evaluated on the union crop of the
    frame's ROIs instead of the whole frame -- element-wise, so the values are the same
    (tests/test_oracle.py::test_streaming_trace_equals_clip_form).
    ROIs: ``rects`` (T,K,4) half-open rectangles, or polygons ``polys`` (T,K,V,2) + ``nverts`` (T,K)
    rasterised by ``oracle.roi.poly_mask`` (masks are cached per distinct polygon) and averaged with
    ``oracle.roi.masked_mean``.  -> (level, filtered, means (T,K,3)[, counts (T,K)])."""
    import cv2
    from . import roi as oroi
    dims = pyr_dims(W, H, levels)
    wl, hl = dims[-1]
    use_poly = polys is not None
    K = polys.shape[1] if use_poly else rects.shape[1]
    lv = np.empty((T, hl, wl, 3), dtype=np.float32)
    crops, boxes, masks = [], [], []
    cache = {}
    t = 0
    for chunk in frame_chunks():
        for f in chunk:
            x = f.astype(np.float32)
            for _ in range(levels):
                x = cv2.pyrDown(x)
            lv[t] = x
            if use_poly:
                mk = []
                for k in range(K):
                    pts = np.ascontiguousarray(polys[t, k, :nverts[t, k]], dtype=np.int64)
                    key = pts.tobytes()
                    if key not in cache:
                        m = oroi.poly_mask(H, W, pts)
                        ys, xs = np.nonzero(m)
                        cache[key] = (m, (int(xs.min()), int(ys.min()), int(xs.max()) + 1, int(ys.max()) + 1) if len(xs) else (0, 0, 0, 0))
                    mk.append(cache[key])
                nz = [b for (_, b) in mk if b[2] > b[0]]
                box = (min(b[0] for b in nz), min(b[1] for b in nz), max(b[2] for b in nz), max(b[3] for b in nz)) if nz else (0, 0, 0, 0)
                masks.append([m for (m, _) in mk])
            else:
                r = rects[t]
                box = (int(r[:, 0].min()), int(r[:, 1].min()), int(r[:, 2].max()), int(r[:, 3].max()))
            boxes.append(box)
            crops.append(f[box[1]:box[3], box[0]:box[2]].copy())
            t += 1
    assert t == T
    filt = (alpha * ideal_bandpass(lv, fps, f_lo, f_hi)).astype(np.float32)
    means = np.full((T, K, 3), np.nan)
    counts = np.zeros((T, K), dtype=np.int64)
    for t in range(T):
        x = filt[t]
        for l in range(levels, 0, -1):
            x = cv2.pyrUp(x, dstsize=dims[l - 1])
        bx1, by1, bx2, by2 = boxes[t]
        o = crops[t].astype(np.float32) + x[by1:by2, bx1:bx2]
        for k in range(K):
            if use_poly:
                m = masks[t][k][by1:by2, bx1:bx2]
                counts[t, k] = int(m.sum())
                means[t, k] = oroi.masked_mean(o, m)
            else:
                x1, y1, x2, y2 = (int(v) for v in rects[t, k])
                roi = o[y1 - by1:y2 - by1, x1 - bx1:x2 - bx1]
                if roi.size:
                    means[t, k] = roi.reshape(-1, 3).mean(axis=0, dtype=np.float64)
    return (lv, filt, means, counts) if use_poly else (lv, filt, means)
