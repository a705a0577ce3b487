"""ROI geometry and reductions -- NumPy restatement (oracle; test infrastructure).

Reference-pinned part (rectangles): restates, function by function,
``rppg_VIDEO.py:49-55`` (get_roi_coords), ``:60-66`` (get_avg), ``:91-110``
(process_frame; twin ``rppg_LIVESTREAM.py:94-112``), ``analysis/utils/roi.py:43-50``
(_bbox_from_landmarks) and ``:53-59`` (_cheek_roi_from_bbox).  Pinned by executing those
bodies verbatim (``oracle/ref_loader.py``) -- see ``tests/golden/make_golden.py``.

Unpinned part (polygons): the reference has no polygon ROI.  The
rule frozen here is exact-integer, boundary-inclusive even-odd at pixel centres; the
tests report (not require) its pixel delta against ``cv2.fillPoly``.
"""
from __future__ import annotations

import numpy as np

# ratios: rppg_VIDEO.py:102-103 ; analysis/utils/roi.py:13-15
FOREHEAD = (0.25, 0.00, 0.25)   # horizontal_ratio, top_ratio, bottom_ratio
CHEEK = (0.15, 0.40, 0.65)


# ------------------------------------------------------------------ VIDEO / LIVE variant
def bbox_from_landmarks_video(xs, ys, w: int, h: int):
    """rppg_VIDEO.py:93-98 -- ``int()`` truncation toward zero, NO clamping."""
    return (int(min(xs) * w), int(min(ys) * h), int(max(xs) * w), int(max(ys) * h))


def roi_coords(bb, horizontal_ratio, top_ratio, bottom_ratio):
    """rppg_VIDEO.py:49-53,55 -> (x1, y1, x2, y2) (the cv.rectangle side effect at :54
    is modelled separately by ``outline_mask``)."""
    bb_x1, bb_y1, bb_x2, bb_y2 = bb
    roi_y1 = int(bb_y1 + top_ratio * (bb_y2 - bb_y1))
    roi_y2 = int(bb_y1 + bottom_ratio * (bb_y2 - bb_y1))
    roi_x1 = int(bb_x1 + horizontal_ratio * (bb_x2 - bb_x1))
    roi_x2 = int(bb_x2 - horizontal_ratio * (bb_x2 - bb_x1))
    return roi_x1, roi_y1, roi_x2, roi_y2


def py_slice(lo: int, hi: int, n: int):
    """Bounds of the NumPy basic slice ``a[lo:hi]`` on an axis of length n (negative
    indices wrap once, then clamp) -- what ``frame_bgr[c_y1:c_y2, c_x1:c_x2]``
    (rppg_VIDEO.py:106) does with the unclamped coordinates."""
    lo, hi, _ = slice(lo, hi).indices(n)
    return lo, max(lo, hi)


def outline_mask(h: int, w: int, rect) -> np.ndarray:
    """Pixels ``cv.rectangle(img, (x1,y1), (x2,y2), colour, 2)`` overwrites (thickness 2
    = three-pixel bands with the four outer corner pixels missing; pinned by probe
    against cv2 4.13.0 on 5000 random rectangles, ``tests/test_oracle_roi.py``)."""
    x1, y1, x2, y2 = rect
    x1, x2 = min(x1, x2), max(x1, x2)
    y1, y2 = min(y1, y2), max(y1, y2)
    yy, xx = np.mgrid[0:h, 0:w]
    horiz = ((np.abs(yy - y1) <= 1) | (np.abs(yy - y2) <= 1)) & (xx >= x1) & (xx <= x2)
    vert = ((np.abs(xx - x1) <= 1) | (np.abs(xx - x2) <= 1)) & (yy >= y1) & (yy <= y2)
    return horiz | vert


def process_frame_rects(xs, ys, w: int, h: int):
    """The three rectangles rppg_VIDEO.py:93-103 derives: bbox, forehead, cheek."""
    bb = bbox_from_landmarks_video(xs, ys, w, h)
    return bb, roi_coords(bb, *FOREHEAD), roi_coords(bb, *CHEEK)


def process_frame_green(frame: np.ndarray, xs, ys, channel: int = 1, overdraw: bool = True,
                        paint=((0, 255, 0), (255, 0, 0), (255, 0, 0))) -> float:
    """Value ``process_frame`` appends to ``green_signal_cheek`` (rppg_VIDEO.py:110):
    mean of ``channel`` over the cheek slice of the frame AFTER the bbox / forehead /
    cheek outlines were drawn into it (:100-103, quirk a2' in SURVEY.md).  ``paint`` are
    the three colours in the frame's channel order."""
    h, w = frame.shape[:2]
    bb, fh, ck = process_frame_rects(xs, ys, w, h)
    img = frame
    if overdraw:
        img = frame.copy()
        for rect, col in zip((bb, fh, ck), paint):
            img[outline_mask(h, w, rect)] = col
    ya, yb = py_slice(ck[1], ck[3], h)
    xa, xb = py_slice(ck[0], ck[2], w)
    roi = img[ya:yb, xa:xb]
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return float(np.mean(roi[:, :, channel]))


# ------------------------------------------------------------------------ analysis variant
def bbox_from_landmarks_clamped(xs, ys, w: int, h: int):
    """analysis/utils/roi.py:43-50."""
    x1 = int(max(0, min(xs) * w))
    y1 = int(max(0, min(ys) * h))
    x2 = int(min(w - 1, max(xs) * w))
    y2 = int(min(h - 1, max(ys) * h))
    return x1, y1, x2, y2


def cheek_roi_from_bbox(bb, w: int, h: int):
    """analysis/utils/roi.py:53-59."""
    x1, y1, x2, y2 = bb
    hr, top, bot = CHEEK
    roi_y1 = int(np.clip(y1 + top * (y2 - y1), 0, h - 1))
    roi_y2 = int(np.clip(y1 + bot * (y2 - y1), 0, h))
    roi_x1 = int(np.clip(x1 + hr * (x2 - x1), 0, w - 1))
    roi_x2 = int(np.clip(x2 - hr * (x2 - x1), 0, w))
    return roi_x1, roi_y1, roi_x2, roi_y2


def rect_mean(frame: np.ndarray, rect) -> np.ndarray:
    """Per-channel float64 mean over ``frame[y1:y2, x1:x2]`` (NaN if empty) --
    ``np.mean(roi[:, :, c])`` of rppg_VIDEO.py:66 / green_avg.py:34 for every c."""
    x1, y1, x2, y2 = (int(v) for v in rect)
    h, w = frame.shape[:2]
    ya, yb = py_slice(y1, y2, h)
    xa, xb = py_slice(x1, x2, w)
    roi = frame[ya:yb, xa:xb]
    if roi.shape[0] == 0 or roi.shape[1] == 0:
        return np.full(frame.shape[2], np.nan)
    return np.array([np.mean(roi[:, :, c], dtype=np.float64) for c in range(frame.shape[2])])


# ------------------------------------------------------------------------------- polygons
def poly_mask(h: int, w: int, pts) -> np.ndarray:
    """Exact-integer polygon rasterisation (frozen spec, parity unpinned):

    pixel (x, y) at integer coordinates is inside iff it lies ON any edge (closed
    segment, int64 cross product == 0 within the segment's box) OR the even-odd rule
    holds with half-open edge spans ``(y0 <= y) != (y1 <= y)`` and the pixel strictly
    left of the crossing (sign of an int64 cross product)."""
    pts = np.asarray(pts, dtype=np.int64).reshape(-1, 2)
    mask = np.zeros((h, w), dtype=bool)
    V = pts.shape[0]
    if V == 0:
        return mask
    yy, xx = np.mgrid[0:h, 0:w]
    yy = yy.astype(np.int64)
    xx = xx.astype(np.int64)
    parity = np.zeros((h, w), dtype=bool)
    for i in range(V):
        x0, y0 = pts[i]
        x1, y1 = pts[(i + 1) % V]
        cross = (x1 - x0) * (yy - y0) - (y1 - y0) * (xx - x0)
        on = ((cross == 0) & (xx >= min(x0, x1)) & (xx <= max(x0, x1))
              & (yy >= min(y0, y1)) & (yy <= max(y0, y1)))
        mask |= on
        dy = y1 - y0
        if dy != 0:
            strad = (y0 <= yy) != (y1 <= yy)
            tt = (xx - x0) * dy - (x1 - x0) * (yy - y0)
            left = (tt < 0) if dy > 0 else (tt > 0)
            parity ^= (strad & left)
    return mask | parity


def masked_mean(frame: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """Per-channel float64 mean of ``frame`` over ``mask`` (NaN if the mask is empty)."""
    n = int(mask.sum())
    if n == 0:
        return np.full(frame.shape[2], np.nan)
    return frame[mask].astype(np.float64).sum(axis=0) / n


def face_polygons(x0: int, y0: int, x1: int, y1: int):
    """Synthetic 'landmark polygons' (forehead, left cheek, right cheek) inside a face
    rectangle -- convex and concave int32 outlines used by the synthetic configs."""
    fw, fh = x1 - x0, y1 - y0

    def P(*uv):
        return np.array([[x0 + int(u * fw), y0 + int(v * fh)] for u, v in uv], dtype=np.int32)

    forehead = P((0.25, 0.06), (0.40, 0.03), (0.60, 0.03), (0.75, 0.06), (0.78, 0.20),
                 (0.60, 0.24), (0.50, 0.21), (0.40, 0.24), (0.22, 0.20))
    lcheek = P((0.14, 0.45), (0.30, 0.42), (0.40, 0.52), (0.36, 0.66), (0.24, 0.70), (0.15, 0.60))
    rcheek = P((0.86, 0.45), (0.70, 0.42), (0.60, 0.52), (0.64, 0.66), (0.76, 0.70), (0.85, 0.60))
    return [forehead, lcheek, rcheek]
