"""Host-side logic of the rPPG path: ROI geometry from landmarks and window bookkeeping.

These are the O(T) scalar parts of the reference that stay on the host (landmarks are
per-frame inputs, BASELINE.json north_star); they are vectorised over frames and repeat the
reference's float64 -> ``int()`` truncation exactly.  Everything that touches pixels or
spectra runs in the CUDA library.
"""
from __future__ import annotations

import numpy as np

# ratios: rppg_VIDEO.py:102-103 ; analysis/utils/roi.py:13-15
FOREHEAD = (0.25, 0.00, 0.25)   # horizontal_ratio, top_ratio, bottom_ratio
CHEEK = (0.15, 0.40, 0.65)
REUSE_LANDMARKS_FOR = 15        # analysis/utils/roi.py:10


def _trunc(a):
    return np.trunc(a).astype(np.int64)        # Python int(): toward zero


def landmarks_minmax(landmarks):
    """landmarks (T,N,2) normalised (x,y) -> (xmin, ymin, xmax, ymax), each (T,)."""
    lm = np.asarray(landmarks, dtype=np.float64)
    if lm.ndim == 2:
        lm = lm[None]
    return lm[..., 0].min(1), lm[..., 1].min(1), lm[..., 0].max(1), lm[..., 1].max(1)


def bbox_video(landmarks, w: int, h: int):
    """rppg_VIDEO.py:93-98 (twin rppg_LIVESTREAM.py:96-101): unclamped, truncated -> (T,4)."""
    xmin, ymin, xmax, ymax = landmarks_minmax(landmarks)
    return np.stack([_trunc(xmin * w), _trunc(ymin * h), _trunc(xmax * w), _trunc(ymax * h)], 1)


def roi_coords(bb, horizontal_ratio, top_ratio, bottom_ratio):
    """rppg_VIDEO.py:49-53 vectorised over (T,4) boxes -> (T,4) [x1,y1,x2,y2]."""
    bb = np.asarray(bb, dtype=np.int64).reshape(-1, 4)
    x1, y1, x2, y2 = (bb[:, i] for i in range(4))
    ry1 = _trunc(y1 + top_ratio * (y2 - y1))
    ry2 = _trunc(y1 + bottom_ratio * (y2 - y1))
    rx1 = _trunc(x1 + horizontal_ratio * (x2 - x1))
    rx2 = _trunc(x2 - horizontal_ratio * (x2 - x1))
    return np.stack([rx1, ry1, rx2, ry2], 1)


def bbox_clamped(landmarks, w: int, h: int):
    """analysis/utils/roi.py:43-50 -> (T,4)."""
    xmin, ymin, xmax, ymax = landmarks_minmax(landmarks)
    return np.stack([_trunc(np.maximum(0, xmin * w)), _trunc(np.maximum(0, ymin * h)),
                     _trunc(np.minimum(w - 1, xmax * w)), _trunc(np.minimum(h - 1, ymax * h))], 1)


def cheek_roi_clamped(bb, w: int, h: int):
    """analysis/utils/roi.py:53-59 -> (T,4)."""
    bb = np.asarray(bb, dtype=np.int64).reshape(-1, 4)
    x1, y1, x2, y2 = (bb[:, i] for i in range(4))
    hr, top, bot = CHEEK
    ry1 = _trunc(np.clip(y1 + top * (y2 - y1), 0, h - 1))
    ry2 = _trunc(np.clip(y1 + bot * (y2 - y1), 0, h))
    rx1 = _trunc(np.clip(x1 + hr * (x2 - x1), 0, w - 1))
    rx2 = _trunc(np.clip(x2 - hr * (x2 - x1), 0, w))
    return np.stack([rx1, ry1, rx2, ry2], 1)


# ---------------------------------------------------------------------------------- polygons
# The reference's ROIs are the ratio rectangles above (rppg_VIDEO.py:102-103); BASELINE.json's
# north_star asks for landmark POLYGONS (forehead and cheeks).  Two sources of outlines:
#   * ``ratio_polygons``     V-gons inscribed in the reference's own ratio rectangles of the landmark
#                            bounding box (forehead; the cheek rectangle split into a left and a right
#                            cheek); with ``shape="rect"`` the 4-gon IS the rectangle (same pixel set);
#   * ``landmark_polygons``  outlines through chosen landmarks of the mesh (index lists).
FACE_PARTS = (
    # name, (horizontal_ratio, top_ratio, bottom_ratio) of rppg_VIDEO.py:102-103, (x_lo, x_hi) share of that rectangle
    ("forehead", FOREHEAD, (0.0, 1.0)),
    ("cheek_l", CHEEK, (0.0, 0.4)),
    ("cheek_r", CHEEK, (0.6, 1.0)),
)
# MediaPipe FaceMesh outlines commonly used for rPPG (478-landmark topology).  mediapipe is not
# installed in this image, so these lists could not be run against the reference's landmarker here.
MESH_FOREHEAD = (109, 10, 338, 337, 336, 9, 107, 108)
MESH_CHEEK_L = (117, 118, 101, 36, 205, 187, 123)
MESH_CHEEK_R = (346, 347, 330, 266, 425, 411, 352)


def ratio_polygons(bb, n_vertices: int = 36, parts=FACE_PARTS, shape: str = "ellipse"):
    """(T,4) bounding boxes -> (polys int32 (T,K,V,2), nverts int32 (T,K)).  Part rectangles come
    from ``roi_coords`` (the reference's float64 -> int() arithmetic); ``shape="ellipse"`` inscribes
    an ellipse sampled at ``n_vertices`` angles (vertex = floor of the real point), ``shape="rect"``
    gives the closed 4-gon (x1,y1) .. (x2-1,y2-1), whose mask is exactly the half-open rectangle."""
    bb = np.asarray(bb, dtype=np.int64).reshape(-1, 4)
    T = bb.shape[0]
    V = 4 if shape == "rect" else int(n_vertices)
    polys = np.zeros((T, len(parts), V, 2), dtype=np.int32)
    ang = 2.0 * np.pi * np.arange(V, dtype=np.float64) / V
    for k, (_, ratios, (lo, hi)) in enumerate(parts):
        r = roi_coords(bb, *ratios).astype(np.float64)
        x1 = r[:, 0] + lo * (r[:, 2] - r[:, 0])
        x2 = r[:, 0] + hi * (r[:, 2] - r[:, 0])
        y1, y2 = r[:, 1], r[:, 3]
        if shape == "rect":
            xa, xb, ya, yb = _trunc(x1), _trunc(x2) - 1, _trunc(y1), _trunc(y2) - 1
            polys[:, k, :, 0] = np.stack([xa, xb, xb, xa], 1)
            polys[:, k, :, 1] = np.stack([ya, ya, yb, yb], 1)
        else:
            cx, cy = 0.5 * (x1 + x2 - 1), 0.5 * (y1 + y2 - 1)
            rx, ry = 0.5 * (x2 - 1 - x1), 0.5 * (y2 - 1 - y1)
            polys[:, k, :, 0] = np.floor(cx[:, None] + np.maximum(rx, 0)[:, None] * np.cos(ang)[None] + 0.5)
            polys[:, k, :, 1] = np.floor(cy[:, None] + np.maximum(ry, 0)[:, None] * np.sin(ang)[None] + 0.5)
    return polys, np.full((T, len(parts)), V, dtype=np.int32)


def landmark_polygons(landmarks, w: int, h: int, index_lists=(MESH_FOREHEAD, MESH_CHEEK_L, MESH_CHEEK_R)):
    """Outlines through mesh landmarks: vertex = (int(x * w), int(y * h)), the truncation the
    reference applies to landmark coordinates (rppg_VIDEO.py:95-98).  landmarks (T,N,2) ->
    (polys int32 (T,K,Vmax,2), nverts int32 (T,K))."""
    lm = np.asarray(landmarks, dtype=np.float64)
    if lm.ndim == 2:
        lm = lm[None]
    T = lm.shape[0]
    V = max(len(ix) for ix in index_lists)
    polys = np.zeros((T, len(index_lists), V, 2), dtype=np.int32)
    nverts = np.zeros((T, len(index_lists)), dtype=np.int32)
    for k, ix in enumerate(index_lists):
        ix = np.asarray(ix, dtype=np.int64)
        polys[:, k, :len(ix), 0] = _trunc(lm[:, ix, 0] * w)
        polys[:, k, :len(ix), 1] = _trunc(lm[:, ix, 1] * h)
        nverts[:, k] = len(ix)
    return polys, nverts


def face_polygons(landmarks, w: int, h: int, n_vertices: int = 36):
    """The forehead + two-cheek outlines of a landmark track: through the mesh landmarks when the
    track has the FaceMesh topology (>= 468 points), otherwise ellipses in the reference's ratio
    rectangles of the clamped landmark bounding box."""
    lm = np.asarray(landmarks, dtype=np.float64)
    if lm.ndim == 2:
        lm = lm[None]
    if lm.shape[1] >= 468:
        return landmark_polygons(lm, w, h)
    return ratio_polygons(bbox_clamped(lm, w, h), n_vertices)


def slice_rects(rects, w: int, h: int):
    """Apply NumPy basic-slice semantics of ``frame[y1:y2, x1:x2]`` (rppg_VIDEO.py:106;
    roi.py:104) to (T,4) coordinates: negative indices wrap once, then clamp; reversed
    bounds give an empty rectangle.  -> int32 (T,4) with 0<=x1<=x2<=w, 0<=y1<=y2<=h."""
    r = np.asarray(rects, dtype=np.int64).reshape(-1, 4)

    def norm(lo, hi, n):
        lo = np.where(lo < 0, np.maximum(lo + n, 0), np.minimum(lo, n))
        hi = np.where(hi < 0, np.maximum(hi + n, 0), np.minimum(hi, n))
        return lo, np.maximum(lo, hi)

    xa, xb = norm(r[:, 0], r[:, 2], w)
    ya, yb = norm(r[:, 1], r[:, 3], h)
    return np.stack([xa, ya, xb, yb], 1).astype(np.int32)


def hold_landmarks(landmarks, valid, reuse_for: int = REUSE_LANDMARKS_FOR):
    """Landmark drop-out policy of analysis/utils/roi.py:85-98: a frame without detection
    reuses the last landmarks for up to ``reuse_for`` frames.  landmarks (T,N,2), valid (T,)
    bool -> (landmarks_held (T,N,2), usable (T,) bool).  Frames that are not usable yield
    no ROI (their mean is NaN and no BPM row is produced)."""
    lm = np.array(landmarks, dtype=np.float64, copy=True)
    valid = np.asarray(valid, dtype=bool)
    usable = np.zeros(len(valid), dtype=bool)
    last, left = -1, 0
    for i, v in enumerate(valid):
        if v:
            last, left = i, reuse_for
            usable[i] = True
        elif last >= 0 and left > 0:
            left -= 1
            lm[i] = lm[last]
            usable[i] = True
        elif last >= 0:
            # roi.py:96-109: an empty ROI is yielded and then, because last_landmarks is
            # still set, the stale ROI is yielded too; we keep the stale landmarks only
            lm[i] = lm[last]
            usable[i] = False
    return lm, usable


# ---------------------------------------------------------------------------------- windows
def green_avg_windows(n_samples: int, fps: float, window_s: float = 30.0, acq_s: float = 10.0):
    """Window list of analysis/measurement/green_avg.py:24-39: after sample i (0-based) the
    deque holds the last min(i+1, window_len) samples; a BPM is estimated once it holds
    >= acquisition_len.  -> (frame_index (M,), start (M,), length (M,)) int32."""
    window_len = int(window_s * fps)
    acq = int(acq_s * fps)
    i = np.arange(n_samples, dtype=np.int64)
    length = np.minimum(i + 1, window_len)
    keep = length >= acq
    if window_len <= 0:
        keep[:] = False
    i, length = i[keep], length[keep]
    return i.astype(np.int32), (i + 1 - length).astype(np.int32), length.astype(np.int32)


def video_windows(n_samples: int, fps: float, window_seconds: float = 10, maxlen: int = 1000):
    """Window list of rppg_VIDEO.py:392-399: deque(maxlen=1000); once ``len > window_len``
    the last ``window_len`` samples are analysed, every frame."""
    window_len = int(fps * window_seconds)
    i = np.arange(n_samples, dtype=np.int64)
    have = np.minimum(i + 1, maxlen)
    keep = have > window_len
    if window_len <= 0:
        keep[:] = False
    i = i[keep]
    return i.astype(np.int32), (i + 1 - window_len).astype(np.int32), np.full(i.shape, window_len, dtype=np.int32)


def sliding_windows(n_samples: int, window_len: int, hop: int):
    """Config-3 style windows: length ``window_len`` every ``hop`` samples."""
    starts = np.arange(0, max(0, n_samples - window_len) + 1, hop, dtype=np.int32)
    if n_samples < window_len:
        starts = starts[:0]
    return starts, np.full(starts.shape, window_len, dtype=np.int32)
