"""Measurement plugin for the reference's analysis harness (drop-in for
``analysis/measurement/green_avg.py``): copy or symlink this file into the harness's
``measurement/`` directory and run ``python main.py --methods green_avg_b200 ...``
(``analysis/main.py:29-31`` imports ``measurement.<name>`` and calls ``measure(video_path)``).

    measure(video_path) -> np.ndarray (N, 2) float64: [timestamp_seconds, bpm]

Same windowing (30 s rolling, 10 s acquisition), same float32 detrend, same FFT-peak rule;
the ROI mean and the spectra are computed by the CUDA library.

Landmarks are an INPUT of the B200 path (BASELINE.json north_star).  They are taken, in order,
from (1) a ``<video>.landmarks.npy`` side-car -- float (T, N, 2) normalised (x, y), optionally
with ``<video>.landmarks_valid.npy`` bool (T,) -- or (2) MediaPipe's FaceLandmarker run exactly
as ``analysis/utils/roi.py:69-90`` does, when mediapipe is installed.
"""
from __future__ import annotations

import os

import numpy as np


def read_video(video_path: str):
    """All BGR frames + fps.  Inside the reference's harness (cwd = analysis/) this IS the harness's
    own reader, ``utils.video_io.read_video`` (analysis/utils/video_io.py:8-33); the few lines below
    are only the stand-alone fallback with the same behaviour (decode is outside the B200 path)."""
    try:
        from utils.video_io import read_video as harness_read_video      # the harness's reader, when installed there
        return harness_read_video(video_path)
    except ImportError:
        pass
    import cv2
    if not os.path.exists(video_path):
        raise FileNotFoundError(f"Video not found: {video_path}")
    cap = cv2.VideoCapture(video_path)
    if not cap.isOpened():
        raise IOError(f"Failed to open video: {video_path}")
    fps = cap.get(cv2.CAP_PROP_FPS)
    frames = []
    while True:
        ret, frame = cap.read()
        if not ret:
            break
        frames.append(frame)
    cap.release()
    return frames, fps


def load_landmarks(video_path: str, frames, fps: float):
    side = os.path.splitext(video_path)[0] + ".landmarks.npy"
    if os.path.exists(side):
        lm = np.load(side)
        vpath = os.path.splitext(video_path)[0] + ".landmarks_valid.npy"
        valid = np.load(vpath) if os.path.exists(vpath) else None
        return lm, valid
    try:
        import mediapipe as mp
    except ImportError as e:
        raise RuntimeError(f"no landmark side-car ({side}) and mediapipe is not installed") from e
    import cv2
    model = os.environ.get("VHR_FACE_LANDMARKER", os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "utils",
                                                               "face_landmarker.task"))
    options = mp.tasks.vision.FaceLandmarkerOptions(
        base_options=mp.tasks.BaseOptions(model_asset_path=str(model)),
        running_mode=mp.tasks.vision.RunningMode.VIDEO, num_faces=1)
    lms, valid = [], []
    timestamps = np.arange(len(frames), dtype=float) / float(fps)
    with mp.tasks.vision.FaceLandmarker.create_from_options(options) as landmarker:
        for i, bgr in enumerate(frames):
            rgb = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)
            res = landmarker.detect_for_video(mp.Image(image_format=mp.ImageFormat.SRGB, data=rgb),
                                              int(timestamps[i] * 1000.0))
            if res and res.face_landmarks:
                lms.append([[p.x, p.y] for p in res.face_landmarks[0]])
                valid.append(True)
            else:
                lms.append(None)
                valid.append(False)
    n = max((len(l) for l in lms if l is not None), default=1)
    arr = np.zeros((len(frames), n, 2))
    for i, l in enumerate(lms):
        if l is not None:
            arr[i] = l
    return arr, np.asarray(valid)


def measure(video_path: str) -> np.ndarray:
    from video_heart_rate_b200 import default_engine
    from video_heart_rate_b200.pipeline import green_avg_measure
    frames, fps = read_video(video_path)
    if not frames:
        return np.zeros((0, 2))
    lm, valid = load_landmarks(video_path, frames, fps)
    return green_avg_measure(default_engine(), np.stack(frames), fps, lm, valid)
