"""ICA measurement -- restatement of analysis/measurement/ica.py:24-76 (oracle; test infrastructure).

``ica_series`` replays the reference's per-frame loop on a mean-BGR trace with the reference's own calls:
``sklearn.decomposition.FastICA`` (third-party, unpinned in requirements.txt; 1.9.0 in this image) with the
arguments of ica.py:36-44, the float32 / ddof=1 std normalisation of :53-61, the skip on
ConvergenceWarning of :64-69 and ``estimate_bpm`` over the three sources (:72).  Pinned by
tests/golden/make_golden.py:gen_ica, which runs the same loop with the reference's ``estimate_bpm`` executed
verbatim.  ``fastica_f64`` is a float64 NumPy statement of the same algorithm (what csrc/ica.cu computes),
used to show how far float32 LAPACK rounding moves the result."""
from __future__ import annotations

import warnings
from collections import deque

import numpy as np

from .bpm import estimate_bpm_analysis

WINDOW_SIZE = 10.0       # ica.py:10
ACQUISITION_TIME = 5.0   # ica.py:11


def ica_series(bgr, fps: float, estimate=None):
    """-> list of (frame index i, converged, bpm | None, bin) for every frame whose deque holds
    >= acquisition_len samples (ica.py:46-76; rows are emitted only where converged and bpm is not None)."""
    from sklearn.decomposition import FastICA
    from sklearn.exceptions import ConvergenceWarning
    window_len = int(WINDOW_SIZE * fps)
    acquisition_len = int(ACQUISITION_TIME * fps)
    dq = deque(maxlen=window_len)
    ica = FastICA(n_components=3, algorithm="parallel", fun="logcosh", max_iter=300, tol=1e-6,
                  whiten="unit-variance", random_state=0)
    out = []
    for i, v in enumerate(np.asarray(bgr)):
        dq.append(v)
        if len(dq) < acquisition_len:
            continue
        signal = np.asarray(dq, dtype=np.float32)
        std_vals = np.std(signal, axis=0, ddof=1)
        std_vals[std_vals == 0] = 1.0
        signal = signal / std_vals
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            sources = ica.fit_transform(signal)
            conv = not any(issubclass(wi.category, ConvergenceWarning) for wi in w)
        if estimate is not None:
            bpm, k = estimate(sources, fps), -1
        else:
            bpm, k, _ = estimate_bpm_analysis(sources, fps)
        out.append((i, conv, bpm, k))
    return out


def w_init_reference() -> np.ndarray:
    """FastICA(random_state=0): check_random_state(0).normal(size=(3, 3)) at every fit."""
    return np.random.RandomState(0).normal(size=(3, 3))


def _symdecor(W):
    s, u = np.linalg.eigh(W @ W.T)
    s = np.clip(s, np.finfo(np.float64).tiny, None)
    return (u * (1.0 / np.sqrt(s))) @ u.T @ W


def fastica_f64(signal32: np.ndarray, max_iter: int = 300, tol: float = 1e-6):
    """float64 FastICA on an (n,3) float32 window already divided by its std: -> (sources (n,3), n_iter, converged)."""
    XT = signal32.astype(np.float32).T.copy()
    XT = (XT - XT.mean(axis=-1, dtype=np.float64).astype(np.float32)[:, None]).astype(np.float64)
    n = XT.shape[1]
    ev, u = np.linalg.eigh(XT @ XT.T)
    order = np.argsort(ev)[::-1]
    d, u = np.sqrt(ev[order]), u[:, order]
    u = u * np.where(u[0] < 0, -1.0, 1.0)
    K = (u / d).T
    X1 = K @ XT * np.sqrt(n)
    W = _symdecor(w_init_reference().astype(np.float32).astype(np.float64))
    conv, it = False, 0
    for it in range(1, max_iter + 1):
        t = np.tanh(W @ X1)
        W1 = _symdecor(t @ X1.T / n - (1.0 - t ** 2).mean(axis=1)[:, None] * W)
        lim = np.max(np.abs(np.abs(np.einsum("ij,ij->i", W1, W)) - 1))
        W = W1
        if lim < tol:
            conv = True
            break
    S = (W @ K @ XT).T
    return S / S.std(axis=0, keepdims=True), it, conv
