"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN FUNCTIONS verbatim.

Run in the build container (``/root/reference`` present):  ``python tests/golden/make_golden.py``

The reference has no tests, fixtures or golden vectors, so these files
are the pin: each array below is the output of an unmodified reference function body
(AST-extracted by ``oracle/ref_loader.py``) on a seeded input that is stored next to it.
Both the oracle restatement (``-m "not gpu"``) and the CUDA path (``-m gpu``) are checked
against them.  numpy 2.3.5 / scipy 1.18.1 / cv2 4.13.0.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader  # noqa: E402


def trace(rng, n, fps, f_hz, amp=1.0, noise=0.5, drift=0.0, base=150.0):
    t = np.arange(n) / fps
    return (base + amp * np.sin(2 * np.pi * f_hz * t + rng.uniform(0, 6.28))
            + drift * t + noise * rng.standard_normal(n))


def gen_roi(out):
    """Rectangle geometry + ROI mean: rppg_VIDEO.py get_roi_coords/process_frame/get_avg,
    analysis/utils/roi.py _bbox_from_landmarks/_cheek_roi_from_bbox."""
    V = ref_loader.load_functions("rppg_VIDEO.py", ["get_roi_coords", "get_avg", "process_frame"],
                                  extra_ns={"green_signal_cheek": []})
    R = ref_loader.load_functions("analysis/utils/roi.py", ["_bbox_from_landmarks", "_cheek_roi_from_bbox"],
                                  extra_ns={"CHEEK_HR": 0.15, "CHEEK_TOP": 0.40, "CHEEK_BOT": 0.65})
    rng = np.random.default_rng(1234)

    def landmarks(ci):
        n_lm = 12
        cx, cy = rng.uniform(0.3, 0.7), rng.uniform(0.3, 0.7)
        sx, sy = rng.uniform(0.1, 0.45), rng.uniform(0.1, 0.45)
        if ci % 7 == 3:   # face partly outside the frame -> negative / >1 landmarks
            cx = rng.choice([0.02, 0.98])
        if ci % 11 == 5:
            cy = rng.choice([0.03, 0.97])
        return cx + sx * rng.uniform(-1, 1, n_lm), cy + sy * rng.uniform(-1, 1, n_lm)

    # (1) geometry only, all resolutions of the degradation grid
    sizes = [(144, 256), (240, 426), (360, 640), (480, 640), (720, 1280), (1080, 1920), (97, 131)]
    geo = []
    for ci in range(210):
        h, w = sizes[ci % len(sizes)]
        xs, ys = landmarks(ci)
        lms = [ref_loader.Landmark(x, y) for x, y in zip(xs, ys)]
        bbc = R["_bbox_from_landmarks"](lms, w, h)
        ckc = R["_cheek_roi_from_bbox"](bbc, w, h)
        bbv = (int(min(xs) * w), int(min(ys) * h), int(max(xs) * w), int(max(ys) * h))  # rppg_VIDEO.py:95-98 inline expr
        dummy = np.zeros((h, w, 3), np.uint8)
        fhv = V["get_roi_coords"](*bbv, 0.25, 0.00, 0.25, dummy)
        ckv = V["get_roi_coords"](*bbv, 0.15, 0.4, 0.65, dummy)
        geo.append((h, w, xs, ys, bbc, ckc, bbv, fhv, ckv))
    rec = dict(geo_hw=np.array([[g[0], g[1]] for g in geo]),
               geo_xs=np.array([g[2] for g in geo]), geo_ys=np.array([g[3] for g in geo]),
               geo_bb_clamped=np.array([g[4] for g in geo]), geo_cheek_clamped=np.array([g[5] for g in geo]),
               geo_bb_video=np.array([g[6] for g in geo]), geo_forehead_video=np.array([g[7] for g in geo]),
               geo_cheek_video=np.array([g[8] for g in geo]))
    # (2) pixel cases on small frames (frame stored)
    small = [(144, 256), (97, 131), (120, 160), (61, 67)]
    n_px = 28
    for ci in range(n_px):
        h, w = small[ci % len(small)]
        xs, ys = landmarks(ci)
        lms = [ref_loader.Landmark(x, y) for x, y in zip(xs, ys)]
        frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        bb = R["_bbox_from_landmarks"](lms, w, h)
        ck = R["_cheek_roi_from_bbox"](bb, w, h)
        roi = frame[ck[1]:ck[3], ck[0]:ck[2]]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mean_clean = np.array([np.mean(roi[:, :, c]) for c in range(3)])   # green_avg.py:34 per channel
            sig = V["__ns__"]["green_signal_cheek"]
            sig.clear()
            fr2 = frame.copy()
            V["process_frame"](fr2, lms)                                        # rppg_VIDEO.py:91-110
        rec[f"px_frame_{ci}"] = frame
        rec[f"px_xs_{ci}"] = xs
        rec[f"px_ys_{ci}"] = ys
        rec[f"px_mean_clean_{ci}"] = mean_clean
        rec[f"px_video_green_{ci}"] = np.float64(sig[0])
        rec[f"px_drawn_{ci}"] = np.any(fr2 != frame, axis=2)
    rec["n_px"] = n_px
    np.savez_compressed(os.path.join(out, "roi_rect.npz"), **rec)
    print("roi_rect:", len(geo), "geometry cases,", n_px, "pixel cases")


def gen_bpm(out):
    """estimate_bpm (analysis), estimate_bpm / estimate_bpm_welch / bandpass_* (VIDEO)."""
    import types
    plt_stub = types.SimpleNamespace()
    A = ref_loader.load_functions("analysis/utils/estimate_bpm.py", ["estimate_bpm"], extra_ns={"plt": plt_stub})
    V = ref_loader.load_functions("rppg_VIDEO.py", ["estimate_bpm", "estimate_bpm_welch", "bandpass_butterworth",
                                                    "bandpass_cheby2", "bandpass_fir"])
    L = ref_loader.load_functions("rppg_LIVESTREAM.py", ["estimate_bpm_welch", "live_sos_init", "live_sos_push",
                                                         "live_sos_reset", "bandpass_butterworth_eqn"],
                                  extra_ns={"_live_sos": None, "_live_zi": None})
    rng = np.random.default_rng(4321)
    rec = {}
    i = 0
    for fps in (5.0, 10.0, 15.0, 25.0, 29.97, 30.0):
        for f_hz in (0.8, 1.2, 1.9, 2.6):
            for n_s in (10, 17, 30):
                n = int(n_s * fps)
                x = trace(rng, n, fps, f_hz, amp=rng.uniform(0.2, 2.0), noise=rng.uniform(0.05, 1.0),
                          drift=rng.uniform(-0.2, 0.2))
                # analysis: float32 detrend then estimate (green_avg.py:42-44)
                sig = np.asarray(x, dtype=np.float32)
                sig = sig - np.mean(sig)
                bpm_a = A["estimate_bpm"](sig, fps)
                # VIDEO: detrend, three filters, Welch (rppg_VIDEO.py:398-409); fft estimator too
                w = x - np.mean(x)
                bpm_fft_v = V["estimate_bpm"](w, fps)
                fb = V["bandpass_butterworth"](w, fps, 0.7, 2, order=2)
                fc = V["bandpass_cheby2"](w, fps, 0.7, 2, order=4)
                try:
                    ff = V["bandpass_fir"](w, fps, 0.7, 2)
                    bpm_wf = V["estimate_bpm_welch"](ff, fps)
                except ValueError:
                    ff = np.zeros(0)
                    bpm_wf = None
                bpm_wb = V["estimate_bpm_welch"](fb, fps)
                bpm_wc = V["estimate_bpm_welch"](fc, fps)
                # LIVE Welch band on the raw detrended window (rppg_LIVESTREAM.py:347)
                bpm_wl = L["estimate_bpm_welch"](w, fps)
                nan = float("nan")
                rec[f"x_{i}"] = x
                rec[f"fps_{i}"] = np.float64(fps)
                rec[f"bpm_{i}"] = np.array([bpm_a if bpm_a is not None else nan,
                                            bpm_fft_v if bpm_fft_v is not None else nan,
                                            bpm_wb if bpm_wb is not None else nan,
                                            bpm_wc if bpm_wc is not None else nan,
                                            bpm_wf if bpm_wf is not None else nan,
                                            bpm_wl if bpm_wl is not None else nan])
                rec[f"fb_{i}"] = fb
                rec[f"fc_{i}"] = fc
                rec[f"ff_{i}"] = ff
                i += 1
    rec["n"] = i
    # live causal SOS (rppg_LIVESTREAM.py:207-251) at 30 and 15 fps
    for j, fps in enumerate((30.0, 15.0)):
        sos = L["bandpass_butterworth_eqn"](None, fps, 40 / 60, 150 / 60, 4)
        L["__ns__"]["live_sos_init"](sos)
        x = trace(rng, 400, fps, 1.3)
        y = np.array([L["__ns__"]["live_sos_push"](v) for v in x])
        rec[f"live_x_{j}"] = x
        rec[f"live_y_{j}"] = y
        rec[f"live_sos_{j}"] = sos
        rec[f"live_fps_{j}"] = np.float64(fps)
    np.savez_compressed(os.path.join(out, "bpm.npz"), **rec)
    print("bpm:", i, "traces")


def gen_metrics(out):
    """colour_quantisation.quantise_colour, video_io.interpolate_hr_to_frames, mae.py:32-36."""
    import pandas as pd
    Q = ref_loader.load_functions("analysis/degradation/colour_quantisation.py", ["quantise_colour"])
    Vio = ref_loader.load_functions("analysis/utils/video_io.py", ["interpolate_hr_to_frames"], extra_ns={"pd": pd})
    rng = np.random.default_rng(99)
    rec = {}
    frame = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    rec["q_frame"] = frame
    for bits in (9, 8, 7, 6, 5, 4):                      # COLOUR_DEPTHS, colour_quantisation.py:9
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rec[f"q_bits_{bits}"] = Q["quantise_colour"](frame, bits)
    for j in range(4):
        n, m = int(rng.integers(5, 60)), int(rng.integers(10, 400))
        tt = np.sort(rng.uniform(0, 60, n))
        if j == 1:
            tt = tt + 5.0                                   # measurements before the first truth sample
        th = rng.uniform(50, 120, n)
        meas = np.column_stack([np.sort(rng.uniform(0, 70, m)), rng.uniform(40, 200, m)])
        truth = pd.DataFrame({"timestamp": tt, "heart_rate": th})
        aligned = Vio["interpolate_hr_to_frames"](truth, meas)
        mae = float(np.mean(np.abs(meas[:, 1].astype(float) - aligned[:, 1].astype(float))))   # mae.py:32-36
        rec[f"m_tt_{j}"] = tt; rec[f"m_th_{j}"] = th; rec[f"m_meas_{j}"] = meas
        rec[f"m_aligned_{j}"] = aligned; rec[f"m_mae_{j}"] = np.float64(mae)
    rec["n_m"] = 4
    np.savez_compressed(os.path.join(out, "metrics.npz"), **rec)
    print("metrics: quantise x6, alignment x4")


def gen_bpm_extra(out):
    """Pins that round 1 left open (bpm_extra.npz):
    * green_avg_psd_plot.py:34-63 ``_bandpass_butterworth`` / ``_estimate_bpm`` executed on the float32
      z-scored window of :173-176;
    * the (T,C) best-column branch of analysis/utils/estimate_bpm.py:59-64 (C = 3);
    * rppg_VIDEO.py:129-147 ``estimate_bpm`` with FREQ_LOW <= 0 (signed fftfreq mask: DC and the
      negative-frequency bins become candidates)."""
    import types
    P = ref_loader.load_functions("analysis/measurement/green_avg_psd_plot.py", ["_bandpass_butterworth", "_estimate_bpm"])
    A = ref_loader.load_functions("analysis/utils/estimate_bpm.py", ["estimate_bpm"], extra_ns={"plt": types.SimpleNamespace()})
    rng = np.random.default_rng(2468)
    rec = {}
    i = 0
    for fps in (5.0, 15.0, 29.97, 30.0):
        for f_hz in (0.9, 1.45, 2.3):
            for n_s in (10.0, 12.5):
                n = int(round(n_s * fps))
                x = trace(rng, n, fps, f_hz, amp=rng.uniform(0.3, 1.5), noise=rng.uniform(0.05, 0.6), drift=rng.uniform(-0.1, 0.1))
                sig = np.asarray(x, dtype=np.float32)
                sig = (sig - np.mean(sig)) / np.std(sig)                               # :174-175
                filt = P["_bandpass_butterworth"](sig, fps, 40 / 60, 200 / 60, 2)     # :176 (FILTER_ORDER = 2)
                r = P["_estimate_bpm"](filt, fps)
                rec[f"psd_x_{i}"] = x
                rec[f"psd_fps_{i}"] = np.float64(fps)
                rec[f"psd_bpm_{i}"] = np.float64(r[0] if isinstance(r, tuple) else np.nan)
                rec[f"psd_filt_{i}"] = np.asarray(filt, dtype=np.float64)
                i += 1
    rec["n_psd"] = i
    j = 0
    for fps in (5.0, 25.0, 30.0):
        for n in (int(10 * fps), int(17.3 * fps)):
            for rep in range(3):
                amp = rng.uniform(0.2, 2.0, 3)
                X = np.stack([trace(rng, n, fps, f, amp=a, noise=0.4, base=0.0) for f, a in zip(rng.uniform(0.8, 3.0, 3), amp)], 1)
                X = X - X.mean(0)
                bpm = A["estimate_bpm"](X, fps)
                rec[f"mc_x_{j}"] = X
                rec[f"mc_fps_{j}"] = np.float64(fps)
                rec[f"mc_bpm_{j}"] = np.float64(np.nan if bpm is None else bpm)
                j += 1
    rec["n_mc"] = j
    k = 0
    for (lo, hi) in ((-0.5, 2.0), (0.0, 1.0), (-3.0, -0.7), (0.0, 0.0)):
        V = ref_loader.load_functions("rppg_VIDEO.py", ["estimate_bpm"])
        V["__ns__"]["FREQ_LOW"], V["__ns__"]["FREQ_HIGH"] = lo, hi
        for fps, n in ((30.0, 300), (5.0, 51)):
            x = trace(rng, n, fps, 1.2, base=rng.choice([0.0, 3.0]))                 # with and without a DC term
            bpm = V["estimate_bpm"](x, fps)
            rec[f"vf_x_{k}"] = x
            rec[f"vf_fps_{k}"] = np.float64(fps)
            rec[f"vf_band_{k}"] = np.array([lo, hi])
            rec[f"vf_bpm_{k}"] = np.float64(np.nan if bpm is None else bpm)
            k += 1
    rec["n_vf"] = k
    np.savez_compressed(os.path.join(out, "bpm_extra.npz"), **rec)
    print("bpm_extra:", i, "psd windows,", j, "multi-column signals,", k, "signed-band cases")


def gen_ica(out):
    """analysis/measurement/ica.py:24-76 replayed on synthetic mean-BGR traces (ica.npz): the loop of
    oracle.ica.ica_series with the reference's own ``estimate_bpm`` (executed verbatim) and scikit-learn's FastICA
    with the reference's arguments.  Per analysed frame: converged flag, BPM."""
    import types
    from oracle import ica as oica
    A = ref_loader.load_functions("analysis/utils/estimate_bpm.py", ["estimate_bpm"], extra_ns={"plt": types.SimpleNamespace()})
    rng = np.random.default_rng(1357)
    rec = {}
    cases = [(30.0, 420, 1.3, 0.6), (30.0, 360, 1.9, 0.25), (10.0, 200, 1.1, 0.4), (25.0, 300, 2.4, 1.0)]
    for j, (fps, T, f_hz, noise) in enumerate(cases):
        t = np.arange(T) / fps
        pulse = np.sin(2 * np.pi * f_hz * t + rng.uniform(0, 6.28))
        mix = rng.uniform(0.3, 1.0, 3)
        bgr = np.stack([110 + 20 * c + mix[c] * pulse + noise * rng.standard_normal(T) + 0.4 * np.sin(2 * np.pi * 0.25 * t + c)
                        for c in range(3)], 1)
        res = oica.ica_series(bgr, fps, estimate=lambda s, fs: A["estimate_bpm"](s, fs=fs))
        rec[f"ica_bgr_{j}"] = bgr
        rec[f"ica_fps_{j}"] = np.float64(fps)
        rec[f"ica_frame_{j}"] = np.array([r[0] for r in res], dtype=np.int32)
        rec[f"ica_conv_{j}"] = np.array([r[1] for r in res], dtype=bool)
        rec[f"ica_bpm_{j}"] = np.array([np.nan if r[2] is None else r[2] for r in res], dtype=np.float64)
    rec["n_ica"] = len(cases)
    np.savez_compressed(os.path.join(out, "ica.npz"), **rec)
    print("ica:", len(cases), "traces,", sum(len(rec[f"ica_frame_{j}"]) for j in range(len(cases))), "windows,",
          sum(int(rec[f"ica_conv_{j}"].sum()) for j in range(len(cases))), "converged in scikit-learn")


def main():
    if not ref_loader.available():
        raise SystemExit("reference tree not found; golden vectors can only be made in the build container")
    gen_roi(HERE)
    gen_bpm(HERE)
    gen_metrics(HERE)
    gen_bpm_extra(HERE)
    gen_ica(HERE)


if __name__ == "__main__":
    main()
