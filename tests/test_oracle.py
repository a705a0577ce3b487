"""CPU tests: the oracle restatement against (a) the golden vectors made by executing the
reference's own functions (tests/golden/make_golden.py), (b) the third-party arithmetic
it restates (cv2 / numpy / scipy), (c) the reference itself when /root/reference exists.
"""
import os
import warnings

import numpy as np
import pytest

from oracle import bpm as obpm
from oracle import evm as oevm
from oracle import roi as oroi
from oracle import synth as osynth
from oracle import ref_loader


# ----------------------------------------------------------------------------- rectangles
def test_rect_geometry_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "roi_rect.npz"))
    for i in range(g["geo_hw"].shape[0]):
        h, w = (int(v) for v in g["geo_hw"][i])
        xs, ys = g["geo_xs"][i], g["geo_ys"][i]
        bbc = oroi.bbox_from_landmarks_clamped(xs, ys, w, h)
        assert tuple(g["geo_bb_clamped"][i]) == bbc
        assert tuple(g["geo_cheek_clamped"][i]) == oroi.cheek_roi_from_bbox(bbc, w, h)
        bb, fh, ck = oroi.process_frame_rects(xs, ys, w, h)
        assert tuple(g["geo_bb_video"][i]) == bb
        assert tuple(g["geo_forehead_video"][i]) == fh
        assert tuple(g["geo_cheek_video"][i]) == ck


def test_rect_means_match_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "roi_rect.npz"))
    for i in range(int(g["n_px"])):
        frame = g[f"px_frame_{i}"]
        xs, ys = g[f"px_xs_{i}"], g[f"px_ys_{i}"]
        h, w = frame.shape[:2]
        ck = oroi.cheek_roi_from_bbox(oroi.bbox_from_landmarks_clamped(xs, ys, w, h), w, h)
        np.testing.assert_array_equal(oroi.rect_mean(frame, ck), g[f"px_mean_clean_{i}"])  # bit-exact incl. NaN
        got = oroi.process_frame_green(frame, xs, ys)
        np.testing.assert_array_equal(np.float64(got), g[f"px_video_green_{i}"])


def test_outline_model_matches_cv2_rectangle():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for i in range(1500):
        h, w = int(rng.integers(5, 80)), int(rng.integers(5, 80))
        x1, x2 = (int(v) for v in rng.integers(-10, w + 10, 2))
        y1, y2 = (int(v) for v in rng.integers(-10, h + 10, 2))
        if i % 10 == 0:
            x2 = x1
        img = np.zeros((h, w, 3), np.uint8)
        cv2.rectangle(img, (x1, y1), (x2, y2), (255, 0, 0), 2)
        np.testing.assert_array_equal(oroi.outline_mask(h, w, (x1, y1, x2, y2)), img[:, :, 0] > 0)


# --------------------------------------------------------------------------------- polygons
def test_poly_mask_rectangle_and_triangle():
    m = oroi.poly_mask(10, 12, [(2, 3), (7, 3), (7, 6), (2, 6)])
    ref = np.zeros((10, 12), bool)
    ref[3:7, 2:8] = True
    np.testing.assert_array_equal(m, ref)
    t = oroi.poly_mask(8, 8, [(0, 0), (6, 0), (0, 6)])
    yy, xx = np.mgrid[0:8, 0:8]
    np.testing.assert_array_equal(t, (xx + yy) <= 6)
    assert oroi.poly_mask(5, 5, []).sum() == 0
    assert oroi.poly_mask(5, 5, [(2, 2)]).sum() == 1
    assert oroi.poly_mask(5, 5, [(0, 0), (4, 4)]).sum() == 5


def test_poly_mask_close_to_fillpoly():
    """Reported, not required: our exact rule vs cv2.fillPoly."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    tot = diff = 0
    for _ in range(100):
        n = int(rng.integers(3, 10))
        ang = np.sort(rng.uniform(0, 2 * np.pi, n))
        r = rng.uniform(10, 40, n)
        pts = np.stack([50 + r * np.cos(ang), 50 + r * np.sin(ang)], 1).astype(np.int32)
        ours = oroi.poly_mask(100, 100, pts)
        m = np.zeros((100, 100), np.uint8)
        cv2.fillPoly(m, [pts], 1)
        tot += int(m.sum())
        diff += int((ours != (m > 0)).sum())
        assert not np.any(ours & ~(m > 0))       # fillPoly ORs drawn edges: superset of ours
    assert diff / tot < 0.10       # boundary pixels only (about 6 % on these small polygons)


# -------------------------------------------------------------------------------------- EVM
@pytest.mark.parametrize("hw", [(144, 256), (135, 241), (68, 121), (9, 16), (5, 7)])
def test_pyrdown_pyrup_match_cv2(hw):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(hw[0])
    a = rng.integers(0, 256, hw + (3,)).astype(np.float32)
    d = cv2.pyrDown(a)
    np.testing.assert_array_equal(oevm.pyrdown(a).astype(np.float32), d)       # exact on integer data
    u = cv2.pyrUp(d, dstsize=(hw[1], hw[0]))
    ours = oevm.pyrup(d, (hw[1], hw[0]))
    assert np.abs(ours - u).max() <= 1e-4 * 255


def test_pyr_dims_chain():
    assert oevm.pyr_dims(1920, 1080, 4) == [(1920, 1080), (960, 540), (480, 270), (240, 135), (120, 68)]
    assert oevm.pyr_dims(256, 144, 4)[-1] == (16, 9)


def test_band_bins_counts():
    # c1 k=21..75, c2 k=42..240, c3 k=7..40
    b = oevm.band_bins(150, 5.0, 0.7, 4.0)
    assert (b[0], b[-1], len(b)) == (21, 75, 55)
    b = oevm.band_bins(1800, 30.0, 0.7, 4.0)
    assert (b[0], b[-1], len(b)) == (42, 240, 199)
    b = oevm.band_bins(300, 30.0, 0.7, 4.0)
    assert (b[0], b[-1], len(b)) == (7, 40, 34)


def test_ideal_bandpass_is_projection():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((150, 7))
    y = oevm.ideal_bandpass(x, 5.0, 0.7, 4.0)
    np.testing.assert_allclose(oevm.ideal_bandpass(y, 5.0, 0.7, 4.0), y, atol=1e-12)
    t = np.arange(150) / 5.0
    s = np.sin(2 * np.pi * 1.2 * t)                       # bin 36 exactly
    np.testing.assert_allclose(oevm.ideal_bandpass(s + 3.0, 5.0, 0.7, 4.0), s, atol=1e-12)


def test_evm_numpy_vs_cv2_pipeline():
    pytest.importorskip("cv2")
    p = osynth.SynthParams(T=40, H=72, W=128, fps=5.0, pulse_hz=1.2)
    fr = osynth.synth_frames(p)
    lv, filt, out = oevm.evm_clip(fr, 5.0, levels=3)
    lv2, filt2, out2, _ = oevm.evm_clip_cv2(fr, 5.0, levels=3, keep_out=True)
    assert np.abs(lv - lv2).max() <= 1e-4 * np.abs(lv).max()
    assert np.abs(filt - filt2).max() <= 1e-4 * max(1.0, np.abs(filt).max())
    assert np.abs(out - out2).max() <= 1e-4 * np.abs(out).max()


# ------------------------------------------------------------------------------------ synth
def test_synth_deterministic_and_pulse_recoverable():
    p = osynth.SynthParams(T=150, H=36, W=64, fps=5.0, pulse_hz=1.2, seed=0)
    a = osynth.synth_frames(p)
    b = osynth.synth_frames(p, 10, 20)
    np.testing.assert_array_equal(a[10:20], b)
    x0, y0, x1, y1 = p.face_rect()
    g = a[:, y0:y1, x0:x1, 1].reshape(150, -1).mean(1)
    bpm, k, _ = obpm.estimate_bpm_analysis(g - g.mean(), 5.0)
    assert bpm == pytest.approx(72.0)


# -------------------------------------------------------------------------------------- BPM
def test_bpm_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "bpm.npz"))
    nan = float("nan")
    for i in range(int(g["n"])):
        x, fps = g[f"x_{i}"], float(g[f"fps_{i}"])
        exp = g[f"bpm_{i}"]
        sig = np.asarray(x, dtype=np.float32)
        sig = sig - np.mean(sig)
        a = obpm.estimate_bpm_analysis(sig, fps)[0]
        w = x - np.mean(x)
        v = obpm.estimate_bpm_video_fft(w, fps)[0]
        fb = obpm.bandpass_butterworth(w, fps, 0.7, 2, 2)
        fc = obpm.bandpass_cheby2(w, fps, 0.7, 2, 4)
        np.testing.assert_array_equal(fb, g[f"fb_{i}"])
        np.testing.assert_array_equal(fc, g[f"fc_{i}"])
        wb = obpm.estimate_bpm_welch(fb, fps)[0]
        wc = obpm.estimate_bpm_welch(fc, fps)[0]
        try:
            ff = obpm.bandpass_fir(w, fps, 0.7, 2)
            np.testing.assert_array_equal(ff, g[f"ff_{i}"])
            wf = obpm.estimate_bpm_welch(ff, fps)[0]
        except ValueError:
            assert g[f"ff_{i}"].size == 0
            wf = None
        wl = obpm.estimate_bpm_welch(w, fps, obpm.LIVE_BAND)[0]
        got = np.array([nan if q is None else q for q in (a, v, wb, wc, wf, wl)])
        np.testing.assert_array_equal(got, exp)


def test_live_sos_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "bpm.npz"))
    for j in range(2):
        fps = float(g[f"live_fps_{j}"])
        sos = obpm.live_sos_design(fps)
        np.testing.assert_array_equal(sos, g[f"live_sos_{j}"])
        f = obpm.LiveSOS(sos)
        y = np.array([f.push(v) for v in g[f"live_x_{j}"]])
        np.testing.assert_array_equal(y, g[f"live_y_{j}"])


def test_known_answers_from_survey():
    # clean 1.2 Hz pulse -> Welch 73.33 BPM, FFT estimators 72.0 BPM
    for fps in (5.0, 30.0):
        n = int(10 * fps)
        t = np.arange(n) / fps
        x = np.sin(2 * np.pi * 1.2 * t)
        assert obpm.estimate_bpm_welch(x, fps)[0] == pytest.approx(73.3333, abs=1e-3)
        assert obpm.estimate_bpm_video_fft(x, fps)[0] == pytest.approx(72.0)
        assert obpm.estimate_bpm_analysis(x, fps)[0] == pytest.approx(72.0)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the build container")
def test_green_avg_series_matches_reference_loop():
    """green_avg.measure's loop body (green_avg.py:32-50) replayed with the reference's
    own estimate_bpm on a synthetic trace == oracle.green_avg_series."""
    import types
    A = ref_loader.load_functions("analysis/utils/estimate_bpm.py", ["estimate_bpm"],
                                  extra_ns={"plt": types.SimpleNamespace()})
    from collections import deque
    rng = np.random.default_rng(9)
    for fps in (5.0, 29.97):
        n = int(45 * fps)
        t = np.arange(n) / fps
        green = 150 + np.sin(2 * np.pi * 1.4 * t) + 0.3 * rng.standard_normal(n)
        window_len, acq = int(30.0 * fps), int(10.0 * fps)
        dq = deque(maxlen=window_len)
        ts, bp = [], []
        for i, gval in enumerate(green):
            dq.append(float(gval))
            if len(dq) < acq:
                continue
            sig = np.asarray(dq, dtype=np.float32)
            sig = sig - np.mean(sig)
            b = A["estimate_bpm"](sig, fps)
            if b is not None:
                ts.append(i * (1 / fps))
                bp.append(b)
        ours, _ = obpm.green_avg_series(green, fps)
        np.testing.assert_array_equal(ours, np.column_stack([ts, bp]))


# ------------------------------------------------------------------- degradations / metric
def test_degrade_and_metric_match_reference_golden(golden_dir):
    from oracle import degrade as odeg
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    for bits in (9, 8, 7, 6, 5, 4):
        np.testing.assert_array_equal(odeg.quantise_colour(g["q_frame"], bits), g[f"q_bits_{bits}"])
    for j in range(int(g["n_m"])):
        al = odeg.align_truth(g[f"m_tt_{j}"], g[f"m_th_{j}"], g[f"m_meas_{j}"])
        np.testing.assert_array_equal(al, g[f"m_aligned_{j}"])
        assert odeg.mae(g[f"m_tt_{j}"], g[f"m_th_{j}"], g[f"m_meas_{j}"]) == float(g[f"m_mae_{j}"])


def test_noise_statistics_match_reference_distribution():
    """Our counter-based 12-term Irwin-Hall noise has the mean / std / tail mass the reference's
    np.random.normal(0, sigma) has (colour_noise.py:22): tails are checked against the normal law
    out to 4 sigma on 1.5 M draws (the previous four-byte sum stopped at 3.45 sigma)."""
    import math
    from oracle import degrade as odeg
    fr = np.full((32, 128, 128, 3), 128, dtype=np.uint8)
    for sigma in (5, 10, 20, 40):                       # NOISE_LEVELS, colour_noise.py:8
        d = odeg.add_noise(fr, sigma, seed=1).astype(np.float64) - 128.0
        if sigma <= 20:                                 # sigma 40 clips at 0 / 255
            assert abs(d.mean() + 0.5) < 0.3            # astype(uint8) truncation biases by -0.5
            assert abs(d.std() - sigma) < 0.03 * sigma + 0.3
        for z, lo in ((2.0, 0.8), (3.0, 0.6), (3.6, 0.3)):       # Irwin-Hall(12) tails are a little lighter than normal
            if (z + 0.1) * sigma > 120 or (z > 2.5 and sigma < 10):
                continue
            frac = float(np.mean(np.abs(d + 0.5) > z * sigma))
            expect = math.erfc(z / math.sqrt(2.0))
            assert lo * expect < frac < 1.3 * expect, (sigma, z, frac, expect)
    d = odeg.add_noise(fr, 20, seed=2).astype(np.float64) - 128.0
    assert np.abs(d).max() > 4.0 * 20


# ------------------------------------------------------------------ round-2 pins (bpm_extra.npz)
def test_psd_plot_oracle_pinned_by_reference(golden_dir):
    """oracle.bpm.psd_plot_estimate against green_avg_psd_plot.py:34-63 EXECUTED (make_golden.gen_bpm_extra)."""
    g = np.load(os.path.join(golden_dir, "bpm_extra.npz"))
    for i in range(int(g["n_psd"])):
        bpm, _ = obpm.psd_plot_estimate(g[f"psd_x_{i}"], float(g[f"psd_fps_{i}"]))
        exp = float(g[f"psd_bpm_{i}"])
        assert bpm == exp or (np.isnan(bpm) and np.isnan(exp)), (i, bpm, exp)


def test_multicolumn_and_signed_band_oracle_pinned(golden_dir):
    """(T,3) best-column branch (estimate_bpm.py:59-64) and the signed-frequency mask of
    rppg_VIDEO.py:137-140 with FREQ_LOW <= 0, against the reference executed."""
    g = np.load(os.path.join(golden_dir, "bpm_extra.npz"))
    for j in range(int(g["n_mc"])):
        bpm, _, _ = obpm.estimate_bpm_analysis(g[f"mc_x_{j}"], float(g[f"mc_fps_{j}"]))
        assert bpm == float(g[f"mc_bpm_{j}"])
    for k in range(int(g["n_vf"])):
        lo, hi = g[f"vf_band_{k}"]
        bpm, _ = obpm.estimate_bpm_video_fft(g[f"vf_x_{k}"], float(g[f"vf_fps_{k}"]), (float(lo), float(hi)))
        exp = float(g[f"vf_bpm_{k}"])
        assert (bpm is None and np.isnan(exp)) or bpm == exp


def test_c_helpers_equal_numpy_forms():
    """oracle/csrc/synth_ref.c (gcc) reproduces oracle.synth.synth_frames and oracle.degrade.add_noise bit for bit."""
    from oracle import degrade as odeg, fast
    if fast._lib() is None:
        pytest.skip("gcc not available: the C helpers were not built")
    for kw in (dict(T=7, H=37, W=52, fps=5.0, pulse_hz=1.2, seed=3, clip=2), dict(T=4, H=72, W=128, fps=30.0, pulse_hz=2.1, seed=9, clip=40, noise_sigma=3.5)):
        p = osynth.SynthParams(**kw)
        np.testing.assert_array_equal(fast.synth_frames(p, 1, p.T), osynth.synth_frames(p, 1, p.T))
        fr = osynth.synth_frames(p)
        for sigma in (5, 40):
            np.testing.assert_array_equal(fast.add_noise(fr, sigma, seed=p.seed, clip=p.clip, t0=3), odeg.add_noise(fr, sigma, seed=p.seed, clip=p.clip, t0=3))


def test_streaming_trace_equals_clip_form():
    """oracle.evm.evm_roi_trace_streaming (used for the configuration goldens) returns exactly
    evm_clip_cv2's ROI means."""
    p = osynth.SynthParams(T=40, H=90, W=160, fps=10.0, pulse_hz=1.2, seed=1)
    fr = osynth.synth_frames(p)
    rects = np.tile(np.array([[30, 20, 100, 60], [50, 30, 70, 88]], dtype=np.int32), (40, 1, 1))
    a = oevm.evm_clip_cv2(fr, 10.0, 3, 0.7, 4.0, 50.0, rects=rects)[3]
    b = oevm.evm_roi_trace_streaming(lambda: (fr[i:i + 16] for i in range(0, 40, 16)), 40, 90, 160, 10.0, rects, 3)[2]
    np.testing.assert_array_equal(a, b)


def test_config_goldens_are_sane(golden_dir):
    """tests/golden/configs.npz (make_config_golden.py): every oracle BPM is the bin nearest the injected pulse
    for the clean configurations; the file covers all 64 c4 clips, c2, the 51 c3 windows and the 512 c5 windows."""
    g = np.load(os.path.join(golden_dir, "configs.npz"))
    assert len(g["c4_ids"]) == 64 and g["c4_trace"].shape == (64, 1800)
    assert len(g["c3_ids"]) == 51 and len(g["c5_ids"]) == 512 and g["c2_trace"].shape == (1, 1800)
    for c in range(64):
        f = 0.8 + c * (2.4 / 63.0)
        assert int(g["c4_bin"][c]) == int(round(f * 60.0))                # 1800 samples at 30 FPS: 1 bin = 1/60 Hz
    assert float(g["c2_bpm"][0]) == 72.0 and set(np.round(g["c3_bpm"], 6)) == {84.0}
