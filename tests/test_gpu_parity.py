"""GPU parity tests (``-m gpu``): the CUDA path, called through the C ABI (ctypes ->
libvhr_b200.so), against the CPU oracle and the committed golden vectors.

Tolerances (stated up front, SURVEY.md section 7 "hard parts"):
  * integer / index / byte work (synthetic clips, masks, rectangle means of uint8 frames,
    pyramid levels 1-2, chosen spectral bin, BPM value): bit-exact;
  * float32 pixels (pyramid level >= 3, filtered/amplified level, magnified frames) and ROI
    traces: max |a - b| <= 1e-4 * max |b| (relative to the tensor's scale).
"""
import os
import sys

import numpy as np
import pytest

from oracle import bpm as obpm
from oracle import evm as oevm
from oracle import roi as oroi
from oracle import synth as osynth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REL = 1e-4


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="module")
def eng(vhr):
    e = vhr.Engine(0)
    yield e
    e.close()


def spec_pair(vhr, **kw):
    return vhr.SynthSpec(**kw), osynth.SynthParams(**kw)


# ------------------------------------------------------------------------------------ synth
@pytest.mark.parametrize("shape", [(7, 36, 64), (3, 61, 67), (5, 144, 256)])
def test_synth_bit_identical(vhr, eng, shape):
    T, H, W = shape
    s, o = spec_pair(vhr, T=T, H=H, W=W, fps=5.0, pulse_hz=1.2, seed=3, clip=2, noise_sigma=2.0)
    np.testing.assert_array_equal(s.pulse_table(), o.pulse_table())
    got = eng.synth_clip(s).cpu().numpy()
    np.testing.assert_array_equal(got, osynth.synth_frames(o))
    part = eng.synth_clip(s, t0=2, t1=T).cpu().numpy()
    np.testing.assert_array_equal(part, got[2:])


# ---------------------------------------------------------------------------------- pyramid
@pytest.mark.parametrize("hw", [(144, 256), (480, 640), (90, 160), (67, 1920), (135, 248), (97, 131), (61, 67), (9, 16), (5, 7)])
@pytest.mark.parametrize("levels", [1, 2, 3, 4])
def test_pyrdown_cascade(vhr, eng, hw, levels):
    import torch
    H, W = hw
    rng = np.random.default_rng(H * 31 + W + levels)
    fr = rng.integers(0, 256, (3, H, W, 3), dtype=np.uint8)
    got = eng.pyrdown(torch.as_tensor(fr, device=eng.tdev), levels).cpu().numpy()
    ref = oevm.pyrdown_cascade(fr, levels)
    assert got.shape == ref.shape                      # bit-exact level indexing / sizes
    if levels <= 2:
        np.testing.assert_array_equal(got, ref.astype(np.float32))    # exact integers * 2^-8l
    else:
        assert rel_err(got, ref) <= REL


def test_pyrdown_generic_kernel_on_aligned_shapes(vhr, eng, monkeypatch):
    """The generic kernel (used for W % 16 != 0) gives the same values as the streaming kernel."""
    import torch
    monkeypatch.setenv("VHR_PYRDOWN_GENERIC", "1")
    rng = np.random.default_rng(77)
    fr = rng.integers(0, 256, (2, 144, 256, 3), dtype=np.uint8)
    frd = torch.as_tensor(fr, device=eng.tdev)
    gen = eng.pyrdown(frd, 4).cpu().numpy()
    monkeypatch.delenv("VHR_PYRDOWN_GENERIC")
    fast = eng.pyrdown(frd, 4).cpu().numpy()
    assert rel_err(gen, oevm.pyrdown_cascade(fr, 4)) <= REL
    assert rel_err(fast, gen) <= 1e-6


def test_pyrdown_many_frames_persistent_split(vhr, eng):
    """More frames than CTAs can take whole: shares start and end mid-frame."""
    import torch
    rng = np.random.default_rng(11)
    fr = rng.integers(0, 256, (700, 40, 64, 3), dtype=np.uint8)
    got = eng.pyrdown(torch.as_tensor(fr, device=eng.tdev), 2).cpu().numpy()
    np.testing.assert_array_equal(got, oevm.pyrdown_cascade(fr, 2).astype(np.float32))
    fr = rng.integers(0, 256, (400, 70, 96, 3), dtype=np.uint8)          # odd level heights, 4 levels
    got = eng.pyrdown(torch.as_tensor(fr, device=eng.tdev), 4).cpu().numpy()
    assert rel_err(got, oevm.pyrdown_cascade(fr, 4)) <= REL


@pytest.mark.parametrize("case", [(400, 70, 128, 4), (300, 70, 128, 3), (2, 40, 3840, 4), (2, 33, 2048, 3),
                                  (3, 200, 256, 5), (2, 130, 320, 6), (5, 1080, 1920, 4), (40, 36, 64, 2),
                                  (3, 90, 720, 4), (6, 61, 1296, 3), (2, 48, 3072, 4), (30, 40, 96, 5), (4, 135, 240, 1),
                                  (700, 40, 64, 2), (9, 270, 480, 4), (3, 135, 2048, 4), (5, 97, 176, 2), (7, 67, 1280, 4)])
def test_pyrdown_streaming_kernels(vhr, eng, case, monkeypatch):
    """The CUDA-core kernels on shapes that exercise shares crossing frames, wide frames, 5 and 6 levels, odd level heights
    and widths the streaming kernel does not take (W % 64 != 0 with >= 3 levels: the generic kernel steps in):
    pyrdown_stream.cu (registers + shuffles for levels 1-2, per-warp TMA input rings, upper levels by one warp in turn) and
    the generic kernel of pyrdown.cu.  Each is held to the oracle; they agree bit for bit on levels 1-2 and to 1e-6 above."""
    import torch
    T, H, W, levels = case
    rng = np.random.default_rng(T + H + W + levels)
    fr = rng.integers(0, 256, (T, H, W, 3), dtype=np.uint8)
    frd = torch.as_tensor(fr, device=eng.tdev)
    n_ref = min(T, 6)                                         # the oracle on the first and last frames
    sel = np.r_[0:n_ref // 2, T - (n_ref - n_ref // 2):T]
    ref = oevm.pyrdown_cascade(fr[sel], levels)
    outs = {}
    for impl, var, val in (("stream", "VHR_PYRDOWN_IMPL", "stream"), ("generic", "VHR_PYRDOWN_GENERIC", "1")):
        monkeypatch.setenv(var, val)                          # (the tcgen05 kernel has its own tests below)
        got = eng.pyrdown(frd, levels).cpu().numpy()
        monkeypatch.delenv(var)
        assert got.shape[1:] == ref.shape[1:]
        if levels <= 2:
            np.testing.assert_array_equal(got[sel], ref.astype(np.float32))
        else:
            assert rel_err(got[sel], ref) <= REL
        outs[impl] = got
    if levels <= 2:
        np.testing.assert_array_equal(outs["stream"], outs["generic"])
    else:
        assert rel_err(outs["stream"], outs["generic"]) <= 1e-6   # every frame, every share boundary


@pytest.mark.parametrize("case", [(5, 1080, 1920), (4, 720, 1280), (7, 480, 640), (300, 67, 1280), (9, 1081, 160), (3, 2160, 320),
                                  (400, 90, 720), (2, 1079, 160), (160, 64, 160), (3, 135, 240), (2, 1440, 2560)])
def test_pyrdown_tensor_core_kernel(vhr, eng, case, monkeypatch):
    """csrc/pyrdown_umma.cu (tcgen05.mma kind::i8 vertical composite, TMEM accumulators; the default for 4 levels and
    W % 80 == 0) against the oracle on the first and last frames and against the streaming kernel on every frame: several
    row tiles, a single tile, odd heights at every level, more items than CTAs, widths from 2 to 32 strips."""
    import torch
    T, H, W = case
    rng = np.random.default_rng(T + H + W)
    fr = rng.integers(0, 256, (T, H, W, 3), dtype=np.uint8)
    frd = torch.as_tensor(fr, device=eng.tdev)
    before = eng.launch_count()
    monkeypatch.setenv("VHR_PYRDOWN_IMPL", "umma")
    got = eng.pyrdown(frd, 4).cpu().numpy()
    assert eng.launch_count() == before + 1
    monkeypatch.setenv("VHR_PYRDOWN_IMPL", "stream")          # (W % 64 != 0: the generic kernel)
    other = eng.pyrdown(frd, 4).cpu().numpy()
    monkeypatch.delenv("VHR_PYRDOWN_IMPL")
    n_ref = min(T, 4)
    sel = np.r_[0:n_ref // 2, T - (n_ref - n_ref // 2):T]
    assert rel_err(got[sel], oevm.pyrdown_cascade(fr[sel], 4)) <= 2e-6      # measured 2e-7 (float32 levels 3-4); the contract is REL
    assert rel_err(got, other) <= 2e-6
    # the kernel is a pure function of the frame: frames repeated at other positions of the clip (other CTAs, other
    # accumulator buffers, other pipeline phases) give the same bits
    if T >= 4:
        fr2 = np.concatenate([fr[-2:], fr[:-2]])
        got2 = eng.pyrdown(torch.as_tensor(fr2, device=eng.tdev), 4).cpu().numpy()
        np.testing.assert_array_equal(got2, np.concatenate([got[-2:], got[:-2]]))


@pytest.mark.parametrize("case", [(2, 1080, 1920), (3, 720, 1280), (2, 90, 720)])
def test_pyrdown_tensor_core_accumulators_are_exact(vhr, eng, case):
    """The MMA stage alone: the raw TMEM accumulators of a few (item, strip) pairs -- first / interior / flush strip,
    top / bottom tile -- equal the integer product of the plan's weight slices with the image rows, bit for bit (TMA box
    layout, 128-byte swizzle, shared-memory descriptors, band offsets, baked reflect-101 borders)."""
    import sys
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tools", "probes"))
    import umma_emulate as em
    T, H, W = case
    rng = np.random.default_rng(H + W)
    fr = rng.integers(0, 256, (T, H, W, 3), dtype=np.uint8)
    frd = torch.as_tensor(fr, device=eng.tdev)
    plan = em.make_plan(H, W)
    nt, S = len(plan["tiles"]), plan["nstrips"]
    for item, strip in ((0, 0), (0, 1), (nt - 1, S - 2), (T * nt - 1, S - 1), (nt, S // 2)):
        _, acc = eng.pyrdown_tc_accumulators(frd, item, strip)
        f, t = divmod(item, nt)
        tile = plan["tiles"][t]
        img = fr[f].reshape(H, W * 3).astype(np.int64)
        D = np.zeros((128, 240), dtype=np.int64)
        for ks in range(tile["nks"]):
            rows = tile["i0"] + 32 * ks + np.arange(32)
            B = np.zeros((32, 240), dtype=np.int64)
            ok = (rows >= 0) & (rows < H)
            seg = img[rows[ok], 240 * strip: 240 * strip + 240]
            B[ok, :seg.shape[1]] = seg
            D += tile["slices"][ks] @ B
        np.testing.assert_array_equal(acc.cpu().numpy()[:tile["nr"]], D[:tile["nr"]])


def test_pyrdown_tensor_core_kernel_many_shapes_two_streams(vhr, eng, monkeypatch):
    """The weight blobs of the tensor-core kernel are cached per context and frame shape (8 slots) and are read by kernels
    on whatever stream the call was made on: twelve shapes interleaved on two streams, twice over (the second round evicts
    and reloads blobs), give the bits of a plain single-stream run."""
    import torch
    rng = np.random.default_rng(5)
    shapes = [(64 + 16 * k, 160 + 80 * (k % 4)) for k in range(12)]
    clips = [torch.as_tensor(rng.integers(0, 256, (3, h, w, 3), dtype=np.uint8), device=eng.tdev) for h, w in shapes]
    monkeypatch.setenv("VHR_PYRDOWN_IMPL", "stream")
    refs = [eng.pyrdown(c, 4).clone() for c in clips]
    monkeypatch.delenv("VHR_PYRDOWN_IMPL")
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for rnd in range(2):
        outs = []
        for k, c in enumerate(clips):
            with torch.cuda.stream(s1 if (k + rnd) % 2 == 0 else s2):
                outs.append(eng.pyrdown(c, 4))
        torch.cuda.synchronize()
        for got, ref in zip(outs, refs):
            assert rel_err(got.cpu().numpy(), ref.cpu().numpy()) <= 2e-6
        if rnd == 0:
            first = [o.clone() for o in outs]
        else:
            for a, b in zip(outs, first):
                assert torch.equal(a, b)


def test_pyrdown_tensor_core_kernel_constant_and_extremes(vhr, eng):
    """All-255 frames exercise the largest accumulators (255 * 256 per column, 255 * 65536 per level-2 value): every level
    of a constant image is that constant, exactly; a single bright pixel checks every weight of the composite filters."""
    import torch
    fr = np.full((2, 256, 320, 3), 255, dtype=np.uint8)
    got = eng.pyrdown(torch.as_tensor(fr, device=eng.tdev), 4).cpu().numpy()
    np.testing.assert_array_equal(got, np.full_like(got, 255.0))
    fr = np.zeros((6, 256, 320, 3), dtype=np.uint8)
    for k, (y, x) in enumerate([(0, 0), (255, 319), (128, 7), (31, 160), (100, 239), (1, 318)]):
        fr[k, y, x, k % 3] = 255
    got = eng.pyrdown(torch.as_tensor(fr, device=eng.tdev), 4).cpu().numpy()
    ref = oevm.pyrdown_cascade(fr, 4)
    np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-9)


# --------------------------------------------------------------------------------- bandpass
@pytest.mark.parametrize("case", [(150, 5.0, 432), (300, 30.0, 777), (299, 29.97, 64), (64, 10.0, 5), (1800, 30.0, 96)])
def test_temporal_bandpass(vhr, eng, case):
    import torch
    T, fps, P = case
    rng = np.random.default_rng(T)
    t = np.arange(T) / fps
    x = (150 + 20 * rng.standard_normal((1, P)) + rng.standard_normal((T, P))
         + 1.5 * np.sin(2 * np.pi * 1.2 * t)[:, None]).astype(np.float32)
    got = eng.bandpass(torch.as_tensor(x, device=eng.tdev), fps, 0.7, 4.0, gain=50.0).cpu().numpy()
    ref = 50.0 * oevm.ideal_bandpass(x, fps, 0.7, 4.0)
    assert rel_err(got, ref) <= REL
    n, k0, k1 = eng.band_bins(T, fps, 0.7, 4.0)
    bins = oevm.band_bins(T, fps, 0.7, 4.0)
    assert (n, k0, k1) == (len(bins), int(bins[0]), int(bins[-1]))


def test_bandpass_in_place_and_empty_band(vhr, eng):
    import torch
    rng = np.random.default_rng(2)
    x = rng.standard_normal((100, 33)).astype(np.float32)
    xd = torch.as_tensor(x, device=eng.tdev)
    ref = oevm.ideal_bandpass(x, 10.0, 0.7, 4.0)
    eng.bandpass(xd, 10.0, 0.7, 4.0, 1.0, out=xd)
    assert rel_err(xd.cpu().numpy(), ref) <= REL
    z = eng.bandpass(torch.as_tensor(x, device=eng.tdev), 10.0, 6.0, 7.0, 1.0).cpu().numpy()   # above Nyquist
    assert np.all(z == 0)


# --------------------------------------------------------------------------------- collapse
@pytest.mark.parametrize("hw", [(144, 256), (135, 248), (97, 131), (480, 640), (33, 700)])
@pytest.mark.parametrize("levels", [1, 3, 4])
def test_collapse_addback(vhr, eng, hw, levels):
    import torch
    H, W = hw
    rng = np.random.default_rng(H + W + levels)
    T = 2
    fr = rng.integers(0, 256, (T, H, W, 3), dtype=np.uint8)
    wl, hl = oevm.pyr_dims(W, H, levels)[-1]
    lv = (40 * rng.standard_normal((T, hl, wl, 3))).astype(np.float32)
    rects = np.array([[[W // 4, H // 3, W // 4 + max(1, W // 3), H // 3 + max(1, H // 4)],
                       [0, 0, W, H], [5, 2, 5, 9]]] * T, dtype=np.int32)
    o32, o8, means = eng.collapse(torch.as_tensor(lv, device=eng.tdev), torch.as_tensor(fr, device=eng.tdev),
                                  levels, out_f32=True, out_u8=True, rects=rects)
    ref = oevm.collapse_addback(lv, fr, levels)
    assert rel_err(o32.cpu().numpy(), ref) <= REL
    # u8 output: the rounding rule is bit-exact wherever the float value is not within
    # tolerance of a .5 boundary
    ref8 = oevm.to_u8(ref)
    got8 = o8.cpu().numpy()
    frac = np.abs((np.clip(ref, 0, 255) + 0.5) % 1.0)
    safe = (frac > 1e-3) & (frac < 1 - 1e-3)
    np.testing.assert_array_equal(got8[safe], ref8[safe])
    assert np.abs(got8.astype(int) - ref8.astype(int)).max() <= 1
    m = means.cpu().numpy()
    for t in range(T):
        for k in range(2):
            x1, y1, x2, y2 = rects[t, k]
            exp = ref[t, y1:y2, x1:x2].reshape(-1, 3).mean(0)
            assert np.abs(m[t, k] - exp).max() <= REL * np.abs(exp).max()
        assert np.isnan(m[t, 2]).all()                 # empty rectangle -> NaN like np.mean([])


def test_full_size_properties_1080p(vhr, eng):
    """BASELINE's full frame size (1920x1080, 4 levels, 0.7-4 Hz, alpha 50), checked through properties
    that need no oracle: (1) a static clip has no energy in the band, so the magnified output is the input,
    bit for bit; (2) the ideal filter is circular in time, so rolling the clip rolls the output; (3) pyramid
    level 4 of the whole cascade has the chain's size (68 x 120); (4) the fused ROI sums over the whole
    frame equal the mean of the output tensor; (5) rectangle and polygon ROI means of the same axis-aligned
    region agree."""
    import torch
    T, H, W = 60, 1080, 1920
    g = torch.Generator(device="cpu").manual_seed(3)
    one = torch.randint(0, 256, (1, H, W, 3), dtype=torch.uint8, generator=g).to(eng.tdev)
    static = one.expand(T, H, W, 3).contiguous()
    r = eng.evm(static, 30.0, 4, 0.7, 4.0, 50.0, out_f32=True)
    assert torch.equal(r["out_f32"], static.float())
    del r, static
    spec = vhr.SynthSpec(T=T, H=H, W=W, fps=30.0, pulse_hz=1.5, seed=7)
    clip = eng.synth_clip(spec)
    full = np.tile(np.array([0, 0, W, H], dtype=np.int32), (T, 1, 1))
    lvl = eng.pyrdown(clip, 4)
    assert tuple(lvl.shape) == (T, 68, 120, 3)
    a = eng.evm(clip, 30.0, 4, 0.7, 4.0, 50.0, rects=full, out_f32=True)
    out = a["out_f32"]
    got = a["roi_mean"][:, 0, :].cpu().numpy()
    exp = out.double().mean(dim=(1, 2)).cpu().numpy()
    assert np.abs(got - exp).max() <= REL * np.abs(exp).max()
    k = 17
    b = eng.evm(torch.roll(clip, k, dims=0), 30.0, 4, 0.7, 4.0, 50.0, out_f32=True)["out_f32"]
    d = (torch.roll(out, k, dims=0) - b).abs().max().item()
    assert d <= REL * out.abs().max().item()
    del b
    x0, y0, x1, y1 = 700, 400, 1200, 650
    rect = np.tile(np.array([x0, y0, x1, y1], dtype=np.int32), (T, 1, 1))
    poly = np.tile(np.array([[x0, y0], [x1 - 1, y0], [x1 - 1, y1 - 1], [x0, y1 - 1]], dtype=np.int32), (T, 1, 1, 1))
    nv = np.full((T, 1), 4, dtype=np.int32)
    mr = eng.roi_mean_rect(clip, rect).cpu().numpy()
    mp_, cnt = eng.roi_mean_poly(clip, poly, nv)
    np.testing.assert_array_equal(mr, mp_.cpu().numpy())                 # both exact integer sums / count
    assert int(cnt[0, 0]) == (x1 - x0) * (y1 - y0)


def test_evm_end_to_end_c1(vhr, eng):
    """Config c1 (256x144, 5 FPS, 30 s, 1.2 Hz): EVM + ROI + BPM against the oracle."""
    s, o = spec_pair(vhr, T=150, H=144, W=256, fps=5.0, pulse_hz=1.2, seed=0)
    fr = osynth.synth_frames(o)
    frd = eng.synth_clip(s)
    np.testing.assert_array_equal(frd.cpu().numpy(), fr)
    lm = o.landmarks()
    rect = oroi.cheek_roi_from_bbox(oroi.bbox_from_landmarks_clamped(lm[:, 0], lm[:, 1], 256, 144), 256, 144)
    rects = np.tile(np.array(rect, dtype=np.int32), (150, 1, 1))
    r = eng.evm(frd, 5.0, 4, 0.7, 4.0, 50.0, rects=rects, out_f32=True, keep_levels=True)
    lv, filt, out = oevm.evm_clip(fr, 5.0, 4, 0.7, 4.0, 50.0)
    assert rel_err(r["level"].cpu().numpy(), lv) <= REL
    assert rel_err(r["filtered"].cpu().numpy(), filt) <= REL
    assert rel_err(r["out_f32"].cpu().numpy(), out) <= REL
    x1, y1, x2, y2 = rect
    trace_ref = out[:, y1:y2, x1:x2, :].reshape(150, -1, 3).mean(1)
    trace = r["roi_mean"].cpu().numpy()[:, 0]
    assert rel_err(trace, trace_ref) <= REL
    # BPM: identical bin, identical value
    from video_heart_rate_b200.pipeline import ANALYSIS_BAND
    bpm, kbin = eng.bpm_fft(r["roi_mean"][:, 0, 1].contiguous(), [0], [150], 5.0, ANALYSIS_BAND,
                            detrend=vhr.DETREND_F32)
    g32 = trace_ref[:, 1].astype(np.float32)
    exp_bpm, exp_k, _ = obpm.estimate_bpm_analysis(g32 - np.mean(g32), 5.0)
    assert int(kbin[0]) == exp_k and float(bpm[0]) == exp_bpm == 72.0


# --------------------------------------------------------------------------------------- ROI
def test_rect_means_golden(vhr, eng, golden_dir):
    """Bit-exact against the reference's own process_frame / np.mean (golden vectors)."""
    import torch
    from video_heart_rate_b200 import host
    from video_heart_rate_b200.pipeline import video_trace, green_avg_trace
    g = np.load(os.path.join(golden_dir, "roi_rect.npz"))
    for i in range(int(g["n_px"])):
        frame = g[f"px_frame_{i}"]
        lm = np.stack([g[f"px_xs_{i}"], g[f"px_ys_{i}"]], 1)
        fr = torch.as_tensor(frame[None], device=eng.tdev)
        means, _ = green_avg_trace(eng, fr, lm[None])
        np.testing.assert_array_equal(means.cpu().numpy()[0], g[f"px_mean_clean_{i}"])
        v = video_trace(eng, fr, lm[None], overdraw=True)
        np.testing.assert_array_equal(v.cpu().numpy()[0], g[f"px_video_green_{i}"])


@pytest.mark.parametrize("shape", [(5, 7), (61, 67), (144, 256), (90, 1919)])
def test_rect_rows_kernel_equals_per_pixel(vhr, eng, shape, monkeypatch):
    """The row-wise rectangle mean (aligned words + IDP4A per channel, head/tail masks) against the
    per-pixel kernel and NumPy: every start phase, widths 1..W, rectangles touching the clip's last byte,
    empty and out-of-frame rectangles (NaN)."""
    import torch
    H, W = shape
    rng = np.random.default_rng(H + W)
    T, K = 5, 6
    fr = rng.integers(0, 256, (T, H, W, 3), dtype=np.uint8)
    rects = np.zeros((T, K, 4), dtype=np.int32)
    for t in range(T):
        for k in range(K):
            x1, x2 = np.sort(rng.integers(0, W + 1, 2)); y1, y2 = np.sort(rng.integers(0, H + 1, 2))
            rects[t, k] = (x1, y1, x2, y2)
    rects[-1, 0] = (0, 0, W, H)                       # whole frame, ends on the clip's last byte
    rects[-1, 1] = (W - 1, H - 1, W, H)               # the last pixel
    rects[0, 2] = (3, 2, 3, 4)                        # empty
    rects[0, 3] = (0, 0, W + 1, H)                    # out of frame
    frd = torch.as_tensor(fr, device=eng.tdev)
    rows = eng.roi_mean_rect(frd, rects).cpu().numpy()
    monkeypatch.setenv("VHR_RECT_PIXEL", "1")
    pix = eng.roi_mean_rect(frd, rects).cpu().numpy()
    monkeypatch.delenv("VHR_RECT_PIXEL")
    np.testing.assert_array_equal(rows, pix)
    for t in range(T):
        for k in range(K):
            x1, y1, x2, y2 = rects[t, k]
            if x2 > x1 and y2 > y1 and x2 <= W:
                np.testing.assert_array_equal(rows[t, k], fr[t, y1:y2, x1:x2].reshape(-1, 3).mean(0))
            else:
                assert np.isnan(rows[t, k]).all()


def test_poly_mask_and_means(vhr, eng):
    import torch
    rng = np.random.default_rng(5)
    H, W, T = 90, 120, 3
    polys = np.zeros((T, 4, 12, 2), dtype=np.int32)
    nv = np.zeros((T, 4), dtype=np.int32)
    for t in range(T):
        for k in range(4):
            n = int(rng.integers(3, 12))
            if k == 3:
                n = int(rng.integers(0, 3))                      # degenerate: empty / point / segment
            ang = np.sort(rng.uniform(0, 2 * np.pi, n))
            r = rng.uniform(5, 60, n)
            if k == 1:
                r[::2] *= 0.4                                     # concave star
            pts = np.stack([60 + r * np.cos(ang), 45 + r * np.sin(ang)], 1).astype(np.int32)   # partly off-frame
            polys[t, k, :n] = pts
            nv[t, k] = n
    mask = eng.poly_mask(T, H, W, polys, nv).cpu().numpy()
    fr8 = rng.integers(0, 256, (T, H, W, 3), dtype=np.uint8)
    fr32 = (rng.standard_normal((T, H, W, 3)) * 50 + 120).astype(np.float32)
    m8, c8 = eng.roi_mean_poly(torch.as_tensor(fr8, device=eng.tdev), polys, nv)
    m32, c32 = eng.roi_mean_poly(torch.as_tensor(fr32, device=eng.tdev), polys, nv)
    for t in range(T):
        for k in range(4):
            ref = oroi.poly_mask(H, W, polys[t, k, :nv[t, k]])
            np.testing.assert_array_equal(mask[t, k].astype(bool), ref)          # bit-exact mask
            assert int(c8[t, k]) == int(ref.sum()) == int(c32[t, k])
            np.testing.assert_array_equal(m8[t, k].cpu().numpy(), oroi.masked_mean(fr8[t], ref))
            e = oroi.masked_mean(fr32[t], ref)
            if ref.sum():
                assert np.abs(m32[t, k].cpu().numpy() - e).max() <= REL * np.abs(e).max()
            else:
                assert np.isnan(m32[t, k].cpu().numpy()).all()


@pytest.mark.parametrize("shape", [(64, 96), (40, 2300), (300, 520)])
def test_poly_scanline_equals_per_pixel(vhr, eng, shape, monkeypatch):
    """The scanline rasteriser (toggle bits + prefix-XOR per row) and the per-pixel form of the frozen rule
    give the same pixel sets: random simple, self-intersecting and degenerate polygons, horizontal edges,
    repeated vertices, vertices outside the frame, rows wider than 2048 pixels."""
    import torch
    H, W = shape
    rng = np.random.default_rng(H * 7 + W)
    T, K, V = 6, 5, 16
    polys = np.zeros((T, K, V, 2), dtype=np.int32)
    nv = np.zeros((T, K), dtype=np.int32)
    for t in range(T):
        for k in range(K):
            n = int(rng.integers(3, V + 1))
            if k == 0:                                            # random vertex order: self-intersecting
                pts = np.stack([rng.integers(-20, W + 20, n), rng.integers(-10, H + 10, n)], 1)
            elif k == 1:                                          # axis-aligned pieces: horizontal / vertical edges
                xs = np.sort(rng.integers(0, W, 4)); ys = np.sort(rng.integers(0, H, 4))
                pts = np.array([[xs[0], ys[0]], [xs[3], ys[0]], [xs[3], ys[2]], [xs[2], ys[2]], [xs[2], ys[3]], [xs[0], ys[3]]])
                n = 6
            elif k == 2:                                          # repeated vertices and collinear runs
                base = np.stack([rng.integers(0, W, 4), rng.integers(0, H, 4)], 1)
                pts = np.repeat(base, 2, axis=0)
                n = 8
            elif k == 3:                                          # star around the centre, partly off-frame
                ang = np.sort(rng.uniform(0, 2 * np.pi, n))
                r = rng.uniform(0.1, 0.8, n) * max(H, W) / 2
                r[::2] *= 0.4
                pts = np.stack([W / 2 + r * np.cos(ang), H / 2 + r * np.sin(ang)], 1)
            else:                                                 # empty / point / segment
                n = int(rng.integers(0, 3))
                pts = np.stack([rng.integers(0, W, n), rng.integers(0, H, n)], 1) if n else np.zeros((0, 2))
            polys[t, k, :n] = np.asarray(pts, dtype=np.int64).astype(np.int32)[:n]
            nv[t, k] = n
    fr8 = torch.as_tensor(rng.integers(0, 256, (T, H, W, 3), dtype=np.uint8), device=eng.tdev)
    m_scan, c_scan = eng.roi_mean_poly(fr8, polys, nv)
    monkeypatch.setenv("VHR_POLY_PIXEL", "1")
    m_pix, c_pix = eng.roi_mean_poly(fr8, polys, nv)
    monkeypatch.delenv("VHR_POLY_PIXEL")
    np.testing.assert_array_equal(c_scan.cpu().numpy(), c_pix.cpu().numpy())
    np.testing.assert_array_equal(m_scan.cpu().numpy(), m_pix.cpu().numpy())          # NaN == NaN for empty sets
    mask = eng.poly_mask(T, H, W, polys, nv)
    np.testing.assert_array_equal(mask.sum(dim=(2, 3)).cpu().numpy(), c_scan.cpu().numpy())
    assert int(c_scan.sum()) > 0


# --------------------------------------------------------------------------------------- BPM
def test_bpm_golden(vhr, eng, golden_dir):
    """Identical BPM (hence identical spectral-peak bin) as the reference's estimators on the
    golden traces: analysis FFT, VIDEO FFT, Butterworth/Chebyshev/FIR + Welch, LIVE Welch."""
    from video_heart_rate_b200.pipeline import ANALYSIS_BAND, LIVE_BAND, VIDEO_BAND, design_filters
    g = np.load(os.path.join(golden_dir, "bpm.npz"))
    bad = []
    for i in range(int(g["n"])):
        x, fps = g[f"x_{i}"], float(g[f"fps_{i}"])
        exp = g[f"bpm_{i}"]
        n = len(x)
        a, _ = eng.bpm_fft(x, [0], [n], fps, ANALYSIS_BAND, detrend=vhr.DETREND_F32, mode=vhr.FFT_ANALYSIS)
        v, _ = eng.bpm_fft(x, [0], [n], fps, VIDEO_BAND, detrend=vhr.DETREND_F64, mode=vhr.FFT_VIDEO)
        f = design_filters(fps, VIDEO_BAND)
        wb, _, fb = eng.bpm_welch(x, [0], [n], fps, VIDEO_BAND, vhr.DETREND_F64, vhr.FILT_SOS, f["butter"], want_filtered=True)
        wc, _, fc = eng.bpm_welch(x, [0], [n], fps, VIDEO_BAND, vhr.DETREND_F64, vhr.FILT_SOS, f["cheby2"], want_filtered=True)
        wf, _, ff = eng.bpm_welch(x, [0], [n], fps, VIDEO_BAND, vhr.DETREND_F64, vhr.FILT_FIR, f["fir"], want_filtered=True)
        wl, _, _ = eng.bpm_welch(x, [0], [n], fps, LIVE_BAND, vhr.DETREND_F64, vhr.FILT_NONE)
        got = np.array([float(q[0]) for q in (a, v, wb, wc, wf, wl)])
        if not np.array_equal(got, exp, equal_nan=True):
            bad.append((i, fps, n, got, exp))
        # filtered traces: float64 recurrences, same operation order as scipy
        assert rel_err(fb.cpu().numpy()[0], g[f"fb_{i}"]) <= 1e-9
        assert rel_err(fc.cpu().numpy()[0], g[f"fc_{i}"]) <= 1e-9
        if g[f"ff_{i}"].size:
            assert rel_err(ff.cpu().numpy()[0], g[f"ff_{i}"]) <= 1e-9
        else:
            assert np.isnan(float(wf[0]))
    assert not bad, f"{len(bad)} traces differ, first: {bad[:3]}"


def test_live_sos_golden(vhr, eng, golden_dir):
    import torch
    g = np.load(os.path.join(golden_dir, "bpm.npz"))
    for j in range(2):
        sos = g[f"live_sos_{j}"]
        x = g[f"live_x_{j}"]
        state = torch.zeros((sos.shape[0], 2), dtype=torch.float64, device=eng.tdev)
        y1 = eng.sos_causal(x[:150], sos, state).cpu().numpy()
        y2 = eng.sos_causal(x[150:], sos, state).cpu().numpy()          # carried state
        assert rel_err(np.concatenate([y1, y2]), g[f"live_y_{j}"]) <= 1e-12


def test_green_avg_measure_matches_oracle(vhr, eng):
    """analysis green_avg.measure() body on a synthetic clip: identical (M,2) array."""
    from video_heart_rate_b200.pipeline import green_avg_measure
    for fps, T in ((5.0, 150), (30.0, 420)):
        s, o = spec_pair(vhr, T=T, H=72, W=128, fps=fps, pulse_hz=1.3, seed=4)
        fr = osynth.synth_frames(o)
        lm = o.landmarks()
        got = green_avg_measure(eng, fr, fps, lm)
        rect = oroi.cheek_roi_from_bbox(oroi.bbox_from_landmarks_clamped(lm[:, 0], lm[:, 1], 128, 72), 128, 72)
        green = [oroi.rect_mean(f, rect)[1] for f in fr]
        exp, _ = obpm.green_avg_series(green, fps)
        np.testing.assert_array_equal(got, exp)


def test_measurement_plugins_through_a_video_file(vhr, eng, tmp_path, monkeypatch):
    """The reference-facing boundary itself: `analysis/main.py:29-31` does
    importlib.import_module(f"measurement.{name}").measure(video_path).  A synthetic clip is written
    losslessly (FFV1) with a landmark side-car; the plugins are imported from a harness-like directory
    the same way; green_avg_b200 must return the reference loop's (N,2) array (oracle: green_avg.py
    restated, pinned by the golden vectors) on the DECODED frames, evm_b200 the injected pulse."""
    import importlib
    import shutil
    import cv2
    fps, T, H, W = 30.0, 330, 144, 256
    s, o = spec_pair(vhr, T=T, H=H, W=W, fps=fps, pulse_hz=1.4, seed=11)
    frames = osynth.synth_frames(o)
    path = str(tmp_path / "subject.mkv")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"FFV1"), fps, (W, H))
    if not wr.isOpened():
        pytest.skip("cv2 build has no FFV1 writer")
    for f in frames:
        wr.write(f)
    wr.release()
    lm = o.landmarks()
    np.save(str(tmp_path / "subject.landmarks.npy"), np.tile(lm[None], (T, 1, 1)))
    harness = tmp_path / "analysis" / "measurement"
    harness.mkdir(parents=True)
    src = os.path.join(ROOT, "video-heart-rate_b200", "analysis", "measurement")
    for name in ("green_avg_b200.py", "evm_b200.py", "green_avg_psd_b200.py"):
        shutil.copy(os.path.join(src, name), str(harness / name))
    (harness / "__init__.py").write_text("")
    monkeypatch.syspath_prepend(str(tmp_path / "analysis"))
    monkeypatch.chdir(str(tmp_path / "analysis"))
    for m in [k for k in sys.modules if k == "measurement" or k.startswith("measurement.")]:
        monkeypatch.delitem(sys.modules, m)
    ga = importlib.import_module("measurement.green_avg_b200")
    got = ga.measure(path)
    cap = cv2.VideoCapture(path)
    dec = []
    while True:
        ok, f = cap.read()
        if not ok:
            break
        dec.append(f)
    cap.release()
    assert len(dec) == T and np.array_equal(np.stack(dec), frames)          # FFV1 is lossless
    rect = oroi.cheek_roi_from_bbox(oroi.bbox_from_landmarks_clamped(lm[:, 0], lm[:, 1], W, H), W, H)
    green = [oroi.rect_mean(f, rect)[1] for f in dec]
    exp, _ = obpm.green_avg_series(green, fps)
    assert got.dtype == np.float64 and got.shape == exp.shape and got.shape[1] == 2
    np.testing.assert_array_equal(got, exp)
    ev = importlib.import_module("measurement.evm_b200").measure(path)
    assert ev.shape[1] == 2 and len(ev) == T - int(10 * fps) + 1
    assert np.all(np.abs(ev[:, 1] - 84.0) <= 6.0)                           # 1.4 Hz pulse, 6 BPM bins at 10 s
    with pytest.raises(FileNotFoundError):
        ga.measure(str(tmp_path / "missing.mkv"))                            # video_io.py:10-11 behaviour


def test_video_bpm_series_matches_oracle(vhr, eng):
    """rppg_VIDEO.py sliding-window block over a trace: identical BPM triplets."""
    from video_heart_rate_b200.pipeline import video_bpm_series
    rng = np.random.default_rng(8)
    for fps, n in ((30.0, 420), (5.0, 90)):
        t = np.arange(n) / fps
        g = 120 + np.sin(2 * np.pi * 1.25 * t) + 0.2 * rng.standard_normal(n)
        got = video_bpm_series(eng, g, fps)
        exp = obpm.video_window_bpm(g, fps)
        assert len(exp) == len(got["frame"])
        for j, (i, b1, b2, b3, _) in enumerate(exp):
            assert got["frame"][j] == i
            assert got["butter"][j] == b1 and got["cheby2"][j] == b2
            assert (np.isnan(got["fir"][j]) and b3 is None) or got["fir"][j] == b3


def test_dropin_functions(vhr, eng):
    """Reference-named functions (rppg.py) behave like the reference's on one window."""
    from video_heart_rate_b200 import rppg
    rng = np.random.default_rng(3)
    fps = 30.0
    t = np.arange(300) / fps
    x = np.sin(2 * np.pi * 1.2 * t) + 0.1 * rng.standard_normal(300)
    for ours, ref in ((rppg.bandpass_butterworth(x, fps, 0.7, 2, 2), obpm.bandpass_butterworth(x, fps, 0.7, 2, 2)),
                      (rppg.bandpass_cheby2(x, fps, 0.7, 2), obpm.bandpass_cheby2(x, fps, 0.7, 2)),
                      (rppg.bandpass_fir(x, fps, 0.7, 2), obpm.bandpass_fir(x, fps, 0.7, 2))):
        assert rel_err(ours, ref) <= 1e-9
    assert rppg.estimate_bpm_welch(x, fps) == obpm.estimate_bpm_welch(x, fps)[0]
    assert rppg.estimate_bpm(x, fps) == obpm.estimate_bpm_video_fft(x, fps)[0]
    with pytest.raises(ValueError):
        rppg.bandpass_fir(x[:50], 5.0, 0.7, 2)              # reference raises at 5 FPS (padlen 123)
    roi = rng.integers(0, 256, (25, 89, 3), dtype=np.uint8)
    assert rppg.get_avg(roi, 1) == float(np.mean(roi[:, :, 1]))
    # what np.mean accepts beyond the scripts' uint8 ROIs: other dtypes / channel counts (float32 kernel, 1e-6)
    rf = rng.standard_normal((17, 33, 4)) * 40 + 100
    assert abs(rppg.get_avg(rf, 3) - float(np.mean(rf[:, :, 3]))) <= 1e-5 * 100
    assert abs(rppg.get_avg(roi.astype(np.int16), 2) - float(np.mean(roi[:, :, 2]))) <= 1e-6 * 255
    # (T, N) signals are filtered column by column, like sosfiltfilt / filtfilt along axis 0
    X = np.stack([x, np.roll(x, 7) * 0.5, x[::-1]], 1)
    for ours, ref in ((rppg.bandpass_butterworth(X, fps, 0.7, 2, 2), obpm.bandpass_butterworth(X, fps, 0.7, 2, 2)),
                      (rppg.bandpass_fir(X, fps, 0.7, 2), obpm.bandpass_fir(X, fps, 0.7, 2))):
        assert ours.shape == ref.shape == (300, 3) and rel_err(ours, ref) <= 1e-9


def test_evm_roi_host_matches_device_path(vhr, eng):
    s, o = spec_pair(vhr, T=60, H=72, W=128, fps=10.0, pulse_hz=1.5, seed=9)
    fr = osynth.synth_frames(o)
    rects = np.tile(np.array([[40, 30, 90, 50]], dtype=np.int32), (60, 1, 1))
    means = eng.evm_roi_host(fr, 10.0, rects, levels=3)
    import torch
    r = eng.evm(torch.as_tensor(fr, device=eng.tdev), 10.0, 3, 0.7, 4.0, 50.0, rects=rects)
    np.testing.assert_array_equal(means, r["roi_mean"].cpu().numpy())       # same kernels, deterministic


# ------------------------------------------------------------------- degradations / metric
def test_degradations_and_mae(vhr, eng, golden_dir):
    import torch
    from oracle import degrade as odeg
    g = np.load(os.path.join(golden_dir, "metrics.npz"))
    fr = torch.as_tensor(g["q_frame"][None], device=eng.tdev)
    for bits in (9, 8, 7, 6, 5, 4):
        np.testing.assert_array_equal(eng.degrade_quantise(fr, bits).cpu().numpy()[0], g[f"q_bits_{bits}"])
    rng = np.random.default_rng(12)
    frames = rng.integers(0, 256, (5, 33, 47, 3), dtype=np.uint8)
    fd = torch.as_tensor(frames, device=eng.tdev)
    for sigma in (5, 10, 20, 40):
        got = eng.degrade_noise(fd, sigma, seed=3, clip=1, t0=2).cpu().numpy()
        np.testing.assert_array_equal(got, odeg.add_noise(frames, sigma, seed=3, clip=1, t0=2))
    for j in range(int(g["n_m"])):
        aligned, mae = eng.align_mae(g[f"m_tt_{j}"], g[f"m_th_{j}"], g[f"m_meas_{j}"])
        np.testing.assert_array_equal(aligned.cpu().numpy(), g[f"m_aligned_{j}"][:, 1])
        assert mae == float(g[f"m_mae_{j}"])


def test_green_avg_psd_variant(vhr, eng):
    """green_avg_psd_plot.py variant (z-score + Butterworth sosfiltfilt + periodogram): identical BPM
    per frame against the oracle restatement of :173-183 / :34-63."""
    from video_heart_rate_b200.pipeline import green_avg_psd_series
    rng = np.random.default_rng(21)
    for fps, n in ((30.0, 400), (5.0, 80), (29.97, 330)):
        t = np.arange(n) / fps
        g = 140 + 0.8 * np.sin(2 * np.pi * 1.45 * t) + 0.3 * rng.standard_normal(n) + 0.05 * t
        got = green_avg_psd_series(eng, g, fps)
        wl = int(round(10.0 * fps))
        exp = np.full(n, np.nan)
        for i in range(n):
            w = g[max(0, i + 1 - wl): i + 1]
            if len(w) >= wl:
                exp[i] = obpm.psd_plot_estimate(w, fps)[0]
        np.testing.assert_array_equal(got[:, 1], exp)
        np.testing.assert_array_equal(got[:, 0], np.arange(n) * (1 / fps))


# ------------------------------------------------------------------- BASELINE configs (reduced)
def test_config_runners_reduced(vhr, eng):
    """tools/run_configs.py: c1 in full (CPU-path config: bit-exact trace, identical BPM triplets,
    FIR raises like the reference) and a 20-window slice of the c5 sweep (MAE within one bin)."""
    from tools import run_configs as rc
    r1 = rc.run_c1(eng, vhr)
    assert r1["trace_bit_exact"] and r1["bpm_identical"] and r1["windows"] == 100
    assert r1["bpm_butter_last"] == pytest.approx(73.3333, abs=1e-3) and r1["evm_bpm"] == 72.0
    r5 = rc.run_c5(eng, vhr, n=20)
    assert r5["windows"] == 20 and r5["nan_windows"] == 0 and r5["mae_bpm"] <= 3.0


def test_bench_gpu_arm_prints_the_contract_line():
    """`bench.py` (GPU arm) on the small c1 workload: the JSON line carries the contract's keys
    (roofline, cpu_baseline, e2e with transfer bytes, gpu_launches, clocks) and a correct BPM."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "c1", "--steps", "4", "--warmup", "3",
                          "--e2e-steps", "2"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["steps"] == 4 and line["value"] > 0 and line["bpm_ok"] is True
    assert line["gpu_launches"] == 5 * 4                                   # pyrdown, bandpass, collapse, ROI finalize, BPM
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    e = line["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] >= 150 * 144 * 256 * 3 and e["d2h_bytes_per_step"] > 0
    c = line["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0
