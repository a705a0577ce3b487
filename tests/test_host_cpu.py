"""CPU tests (no GPU): host-side logic of the product package against the oracle / golden
vectors, the C-ABI library's exported symbols against include/vhr_b200.h, and the
world_size-2 gloo path of the clip sharding."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import bpm as obpm
from oracle import roi as oroi
from oracle import synth as osynth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pkg():
    import video_heart_rate_b200 as p
    return p


def test_import_has_no_oracle_dependency(pkg):
    """The product never imports the oracle (it is test infrastructure only)."""
    code = ("import sys, video_heart_rate_b200, video_heart_rate_b200.pipeline, video_heart_rate_b200.rppg, "
            "video_heart_rate_b200.parallel; print(any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules))")
    out = subprocess.check_output([sys.executable, "-c", code], cwd=ROOT, text=True).strip()
    assert out == "False"
    for dirpath, _, files in os.walk(os.path.join(ROOT, "video-heart-rate_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_engine_fails_loudly_without_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.VhrError):
        pkg.Engine(0)


def test_geometry_matches_reference_golden(pkg, golden_dir):
    g = np.load(os.path.join(golden_dir, "roi_rect.npz"))
    host = pkg.host
    for i in range(g["geo_hw"].shape[0]):
        h, w = (int(v) for v in g["geo_hw"][i])
        lm = np.stack([g["geo_xs"][i], g["geo_ys"][i]], 1)[None]
        bbc = host.bbox_clamped(lm, w, h)
        np.testing.assert_array_equal(bbc[0], g["geo_bb_clamped"][i])
        np.testing.assert_array_equal(host.cheek_roi_clamped(bbc, w, h)[0], g["geo_cheek_clamped"][i])
        bbv = host.bbox_video(lm, w, h)
        np.testing.assert_array_equal(bbv[0], g["geo_bb_video"][i])
        np.testing.assert_array_equal(host.roi_coords(bbv, *host.FOREHEAD)[0], g["geo_forehead_video"][i])
        ck = host.roi_coords(bbv, *host.CHEEK)
        np.testing.assert_array_equal(ck[0], g["geo_cheek_video"][i])
        s = host.slice_rects(ck, w, h)[0]
        ya, yb = oroi.py_slice(int(ck[0, 1]), int(ck[0, 3]), h)
        xa, xb = oroi.py_slice(int(ck[0, 0]), int(ck[0, 2]), w)
        assert tuple(s) == (xa, ya, xb, yb)


def test_slice_rects_is_numpy_slicing(pkg):
    rng = np.random.default_rng(0)
    a = np.arange(37)
    for _ in range(500):
        lo, hi = (int(v) for v in rng.integers(-60, 60, 2))
        s = pkg.host.slice_rects([[lo, 0, hi, 1]], 37, 5)[0]
        assert len(a[lo:hi]) == s[2] - s[0]
        if s[2] > s[0]:
            assert a[lo:hi][0] == s[0]


def test_window_lists_match_reference_loops(pkg):
    host = pkg.host
    for fps in (5.0, 29.97, 30.0):
        n = int(47 * fps)
        g = np.random.default_rng(1).standard_normal(n)
        fi, st, ln = host.green_avg_windows(n, fps)
        from collections import deque
        dq = deque(maxlen=int(30.0 * fps))
        exp = []
        for i in range(n):
            dq.append(i)
            if len(dq) < int(10.0 * fps):
                continue
            exp.append((i, dq[0], len(dq)))
        assert [(int(a), int(b), int(c)) for a, b, c in zip(fi, st, ln)] == exp
        fi, st, ln = host.video_windows(n, fps)
        dq = deque(maxlen=1000)
        exp = []
        wl = int(fps * 10)
        for i in range(n):
            dq.append(i)
            if len(dq) > wl:
                w = list(dq)[-wl:]
                exp.append((i, w[0], len(w)))
        assert [(int(a), int(b), int(c)) for a, b, c in zip(fi, st, ln)] == exp


def test_synth_spec_matches_oracle(pkg):
    kw = dict(T=150, H=144, W=256, fps=5.0, pulse_hz=1.2, seed=7, clip=3, noise_sigma=5.0)
    s, o = pkg.SynthSpec(**kw), osynth.SynthParams(**kw)
    np.testing.assert_array_equal(s.pulse_table(), o.pulse_table())
    np.testing.assert_array_equal(s.base_q8(), o.base_q8())
    assert s.face_rect() == o.face_rect() and s.noise_gain() == o.noise_gain()
    np.testing.assert_array_equal(s.landmarks(), o.landmarks())
    polys, nv = s.polygons()
    for k, p in enumerate(oroi.face_polygons(*o.face_rect())):
        np.testing.assert_array_equal(polys[k, :nv[k]], p)


def test_hold_landmarks_policy(pkg):
    lm = np.arange(40 * 2 * 2, dtype=float).reshape(40, 2, 2)
    valid = np.ones(40, bool)
    valid[5:25] = False
    held, usable = pkg.host.hold_landmarks(lm, valid)
    assert usable[:5].all() and usable[5:20].all() and not usable[20:25].any() and usable[25:].all()
    np.testing.assert_array_equal(held[19], lm[4])
    np.testing.assert_array_equal(held[30], lm[30])


def test_library_exports_every_declared_symbol(pkg):
    """The built .so loads and exports exactly the entry points include/vhr_b200.h declares
    (no compute call is made: there is no GPU here)."""
    if not os.path.exists(pkg.LIB_PATH):
        pytest.skip("libvhr_b200.so not built (run python __graft_entry__.py)")
    hdr = open(os.path.join(ROOT, "include", "vhr_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t|const char\*)\s+(vhr_\w+)\s*\(", hdr, re.M))
    assert len(declared) >= 18
    lib = ctypes.CDLL(pkg.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    from video_heart_rate_b200 import _lib
    assert set(_lib.SIGNATURES) == declared
    assert lib.vhr_abi_version() == 2
    w = (ctypes.c_int32 * 5)()
    h = (ctypes.c_int32 * 5)()
    assert lib.vhr_pyr_dims(1920, 1080, 4, w, h) == 0
    assert (list(w), list(h)) == ([1920, 960, 480, 240, 120], [1080, 540, 270, 135, 68])
    lib.vhr_band_bins.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                  ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    k0, k1 = ctypes.c_int(), ctypes.c_int()
    for T, fps in ((150, 5.0), (1800, 30.0), (300, 30.0), (299, 29.97), (64, 10.0)):
        n = lib.vhr_band_bins(T, fps, 0.7, 4.0, ctypes.byref(k0), ctypes.byref(k1))
        from oracle import evm as oevm
        b = oevm.band_bins(T, fps, 0.7, 4.0)
        assert (n, k0.value, k1.value) == (len(b), int(b[0]), int(b[-1]))


def test_shard_clips_gloo_world2():
    """N>1 path: clip round-robin + final gather over gloo, world_size 2, on CPU."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29671")
    procs = [subprocess.Popen([sys.executable, script, str(r), "2"], env=env, cwd=ROOT, stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "OK rank 0" in outs[0] and "OK rank 1" in outs[1]


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the
    contract's keys, the same metric / unit as the GPU arm, kind "port", zero transfer bytes."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--workload", "c1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("1080p30 frames/sec")
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["value"] > 0 and line["vs_baseline"] is None


@pytest.mark.parametrize("hw", [(1080, 1920), (720, 1280), (480, 640), (67, 1280), (1081, 160), (2160, 320), (90, 720), (64, 160)])
def test_tensor_core_pyrdown_plan_matches_the_emulation(pkg, hw):
    """csrc/pyrdown_umma.cu bakes cv2's reflect-101 borders into weight slices on the host.  The plan it builds for a
    frame shape (vhr_pyrdown_umma_plan: tiles, per-k-step slices read back through the same descriptor offsets the MMA
    uses, border weights of the horizontal pass) equals the one of tools/probes/umma_emulate.py, and that emulation of
    the kernel's data flow (tiles, lagged strips, register carries, border patches) reproduces the oracle exactly."""
    if not os.path.exists(pkg.LIB_PATH):
        pytest.skip("libvhr_b200.so not built (run python __graft_entry__.py)")
    sys.path.insert(0, os.path.join(ROOT, "tools", "probes"))
    import umma_emulate as em
    from oracle import evm as oevm
    H, W = hw
    lib = ctypes.CDLL(pkg.LIB_PATH)
    tiles = np.zeros((12, 8), np.int32)
    codes = np.zeros((12, 18), np.int32)
    wsp = np.zeros((3, 13), np.int32)
    meta = np.zeros(4, np.int32)
    blob = np.zeros(8192 + 8 * 4096, np.uint8)
    vp = ctypes.c_void_p
    lib.vhr_pyrdown_umma_plan.argtypes = [ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp, ctypes.c_int]
    assert lib.vhr_pyrdown_umma_plan(H, W, tiles.ctypes.data, codes.ctypes.data, wsp.ctypes.data, meta.ctypes.data,
                                     blob.ctypes.data, blob.size) == 0
    plan = em.make_plan(H, W)
    assert meta[0] == len(plan["tiles"]) and meta[1] == plan["nstrips"]

    def canon(m, k):                      # UMMA K-major core-matrix layout, no swizzle
        return (m >> 3) * 256 + (k >> 4) * 128 + (m & 7) * 16 + (k & 15)

    mm, kk = np.meshgrid(np.arange(128), np.arange(32), indexing="ij")
    for t, tl in enumerate(plan["tiles"]):
        assert list(tiles[t]) == [tl[k] for k in ("a", "n4", "g0", "n3", "r0", "nr", "i0", "nks")]
        assert tl["nks"] % 2 == 0
        for ks in range(tl["nks"]):
            sl = blob[codes[t, ks] * 16 + canon(mm, kk)]
            np.testing.assert_array_equal(sl[:tl["nr"]], tl["slices"][ks][:tl["nr"]])
    w2 = plan["w"][2]
    for k, x in enumerate((0, 1, w2 - 1)):
        np.testing.assert_array_equal(wsp[k], plan["special"][x])
    if H * W <= 720 * 1280:
        fr = np.random.default_rng(H + W).integers(0, 256, (H, W, 3), dtype=np.uint8)
        np.testing.assert_array_equal(em.emulate(fr, plan), oevm.pyrdown_cascade(fr[None], 4)[0])
    assert lib.vhr_pyrdown_umma_plan(1080, 1000, tiles.ctypes.data, codes.ctypes.data, wsp.ctypes.data, meta.ctypes.data,
                                     None, 0) == -4           # W % 80 != 0: not eligible
