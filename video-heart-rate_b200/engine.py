"""Engine: thin Python host over the C ABI (include/vhr_b200.h).

PyTorch is used for device memory, streams and ``torch.distributed`` only; every number is
produced by the hand-written sm_100a kernels in ``csrc/``.  There is no CPU fallback: an
Engine cannot be constructed without the built library and a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .synth import SynthSpec

DETREND_NONE, DETREND_F64, DETREND_F32, DETREND_ZSCORE_F32 = 0, 1, 2, 3
FFT_ANALYSIS, FFT_VIDEO = 0, 1
FILT_NONE, FILT_SOS, FILT_FIR = 0, 1, 2


def _torch():
    import torch
    return torch


class Engine:
    """One context per (device, host thread)."""

    def __init__(self, device: int = 0):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.VhrError("no CUDA device: the rPPG path has no CPU fallback")
        self.lib = _lib.load()
        self.device = int(device)
        self.tdev = torch.device("cuda", self.device)
        torch.cuda.set_device(self.device)
        ctx = C.c_void_p()
        rc = self.lib.vhr_create(C.byref(ctx), self.device)
        _lib.check(self.lib, None, rc, "vhr_create")
        self.ctx = ctx

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "ctx", None):
            self.lib.vhr_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.tdev).cuda_stream)

    def _check(self, rc, what):
        _lib.check(self.lib, self.ctx, rc, what)

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

    def launch_count(self) -> int:
        return int(self.lib.vhr_launch_count(self.ctx))

    def _dev(self, a, dtype):
        """numpy / tensor -> contiguous device tensor of dtype."""
        torch = _torch()
        if isinstance(a, torch.Tensor):
            return a.to(device=self.tdev, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=self.tdev)

    # ------------------------------------------------------------------ synthetic clips
    def synth_clip(self, spec: SynthSpec, out=None, t0: int = 0, t1: Optional[int] = None):
        """uint8 (t1-t0,H,W,3) frames of ``spec`` generated on the device; bit-identical to
        the NumPy generator of the oracle."""
        torch = _torch()
        t1 = spec.T if t1 is None else t1
        n = t1 - t0
        if out is None:
            out = torch.empty((n, spec.H, spec.W, 3), dtype=torch.uint8, device=self.tdev)
        assert out.is_contiguous() and out.dtype == torch.uint8 and out.numel() == n * spec.H * spec.W * 3
        pulse = self._dev(spec.pulse_table(), torch.int32)
        p = _lib.SynthParams()
        p.seed, p.clip = spec.seed & 0xFFFFFFFF, spec.clip & 0xFFFFFFFF
        p.T, p.H, p.W, p.t0 = n, spec.H, spec.W, t0
        for i, v in enumerate(spec.face_rect()):
            p.face[i] = v
        bq = spec.base_q8()
        for i in range(2):
            for c in range(3):
                p.base_q8[i][c] = int(bq[i][c])
        p.noise_gain = spec.noise_gain()
        self._check(self.lib.vhr_synth_clip(self.ctx, C.byref(p), self._p(pulse), self._p(out), self._stream()),
                    "vhr_synth_clip")
        return out

    # ------------------------------------------------------------------ EVM stages
    def pyr_dims(self, W: int, H: int, levels: int):
        w = (C.c_int32 * (levels + 1))()
        h = (C.c_int32 * (levels + 1))()
        rc = self.lib.vhr_pyr_dims(W, H, levels, w, h)
        self._check(rc, "vhr_pyr_dims")
        return [(int(w[i]), int(h[i])) for i in range(levels + 1)]

    def pyrdown(self, frames, levels: int, out=None):
        torch = _torch()
        assert frames.dtype == torch.uint8 and frames.is_cuda and frames.is_contiguous() and frames.shape[-1] == 3
        T, H, W, _ = frames.shape
        wl, hl = self.pyr_dims(W, H, levels)[-1]
        if out is None:
            out = torch.empty((T, hl, wl, 3), dtype=torch.float32, device=self.tdev)
        self._check(self.lib.vhr_pyrdown_cascade(self.ctx, self._p(frames), T, H, W, levels, self._p(out),
                                                 self._stream()), "vhr_pyrdown_cascade")
        return out

    def pyrdown_tc_accumulators(self, frames, item: int, strip: int):
        """Diagnostics: the 4-level cascade through the tensor-core kernel plus the raw TMEM accumulators (128, 240) int32 of
        one (item, strip) -- include/vhr_b200.h:vhr_pyrdown_umma_accumulators.  Raises on shapes the kernel does not take."""
        torch = _torch()
        assert frames.dtype == torch.uint8 and frames.is_cuda and frames.is_contiguous() and frames.shape[-1] == 3
        T, H, W, _ = frames.shape
        wl, hl = self.pyr_dims(W, H, 4)[-1]
        out = torch.empty((T, hl, wl, 3), dtype=torch.float32, device=self.tdev)
        acc = torch.zeros((128, 240), dtype=torch.int32, device=self.tdev)
        self._check(self.lib.vhr_pyrdown_umma_accumulators(self.ctx, self._p(frames), T, H, W, self._p(out), int(item), int(strip),
                                                           self._p(acc), self._stream()), "vhr_pyrdown_umma_accumulators")
        return out, acc

    def band_bins(self, T: int, fps: float, f_lo: float, f_hi: float):
        k0, k1 = C.c_int(-1), C.c_int(-1)
        n = self.lib.vhr_band_bins(T, fps, f_lo, f_hi, C.byref(k0), C.byref(k1))
        return n, k0.value, k1.value

    def bandpass(self, level, fps: float, f_lo: float, f_hi: float, gain: float = 1.0, out=None):
        torch = _torch()
        assert level.dtype == torch.float32 and level.is_cuda and level.is_contiguous()
        T = level.shape[0]
        P = level.numel() // T
        if out is None:
            out = torch.empty_like(level)
        self._check(self.lib.vhr_temporal_bandpass(self.ctx, self._p(level), self._p(out), T, P, float(fps),
                                                   float(f_lo), float(f_hi), float(gain), self._stream()),
                    "vhr_temporal_bandpass")
        return out

    def collapse(self, level, frames, levels: int, out_f32=True, out_u8=False, rects=None, polys=None, nverts=None,
                 want_counts: bool = False):
        """-> (out_f32 | None, out_u8 | None, roi_mean (T,K,3) float64 | None[, counts (T,K) int64]).
        ``out_f32`` / ``out_u8`` may be True (allocate), False (skip) or a preallocated tensor.
        ROIs: ``rects`` (T,K,4) rectangles, or ``polys`` (T,K,Vmax,2) + ``nverts`` (T,K) landmark
        polygons (rasterised by the frozen exact-integer rule).  With both outputs False only the
        image parts under a ROI are evaluated (ROI-only mode)."""
        torch = _torch()
        T, H, W, _ = frames.shape
        o32 = torch.empty((T, H, W, 3), dtype=torch.float32, device=self.tdev) if out_f32 is True else (
            None if out_f32 is False else out_f32)
        o8 = torch.empty((T, H, W, 3), dtype=torch.uint8, device=self.tdev) if out_u8 is True else (
            None if out_u8 is False else out_u8)
        if polys is not None:
            assert rects is None, "give rectangles or polygons, not both"
            p, nv, K, V = self._poly_args(polys, nverts, T)
            means = torch.empty((T, K, 3), dtype=torch.float64, device=self.tdev)
            counts = torch.empty((T, K), dtype=torch.int64, device=self.tdev) if want_counts else None
            self._check(self.lib.vhr_collapse_addback_poly(self.ctx, self._p(level), self._p(frames), T, H, W, levels,
                                                           self._p(o32), self._p(o8), self._p(p), self._p(nv), K, V,
                                                           self._p(means), self._p(counts), self._stream()),
                        "vhr_collapse_addback_poly")
            return (o32, o8, means, counts) if want_counts else (o32, o8, means)
        K, r_dev, means = 0, None, None
        if rects is not None:
            r_dev = self._dev(rects, torch.int32).reshape(T, -1, 4)
            K = r_dev.shape[1]
            means = torch.empty((T, K, 3), dtype=torch.float64, device=self.tdev)
        self._check(self.lib.vhr_collapse_addback_roi(self.ctx, self._p(level), self._p(frames), T, H, W, levels,
                                                      self._p(o32), self._p(o8), self._p(r_dev), K, self._p(means),
                                                      self._stream()), "vhr_collapse_addback_roi")
        return (o32, o8, means, None) if want_counts else (o32, o8, means)

    def evm(self, frames, fps: float, levels: int = 4, f_lo: float = 0.7, f_hi: float = 4.0, alpha: float = 50.0,
            rects=None, out_f32=True, out_u8=False, keep_levels: bool = False, polys=None, nverts=None):
        """Whole EVM path on device-resident frames: pyrDown cascade -> ideal bandpass (x alpha)
        -> collapse + add-back (+ fused rectangle / polygon ROI means)."""
        lvl = self.pyrdown(frames, levels)
        filt = self.bandpass(lvl, fps, f_lo, f_hi, alpha, out=None if keep_levels else lvl)
        o32, o8, means, counts = self.collapse(filt, frames, levels, out_f32=out_f32, out_u8=out_u8, rects=rects,
                                               polys=polys, nverts=nverts, want_counts=True)
        return {"level": lvl if keep_levels else None, "filtered": filt, "out_f32": o32, "out_u8": o8,
                "roi_mean": means, "roi_count": counts}

    def evm_roi_host(self, frames_np: np.ndarray, fps: float, rects_np: Optional[np.ndarray] = None, levels: int = 4,
                     f_lo: float = 0.7, f_hi: float = 4.0, alpha: float = 50.0, out: Optional[np.ndarray] = None,
                     polys_np: Optional[np.ndarray] = None, nverts_np: Optional[np.ndarray] = None,
                     want_counts: bool = False):
        """Host-buffer call (NumPy in, NumPy out): (T,H,W,3) uint8 -> (T,K,3) float64 ROI means
        (and (T,K) int64 mask pixel counts with ``want_counts`` for polygons).  H2D and D2H are
        inside the call (vhr_evm_roi_host / vhr_evm_poly_host).  ``frames_np`` should be
        page-locked (e.g. ``torch.empty(..., pin_memory=True).numpy()``); pageable memory works
        but its copies do not overlap the kernels.  Without ``out`` the magnified frames are
        never materialised (ROI-only collapse)."""
        assert frames_np.dtype == np.uint8 and frames_np.flags.c_contiguous and frames_np.ndim == 4
        T, H, W, _ = frames_np.shape
        optr = C.c_void_p(0)
        if out is not None:
            assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == frames_np.shape
            optr = C.c_void_p(out.ctypes.data)
        if polys_np is not None:
            polys_np = np.ascontiguousarray(polys_np, dtype=np.int32)
            assert polys_np.ndim == 4 and polys_np.shape[0] == T and polys_np.shape[3] == 2
            K, V = polys_np.shape[1], polys_np.shape[2]
            nverts_np = np.ascontiguousarray(nverts_np, dtype=np.int32).reshape(T, K)
            means = np.empty((T, K, 3), dtype=np.float64)
            counts = np.empty((T, K), dtype=np.int64) if want_counts else None
            self._check(self.lib.vhr_evm_poly_host(self.ctx, C.c_void_p(frames_np.ctypes.data), T, H, W, levels,
                                                   float(fps), float(f_lo), float(f_hi), float(alpha),
                                                   C.c_void_p(polys_np.ctypes.data), C.c_void_p(nverts_np.ctypes.data), K, V,
                                                   C.c_void_p(means.ctypes.data),
                                                   C.c_void_p(counts.ctypes.data if want_counts else 0), optr),
                        "vhr_evm_poly_host")
            return (means, counts) if want_counts else means
        rects_np = np.ascontiguousarray(rects_np, dtype=np.int32).reshape(T, -1, 4)
        K = rects_np.shape[1]
        means = np.empty((T, K, 3), dtype=np.float64)
        self._check(self.lib.vhr_evm_roi_host(self.ctx, C.c_void_p(frames_np.ctypes.data), T, H, W, levels,
                                              float(fps), float(f_lo), float(f_hi), float(alpha),
                                              C.c_void_p(rects_np.ctypes.data), K, C.c_void_p(means.ctypes.data),
                                              optr), "vhr_evm_roi_host")
        return (means, None) if want_counts else means

    def trim(self):
        """Release the context's cached device buffers (host-path arena, scratch arena)."""
        self._check(self.lib.vhr_trim(self.ctx), "vhr_trim")

    # ------------------------------------------------------------------ ROI
    def roi_mean_rect(self, frames, rects, paint=None, paint_rgb=None):
        """uint8 frames (T,H,W,3) + rects (T,K,4) -> float64 (T,K,3) means (tensor)."""
        torch = _torch()
        assert frames.dtype == torch.uint8 and frames.is_cuda and frames.is_contiguous()
        T, H, W, _ = frames.shape
        r = self._dev(rects, torch.int32).reshape(T, -1, 4)
        K = r.shape[1]
        means = torch.empty((T, K, 3), dtype=torch.float64, device=self.tdev)
        NP, p_dev, rgb = 0, None, None
        if paint is not None:
            p_dev = self._dev(paint, torch.int32).reshape(T, -1, 4)
            NP = p_dev.shape[1]
            rgb_np = np.ascontiguousarray(paint_rgb, dtype=np.uint8).reshape(NP, 3)
            rgb = C.c_void_p(rgb_np.ctypes.data)
        self._check(self.lib.vhr_roi_mean_rect_u8(self.ctx, self._p(frames), T, H, W, self._p(r), K, self._p(p_dev),
                                                  NP, rgb if rgb is not None else C.c_void_p(0), self._p(means),
                                                  self._stream()), "vhr_roi_mean_rect_u8")
        return means

    def _poly_args(self, polys, nverts, T):
        torch = _torch()
        p = self._dev(polys, torch.int32)
        assert p.ndim == 4 and p.shape[0] == T and p.shape[3] == 2, "polys must be (T,K,Vmax,2)"
        nv = self._dev(nverts, torch.int32).reshape(T, p.shape[1])
        return p, nv, p.shape[1], p.shape[2]

    def roi_mean_poly(self, frames, polys, nverts):
        """frames uint8 or float32 (T,H,W,3); polys int32 (T,K,Vmax,2); nverts (T,K)
        -> (means float64 (T,K,3), counts int64 (T,K))."""
        torch = _torch()
        assert frames.is_cuda and frames.is_contiguous()
        T, H, W, _ = frames.shape
        p, nv, K, V = self._poly_args(polys, nverts, T)
        means = torch.empty((T, K, 3), dtype=torch.float64, device=self.tdev)
        counts = torch.empty((T, K), dtype=torch.int64, device=self.tdev)
        fn = {torch.uint8: self.lib.vhr_roi_mean_poly_u8, torch.float32: self.lib.vhr_roi_mean_poly_f32}[frames.dtype]
        self._check(fn(self.ctx, self._p(frames), T, H, W, self._p(p), self._p(nv), K, V, self._p(means),
                       self._p(counts), self._stream()), "vhr_roi_mean_poly")
        return means, counts

    def poly_mask(self, T: int, H: int, W: int, polys, nverts):
        torch = _torch()
        p, nv, K, V = self._poly_args(polys, nverts, T)
        mask = torch.empty((T, K, H, W), dtype=torch.uint8, device=self.tdev)
        self._check(self.lib.vhr_poly_mask(self.ctx, T, H, W, self._p(p), self._p(nv), K, V, self._p(mask),
                                           self._stream()), "vhr_poly_mask")
        return mask

    # ------------------------------------------------------------------ BPM
    def bpm_fft(self, trace, starts, lens, fs: float, band, detrend: int = DETREND_NONE, mode: int = FFT_ANALYSIS,
                max_len: Optional[int] = None):
        """trace float64 (n,) or (n,C); windows (start,len) -> (bpm float64 (n_win), bin int32).
        A device float64 VIEW (e.g. ``roi_mean[:, 0, 1]`` or ``roi_mean[:, :, 1]`` of a (T,K,3) trace) is
        read in place through its row / column strides; with C > 1 columns the result is the peak of
        the column with the largest peak (estimate_bpm.py:59-64)."""
        torch = _torch()
        tr = trace
        if not (isinstance(tr, torch.Tensor) and tr.is_cuda and tr.dtype == torch.float64 and tr.ndim in (1, 2)
                and tr.shape[0] >= 1 and all(st_ >= 1 for st_ in tr.stride())):
            tr = self._dev(trace, torch.float64)
            if tr.ndim == 1:
                tr = tr[:, None].contiguous()
        n, Cc = tr.shape[0], (1 if tr.ndim == 1 else tr.shape[1])
        ld = int(tr.stride(0)) if n > 1 else Cc
        cs = int(tr.stride(1)) if (tr.ndim == 2 and Cc > 1) else 1
        st = self._dev(starts, torch.int32)
        ln = self._dev(lens, torch.int32)
        nw = st.numel()
        bpm = torch.empty(nw, dtype=torch.float64, device=self.tdev)
        kbin = torch.empty(nw, dtype=torch.int32, device=self.tdev)
        if nw == 0:
            return bpm, kbin
        if max_len is None:      # device-resident window lists: pass max_len to avoid a D2H sync
            lens_np = np.asarray(lens if not isinstance(lens, torch.Tensor) else lens.cpu(), dtype=np.int64)
            max_len = int(lens_np.max())
        max_len = int(min(n, max(1, max_len)))
        self._check(self.lib.vhr_bpm_fft(self.ctx, self._p(tr), n, Cc, ld, cs, self._p(st), self._p(ln), nw, max_len,
                                         float(fs), float(band[0]), float(band[1]), detrend, mode, self._p(bpm),
                                         self._p(kbin), self._stream()), "vhr_bpm_fft")
        return bpm, kbin

    def bpm_welch(self, trace, starts, lens, fs: float, band, detrend: int = DETREND_NONE, filt_kind: int = FILT_NONE,
                  coef: Optional[np.ndarray] = None, welch_seconds: float = 9.0, want_filtered: bool = False):
        """-> (bpm, bin, filtered | None).  coef: SOS (n_sec,6) or FIR taps (float64, host)."""
        torch = _torch()
        tr = self._dev(trace, torch.float64).reshape(-1)
        n = tr.numel()
        lens_np = np.asarray(lens if not isinstance(lens, torch.Tensor) else lens.cpu(), dtype=np.int64)
        st = self._dev(starts, torch.int32)
        ln = self._dev(lens, torch.int32)
        nw = st.numel()
        bpm = torch.empty(nw, dtype=torch.float64, device=self.tdev)
        kbin = torch.empty(nw, dtype=torch.int32, device=self.tdev)
        if nw == 0:
            return bpm, kbin, None
        max_len = int(min(n, max(1, lens_np.max())))
        filt = torch.empty((nw, max_len), dtype=torch.float64, device=self.tdev) if want_filtered else None
        n_coef, cptr = 0, C.c_void_p(0)
        if filt_kind != FILT_NONE:
            c_np = np.ascontiguousarray(coef, dtype=np.float64)
            n_coef = c_np.shape[0]
            cptr = C.c_void_p(c_np.ctypes.data)
        self._check(self.lib.vhr_bpm_welch(self.ctx, self._p(tr), n, self._p(st), self._p(ln), nw, float(fs),
                                           float(band[0]), float(band[1]), detrend, filt_kind, cptr, n_coef,
                                           float(welch_seconds), self._p(bpm), self._p(kbin), self._p(filt), max_len,
                                           self._stream()), "vhr_bpm_welch")
        return bpm, kbin, filt

    def ica_fastica(self, trace, starts, lens, w_init: Optional[np.ndarray] = None, max_iter: int = 300, tol: float = 1e-6):
        """Batched FastICA over windows of a mean-BGR trace (analysis/measurement/ica.py:36-69): trace float64
        (n,3), windows (start,len) -> (sources float64 (n_win, max_len, 3) device, n_iter int32 (n_win) device;
        negative where the fixed point did not reach ``tol`` -- the reference skips those windows).  ``w_init``
        defaults to the reference's ``FastICA(random_state=0)`` start matrix."""
        torch = _torch()
        tr = self._dev(trace, torch.float64).reshape(-1, 3)
        n = tr.shape[0]
        lens_np = np.asarray(lens if not isinstance(lens, torch.Tensor) else lens.cpu(), dtype=np.int64)
        st, ln = self._dev(starts, torch.int32), self._dev(lens, torch.int32)
        nw = st.numel()
        max_len = int(min(n, max(1, lens_np.max()))) if nw else 1
        src = torch.empty((nw, max_len, 3), dtype=torch.float64, device=self.tdev)
        nit = torch.empty(nw, dtype=torch.int32, device=self.tdev)
        if nw == 0:
            return src, nit
        if w_init is None:
            w_init = np.random.RandomState(0).normal(size=(3, 3))       # check_random_state(0) at every fit (ica.py:43)
        w = np.ascontiguousarray(w_init, dtype=np.float64).reshape(3, 3)
        self._check(self.lib.vhr_ica_fastica(self.ctx, self._p(tr), n, self._p(st), self._p(ln), nw, max_len,
                                             C.c_void_p(w.ctypes.data), int(max_iter), float(tol), self._p(src), self._p(nit),
                                             self._stream()), "vhr_ica_fastica")
        return src, nit

    def sos_causal(self, x, sos: np.ndarray, state):
        """Causal SOS filter with carried state (n_sec,2) float64 device tensor (updated)."""
        torch = _torch()
        xd = self._dev(x, torch.float64).reshape(-1)
        sos = np.ascontiguousarray(sos, dtype=np.float64)
        y = torch.empty_like(xd)
        assert state.dtype == torch.float64 and state.is_cuda and state.numel() == sos.shape[0] * 2
        self._check(self.lib.vhr_sos_causal(self.ctx, self._p(xd), xd.numel(), C.c_void_p(sos.ctypes.data),
                                            sos.shape[0], self._p(state), self._p(y), self._stream()),
                    "vhr_sos_causal")
        return y


    # ------------------------------------------------------------------ degradations / metric
    def degrade_noise(self, frames, sigma: float, seed: int = 0, clip: int = 0, t0: int = 0, out=None):
        """colour_noise.add_gaussian_noise on device: counter-based 12-term Irwin-Hall draw of std
        ``sigma`` LSB (tails to 5.98 sigma), clip, truncate (colour_noise.py:22-24)."""
        torch = _torch()
        assert frames.dtype == torch.uint8 and frames.is_cuda and frames.is_contiguous()
        T, H, W, _ = frames.shape
        out = torch.empty_like(frames) if out is None else out
        gain = int(round(sigma * 65536.0 / float(np.sqrt(65535.0))))      # Q16 per unit of the twelve-byte sum (std sqrt(65535))
        self._check(self.lib.vhr_degrade_noise_u8(self.ctx, self._p(frames), self._p(out), T, H, W, gain,
                                                  seed & 0xFFFFFFFF, clip & 0xFFFFFFFF, t0, self._stream()),
                    "vhr_degrade_noise_u8")
        return out

    def degrade_quantise(self, frames, bits: int, out=None):
        """colour_quantisation.quantise_colour on device."""
        torch = _torch()
        assert frames.dtype == torch.uint8 and frames.is_cuda and frames.is_contiguous()
        out = torch.empty_like(frames) if out is None else out
        self._check(self.lib.vhr_degrade_quantise_u8(self.ctx, self._p(frames), self._p(out), frames.numel(), int(bits),
                                                     self._stream()), "vhr_degrade_quantise_u8")
        return out

    def align_mae(self, truth_t, truth_hr, measured):
        """Step-hold truth alignment + MAE -> (aligned_hr (m,) tensor, mae float)."""
        torch = _torch()
        tt = self._dev(truth_t, torch.float64).reshape(-1)
        th = self._dev(truth_hr, torch.float64).reshape(-1)
        me = self._dev(measured, torch.float64).reshape(-1, 2)
        m = me.shape[0]
        aligned = torch.empty(m, dtype=torch.float64, device=self.tdev)
        mae = torch.empty(1, dtype=torch.float64, device=self.tdev)
        self._check(self.lib.vhr_align_mae(self.ctx, self._p(tt), self._p(th), tt.numel(), self._p(me), m,
                                           self._p(aligned), self._p(mae), self._stream()), "vhr_align_mae")
        return aligned, float(mae.item())


_default: dict = {}


def default_engine(device: Optional[int] = None) -> Engine:
    """Process-wide engine per device (what the drop-in functions use)."""
    torch = _torch()
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if device not in _default:
        _default[device] = Engine(device)
    return _default[device]
