"""ctypes binding of libvhr_b200.so (include/vhr_b200.h).  No CPU fallback: importing the
package works anywhere (so host logic can be tested on CPU), but the first call that needs
the library raises if the .so is missing or no CUDA device is present."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# VHR_LIB selects another build of the same library (the -DVHR_WATCHDOG one, csrc/build.py --watchdog)
LIB_PATH = os.environ.get("VHR_LIB") or os.path.join(HERE, "csrc", "libvhr_b200.so")

c_void_p, c_int, c_int64, c_double, c_float = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_float


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint32), ("clip", C.c_uint32), ("T", C.c_int32), ("H", C.c_int32),
                ("W", C.c_int32), ("t0", C.c_int32), ("face", C.c_int32 * 4),
                ("base_q8", (C.c_int32 * 3) * 2), ("noise_gain", C.c_int32)]


# name -> (restype, argtypes); mirrors include/vhr_b200.h one to one
SIGNATURES = {
    "vhr_abi_version": (c_int, []),
    "vhr_create": (c_int, [C.POINTER(c_void_p), c_int]),
    "vhr_destroy": (c_int, [c_void_p]),
    "vhr_last_error": (C.c_char_p, [c_void_p]),
    "vhr_launch_count": (c_int64, [c_void_p]),
    "vhr_synth_clip": (c_int, [c_void_p, C.POINTER(SynthParams), c_void_p, c_void_p, c_void_p]),
    "vhr_pyr_dims": (c_int, [c_int, c_int, c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "vhr_pyrdown_cascade": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "vhr_pyrdown_umma_plan": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "vhr_pyrdown_umma_accumulators": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "vhr_temporal_bandpass": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_double, c_double,
                                      c_double, c_float, c_void_p]),
    "vhr_band_bins": (c_int, [c_int, c_double, c_double, c_double, C.POINTER(c_int), C.POINTER(c_int)]),
    "vhr_collapse_addback_roi": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                         c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "vhr_collapse_addback_poly": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "vhr_roi_mean_rect_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int,
                                     c_void_p, c_void_p, c_void_p]),
    "vhr_roi_mean_poly_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                                     c_void_p, c_void_p, c_void_p]),
    "vhr_roi_mean_poly_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                                      c_void_p, c_void_p, c_void_p]),
    "vhr_poly_mask": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "vhr_bpm_fft": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_double, c_double,
                            c_double, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "vhr_bpm_welch": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_double, c_double, c_double,
                              c_int, c_int, c_void_p, c_int, c_double, c_void_p, c_void_p, c_void_p, c_int,
                              c_void_p]),
    "vhr_sos_causal": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "vhr_ica_fastica": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_double,
                                c_void_p, c_void_p, c_void_p]),
    "vhr_degrade_noise_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, C.c_uint32, C.c_uint32,
                                     c_int, c_void_p]),
    "vhr_degrade_quantise_u8": (c_int, [c_void_p, c_void_p, c_void_p, C.c_longlong, c_int, c_void_p]),
    "vhr_align_mae": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "vhr_evm_roi_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double, c_double, c_double,
                                 c_float, c_void_p, c_int, c_void_p, c_void_p]),
    "vhr_evm_poly_host": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double, c_double, c_double,
                                  c_float, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "vhr_trim": (c_int, [c_void_p]),
}

_lib = None


class VhrError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the CUDA library (once).  Raises VhrError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VhrError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                           "(there is no CPU fallback for this path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.vhr_abi_version() != 2:
            raise VhrError("libvhr_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def check(lib, ctx, rc: int, what: str):
    if rc != 0:
        msg = lib.vhr_last_error(ctx)
        raise VhrError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")
