"""ctypes binding of libqsvc_b200.so (C ABI declared in include/qsvc_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is
present, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libqsvc_b200.so")
CSRC = os.path.join(_HERE, "csrc")

QSVC_OK, QSVC_EINVAL, QSVC_ECUDA, QSVC_ENOMEM, QSVC_EDOMAIN = 0, -1, -2, -3, -4


class QsvcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"qsvc_b200 error {code}: {msg}")
        self.code = code


class AnalyzeParams(C.Structure):
    _fields_ = [
        ("pixels_in_x", C.c_int), ("pixels_in_y", C.c_int), ("TRLs", C.c_int),
        ("block_size", C.c_int), ("block_size_min", C.c_int), ("border_size", C.c_int),
        ("block_overlaping", C.c_int), ("search_range", C.c_int),
        ("subpixel_accuracy", C.c_int), ("always_B", C.c_int),
        ("update_factor", C.c_float), ("first_gop_is_global_first", C.c_int),
    ]


u8p, i16p, chp = C.POINTER(C.c_uint8), C.POINTER(C.c_int16), C.c_char_p


class LevelOut(C.Structure):
    _fields_ = [("high", u8p), ("motion", i16p), ("motion_filtered", i16p),
                ("frame_types", C.c_void_p), ("low", u8p)]

_i = C.c_int
# qsvc_tail_fn: (user, level, synthesis, phase, state, bytes) -> int
TAIL_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint8), C.c_longlong)

# qsvc_boundary_fn: (user, level, inverse, phase, data, bytes) -> int
BOUNDARY_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint8), C.c_longlong)

# name -> (restype, argtypes); every symbol include/qsvc_b200.h declares
SIGNATURES = {
    "qsvc_version": (_i, []),
    "qsvc_device_count": (_i, []),
    "qsvc_create": (C.c_void_p, [_i]),
    "qsvc_destroy": (None, [C.c_void_p]),
    "qsvc_last_error": (C.c_char_p, []),
    "qsvc_launch_count": (C.c_longlong, [C.c_void_p]),
    "qsvc_timer_start": (_i, [C.c_void_p]),
    "qsvc_timer_stop": (_i, [C.c_void_p, C.POINTER(C.c_float)]),
    "qsvc_synchronize": (_i, [C.c_void_p]),
    "qsvc_set_tail_exchange": (_i, [C.c_void_p, TAIL_FN, C.c_void_p]),
    "qsvc_set_tail_exchange_device": (_i, [C.c_void_p, TAIL_FN, C.c_void_p]),
    "qsvc_set_boundary_exchange": (_i, [C.c_void_p, BOUNDARY_FN, C.c_void_p]),
    "qsvc_set_overlap": (_i, [C.c_void_p, _i]),
    "qsvc_set_me_mode": (_i, [C.c_void_p, _i]),
    "qsvc_set_mc_mode": (_i, [C.c_void_p, _i]),
    "qsvc_profile_enable": (_i, [C.c_void_p, _i]),
    "qsvc_profile_read": (_i, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_longlong), _i]),
    "qsvc_int_peak": (_i, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "qsvc_motion_estimate": (_i, [C.c_void_p, u8p, u8p, _i, _i, _i, _i, _i, _i, _i, _i, i16p]),
    "qsvc_decorrelate": (_i, [C.c_void_p, u8p, u8p, i16p, _i, _i, _i, _i, _i, _i, _i, _i, u8p,
                              C.c_void_p, i16p, u8p]),
    "qsvc_correlate": (_i, [C.c_void_p, u8p, u8p, i16p, C.c_void_p, _i, _i, _i, _i, _i, _i, _i,
                            u8p, u8p]),
    "qsvc_update": (_i, [C.c_void_p, _i, u8p, u8p, i16p, C.c_void_p, _i, _i, _i, _i, C.c_float,
                         u8p]),
    "qsvc_bidirectional_motion_decorrelate": (_i, [C.c_void_p, _i, i16p, _i, _i, _i, i16p]),
    "qsvc_interlevel_motion_decorrelate": (_i, [C.c_void_p, _i, i16p, _i, i16p, _i, _i, _i, i16p]),
    "qsvc_resident_fetch_motion_residue": (_i, [C.c_void_p, _i, i16p]),
    "qsvc_sse": (_i, [C.c_void_p, u8p, u8p, C.c_longlong, _i, C.POINTER(C.c_ulonglong)]),
    "qsvc_resident_load": (_i, [C.c_void_p, u8p, _i, _i, _i]),
    "qsvc_resident_analyze": (_i, [C.c_void_p, C.POINTER(AnalyzeParams)]),
    "qsvc_analyze": (_i, [C.c_void_p, C.POINTER(AnalyzeParams), u8p, _i, C.POINTER(LevelOut)]),
    "qsvc_host_alloc": (C.c_void_p, [C.c_size_t]),
    "qsvc_host_free": (None, [C.c_void_p]),
    "qsvc_host_register": (_i, [C.c_void_p, C.c_size_t]),
    "qsvc_host_unregister": (_i, [C.c_void_p]),
    "qsvc_resident_fetch": (_i, [C.c_void_p, _i, u8p, i16p, i16p, C.c_void_p, u8p]),
    "qsvc_resident_stats": (_i, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_float),
                                 C.POINTER(C.c_float)]),
    "qsvc_resident_push": (_i, [C.c_void_p, _i, _i, u8p, i16p, C.c_void_p, u8p, _i, _i, _i]),
    "qsvc_resident_synthesize": (_i, [C.c_void_p, C.POINTER(AnalyzeParams)]),
    "qsvc_resident_fetch_low0": (_i, [C.c_void_p, u8p, _i]),
}

_lib = None


def build(force: bool = False) -> str:
    """Compiles the CUDA sources for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".inc"))]
    srcs.append(os.path.join(_HERE, "..", "include", "qsvc_b200.h"))
    stale = force or not os.path.exists(SO_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(SO_PATH) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", CSRC, "-j8"], stdout=subprocess.DEVNULL)
    return SO_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise QsvcError(QSVC_ECUDA, f"{SO_PATH} is missing: run `python -c 'import "
                            "__graft_entry__ as g; g.build()'` (there is no CPU fallback)")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return lib().qsvc_last_error().decode(errors="replace")


def check(rc: int) -> None:
    if rc != QSVC_OK:
        raise QsvcError(rc, last_error())
