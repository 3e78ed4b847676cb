"""ctypes front-end of oracle/libmctf_oracle.so (CPU restatement of the MCTF path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never by qsvc_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libmctf_oracle.so")
    src = os.path.join(_HERE, "mctf_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "lib"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_entropy.restype = C.c_float
    return _LIB


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _fb(X, Y):
    return X * Y + 2 * (X // 2) * (Y // 2)


def motion_estimate(even, odd, X, Y, block_size, search_range, subpixel_accuracy=0, border_size=0):
    """even: (n+1, fb) uint8, odd: (n, fb) uint8 -> (n, 4, by, bx) int16."""
    even = np.ascontiguousarray(even, np.uint8)
    odd = np.ascontiguousarray(odd, np.uint8)
    n = odd.shape[0]
    by, bx = Y // block_size, X // block_size
    mv = np.zeros((n, 4, by, bx), np.int16)
    rc = lib().orc_motion_estimate(_p(even, C.c_uint8), _p(odd, C.c_uint8), 2 * n + 1, X, Y,
                                   block_size, border_size, search_range, subpixel_accuracy,
                                   _p(mv, C.c_int16))
    assert rc == 0
    return mv


def decorrelate(even, odd, mv, X, Y, block_size, search_range, subpixel_accuracy=0,
                block_overlaping=0, always_B=0):
    """-> (high (n,fb) u8, frame_types bytes, mv_out (n,4,by,bx) i16, prediction (n,fb) u8, rc)."""
    even = np.ascontiguousarray(even, np.uint8)
    odd = np.ascontiguousarray(odd, np.uint8)
    mv = np.ascontiguousarray(mv, np.int16)
    n = odd.shape[0]
    high = np.zeros((n, _fb(X, Y)), np.uint8)
    pred = np.zeros((n, _fb(X, Y)), np.uint8)
    types = np.zeros(n, np.uint8)
    mvo = np.zeros_like(mv)
    rc = lib().orc_decorrelate(1, _p(even, C.c_uint8), _p(odd, C.c_uint8), None,
                               _p(mv, C.c_int16), None, 2 * n + 1, X, Y, block_size,
                               block_overlaping, search_range, subpixel_accuracy, always_B,
                               _p(high, C.c_uint8), None, _p(types, C.c_char),
                               _p(mvo, C.c_int16), _p(pred, C.c_uint8))
    assert rc in (0, 1)
    return high, types.tobytes(), mvo, pred, rc


def correlate(even, high, mv, frame_types, X, Y, block_size, search_range, subpixel_accuracy=0,
              block_overlaping=0):
    """-> (odd (n,fb) u8, prediction (n,fb) u8)."""
    even = np.ascontiguousarray(even, np.uint8)
    high = np.ascontiguousarray(high, np.uint8)
    mv = np.ascontiguousarray(mv, np.int16)
    n = high.shape[0]
    types = np.frombuffer(bytes(frame_types), np.uint8).copy()
    odd = np.zeros((n, _fb(X, Y)), np.uint8)
    pred = np.zeros((n, _fb(X, Y)), np.uint8)
    rc = lib().orc_decorrelate(0, _p(even, C.c_uint8), None, _p(high, C.c_uint8),
                               _p(mv, C.c_int16), _p(types, C.c_char), 2 * n + 1, X, Y,
                               block_size, block_overlaping, search_range, subpixel_accuracy, 1,
                               None, _p(odd, C.c_uint8), None, None, _p(pred, C.c_uint8))
    assert rc == 0
    return odd, pred


def bidirectional_motion(fields, inverse=False):
    """bidirectional_motion_decorrelate (inverse: _correlate) on (n, 4, by, bx) int16 fields."""
    fields = np.ascontiguousarray(fields, np.int16)
    n, _, by, bx = fields.shape
    out = np.zeros_like(fields)
    rc = lib().orc_bidirectional_motion(0 if inverse else 1, _p(fields, C.c_int16), n, by, bx, _p(out, C.c_int16))
    assert rc == 0
    return out


def interlevel_motion(fields, reference=None, fields_in_predicted=None, inverse=False):
    """interlevel_motion_decorrelate (inverse: _correlate); reference None = /dev/zero."""
    fields = np.ascontiguousarray(fields, np.int16)
    n, _, by, bx = fields.shape
    if reference is not None:
        reference = np.ascontiguousarray(reference, np.int16)
    out = np.zeros_like(fields)
    wr = lib().orc_interlevel_motion(0 if inverse else 1, _p(fields, C.c_int16), n,
                                     _p(reference, C.c_int16), 0 if reference is None else reference.shape[0],
                                     n if fields_in_predicted is None else fields_in_predicted, by, bx,
                                     _p(out, C.c_int16))
    return out[:wr]


def update(frames_in, high, mv, frame_types, X, Y, block_size, update_factor, inverse=False):
    """update (even->low) or, with inverse=True, un_update (low->even)."""
    frames_in = np.ascontiguousarray(frames_in, np.uint8)
    high = np.ascontiguousarray(high, np.uint8)
    mv = np.ascontiguousarray(mv, np.int16)
    n = high.shape[0]
    types = np.frombuffer(bytes(frame_types), np.uint8).copy()
    out = np.zeros((n + 1, _fb(X, Y)), np.uint8)
    rc = lib().orc_update(0 if inverse else 1, _p(frames_in, C.c_uint8), _p(high, C.c_uint8),
                          _p(mv, C.c_int16), _p(types, C.c_char), 2 * n + 1, X, Y, block_size,
                          C.c_float(update_factor), _p(out, C.c_uint8))
    assert rc == 0
    return out


def dwt53(img, levels, synth=False, y=None, x=None):
    """In-place reference-layout 2-D 5/3 transform on a copy of an int16 image."""
    img = np.ascontiguousarray(img, np.int16).copy()
    y = img.shape[0] if y is None else y
    x = img.shape[1] if x is None else x
    lib().orc_dwt53(_p(img, C.c_int16), img.shape[1], y, x, levels, 1 if synth else 0)
    return img


def bordered_view(luma, y_alloc, x_alloc, b_alloc, b_fill):
    luma = np.ascontiguousarray(luma, np.uint8)
    out = np.zeros((y_alloc + 2 * b_alloc, x_alloc + 2 * b_alloc), np.int16)
    lib().orc_bordered_view(_p(luma, C.c_uint8), luma.shape[0], luma.shape[1], y_alloc, x_alloc,
                            b_alloc, b_fill, _p(out, C.c_int16))
    return out


def entropy(count):
    count = np.ascontiguousarray(count, np.int32)
    return float(lib().orc_entropy(_p(count, C.c_int), count.size))


# ---------------------------------------------------------------- level chains

def analyze(low0, X, Y, TRLs, block_size, search_range, subpixel_accuracy=0, update_factor=0.0,
            always_B=0, block_overlaping=0, border_size=0, block_size_min=32):
    """Python-3 restatement of analyze.py:107-153 + analyze_step.py:115-232 on
    in-memory arrays.  Returns dict of per-level outputs keyed like the files."""
    out = {}
    low = np.ascontiguousarray(low0, np.uint8)
    sr, bs = search_range, block_size
    if bs < block_size_min:
        block_size_min = bs
    for t in range(1, TRLs):
        even, odd = low[0::2], low[1::2]
        mv = motion_estimate(even, odd, X, Y, bs, sr, subpixel_accuracy, border_size)
        high, types, mvf, pred, rc = decorrelate(even, odd, mv, X, Y, bs, sr, subpixel_accuracy,
                                                 block_overlaping, always_B)
        low = update(even, high, mvf, types, X, Y, bs, update_factor)
        out[f"motion_{t}"] = mv
        out[f"motion_filtered_{t}"] = mvf
        out[f"high_{t}"] = high
        out[f"frame_types_{t}"] = types
        out[f"low_{t}"] = low
        out[f"prediction_even_{t}"] = pred
        sr = min(sr * 2, 128)
        bs = max(bs // 2, block_size_min)
    return out
