"""Host-side mirror of the reference's MCTF tool set on top of the C ABI.

One `Context` per GPU.  Methods are named after the reference tools they
replace (split | motion_estimate | decorrelate | update | un_update |
correlate | merge) and take/return the tools' file payloads as numpy arrays:
frames (n, frame_bytes) uint8, motion (n, 4, by, bx) int16, frame types bytes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import AnalyzeParams, LevelOut, check
from .yuv import frame_bytes

SEARCH_RANGE_MAX = 128  # analyze.py:26


def _u8(a):
    return a.ctypes.data_as(_lib.u8p)


def _i16(a):
    return a.ctypes.data_as(_lib.i16p)


def gop_size(TRLs: int) -> int:
    """GOP.py:23-24"""
    return 2 ** (TRLs - 1)


def level_schedule(GOPs, TRLs, block_size, search_range, block_size_min=32):
    """Per-level (t, pictures, search_range, block_size) of analyze.py:107-153."""
    pictures = GOPs * gop_size(TRLs) + 1
    if block_size < block_size_min:
        block_size_min = block_size
    out = []
    for t in range(1, TRLs):
        out.append(dict(t=t, pictures=pictures, pairs=pictures // 2, search_range=search_range,
                        block_size=block_size))
        pictures = (pictures + 1) // 2
        search_range = min(search_range * 2, SEARCH_RANGE_MAX)
        block_size = max(block_size // 2, block_size_min)
    return out


def split(low: np.ndarray):
    """split.cpp:229-341: frame 0 -> even, then alternately odd, even."""
    return low[0::2], low[1::2]


def merge(even: np.ndarray, odd: np.ndarray) -> np.ndarray:
    n = odd.shape[0]
    low = np.empty((2 * n + 1, even.shape[1]), np.uint8)
    low[0::2] = even[: n + 1]
    low[1::2] = odd
    return low


class Context:
    def __init__(self, device: int = 0):
        self._L = _lib.lib()
        self._h = self._L.qsvc_create(device)
        if not self._h:
            raise _lib.QsvcError(_lib.QSVC_ECUDA, _lib.last_error())
        self.device = device
        self._pinned = []
        self._out_cache = {}

    def close(self):
        if getattr(self, "_h", None):
            self._out_cache = {}
            for ptr in self._pinned:
                self._L.qsvc_host_free(ptr)
            self._pinned = []
            self._L.qsvc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- bookkeeping
    @property
    def launches(self) -> int:
        return int(self._L.qsvc_launch_count(self._h))

    def timer_start(self):
        check(self._L.qsvc_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        check(self._L.qsvc_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def synchronize(self):
        check(self._L.qsvc_synchronize(self._h))

    KERNEL_CLASSES = ("image", "dwt_rows", "dwt_cols", "search", "predict", "residue", "update",
                      "search_exact")

    def set_me_mode(self, mode: int):
        """0 automatic, 1 literal (materialised) ME path, 2 fused ME path or fail."""
        check(self._L.qsvc_set_me_mode(self._h, mode))

    def set_overlap(self, on: bool):
        """Resident analysis with update_factor == 0: motion estimation of level t+1 beside the
        decorrelate of level t on a second stream (default on).  Same results either way."""
        check(self._L.qsvc_set_overlap(self._h, 1 if on else 0))

    def set_mc_mode(self, mode: int):
        """0 automatic, 1 literal decorrelate/correlate path, 2 byte-plane fused path or fail."""
        check(self._L.qsvc_set_mc_mode(self._h, mode))

    def set_tail_exchange(self, fn=None, device=False):
        """Installs (or clears) the GOP-shard exchange of the prediction tail state
        (include/qsvc_b200.h): fn(level, synthesis, phase, state) with
        `state` a writable uint8 array; phase 0 returns True after filling in the left
        neighbour's state, phase 1 receives the state to pass to the right.
        device=True: `state` is instead a (device pointer, bytes) tuple of the context's GPU
        (qsvc_set_tail_exchange_device): the state travels GPU to GPU without a host hop."""
        setter = self._L.qsvc_set_tail_exchange_device if device else self._L.qsvc_set_tail_exchange
        if fn is None:
            self._tail_cb = None
            check(self._L.qsvc_set_tail_exchange(self._h, _lib.TAIL_FN(), None))
            return

        def cb(_user, level, synthesis, phase, state, nbytes):
            try:
                if device:
                    a = (C.cast(state, C.c_void_p).value, int(nbytes))
                else:
                    a = np.ctypeslib.as_array(state, shape=(int(nbytes),))
                r = fn(int(level), int(synthesis), int(phase), a)
                return 1 if r else 0
            except Exception:  # noqa: BLE001 -- must not propagate through the C frame
                import traceback
                traceback.print_exc()
                return -1

        self._tail_cb = _lib.TAIL_FN(cb)  # keep the trampoline alive
        check(setter(self._h, self._tail_cb, None))

    def set_boundary_exchange(self, fn=None):
        """Installs (or clears) the GOP-shard exchange of the shared boundary frame for
        update_factor != 0 (include/qsvc_b200.h):
        fn(level, inverse, phase, data) with `data` a writable uint8 view (int16 planes in
        phases 0/1, an I420 frame in phases 2/3); returns True / False as documented there."""
        if fn is None:
            self._boundary_cb = None
            check(self._L.qsvc_set_boundary_exchange(self._h, _lib.BOUNDARY_FN(), None))
            return

        def cb(_user, level, inverse, phase, data, nbytes):
            try:
                a = np.ctypeslib.as_array(data, shape=(int(nbytes),))
                return 1 if fn(int(level), int(inverse), int(phase), a) else 0
            except Exception:  # noqa: BLE001 -- must not propagate through the C frame
                import traceback
                traceback.print_exc()
                return -1

        self._boundary_cb = _lib.BOUNDARY_FN(cb)  # keep the trampoline alive
        check(self._L.qsvc_set_boundary_exchange(self._h, self._boundary_cb, None))

    def profile_enable(self, on=True):
        check(self._L.qsvc_profile_enable(self._h, 1 if on else 0))

    def profile_read(self):
        """{class: (device ms, launches)} since the previous read."""
        n = len(self.KERNEL_CLASSES)
        ms = (C.c_float * n)()
        cnt = (C.c_longlong * n)()
        check(self._L.qsvc_profile_read(self._h, ms, cnt, n))
        return {k: (ms[i], cnt[i]) for i, k in enumerate(self.KERNEL_CLASSES)}

    def int_peak(self):
        """Measured SAD-op/s of this GPU: (packed u8 __vsadu4, 32-bit __sad)."""
        a, b = C.c_double(), C.c_double()
        check(self._L.qsvc_int_peak(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # ---- per-tool calls (host buffers in, host buffers out)
    def motion_estimate(self, even, odd, X, Y, block_size=32, search_range=4, subpixel_accuracy=0,
                        border_size=0, first_pair_is_global_first=True):
        even = np.ascontiguousarray(even, np.uint8)
        odd = np.ascontiguousarray(odd, np.uint8)
        n = odd.shape[0]
        assert even.shape == (n + 1, frame_bytes(X, Y)) and odd.shape[1] == frame_bytes(X, Y)
        mv = np.zeros((n, 4, Y // block_size, X // block_size), np.int16)
        check(self._L.qsvc_motion_estimate(self._h, _u8(even), _u8(odd), n, X, Y, block_size,
                                           border_size, search_range, subpixel_accuracy,
                                           1 if first_pair_is_global_first else 0, _i16(mv)))
        return mv

    def decorrelate(self, even, odd, motion, X, Y, block_size=16, search_range=4,
                    subpixel_accuracy=0, block_overlaping=0, always_B=0, want_prediction=False):
        even = np.ascontiguousarray(even, np.uint8)
        odd = np.ascontiguousarray(odd, np.uint8)
        motion = np.ascontiguousarray(motion, np.int16)
        n = odd.shape[0]
        high = np.zeros((n, frame_bytes(X, Y)), np.uint8)
        pred = np.zeros((n, frame_bytes(X, Y)), np.uint8) if want_prediction else None
        types = C.create_string_buffer(max(n, 1))
        mvo = np.zeros_like(motion)
        check(self._L.qsvc_decorrelate(self._h, _u8(even), _u8(odd), _i16(motion), n, X, Y,
                                       block_size, block_overlaping, search_range,
                                       subpixel_accuracy, always_B, _u8(high), types, _i16(mvo),
                                       _u8(pred) if want_prediction else None))
        return high, types.raw[:n], mvo, pred

    def correlate(self, even, high, motion, frame_types, X, Y, block_size=16, search_range=4,
                  subpixel_accuracy=0, block_overlaping=0, want_prediction=False):
        even = np.ascontiguousarray(even, np.uint8)
        high = np.ascontiguousarray(high, np.uint8)
        motion = np.ascontiguousarray(motion, np.int16)
        n = high.shape[0]
        odd = np.zeros((n, frame_bytes(X, Y)), np.uint8)
        pred = np.zeros((n, frame_bytes(X, Y)), np.uint8) if want_prediction else None
        types = C.create_string_buffer(bytes(frame_types), max(n, 1))
        check(self._L.qsvc_correlate(self._h, _u8(even), _u8(high), _i16(motion), types, n, X, Y,
                                     block_size, block_overlaping, search_range,
                                     subpixel_accuracy, _u8(odd),
                                     _u8(pred) if want_prediction else None))
        return odd, pred

    def bidirectional_motion_decorrelate(self, fields, inverse=False):
        """(n, 4, by, bx) int16 -> same shape: NEXT -= PREV (inverse: += ), the reference's
        bidirectional_motion_decorrelate / _correlate (bidirectional_motion_decorrelate.cpp:25-52)."""
        fields = np.ascontiguousarray(fields, np.int16)
        n, four, by, bx = fields.shape
        assert four == 4
        out = np.zeros_like(fields)
        check(self._L.qsvc_bidirectional_motion_decorrelate(self._h, 1 if inverse else 0, _i16(fields), n,
                                                            by, bx, _i16(out)))
        return out

    def interlevel_motion_decorrelate(self, fields, reference=None, inverse=False):
        """residue[k] = predicted[k] - reference[k // 2] / 2 (inverse: predicted = residue + ...),
        the reference's interlevel_motion_decorrelate / _correlate
        (interlevel_motion_decorrelate.cpp:32-69, 250-297); reference None = the tool's /dev/zero."""
        fields = np.ascontiguousarray(fields, np.int16)
        n, four, by, bx = fields.shape
        assert four == 4
        nref = 0
        if reference is not None:
            reference = np.ascontiguousarray(reference, np.int16)
            nref = reference.shape[0]
            assert reference.shape[1:] == fields.shape[1:]
        out = np.zeros_like(fields)
        check(self._L.qsvc_interlevel_motion_decorrelate(self._h, 1 if inverse else 0, _i16(fields), n,
                                                         _i16(reference) if nref else None, nref, by, bx,
                                                         _i16(out)))
        return out

    def resident_motion_residue(self, t, n_pairs, by, bx):
        """motion_residue_t of the last resident analysis (motion_compress.py:141-182 on the device)."""
        out = np.zeros((n_pairs, 4, by, bx), np.int16)
        check(self._L.qsvc_resident_fetch_motion_residue(self._h, t, _i16(out)))
        return out

    def sse(self, a, b, block_bytes):
        """Sum of squared byte differences per block of block_bytes bytes of two equally long
        uint8 streams (the distortion behind psnr.py:78-90): uint64 array, one entry per block."""
        a = np.ascontiguousarray(a, np.uint8).ravel()
        b = np.ascontiguousarray(b, np.uint8).ravel()
        n = min(a.size, b.size) // int(block_bytes)
        out = np.zeros(n, np.uint64)
        check(self._L.qsvc_sse(self._h, _u8(a), _u8(b), int(block_bytes), n,
                               out.ctypes.data_as(C.POINTER(C.c_ulonglong))))
        return out

    def psnr(self, a, b, block_bytes, peak=255.0):
        """(per-block PSNR in dB, overall PSNR in dB); identical blocks give inf."""
        sse = self.sse(a, b, block_bytes).astype(np.float64)
        with np.errstate(divide="ignore"):
            per = 10.0 * np.log10(peak * peak * float(block_bytes) / sse)
            tot = 10.0 * np.log10(peak * peak * float(block_bytes) * max(len(sse), 1) / sse.sum()) if len(sse) else np.inf
        return per, float(tot)

    def update(self, frames_in, high, motion, frame_types, X, Y, block_size=16,
               update_factor=0.25, inverse=False):
        frames_in = np.ascontiguousarray(frames_in, np.uint8)
        high = np.ascontiguousarray(high, np.uint8)
        motion = np.ascontiguousarray(motion, np.int16)
        n = high.shape[0]
        out = np.zeros((n + 1, frame_bytes(X, Y)), np.uint8)
        types = C.create_string_buffer(bytes(frame_types), max(n, 1))
        check(self._L.qsvc_update(self._h, 1 if inverse else 0, _u8(frames_in), _u8(high),
                                  _i16(motion), types, n, X, Y, block_size,
                                  C.c_float(update_factor), _u8(out)))
        return out

    def un_update(self, low, high, motion, frame_types, X, Y, block_size=16, update_factor=0.25):
        return self.update(low, high, motion, frame_types, X, Y, block_size, update_factor, True)

    # ---- whole-sequence analysis / synthesis with frames resident in HBM
    @staticmethod
    def _params(X, Y, TRLs, block_size, search_range, subpixel_accuracy, update_factor, always_B,
                block_overlaping, border_size, block_size_min, first_global=True):
        return AnalyzeParams(X, Y, TRLs, block_size, block_size_min, border_size,
                             block_overlaping, search_range, subpixel_accuracy, always_B,
                             float(update_factor), 1 if first_global else 0)

    def resident_load(self, low0, X, Y):
        low0 = np.ascontiguousarray(low0, np.uint8)
        assert low0.shape[1] == frame_bytes(X, Y)
        check(self._L.qsvc_resident_load(self._h, _u8(low0), low0.shape[0], X, Y))
        self._geom = (X, Y, low0.shape[0])

    def resident_analyze(self, TRLs, block_size=32, search_range=4, subpixel_accuracy=0,
                         update_factor=0.0, always_B=0, block_overlaping=0, border_size=0,
                         block_size_min=32, first_global=True):
        X, Y, _ = self._geom
        p = self._params(X, Y, TRLs, block_size, search_range, subpixel_accuracy, update_factor,
                         always_B, block_overlaping, border_size, block_size_min, first_global)
        check(self._L.qsvc_resident_analyze(self._h, C.byref(p)))
        self._last = (TRLs, block_size, search_range, block_size_min)

    def resident_stats(self):
        ops, sms, tms = C.c_double(), C.c_float(), C.c_float()
        check(self._L.qsvc_resident_stats(self._h, C.byref(ops), C.byref(sms), C.byref(tms)))
        return dict(sad_ops=ops.value, search_ms=sms.value, total_ms=tms.value)

    def resident_fetch(self, t, n_pairs, block_size, want=("high", "motion", "motion_filtered",
                                                           "frame_types", "low")):
        X, Y, _ = self._geom
        fb = frame_bytes(X, Y)
        by, bx = Y // block_size, X // block_size
        high = np.zeros((n_pairs, fb), np.uint8) if "high" in want else None
        mv = np.zeros((n_pairs, 4, by, bx), np.int16) if "motion" in want else None
        mvf = np.zeros((n_pairs, 4, by, bx), np.int16) if "motion_filtered" in want else None
        low = np.zeros((n_pairs + 1, fb), np.uint8) if "low" in want else None
        types = C.create_string_buffer(max(n_pairs, 1))
        check(self._L.qsvc_resident_fetch(
            self._h, t, _u8(high) if high is not None else None,
            _i16(mv) if mv is not None else None, _i16(mvf) if mvf is not None else None, types,
            _u8(low) if low is not None else None))
        return dict(high=high, motion=mv, motion_filtered=mvf, frame_types=types.raw[:n_pairs],
                    low=low)

    def host_alloc(self, shape, dtype=np.uint8):
        """numpy array in pinned host memory (freed with the context)."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        ptr = self._L.qsvc_host_alloc(max(nbytes, 1))
        if not ptr:
            raise _lib.QsvcError(_lib.QSVC_ENOMEM, _lib.last_error())
        self._pinned.append(ptr)
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def host_register(self, array):
        """Page-locks caller-owned host memory (qsvc_host_register), e.g. a shared mapping that
        several GOP shards write their slices of the gathered files into; False if the driver
        refuses (the copies then go through staging, slower but correct)."""
        rc = self._L.qsvc_host_register(C.c_void_p(array.ctypes.data), array.nbytes)
        return rc == _lib.QSVC_OK

    def host_unregister(self, array):
        self._L.qsvc_host_unregister(C.c_void_p(array.ctypes.data))

    def analyze(self, low0, X, Y, GOPs, TRLs, block_size=32, search_range=4, subpixel_accuracy=0,
                update_factor=0.0, always_B=0, block_overlaping=0, border_size=0,
                block_size_min=32, first_global=True, reuse_buffers=False, out=None):
        """analyze.py equivalent on arrays: returns {file name: payload}.

        One C-ABI call (qsvc_analyze): upload, every temporal level, and the download of
        each level's results overlapped with the next level's compute.  With
        reuse_buffers=True the returned arrays are views of pinned buffers owned by the
        context and are overwritten by the next call with the same geometry.  `out`: caller-
        provided contiguous arrays per file name (high_t, motion_t, motion_filtered_t, low_t) that
        receive the results directly, e.g. slices of a page-locked gathered file (host_register)."""
        low0 = np.ascontiguousarray(low0, np.uint8)
        assert low0.shape == (GOPs * gop_size(TRLs) + 1, frame_bytes(X, Y))
        sched = level_schedule(GOPs, TRLs, block_size, search_range, block_size_min)
        key = (X, Y, GOPs, TRLs, block_size, block_size_min)
        bufs = self._out_cache.get(key) if out is None else out
        if out is not None:
            reuse_buffers = True
            for s in sched:
                t, n, b = s["t"], s["pairs"], s["block_size"]
                want = {f"high_{t}": ((n, fb := frame_bytes(X, Y)), np.uint8),
                        f"motion_{t}": ((n, 4, Y // b, X // b), np.int16),
                        f"motion_filtered_{t}": ((n, 4, Y // b, X // b), np.int16)}
                if not (float(update_factor) == 0.0 and t < TRLs - 1):
                    want[f"low_{t}"] = ((n + 1, fb), np.uint8)
                for k, (shape, dt) in want.items():
                    a = out[k]
                    assert a.shape == shape and a.dtype == dt and a.flags["C_CONTIGUOUS"], k
        if bufs is None:
            bufs = {}
            fb = frame_bytes(X, Y)
            for s in sched:
                t, n, b = s["t"], s["pairs"], s["block_size"]
                bufs[f"high_{t}"] = self.host_alloc((n, fb))
                bufs[f"motion_{t}"] = self.host_alloc((n, 4, Y // b, X // b), np.int16)
                bufs[f"motion_filtered_{t}"] = self.host_alloc((n, 4, Y // b, X // b), np.int16)
                bufs[f"low_{t}"] = self.host_alloc((n + 1, fb))
            self._out_cache[key] = bufs
        outs = (LevelOut * TRLs)()
        types = {}
        # update_factor == 0 (analyze.py's default): update is the identity, low_t == even_t, i.e.
        # frames of the clip the caller already holds.  Only low_{TRLs-1} is downloaded (for symmetry
        # with the files texture_compress.py:183-215 consumes); the lower levels are views of low0.
        host_low = float(update_factor) == 0.0
        for s in sched:
            t = s["t"]
            types[t] = C.create_string_buffer(max(s["pairs"], 1))
            outs[t].high = _u8(bufs[f"high_{t}"])
            outs[t].motion = _i16(bufs[f"motion_{t}"])
            outs[t].motion_filtered = _i16(bufs[f"motion_filtered_{t}"])
            outs[t].frame_types = C.cast(types[t], C.c_void_p)
            if not (host_low and t < TRLs - 1):
                outs[t].low = _u8(bufs[f"low_{t}"])
        p = self._params(X, Y, TRLs, block_size, search_range, subpixel_accuracy, update_factor,
                         always_B, block_overlaping, border_size, block_size_min, first_global)
        check(self._L.qsvc_analyze(self._h, C.byref(p), _u8(low0), low0.shape[0], outs))
        self._geom = (X, Y, low0.shape[0])
        out = {}
        self.last_d2h_bytes = 0
        for s in sched:
            t = s["t"]
            for name in ("high", "motion", "motion_filtered", "low"):
                if name == "low" and host_low and t < TRLs - 1:
                    out[f"low_{t}"] = low0[:: 2 ** t]  # even_t: every 2^t-th frame of the input clip
                    continue
                a = bufs[f"{name}_{t}"]
                self.last_d2h_bytes += a.nbytes
                out[f"{name}_{t}"] = a if reuse_buffers else a.copy()
            out[f"frame_types_{t}"] = types[t].raw[: s["pairs"]]
            self.last_d2h_bytes += s["pairs"]
        return out

    def synthesize(self, subbands, X, Y, GOPs, TRLs, block_size=16, search_range=4,
                   subpixel_accuracy=0, update_factor=0.25, block_overlaping=0, out=None):
        """synthesize.py equivalent on arrays.  `subbands` maps high_t, motion_t,
        frame_types_t (t = 1..TRLs-1) and low_{TRLs-1} to their payloads.  Returns low_0
        (written into `out`, e.g. a host_alloc'd array, when given)."""
        fb = frame_bytes(X, Y)
        for t in range(TRLs - 1, 0, -1):
            high = np.ascontiguousarray(subbands[f"high_{t}"], np.uint8)
            mv = np.ascontiguousarray(subbands[f"motion_{t}"], np.int16)
            types = bytes(subbands[f"frame_types_{t}"])
            n = high.shape[0]
            low_top = None
            if t == TRLs - 1:
                low_top = np.ascontiguousarray(subbands[f"low_{t}"], np.uint8)
                assert low_top.shape == (n + 1, fb)
            tb = C.create_string_buffer(types, max(n, 1))
            check(self._L.qsvc_resident_push(self._h, t, n, _u8(high), _i16(mv), tb,
                                             _u8(low_top) if low_top is not None else None, X, Y,
                                             block_size))
        p = self._params(X, Y, TRLs, block_size, search_range, subpixel_accuracy, update_factor, 1,
                         block_overlaping, 0, block_size)
        check(self._L.qsvc_resident_synthesize(self._h, C.byref(p)))
        frames = GOPs * gop_size(TRLs) + 1
        if out is None:
            out = np.zeros((frames, fb), np.uint8)
        assert out.shape == (frames, fb) and out.dtype == np.uint8 and out.flags["C_CONTIGUOUS"]
        check(self._L.qsvc_resident_fetch_low0(self._h, _u8(out), frames))
        self._geom = (X, Y, frames)
        return out
