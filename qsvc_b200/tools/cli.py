"""Flag-compatible command-line front-ends of the reference's MCTF tools.

    mctf <tool> --flag=value ...        (bin/mctf, same contract as mctf.sh:35-39)
    python -m qsvc_b200.tools <tool> ...

Tools: split merge motion_estimate decorrelate correlate update un_update
       analyze_step analyze synthesize_step synthesize
       bidirectional_motion_decorrelate bidirectional_motion_correlate
       interlevel_motion_decorrelate interlevel_motion_correlate
Flag names, short forms and defaults follow the reference's getopt_long tables
(motion_estimate.cpp:500-530, decorrelate.cpp:209-251, update.cpp:170-205,
split.cpp:50-71) and MCTF_parser.py; unambiguous prefixes are accepted like
getopt_long does (synthesize_step.py:135-137 relies on --even= --low= --odd=).
Files are read from / written to the current directory.  Exit codes: 0 on
success, 1 for --help and for motion_estimate when the motion file already
exists (motion_estimate.cpp:659-682), 134 (abort) when a file cannot be opened.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

from .. import yuv
from ..mctf import Context, SEARCH_RANGE_MAX, gop_size, level_schedule, merge, split

RED, OFF = "\033[1;31m", "\033[0m"


def _error(msg: str):
    sys.stderr.write(f"{RED}{msg}{OFF}")


def _abort(tool: str, msg: str):
    _error(f"{tool}: {msg} ... aborting!\n")
    sys.exit(134)


class _Parser(argparse.ArgumentParser):
    def __init__(self, tool, **kw):
        super().__init__(prog=tool, add_help=False, allow_abbrev=True, **kw)
        self.tool = tool
        self.add_argument("--help", "-?", action="store_true")

    def opt(self, long, short, default, typ=str):
        names = [f"--{long}"] + ([f"-{short}"] if short else [])
        self.add_argument(*names, dest=long, default=default, type=typ)

    def parse(self, argv):
        args, _unknown = self.parse_known_args(argv)
        if args.help:
            self.print_usage()
            sys.exit(1)
        return args

    def error(self, message):  # unknown/ambiguous flags: the reference tools print and go on
        _error(f"{self.tool}: {message}\n")
        sys.exit(134)


def _read_frames(tool, fn, X, Y, count):
    if not os.path.exists(fn):
        _abort(tool, f'unable to read "{fn}"')
    return yuv.read_frames(fn, X, Y, count)


def _ctx():
    return Context(int(os.environ.get("QSVC_DEVICE", "0")))


# ------------------------------------------------------------------- tools

def t_split(argv, inverse=False):
    tool = "merge" if inverse else "split"
    p = _Parser(tool)
    p.opt("even_fn", "e", "even")
    p.opt("low_fn", "l", "low")
    p.opt("odd_fn", "o", "odd")
    p.opt("pictures", "p", 9, int)
    p.opt("pixels_in_x", "x", 352, int)
    p.opt("pixels_in_y", "y", 288, int)
    a = p.parse(argv)
    n = a.pictures // 2
    if not inverse:
        low = _read_frames(tool, a.low_fn, a.pixels_in_x, a.pixels_in_y, 2 * n + 1)
        even, odd = split(low)
        yuv.write_frames(a.even_fn, even)
        yuv.write_frames(a.odd_fn, odd)
    else:
        even = _read_frames(tool, a.even_fn, a.pixels_in_x, a.pixels_in_y, n + 1)
        odd = _read_frames(tool, a.odd_fn, a.pixels_in_x, a.pixels_in_y, n)
        yuv.write_frames(a.low_fn, merge(even, odd))
    return 0


def t_motion_estimate(argv):
    tool = "motion_estimate"
    p = _Parser(tool)
    p.opt("block_size", "b", 32, int)
    p.opt("border_size", "d", 0, int)
    p.opt("even_fn", "e", "even")
    p.opt("imotion_fn", "i", "imotion")  # opened but never read by the reference (:825)
    p.opt("motion_fn", "m", "motion")
    p.opt("odd_fn", "o", "odd")
    p.opt("pictures", "p", 9, int)
    p.opt("pixels_in_x", "x", 352, int)
    p.opt("pixels_in_y", "y", 288, int)
    p.opt("search_range", "s", 4, int)
    p.opt("subpixel_accuracy", "a", 0, int)
    a = p.parse(argv)
    if os.path.exists(a.motion_fn):
        return 1  # "reusing motion information": exit(1) without computing (:659-682)
    n = a.pictures // 2
    open(a.motion_fn, "wb").close()
    even = _read_frames(tool, a.even_fn, a.pixels_in_x, a.pixels_in_y, n + 1)
    odd = _read_frames(tool, a.odd_fn, a.pixels_in_x, a.pixels_in_y, n)
    with _ctx() as c:
        mv = c.motion_estimate(even, odd, a.pixels_in_x, a.pixels_in_y, a.block_size,
                               a.search_range, a.subpixel_accuracy, a.border_size)
    yuv.write_motion(a.motion_fn, mv)
    return 0


def t_decorrelate(argv, inverse=False):
    tool = "correlate" if inverse else "decorrelate"
    p = _Parser(tool)
    p.opt("block_overlaping", "v", 0, int)
    p.opt("block_size", "b", 16, int)
    p.opt("even_fn", "e", "even")
    p.opt("frame_types_fn", "f", "frame_types")
    p.opt("high_fn", "h", "high")
    p.opt("motion_in_fn", "i", "motion_in")
    if not inverse:
        p.opt("motion_out_fn", "t", "motion_out")
    p.opt("odd_fn", "o", "odd")
    p.opt("pictures", "p", 33, int)
    p.opt("pixels_in_x", "x", 352, int)
    p.opt("pixels_in_y", "y", 288, int)
    p.opt("search_range", "s", 4, int)
    p.opt("subpixel_accuracy", "a", 0, int)
    p.opt("always_B", "B", 0, int)
    a = p.parse(argv)
    X, Y, n = a.pixels_in_x, a.pixels_in_y, a.pictures // 2
    even = _read_frames(tool, a.even_fn, X, Y, n + 1)
    if not os.path.exists(a.motion_in_fn):
        _abort(tool, f'unable to read "{a.motion_in_fn}"')
    mv = yuv.read_motion(a.motion_in_fn, X, Y, a.block_size, n)
    with _ctx() as c:
        if not inverse:
            odd = _read_frames(tool, a.odd_fn, X, Y, n)
            high, types, mvo, pred = c.decorrelate(even, odd, mv, X, Y, a.block_size,
                                                   a.search_range, a.subpixel_accuracy,
                                                   a.block_overlaping, a.always_B,
                                                   want_prediction=True)
            yuv.write_frames(a.high_fn, high)
            open(a.frame_types_fn, "wb").write(types)
            yuv.write_motion(a.motion_out_fn, mvo)
        else:
            high = _read_frames(tool, a.high_fn, X, Y, n)
            if not os.path.exists(a.frame_types_fn):
                _abort(tool, f'unable to read "{a.frame_types_fn}"')
            types = open(a.frame_types_fn, "rb").read()[:n]
            odd, pred = c.correlate(even, high, mv, types, X, Y, a.block_size, a.search_range,
                                    a.subpixel_accuracy, a.block_overlaping, want_prediction=True)
            yuv.write_frames(a.odd_fn, odd)
    yuv.write_frames(f"prediction_{a.even_fn}", pred)  # decorrelate.cpp:454-472
    return 0


def t_update(argv, inverse=False):
    tool = "un_update" if inverse else "update"
    p = _Parser(tool)
    p.opt("block_size", "b", 16, int)
    p.opt("even_fn", "e", "even")
    p.opt("frame_types_fn", "f", "frame_types")
    p.opt("high_fn", "h", "high")
    p.opt("low_fn", "l", "low")
    p.opt("motion_fn", "m", "motion")
    p.opt("pictures", None, 33, int)  # long form only (update.cpp:209)
    p.opt("pixels_in_x", "x", 352, int)
    p.opt("pixels_in_y", "y", 288, int)
    p.opt("subpixel_accuracy", "a", 0, int)  # parsed, unused (update.cpp)
    p.opt("update_factor", "u", 0.25, float)
    a = p.parse(argv)
    X, Y, n = a.pixels_in_x, a.pixels_in_y, a.pictures // 2
    src, dst = (a.low_fn, a.even_fn) if inverse else (a.even_fn, a.low_fn)
    frames = _read_frames(tool, src, X, Y, n + 1)
    high = _read_frames(tool, a.high_fn, X, Y, n)
    for fn in (a.motion_fn, a.frame_types_fn):
        if not os.path.exists(fn):
            _abort(tool, f'unable to read "{fn}"')
    mv = yuv.read_motion(a.motion_fn, X, Y, a.block_size, n)
    types = open(a.frame_types_fn, "rb").read()[:n]
    with _ctx() as c:
        out = c.update(frames, high, mv, types, X, Y, a.block_size,
                       float(np.float32(a.update_factor)), inverse=inverse)
    yuv.write_frames(dst, out)
    return 0


# ----------------------------------------------------------------- drivers

def _driver_parser(tool, lists=False):
    p = _Parser(tool)
    s = str if lists else int
    p.opt("GOPs", None, 1, int)
    p.opt("TRLs", None, 4, int)
    p.opt("always_B", None, 0, int)
    p.opt("block_overlaping", None, 0, int)
    p.opt("block_size", None, None, s)
    p.opt("block_size_min", None, None, int)
    p.opt("border_size", None, 0, int)
    p.opt("pictures", None, None, int)
    p.opt("pixels_in_x", None, "352" if lists else 352, s)
    p.opt("pixels_in_y", None, "288" if lists else 288, s)
    p.opt("search_range", None, 4, int)
    p.opt("subpixel_accuracy", None, "0" if lists else 0, s)
    p.opt("temporal_subband", None, 1, int)
    p.opt("update_factor", None, None, float)
    return p


def _default_bs(X, Y):  # analyze.py:79-83 (evaluated on the default 352x288: always 32)
    return 32


def t_analyze(argv):
    """analyze.py:107-153, fused: frames stay resident in HBM across levels.  Writes every file
    the reference chain leaves except the prediction_even_<t> side files (written by
    decorrelate under GET_PREDICTION, read by nothing; `analyze_step` writes them)."""
    a = _driver_parser("analyze").parse(argv)
    X, Y = a.pixels_in_x, a.pixels_in_y
    bs = a.block_size if a.block_size is not None else _default_bs(X, Y)
    bs_min = a.block_size_min if a.block_size_min is not None else 32
    uf = 0.0 if a.update_factor is None else a.update_factor  # analyze.py default 0
    pictures = a.GOPs * gop_size(a.TRLs) + 1
    low0 = _read_frames("analyze", "low_0", X, Y, pictures)
    # motion_estimate exits 1 without computing when its motion file exists
    # (motion_estimate.cpp:659-682): analyze.py's chain stops at the first such level, after
    # that level's split, and turns the failure into exit -1
    stop_at = next((t for t in range(1, a.TRLs) if os.path.exists(f"motion_{t}")), None)
    with _ctx() as c:
        out = c.analyze(low0, X, Y, a.GOPs, a.TRLs, bs, a.search_range, a.subpixel_accuracy, uf,
                        a.always_B, a.block_overlaping, a.border_size, bs_min)
    low = low0
    for s in level_schedule(a.GOPs, a.TRLs, bs, a.search_range, bs_min):
        t = s["t"]
        even, odd = split(low)
        yuv.write_frames(f"even_{t}", even)
        yuv.write_frames(f"odd_{t}", odd)
        if t == stop_at:
            return 255
        yuv.write_motion(f"motion_{t}", out[f"motion_{t}"])
        yuv.write_motion(f"motion_filtered_{t}", out[f"motion_filtered_{t}"])
        open(f"frame_types_{t}", "wb").write(out[f"frame_types_{t}"])
        yuv.write_frames(f"high_{t}", out[f"high_{t}"])
        yuv.write_frames(f"low_{t}", out[f"low_{t}"])
        low = out[f"low_{t}"]
    return 0


def t_analyze_step(argv):
    """analyze_step.py:115-232: split, motion_estimate, decorrelate, update."""
    a = _driver_parser("analyze_step").parse(argv)
    X, Y, t = a.pixels_in_x, a.pixels_in_y, a.temporal_subband
    bs = a.block_size if a.block_size is not None else _default_bs(X, Y)
    uf = 0.0 if a.update_factor is None else a.update_factor
    pictures = a.pictures if a.pictures is not None else 9
    common = [f"--pictures={pictures}", f"--pixels_in_x={X}", f"--pixels_in_y={Y}"]
    rc = t_split([f"--even_fn=even_{t}", f"--low_fn=low_{t-1}", f"--odd_fn=odd_{t}"] + common)
    rc = rc or t_motion_estimate([f"--block_size={bs}", f"--border_size={a.border_size}",
                                  f"--even_fn=even_{t}", f"--imotion_fn=imotion_{t}",
                                  f"--motion_fn=motion_{t}", f"--odd_fn=odd_{t}",
                                  f"--search_range={a.search_range}",
                                  f"--subpixel_accuracy={a.subpixel_accuracy}"] + common)
    rc = rc or t_decorrelate([f"--block_overlaping={a.block_overlaping}", f"--block_size={bs}",
                              f"--even_fn=even_{t}", f"--frame_types_fn=frame_types_{t}",
                              f"--high_fn=high_{t}", f"--motion_in_fn=motion_{t}",
                              f"--motion_out_fn=motion_filtered_{t}", f"--odd_fn=odd_{t}",
                              f"--search_range={a.search_range}",
                              f"--subpixel_accuracy={a.subpixel_accuracy}",
                              f"--always_B={a.always_B}"] + common)
    rc = rc or t_update([f"--block_size={bs}", f"--even_fn=even_{t}",
                         f"--frame_types_fn=frame_types_{t}", f"--high_fn=high_{t}",
                         f"--low_fn=low_{t}", f"--motion_fn=motion_filtered_{t}",
                         f"--subpixel_accuracy={a.subpixel_accuracy}",
                         f"--update_factor={uf}"] + common)
    return -1 & 0xFF if rc else 0  # the reference drivers turn any failure into sys.exit(-1)


def _per_level(value: str, index: int, typ=int):
    parts = str(value).split(",")
    return typ(parts[index] if index < len(parts) else parts[-1])


def t_synthesize(argv):
    """synthesize.py:95-153.  block_size / pixels_in_x / pixels_in_y /
    subpixel_accuracy are comma lists indexed per temporal level
    (:127-133).  Constant lists run fused with the frames resident in HBM;
    lists that vary per level run step by step through
    files like the reference."""
    a = _driver_parser("synthesize", lists=True).parse(argv)
    T = a.TRLs
    bs_list = a.block_size if a.block_size is not None else "16,16,16,16"
    uf = 0.25 if a.update_factor is None else a.update_factor  # synthesize.py default 1/4
    geo = set()
    for t in range(1, T):
        geo.add((_per_level(bs_list, (T - 1) - t), _per_level(a.pixels_in_x, T - t),
                 _per_level(a.pixels_in_y, T - t), _per_level(a.subpixel_accuracy, T - t)))
    if len(geo) != 1:
        # per-level geometry (spatially scalable decoding, expand.py:150-209): chain the steps with
        # each level's own values exactly like synthesize.py:108-153 does (files in CWD)
        pictures_all = a.GOPs * gop_size(T) + 1
        for t in range(T - 1, 0, -1):
            pictures, sr = pictures_all, a.search_range
            for _ in range(1, t):
                sr = min(sr * 2, SEARCH_RANGE_MAX)
                pictures = (pictures + 1) // 2
            rc = t_synthesize_step([f"--block_overlaping={a.block_overlaping}",
                                    f"--block_size={_per_level(bs_list, (T - 1) - t)}",
                                    f"--pictures={pictures}",
                                    f"--pixels_in_x={_per_level(a.pixels_in_x, T - t)}",
                                    f"--pixels_in_y={_per_level(a.pixels_in_y, T - t)}",
                                    f"--search_range={sr}",
                                    f"--subpixel_accuracy={_per_level(a.subpixel_accuracy, T - t)}",
                                    f"--temporal_subband={t}", f"--update_factor={uf}"])
            if rc:
                return 255
        return 0
    bs, X, Y, acc = geo.pop()
    sub = {}
    for s in level_schedule(a.GOPs, T, bs, a.search_range, bs):
        t, n = s["t"], s["pairs"]
        sub[f"high_{t}"] = _read_frames("synthesize", f"high_{t}", X, Y, n)
        if not os.path.exists(f"motion_{t}") or not os.path.exists(f"frame_types_{t}"):
            _abort("synthesize", f'unable to read "motion_{t}"/"frame_types_{t}"')
        sub[f"motion_{t}"] = yuv.read_motion(f"motion_{t}", X, Y, bs, n)
        sub[f"frame_types_{t}"] = open(f"frame_types_{t}", "rb").read()[:n]
        if t == T - 1:
            sub[f"low_{t}"] = _read_frames("synthesize", f"low_{t}", X, Y, n + 1)
    with _ctx() as c:
        low0 = c.synthesize(sub, X, Y, a.GOPs, T, bs, a.search_range, acc, uf, a.block_overlaping)
    yuv.write_frames("low_0", low0)
    return 0


def t_synthesize_step(argv):
    """synthesize_step.py:84-143: un_update, correlate, merge."""
    a = _driver_parser("synthesize_step").parse(argv)
    X, Y, t = a.pixels_in_x, a.pixels_in_y, a.temporal_subband
    bs = a.block_size if a.block_size is not None else 16
    uf = 0.25 if a.update_factor is None else a.update_factor
    pictures = a.pictures if a.pictures is not None else 33
    common = [f"--pictures={pictures}", f"--pixels_in_x={X}", f"--pixels_in_y={Y}"]
    rc = t_update([f"--block_size={bs}", f"--even_fn=even_{t}", f"--frame_types_fn=frame_types_{t}",
                   f"--high_fn=high_{t}", f"--low_fn=low_{t}", f"--motion_fn=motion_{t}",
                   f"--subpixel_accuracy={a.subpixel_accuracy}", f"--update_factor={uf}"] + common,
                  inverse=True)
    rc = rc or t_decorrelate([f"--block_overlaping={a.block_overlaping}", f"--block_size={bs}",
                              f"--even_fn=even_{t}", f"--frame_types_fn=frame_types_{t}",
                              f"--high_fn=high_{t}", f"--motion_in_fn=motion_{t}",
                              f"--odd_fn=odd_{t}", f"--search_range={a.search_range}",
                              f"--subpixel_accuracy={a.subpixel_accuracy}"] + common, inverse=True)
    rc = rc or t_split([f"--even={f'even_{t}'}", f"--low=low_{t-1}", f"--odd=odd_{t}"] + common,
                       inverse=True)
    return 255 if rc else 0


def _read_fields(tool, fn, by, bx):
    """Every whole field of a motion file, (n, 4, by, bx) int16."""
    if not os.path.exists(fn):
        _abort(tool, f'unable to read "{fn}"')
    data = np.fromfile(fn, dtype="<i2")
    fsz = 4 * by * bx
    n = data.size // fsz if fsz else 0
    return data[: n * fsz].reshape(n, 4, by, bx)


def t_bidirectional(argv, inverse=False):
    """bidirectional_motion_decorrelate.cpp:64-215 (flags :82-91)."""
    tool = "bidirectional_motion_correlate" if inverse else "bidirectional_motion_decorrelate"
    p = _Parser(tool)
    p.opt("blocks_in_x", "x", 11, int)
    p.opt("blocks_in_y", "y", 9, int)
    p.opt("fields", "f", 1, int)
    p.opt("input_fn", "i", "/dev/zero")
    p.opt("output_fn", "o", "/dev/zero")
    a = p.parse(argv)
    if a.input_fn == "/dev/zero":
        fields = np.zeros((a.fields, 4, a.blocks_in_y, a.blocks_in_x), np.int16)
    else:
        fields = _read_fields(tool, a.input_fn, a.blocks_in_y, a.blocks_in_x)[: a.fields]
    with _ctx() as c:
        out = c.bidirectional_motion_decorrelate(fields, inverse=inverse)
    if a.output_fn != "/dev/zero":
        yuv.write_motion(a.output_fn, out)
    return 0


def t_interlevel(argv, inverse=False):
    """interlevel_motion_decorrelate.cpp:77-297 (flags :143-152): the loop reads one reference
    field per iteration and up to two predicted (or residue) fields with it."""
    tool = "interlevel_motion_correlate" if inverse else "interlevel_motion_decorrelate"
    p = _Parser(tool)
    p.opt("blocks_in_x", "x", 11, int)
    p.opt("blocks_in_y", "y", 9, int)
    p.opt("fields_in_predicted", "f", 1, int)
    p.opt("predicted_fn", "p", "/dev/zero")
    p.opt("reference_fn", "r", "/dev/zero")
    p.opt("residue_fn", "e", "/dev/zero")
    a = p.parse(argv)
    src_fn, dst_fn = (a.residue_fn, a.predicted_fn) if inverse else (a.predicted_fn, a.residue_fn)
    if src_fn == "/dev/zero":
        fields = np.zeros((2 * a.fields_in_predicted, 4, a.blocks_in_y, a.blocks_in_x), np.int16)
    else:
        fields = _read_fields(tool, src_fn, a.blocks_in_y, a.blocks_in_x)[: 2 * a.fields_in_predicted]
    ref = None  # a missing reference file reads as zeros (:229-238)
    if a.reference_fn != "/dev/zero" and os.path.exists(a.reference_fn):
        ref = _read_fields(tool, a.reference_fn, a.blocks_in_y, a.blocks_in_x)[: a.fields_in_predicted]
    with _ctx() as c:
        out = c.interlevel_motion_decorrelate(fields, ref, inverse=inverse)
    if dst_fn != "/dev/zero":
        yuv.write_motion(dst_fn, out)
    return 0


def t_demux(argv):
    """`demux <record_bytes> <offset> <length> < in > out`: the byte range [offset, offset +
    length) of every record of the input, the external filter the reference's texture and
    motion coders pipe frame files through to pull one component out of every picture
    (texture_compress_fb_j2k.py:155-163,255-257; motion_compress_j2k.py:103).  Host-side byte
    slicing: its source is not part of the reference tree (contract restated from the call sites)."""
    if len(argv) != 3:
        sys.stderr.write("usage: demux <record_bytes> <offset> <length> < input > output\n")
        return 1
    rec, off, ln = (int(v) for v in argv)
    data = np.frombuffer(sys.stdin.buffer.read(), np.uint8)
    n = data.size // rec
    sys.stdout.buffer.write(np.ascontiguousarray(data[: n * rec].reshape(n, rec)[:, off:off + ln]).tobytes())
    tail = data[n * rec:]  # a trailing partial record contributes what it holds of the range
    if tail.size > off:
        sys.stdout.buffer.write(tail[off:off + ln].tobytes())
    return 0


def t_snr(argv):
    """`snr --type=uchar --peak=255 --file_A=a --file_B=b --block_size=<bytes>`: the external
    distortion meter psnr.py:78-90 greps for its `PSNR ... dB` line (third blank-separated
    field).  Sums of squared differences per block on the GPU (qsvc_sse), PSNR on the host.
    Its source is not part of the reference tree: only the contract of the call site is kept."""
    p = _Parser("snr")
    p.opt("type", "t", "uchar")
    p.opt("peak", "p", 255.0, float)
    p.opt("file_A", "a", "")
    p.opt("file_B", "b", "")
    p.opt("block_size", "s", 0, int)
    a = p.parse(argv)
    if a.type != "uchar":
        sys.stderr.write("snr: only --type=uchar is supported\n")
        return 1
    for fn in (a.file_A, a.file_B):
        if not os.path.exists(fn):
            _abort("snr", f'unable to read "{fn}"')
    A, B = np.fromfile(a.file_A, np.uint8), np.fromfile(a.file_B, np.uint8)
    n = min(A.size, B.size)
    block = a.block_size if a.block_size > 0 else n
    with _ctx() as c:
        sse = c.sse(A[:n], B[:n], block)
    tot = float(sse.sum())
    samples = float(len(sse) * block)
    for k, v in enumerate(sse):
        mse = float(v) / block
        print(f"block {k}: MSE = {mse:.6f} RMSE = {mse ** 0.5:.6f}")
    mse = tot / samples if samples else 0.0
    psnr = float("inf") if mse == 0 else 10.0 * np.log10(a.peak * a.peak / mse)
    print(f"MSE = {mse:.6f}")
    print(f"PSNR = {psnr:.6f} dB")
    return 0


TOOLS = {
    "demux": t_demux,
    "snr": t_snr,
    "bidirectional_motion_decorrelate": t_bidirectional,
    "bidirectional_motion_correlate": lambda argv: t_bidirectional(argv, inverse=True),
    "interlevel_motion_decorrelate": t_interlevel,
    "interlevel_motion_correlate": lambda argv: t_interlevel(argv, inverse=True),
    "split": t_split,
    "merge": lambda argv: t_split(argv, inverse=True),
    "motion_estimate": t_motion_estimate,
    "decorrelate": t_decorrelate,
    "correlate": lambda argv: t_decorrelate(argv, inverse=True),
    "update": t_update,
    "un_update": lambda argv: t_update(argv, inverse=True),
    "analyze": t_analyze,
    "analyze_step": t_analyze_step,
    "synthesize": t_synthesize,
    "synthesize_step": t_synthesize_step,
}


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in TOOLS:
        sys.stderr.write("usage: mctf <" + "|".join(TOOLS) + "> [--flag=value ...]\n")
        return 1
    from .._lib import QsvcError
    try:
        return int(TOOLS[argv[0]](argv[1:]) or 0)
    except QsvcError as e:
        _error(f"{argv[0]}: {e}\n")
        return 2
