"""Generates tests/golden/*.npz from the UNMODIFIED reference tools in oracle/_ref/.

Run in the build container (needs oracle/_ref, i.e. /root/reference):
    python oracle/make_golden.py
Each fixture holds the input clip and every file the reference chain
(analyze.py + synthesize.py restated in oracle/run_ref.py) produced for it.
"""
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import run_ref  # noqa: E402
from qsvc_b200 import yuv  # noqa: E402

# name: X, Y, GOPs, TRLs, bs, sr, a, uf, always_B, flat_every, seed[, block_overlaping]
CASES = {
    "ib_types_a0":   (64, 48, 2, 3, 16, 4, 0, 0.25, 0, 2, 21),   # IBIB frame types, update 1/4
    "quarter_pel":   (64, 48, 1, 4, 16, 8, 2, 0.3, 0, 0, 22),    # border pollution, size-field reads
    "ragged_height": (64, 40, 2, 4, 16, 4, 1, 0.0, 0, 0, 23),    # Y % bs != 0: chained tail rows
    "odd_pyramid":   (64, 60, 2, 3, 8, 32, 0, 0.0, 1, 0, 24),    # non-invertible pyramid descent
    "obmc_half_pel": (64, 48, 1, 3, 16, 4, 1, 0.25, 0, 0, 25, 2),  # block_overlaping=2: per-block DWT + scatter
}


def make(name, X, Y, GOPs, TRLs, bs, sr, a, uf, always_B, flat, seed, ov=0):
    frames = GOPs * 2 ** (TRLs - 1) + 1
    clip = yuv.synthetic_clip(X, Y, frames, seed, max_shift=min(24, 3 * sr), flat_every=flat)
    d = tempfile.mkdtemp(prefix="golden_")
    try:
        yuv.write_frames(os.path.join(d, "low_0"), clip)
        sched = run_ref.analyze(d, X, Y, GOPs, TRLs, bs, sr, a, uf, always_B, block_overlaping=ov, block_size_min=bs)
        out = {"low_0": clip,
               "params": np.array([X, Y, GOPs, TRLs, bs, sr, a, always_B], np.int64),
               "update_factor": np.array([uf], np.float64),
               "block_overlaping": np.array([ov], np.int64)}
        for s in sched:
            t, n = s["t"], s["pictures"] // 2
            out[f"motion_{t}"] = yuv.read_motion(os.path.join(d, f"motion_{t}"), X, Y, bs, n)
            out[f"motion_filtered_{t}"] = yuv.read_motion(os.path.join(d, f"motion_filtered_{t}"), X, Y, bs, n)
            out[f"high_{t}"] = yuv.read_frames(os.path.join(d, f"high_{t}"), X, Y)
            out[f"low_{t}"] = yuv.read_frames(os.path.join(d, f"low_{t}"), X, Y)
            out[f"prediction_even_{t}"] = yuv.read_frames(os.path.join(d, f"prediction_even_{t}"), X, Y)
            out[f"frame_types_{t}"] = np.frombuffer(open(os.path.join(d, f"frame_types_{t}"), "rb").read(), np.uint8)
        # synthesis: the decoder sees motion_filtered_t under the name motion_t
        for s in sched:
            shutil.copy(os.path.join(d, f"motion_filtered_{s['t']}"), os.path.join(d, f"motion_{s['t']}"))
        for t in range(1, TRLs):
            os.remove(os.path.join(d, f"even_{t}"))
            os.remove(os.path.join(d, f"odd_{t}"))
        for t in range(0, TRLs - 1):
            os.remove(os.path.join(d, f"low_{t}"))
        run_ref.synthesize(d, X, Y, GOPs, TRLs, bs, sr, a, uf, block_overlaping=ov)
        for t in range(1, TRLs):
            out[f"syn_even_{t}"] = yuv.read_frames(os.path.join(d, f"even_{t}"), X, Y)
            out[f"syn_odd_{t}"] = yuv.read_frames(os.path.join(d, f"odd_{t}"), X, Y)
        out["syn_low_0"] = yuv.read_frames(os.path.join(d, "low_0"), X, Y)
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path), "bytes;", "types:",
              [bytes(out[f"frame_types_{t}"]).decode() for t in range(1, TRLs)])
    finally:
        shutil.rmtree(d)


def make_level_lists():
    """synthesize.py:127-133 with geometry lists that vary per temporal level (SURVEY.md 8f
    rank 4, expand.py:150-209): the reference runs one synthesize_step per level with that
    level's own --pixels_in_x / --pixels_in_y / --block_size / --subpixel_accuracy, and the
    files handed from level to level are simply re-read with the next level's geometry.  The
    fixture keeps the frame bytes and the block count equal across levels (64x48 <-> 48x64,
    block 16 -> 12 blocks either way) so that every file has a consistent length; inputs are
    the analysis outputs of the 'quarter_pel' geometry plus level-specific vectors."""
    rng = np.random.default_rng(77)
    GOPs, TRLs, bs, sr, uf = 1, 4, 16, 4, 0.25
    geo = {3: (64, 48, 1), 2: (48, 64, 0), 1: (64, 48, 2)}  # per temporal level: X, Y, subpixel_accuracy
    fb = 64 * 48 * 3 // 2
    d = tempfile.mkdtemp(prefix="golden_")
    try:
        out = {"params": np.array([GOPs, TRLs, bs, sr], np.int64), "update_factor": np.array([uf], np.float64),
               "geometry": np.array([[t, *geo[t]] for t in (3, 2, 1)], np.int64)}
        pictures = {1: 9, 2: 5, 3: 3}
        clip = yuv.synthetic_clip(64, 48, 9, 31, max_shift=8)
        out["low_3"] = clip[:2]
        yuv.write_frames(os.path.join(d, "low_3"), out["low_3"])
        for t in (1, 2, 3):
            n = pictures[t] // 2
            X, Y, a = geo[t]
            high = (128 + rng.integers(-20, 21, size=(n, fb))).astype(np.uint8)
            mv = rng.integers(-(sr << a), (sr << a) + 1, size=(n, 4, Y // bs, X // bs)).astype(np.int16)
            types = np.frombuffer(b"B" * n, np.uint8).copy()
            if t == 1:
                types[1] = ord("I")
            out[f"high_{t}"], out[f"motion_{t}"], out[f"frame_types_{t}"] = high, mv, types
            yuv.write_frames(os.path.join(d, f"high_{t}"), high)
            yuv.write_motion(os.path.join(d, f"motion_{t}"), mv)
            open(os.path.join(d, f"frame_types_{t}"), "wb").write(types.tobytes())
        search = {1: sr, 2: 2 * sr, 3: 4 * sr}
        for t in (3, 2, 1):
            X, Y, a = geo[t]
            run_ref.synthesize_step(d, t, pictures[t], X, Y, bs, search[t], a, uf)
        out["syn_low_0"] = yuv.read_frames(os.path.join(d, "low_0"), 64, 48)
        path = os.path.join(ROOT, "tests", "golden", "level_lists.npz")
        np.savez_compressed(path, **out)
        print("level_lists", os.path.getsize(path), "bytes")
    finally:
        shutil.rmtree(d)


if __name__ == "__main__":
    assert run_ref.build(), "oracle/_ref is missing and /root/reference is not available"
    only = sys.argv[1:]
    for k, v in CASES.items():
        if not only or k in only:
            make(k, *v)
    if not only or "level_lists" in only:
        make_level_lists()
