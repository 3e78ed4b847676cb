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


if __name__ == "__main__":
    assert run_ref.build(), "oracle/_ref is missing and /root/reference is not available"
    only = sys.argv[1:]
    for k, v in CASES.items():
        if not only or k in only:
            make(k, *v)
