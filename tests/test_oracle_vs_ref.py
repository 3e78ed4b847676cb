"""CPU: the oracle against the reference tools run live (oracle/_ref, compiled from
/root/reference by oracle/Makefile).  Skipped where the binaries are absent."""
import os
import tempfile

import numpy as np
import pytest

from oracle import oracle as orc, run_ref
from qsvc_b200 import yuv

pytestmark = pytest.mark.skipif(not run_ref.available(), reason="oracle/_ref not built")

CASES = [
    # X, Y, GOPs, TRLs, bs, sr, a, uf, flat
    (96, 64, 1, 4, 16, 4, 0, 0.25, 2),
    (96, 64, 1, 3, 16, 8, 2, 0.3, 0),
    (96, 72, 1, 4, 16, 4, 1, 0.0, 0),
    (64, 60, 1, 3, 8, 64, 0, 0.0, 0),
    (80, 48, 1, 3, 16, 16, 2, 0.25, 0),
    (176, 144, 2, 3, 32, 4, 1, 0.5, 3),     # 32 x 32 blocks, X % bs != 0, two GOPs, I frames
    (128, 72, 2, 4, 16, 4, 2, 0.25, 0),     # Y % bs != 0 at quarter-pel: tail chain across GOPs
    (96, 64, 1, 3, 16, 6, 1, 0.25, 0),      # search range that is not a power of two
]


@pytest.mark.parametrize("X,Y,GOPs,TRLs,bs,sr,a,uf,flat", CASES)
def test_oracle_matches_live_reference(X, Y, GOPs, TRLs, bs, sr, a, uf, flat):
    frames = GOPs * 2 ** (TRLs - 1) + 1
    clip = yuv.synthetic_clip(X, Y, frames, 31, max_shift=min(24, 3 * sr), flat_every=flat)
    with tempfile.TemporaryDirectory() as d:
        yuv.write_frames(os.path.join(d, "low_0"), clip)
        sched = run_ref.analyze(d, X, Y, GOPs, TRLs, bs, sr, a, uf, 0, block_size_min=bs)
        res = orc.analyze(clip, X, Y, TRLs, bs, sr, a, uf, block_size_min=bs)
        for s in sched:
            t, n = s["t"], s["pictures"] // 2
            assert np.array_equal(yuv.read_motion(os.path.join(d, f"motion_{t}"), X, Y, bs, n), res[f"motion_{t}"])
            assert np.array_equal(yuv.read_frames(os.path.join(d, f"high_{t}"), X, Y), res[f"high_{t}"])
            assert np.array_equal(yuv.read_frames(os.path.join(d, f"low_{t}"), X, Y), res[f"low_{t}"])
            assert open(os.path.join(d, f"frame_types_{t}"), "rb").read() == res[f"frame_types_{t}"]


OBMC_CASES = [
    # X, Y, TRLs, bs, sr, a, uf, block_overlaping
    (96, 64, 3, 16, 4, 0, 0.25, 2),
    (64, 48, 3, 16, 4, 1, 0.0, 2),
    (64, 64, 3, 16, 4, 2, 0.25, 4),
    (64, 48, 2, 16, 4, 1, 0.0, 3),
    (96, 72, 4, 16, 4, 0, 0.25, 2),    # ragged pictures: uncovered areas carry the previous pair's leftovers
    (88, 64, 3, 16, 4, 1, 0.25, 2),    #   through the picture synthesis (A.2.6 with A.2.2)
    (88, 72, 3, 16, 4, 2, 0.0, 4),
    (104, 56, 4, 16, 4, 1, 0.0, 3),
]


@pytest.mark.parametrize("X,Y,TRLs,bs,sr,a,uf,ov", OBMC_CASES)
def test_oracle_overlapped_prediction_matches_live_reference(X, Y, TRLs, bs, sr, a, uf, ov):
    """--block_overlaping > 0 (decorrelate.cpp:84-88, 99-172) through the whole analysis chain."""
    frames = 2 ** (TRLs - 1) + 1
    clip = yuv.synthetic_clip(X, Y, frames, 33, max_shift=12)
    with tempfile.TemporaryDirectory() as d:
        yuv.write_frames(os.path.join(d, "low_0"), clip)
        sched = run_ref.analyze(d, X, Y, 1, TRLs, bs, sr, a, uf, 0, block_overlaping=ov, block_size_min=bs)
        res = orc.analyze(clip, X, Y, TRLs, bs, sr, a, uf, block_overlaping=ov, block_size_min=bs)
        for s in sched:
            t = s["t"]
            assert np.array_equal(yuv.read_frames(os.path.join(d, f"high_{t}"), X, Y), res[f"high_{t}"])
            assert np.array_equal(yuv.read_frames(os.path.join(d, f"low_{t}"), X, Y), res[f"low_{t}"])
            assert open(os.path.join(d, f"frame_types_{t}"), "rb").read() == res[f"frame_types_{t}"]
