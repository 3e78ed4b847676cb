"""Oracle jobs of tests/test_gpu_parity_large.py (module-level functions: they run in spawned
worker processes).  Test infrastructure only."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qsvc_b200.mctf import level_schedule  # noqa: E402


def _job_chain(clip, X, Y, bs, sr, a, TRLs, uf, always_B, want_pred):
    """analyze.py's level loop on the oracle; returns every level's outputs."""
    from oracle import oracle as orc
    out, low = {}, clip
    for s in level_schedule((clip.shape[0] - 1) >> (TRLs - 1), TRLs, bs, sr, block_size_min=bs):
        t, r = s["t"], s["search_range"]
        even, odd = low[0::2], low[1::2]
        mv = orc.motion_estimate(even, odd, X, Y, bs, r, a)
        high, types, mvf, pred, rc = orc.decorrelate(even, odd, mv, X, Y, bs, r, a, always_B=always_B)
        assert rc == 0
        low = orc.update(even, high, mvf, types, X, Y, bs, uf)
        out.update({f"motion_{t}": mv, f"high_{t}": high, f"frame_types_{t}": types,
                    f"motion_filtered_{t}": mvf, f"low_{t}": low})
        if want_pred:
            out[f"prediction_{t}"] = pred
    return out


def _job_pair(clip, X, Y, bs, sr, a):
    """One pair through motion_estimate, decorrelate and correlate."""
    from oracle import oracle as orc
    even, odd = clip[0::2], clip[1::2]
    mv = orc.motion_estimate(even, odd, X, Y, bs, sr, a)
    high, types, mvf, pred, rc = orc.decorrelate(even, odd, mv, X, Y, bs, sr, a, always_B=1)
    assert rc == 0
    rec, _ = orc.correlate(even, high, mvf, types, X, Y, bs, sr, a)
    return dict(motion=mv, high=high, types=types, motion_filtered=mvf, prediction=pred, odd=rec)


def _job_synth(sub, X, Y, GOPs, TRLs, bs, sr, a, uf):
    from oracle import oracle as orc
    low = sub[f"low_{TRLs-1}"]
    for s in reversed(level_schedule(GOPs, TRLs, bs, sr, block_size_min=bs)):
        t = s["t"]
        even = orc.update(low, sub[f"high_{t}"], sub[f"motion_{t}"], sub[f"frame_types_{t}"], X, Y, bs, uf,
                          inverse=True)
        odd, _ = orc.correlate(even, sub[f"high_{t}"], sub[f"motion_{t}"], sub[f"frame_types_{t}"], X, Y, bs,
                               s["search_range"], a)
        low = np.empty((2 * odd.shape[0] + 1, even.shape[1]), np.uint8)
        low[0::2], low[1::2] = even, odd
    return low


