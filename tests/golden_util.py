"""Loads the committed reference fixtures (made by oracle/make_golden.py)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    X, Y, GOPs, TRLs, bs, sr, a, always_B = (int(v) for v in z["params"])
    g = {k: z[k] for k in z.files}
    g.update(X=X, Y=Y, GOPs=GOPs, TRLs=TRLs, bs=bs, sr=sr, a=a, always_B=always_B,
             uf=float(z["update_factor"][0]),
             ov=int(z["block_overlaping"][0]) if "block_overlaping" in z.files else 0)
    return g


def schedule(g):
    """(t, search_range) per level: sr doubles up to 128 (analyze.py:144-147)."""
    out, sr = [], g["sr"]
    for t in range(1, g["TRLs"]):
        out.append((t, sr))
        sr = min(2 * sr, 128)
    return out
