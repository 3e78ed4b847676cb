"""Loads the committed reference fixtures (made by oracle/make_golden.py)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# level_lists.npz has its own layout (per-level geometry, synthesis only): see load_level_lists
NAMES = sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
               if n != "level_lists")


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


def load_level_lists():
    """Synthesis with geometry lists that vary per temporal level (oracle/make_golden.py
    make_level_lists): {t: (X, Y, a)}, the sub-band files and the reference's low_0."""
    z = np.load(os.path.join(GOLDEN_DIR, "level_lists.npz"))
    GOPs, TRLs, bs, sr = (int(v) for v in z["params"])
    g = {k: z[k] for k in z.files}
    g.update(GOPs=GOPs, TRLs=TRLs, bs=bs, sr=sr, uf=float(z["update_factor"][0]),
             geo={int(r[0]): (int(r[1]), int(r[2]), int(r[3])) for r in z["geometry"]})
    return g
