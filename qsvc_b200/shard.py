"""GOP sharding of one sequence across GPUs: no data-path collective.

GPU g of N takes GOPs [g0, g1), i.e. input frames [g0*G, g1*G] inclusive (the
boundary frame is read by both neighbours), runs every temporal level locally
and returns its per-level outputs; the host concatenates them in GOP order into
the reference's file layout, dropping the duplicated boundary low frame.

Exact when update_factor == 0 and the picture height/width are multiples of the
block size.  Otherwise a neighbour exchange would be needed (update of the
boundary frame; chained prediction tail rows) and the functions
below refuse unless `allow_inexact=True`.
"""
from __future__ import annotations

import numpy as np

from .mctf import gop_size, level_schedule


def partition(GOPs: int, world: int):
    """Contiguous, balanced GOP ranges per rank; ranks beyond the work get (g, g)."""
    out, g = [], 0
    for r in range(world):
        n = GOPs // world + (1 if r < GOPs % world else 0)
        out.append((g, g + n))
        g += n
    return out


def shard_frames(low0: np.ndarray, TRLs: int, g0: int, g1: int) -> np.ndarray:
    G = gop_size(TRLs)
    return low0[g0 * G : g1 * G + 1]


def check_exact(X, Y, block_size, update_factor, world, allow_inexact=False):
    if world <= 1 or allow_inexact:
        return
    if update_factor != 0:
        raise ValueError("GOP sharding with update_factor != 0 needs the boundary-frame "
                         "exchange of SURVEY.md 8e(1); run on one GPU or pass allow_inexact")
    if Y % block_size or X % block_size:
        raise ValueError("GOP sharding with a picture size that is not a multiple of the block "
                         "size needs the chained tail rows of SURVEY.md A.2.6; run on one GPU "
                         "or pass allow_inexact")


def analyze_shard(ctx, low0, X, Y, GOPs, TRLs, rank, world, block_size=32, search_range=4,
                  subpixel_accuracy=0, update_factor=0.0, always_B=0, block_size_min=32,
                  allow_inexact=False, analyze_fn=None):
    """Runs this rank's GOP range.  `analyze_fn(frames, n_gops, first_global)` defaults
    to ctx.analyze (the CUDA path); tests may inject another callable."""
    check_exact(X, Y, block_size, update_factor, world, allow_inexact)
    g0, g1 = partition(GOPs, world)[rank]
    if g1 == g0:
        return None
    frames = shard_frames(low0, TRLs, g0, g1)
    if analyze_fn is None:
        def analyze_fn(fr, n_gops, first_global):
            return ctx.analyze(fr, X, Y, n_gops, TRLs, block_size, search_range,
                               subpixel_accuracy, update_factor, always_B,
                               block_size_min=block_size_min, first_global=first_global)
    return analyze_fn(frames, g1 - g0, g0 == 0)


def gather(parts, TRLs: int):
    """Concatenates the per-rank outputs (rank order = GOP order; None = idle rank)."""
    parts = [p for p in parts if p is not None]
    out = {}
    for t in range(1, TRLs):
        for name in ("high", "motion", "motion_filtered"):
            out[f"{name}_{t}"] = np.concatenate([p[f"{name}_{t}"] for p in parts], axis=0)
        out[f"frame_types_{t}"] = b"".join(bytes(p[f"frame_types_{t}"]) for p in parts)
        lows = [p[f"low_{t}"] if i == 0 else p[f"low_{t}"][1:] for i, p in enumerate(parts)]
        out[f"low_{t}"] = np.concatenate(lows, axis=0)
    return out


def analyze_distributed(ctx, low0, X, Y, GOPs, TRLs, **kw):
    """torch.distributed front-end: every rank analyses its shard, rank 0 gathers
    (gather_object: host-side gather, no collective on the data path)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    local = analyze_shard(ctx, low0, X, Y, GOPs, TRLs, rank, world, **kw)
    parts = [None] * world if rank == 0 else None
    dist.gather_object(local, parts, dst=0)
    return gather(parts, TRLs) if rank == 0 else None
