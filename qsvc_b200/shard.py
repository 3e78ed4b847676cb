"""GOP sharding of one sequence across GPUs: no data-path collective.

GPU g of N takes GOPs [g0, g1), i.e. input frames [g0*G, g1*G] inclusive (the
boundary frame is read by both neighbours), runs every temporal level locally
and returns its per-level outputs; the host concatenates them in GOP order into
the reference's file layout, dropping the duplicated boundary frame.

What couples neighbouring shards, and how it is handled:
  * picture height not a multiple of the block size (1080 lines, block 16): the
    prediction buffer of decorrelate / correlate carries its uncovered rows from
    pair to pair.  Every shard receives that state from its left
    neighbour and passes its own to the right, once per temporal level (a point-to-
    point message of 3 * rows * (X << a) bytes; `TailRelay` over torch.distributed,
    `LocalTailRelay` when the shards run one after the other on one GPU);
  * update_factor != 0: the boundary frame receives the left shard's NEXT update and
    then the right shard's PREV update.  The left shard passes the
    int16 planes of that frame to the right between the two passes and gets the finished
    frame back, once per temporal level (`BoundaryRelay` over torch.distributed,
    `ThreadBoundaryRelay` for shards driven by threads of one process).  The shards must
    run concurrently;
  * X % block_size != 0 with Y % block_size != 0: the byte-plane path does not apply
    and the literal path has no exchange hook: refused.
"""
from __future__ import annotations

import numpy as np

from .mctf import gop_size, level_schedule


def _dist_ready():
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized()
    except ImportError:
        return False


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


def needs_tail_exchange(Y, block_size, world):
    return world > 1 and Y % block_size != 0


def needs_boundary_exchange(update_factor, world):
    return world > 1 and update_factor != 0


def check_exact(X, Y, block_size, update_factor, world, allow_inexact=False, boundary_relay=None):
    if world <= 1 or allow_inexact:
        return
    if update_factor != 0 and boundary_relay is None:
        raise ValueError("GOP sharding with update_factor != 0 needs the boundary-frame "
                         "exchange of SURVEY.md 8e(1): pass a boundary relay (shards must run "
                         "concurrently), run on one GPU, or pass allow_inexact")
    if Y % block_size and (X % block_size or X % 8):
        raise ValueError("GOP sharding with uncovered rows AND columns has no tail exchange "
                         "(literal decorrelate path); run on one GPU or pass allow_inexact")


class LocalTailRelay:
    """Tail-state hand-over between shards that run one after the other in this
    process (one GPU working through a long sequence GOP range by GOP range)."""

    def __init__(self):
        self.state = {}
        self.first = True

    def next_shard(self):
        self.first = False

    def __call__(self, level, synthesis, phase, state):
        key = (level, synthesis)
        if phase == 0:
            if self.first or key not in self.state:
                return False
            state[:] = self.state[key]
            return True
        self.state[key] = state.copy()
        return False


class _DeviceBytes:
    """A device buffer of the context seen through __cuda_array_interface__ (torch.as_tensor)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class TailRelay:
    """Tail-state hand-over over torch.distributed point-to-point messages: receive from
    the nearest rank on the left that has work, send to the nearest on the right.  With the
    NCCL backend the state stays on the devices (`device` = True: the library hands out a
    device buffer, qsvc_set_tail_exchange_device); with gloo it travels through host arrays."""

    def __init__(self, rank, ranges):
        import torch.distributed as dist
        active = [r for r, (a, b) in enumerate(ranges) if b > a]
        i = active.index(rank)
        self.left = active[i - 1] if i > 0 else None
        self.right = active[i + 1] if i + 1 < len(active) else None
        self.device = dist.get_backend() == "nccl"

    def __call__(self, level, synthesis, phase, state):
        import torch
        import torch.distributed as dist
        if self.device:
            ptr, nbytes = state
            if phase == 0:
                if self.left is None:
                    return False
                t = torch.as_tensor(_DeviceBytes(ptr, nbytes), device="cuda")
                dist.recv(t, src=self.left)
                torch.cuda.current_stream().synchronize()
                return True
            if self.right is not None:
                t = torch.as_tensor(_DeviceBytes(ptr, nbytes), device="cuda")
                dist.send(t, dst=self.right)
                torch.cuda.current_stream().synchronize()
            return False
        if phase == 0:
            if self.left is None:
                return False
            t = torch.empty(state.shape[0], dtype=torch.uint8)
            dist.recv(t, src=self.left)
            state[:] = t.numpy()
            return True
        if self.right is not None:
            dist.send(torch.from_numpy(state.copy()), dst=self.right)
        return False


class BoundaryRelay:
    """Boundary-frame hand-over (update_factor != 0) over torch.distributed point-to-point
    messages between neighbouring ranks that have work."""

    def __init__(self, rank, ranges):
        import torch.distributed as dist
        active = [r for r, (a, b) in enumerate(ranges) if b > a]
        i = active.index(rank)
        self.left = active[i - 1] if i > 0 else None
        self.right = active[i + 1] if i + 1 < len(active) else None
        self.cuda = dist.get_backend() == "nccl"

    def _send(self, data, dst):
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(data.copy())
        dist.send(t.cuda() if self.cuda else t, dst=dst)

    def _recv(self, data, src):
        import torch
        import torch.distributed as dist
        t = torch.empty(data.shape[0], dtype=torch.uint8, device="cuda" if self.cuda else "cpu")
        dist.recv(t, src=src)
        data[:] = t.cpu().numpy()

    def __call__(self, level, inverse, phase, data):
        if phase == 0:
            if self.left is None:
                return False
            self._recv(data, self.left)
            return True
        if phase == 1:
            if self.right is None:
                return False
            self._send(data, self.right)
            return True
        if phase == 2:
            if self.left is not None:
                self._send(data, self.left)
            return False
        self._recv(data, self.right)
        return True


class ThreadBoundaryRelay:
    """The same hand-over between shards driven by threads of one process (one context per
    thread, any number of GPUs): `ThreadBoundaryRelay.make(n)` returns one relay per shard."""

    def __init__(self, to_left, from_left, to_right, from_right):
        self.to_left, self.from_left, self.to_right, self.from_right = to_left, from_left, to_right, from_right

    @staticmethod
    def make(n):
        import queue
        right = [queue.Queue() for _ in range(n - 1)]  # right[i]: shard i -> shard i + 1
        left = [queue.Queue() for _ in range(n - 1)]   # left[i]:  shard i + 1 -> shard i
        return [ThreadBoundaryRelay(left[i - 1] if i > 0 else None, right[i - 1] if i > 0 else None,
                                    right[i] if i < n - 1 else None, left[i] if i < n - 1 else None)
                for i in range(n)]

    def __call__(self, level, inverse, phase, data):
        if phase == 0:
            if self.from_left is None:
                return False
            data[:] = self.from_left.get(timeout=600)
            return True
        if phase == 1:
            if self.to_right is None:
                return False
            self.to_right.put(data.copy())
            return True
        if phase == 2:
            if self.to_left is not None:
                self.to_left.put(data.copy())
            return False
        data[:] = self.from_right.get(timeout=600)
        return True


def analyze_shard(ctx, low0, X, Y, GOPs, TRLs, rank, world, block_size=32, search_range=4,
                  subpixel_accuracy=0, update_factor=0.0, always_B=0, block_size_min=32,
                  allow_inexact=False, analyze_fn=None, relay=None, boundary_relay=None, **analyze_kw):
    """Runs this rank's GOP range.  `analyze_fn(frames, n_gops, first_global)` defaults
    to ctx.analyze (the CUDA path); tests may inject another callable.  `relay` is the
    tail-state hand-over (default: TailRelay when the geometry needs one)."""
    ranges = partition(GOPs, world)
    if (boundary_relay is None and needs_boundary_exchange(update_factor, world) and ctx is not None
            and analyze_fn is None and _dist_ready()):
        boundary_relay = BoundaryRelay(rank, ranges)
    check_exact(X, Y, block_size, update_factor, world, allow_inexact, boundary_relay)
    g0, g1 = ranges[rank]
    if g1 == g0:
        return None
    frames = shard_frames(low0, TRLs, g0, g1)
    if analyze_fn is None:
        def analyze_fn(fr, n_gops, first_global):
            return ctx.analyze(fr, X, Y, n_gops, TRLs, block_size, search_range,
                               subpixel_accuracy, update_factor, always_B,
                               block_size_min=block_size_min, first_global=first_global, **analyze_kw)
    if relay is None and needs_tail_exchange(Y, block_size, world) and ctx is not None:
        relay = TailRelay(rank, ranges)
    if relay is not None and ctx is not None:
        if getattr(relay, "device", False):
            ctx.set_tail_exchange(relay, device=True)
        else:
            ctx.set_tail_exchange(relay)
    if boundary_relay is not None and ctx is not None:
        ctx.set_boundary_exchange(boundary_relay)
    try:
        return analyze_fn(frames, g1 - g0, g0 == 0)
    finally:
        if relay is not None and ctx is not None:
            ctx.set_tail_exchange(None)
        if boundary_relay is not None and ctx is not None:
            ctx.set_boundary_exchange(None)


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


# ------------------------------------------------------------------ synthesis

def shard_subbands(subbands, TRLs: int, g0: int, g1: int):
    """The slices of high_t / motion_t / frame_types_t / low_{TRLs-1} that belong to GOPs
    [g0, g1): level t has 2^(TRLs-1-t) pairs per GOP, low_{TRLs-1} one frame per GOP + 1."""
    out = {}
    for t in range(1, TRLs):
        k = 2 ** (TRLs - 1 - t)
        out[f"high_{t}"] = subbands[f"high_{t}"][g0 * k : g1 * k]
        out[f"motion_{t}"] = subbands[f"motion_{t}"][g0 * k : g1 * k]
        out[f"frame_types_{t}"] = bytes(subbands[f"frame_types_{t}"])[g0 * k : g1 * k]
    out[f"low_{TRLs - 1}"] = subbands[f"low_{TRLs - 1}"][g0 : g1 + 1]
    return out


def synthesize_shard(ctx, subbands, X, Y, GOPs, TRLs, rank, world, block_size=16, search_range=4,
                     subpixel_accuracy=0, update_factor=0.0, allow_inexact=False,
                     synthesize_fn=None, relay=None, boundary_relay=None):
    """Reconstructs this rank's GOP range: low_0 frames [g0*G, g1*G] inclusive."""
    ranges = partition(GOPs, world)
    if (boundary_relay is None and needs_boundary_exchange(update_factor, world) and ctx is not None
            and synthesize_fn is None and _dist_ready()):
        boundary_relay = BoundaryRelay(rank, ranges)
    check_exact(X, Y, block_size, update_factor, world, allow_inexact, boundary_relay)
    g0, g1 = ranges[rank]
    if g1 == g0:
        return None
    sub = shard_subbands(subbands, TRLs, g0, g1)
    if synthesize_fn is None:
        def synthesize_fn(sb, n_gops):
            return ctx.synthesize(sb, X, Y, n_gops, TRLs, block_size, search_range,
                                  subpixel_accuracy, update_factor)
    if relay is None and needs_tail_exchange(Y, block_size, world) and ctx is not None:
        relay = TailRelay(rank, ranges)
    if relay is not None and ctx is not None:
        if getattr(relay, "device", False):
            ctx.set_tail_exchange(relay, device=True)
        else:
            ctx.set_tail_exchange(relay)
    if boundary_relay is not None and ctx is not None:
        ctx.set_boundary_exchange(boundary_relay)
    try:
        return synthesize_fn(sub, g1 - g0)
    finally:
        if relay is not None and ctx is not None:
            ctx.set_tail_exchange(None)
        if boundary_relay is not None and ctx is not None:
            ctx.set_boundary_exchange(None)


def gather_frames(parts):
    """Concatenates reconstructed GOP ranges, dropping each duplicated boundary frame."""
    parts = [p for p in parts if p is not None]
    return np.concatenate([p if i == 0 else p[1:] for i, p in enumerate(parts)], axis=0)


def synthesize_distributed(ctx, subbands, X, Y, GOPs, TRLs, **kw):
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    local = synthesize_shard(ctx, subbands, X, Y, GOPs, TRLs, rank, world, **kw)
    parts = [None] * world if rank == 0 else None
    dist.gather_object(local, parts, dst=0)
    return gather_frames(parts) if rank == 0 else None
