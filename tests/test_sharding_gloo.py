"""CPU, world_size 2 over gloo: the multi-GPU host logic (GOP partition, per-rank
analysis, host-side gather).  The per-shard compute is injected (the CPU oracle
stands in for the CUDA call) so that the test checks what the sharding layer is
responsible for: that concatenating independently analysed GOP ranges reproduces
the single-process result byte for byte."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as orc
from qsvc_b200 import shard, yuv

X, Y, GOPs, TRLs, BS, SR, A = 64, 48, 4, 3, 16, 4, 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    clip = yuv.synthetic_clip(X, Y, GOPs * 2 ** (TRLs - 1) + 1, 5, max_shift=12)

    def fn(frames, n_gops, first_global):
        return orc.analyze(frames, X, Y, TRLs, BS, SR, A, 0.0, block_size_min=BS)

    out = shard.analyze_distributed(None, clip, X, Y, GOPs, TRLs, block_size=BS, search_range=SR,
                                    subpixel_accuracy=A, update_factor=0.0, block_size_min=BS,
                                    analyze_fn=fn)
    if rank == 0:
        full = orc.analyze(clip, X, Y, TRLs, BS, SR, A, 0.0, block_size_min=BS)
        ok = all(np.array_equal(out[k], full[k]) if not isinstance(out[k], bytes) else out[k] == full[k]
                 for k in out)
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_gop_sharded_analysis_equals_single_process():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok


def test_sharding_refuses_inexact_configurations():
    import pytest
    with pytest.raises(ValueError):
        shard.check_exact(64, 48, 16, 0.25, world=2)
    with pytest.raises(ValueError):
        shard.check_exact(60, 40, 16, 0.0, world=2)   # uncovered rows and columns: no exchange hook
    shard.check_exact(64, 40, 16, 0.0, world=2)       # uncovered rows only: tail exchange
    shard.check_exact(64, 40, 16, 0.25, world=1)
    assert shard.needs_tail_exchange(1080, 16, 2) and not shard.needs_tail_exchange(2160, 16, 8)


class _FakeCtx:
    """Stands in for mctf.Context: drives the installed tail callback the way the library
    does (phase 0 before the first pair of a level, phase 1 after the last one)."""

    def __init__(self, rank):
        self.rank, self.fn, self.seen = rank, None, []

    def set_tail_exchange(self, fn):
        self.fn = fn

    def run_levels(self, synthesis):
        levels = range(TRLs - 1, 0, -1) if synthesis else range(1, TRLs)
        for t in levels:
            st = np.zeros(24, np.uint8)
            got = self.fn(t, synthesis, 0, st)
            self.seen.append((t, bool(got), st.copy()))
            st[:] = 100 * self.rank + 10 * synthesis + t
            self.fn(t, synthesis, 1, st)


def _relay_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = _FakeCtx(rank)
    clip = np.zeros((GOPs * 2 ** (TRLs - 1) + 1, 8), np.uint8)

    def fn(frames, n_gops, first_global):
        ctx.run_levels(0)
        return {"rank": rank, "first": first_global, "frames": len(frames)}

    out = shard.analyze_shard(ctx, clip, 64, 40, GOPs, TRLs, rank, world, block_size=16, update_factor=0.0,
                              analyze_fn=fn)
    assert ctx.fn is None  # callback removed again
    sub = {f"high_{t}": np.zeros((GOPs * 2 ** (TRLs - 1 - t), 8), np.uint8) for t in range(1, TRLs)}
    sub.update({f"motion_{t}": np.zeros((GOPs * 2 ** (TRLs - 1 - t), 4, 1, 1), np.int16) for t in range(1, TRLs)})
    sub.update({f"frame_types_{t}": b"B" * (GOPs * 2 ** (TRLs - 1 - t)) for t in range(1, TRLs)})
    sub[f"low_{TRLs - 1}"] = np.arange(GOPs + 1, dtype=np.uint8).reshape(-1, 1).repeat(8, 1)

    def sfn(sb, n_gops):
        ctx.run_levels(1)
        G = 2 ** (TRLs - 1)
        assert sb[f"low_{TRLs - 1}"].shape[0] == n_gops + 1 and sb["high_1"].shape[0] == n_gops * G // 2
        first = int(sb[f"low_{TRLs - 1}"][0, 0]) * G
        return np.arange(first, first + n_gops * G + 1, dtype=np.uint8).reshape(-1, 1)

    rec = shard.synthesize_distributed(ctx, sub, 64, 40, GOPs, TRLs, block_size=16, update_factor=0.0,
                                       synthesize_fn=sfn)
    q.put((rank, out, ctx.seen, None if rec is None else rec[:, 0].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_tail_relay_passes_state_left_to_right():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_relay_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict((r[0], r) for r in (q.get(timeout=120), q.get(timeout=120)))
    for p in procs:
        p.join(timeout=60)
    # rank 0 starts every level from zeros; rank 1 from what rank 0 ended with
    assert all(not got and not st.any() for (_, got, st) in res[0][2])
    for (t, got, st), syn in zip(res[1][2], [0] * (TRLs - 1) + [1] * (TRLs - 1)):
        assert got and (st == 10 * syn + t).all()
    assert res[0][1]["first"] and not res[1][1]["first"]
    assert res[0][3] == list(range(GOPs * 2 ** (TRLs - 1) + 1))  # gathered frames, no duplicates


class _FakeUpdateCtx:
    """Drives the boundary-frame callback the way update_level does: the shard's last frame
    first (phase 1), then its first frame (phases 0 and 2), then the frame coming back (phase 3)."""

    def __init__(self, rank):
        self.rank, self.fn, self.log = rank, None, []

    def set_tail_exchange(self, fn):
        pass

    def set_boundary_exchange(self, fn):
        self.fn = fn

    def run_levels(self):
        for t in range(1, TRLs):
            planes = np.full(32, 10 * self.rank + t, np.uint8)
            has_right = self.fn(t, 0, 1, planes)
            first = np.zeros(32, np.uint8)
            got = self.fn(t, 0, 0, first)
            frame = np.full(16, 100 + 10 * self.rank + t, np.uint8)
            self.fn(t, 0, 2, frame)
            back = np.zeros(16, np.uint8)
            if has_right:
                assert self.fn(t, 0, 3, back)
            self.log.append((t, bool(has_right), bool(got), int(first[0]), int(back[0])))


def _boundary_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = _FakeUpdateCtx(rank)
    clip = np.zeros((GOPs * 2 ** (TRLs - 1) + 1, 8), np.uint8)

    def fn(frames, n_gops, first_global):
        ctx.run_levels()
        return {"rank": rank}

    ranges = shard.partition(GOPs, world)
    shard.analyze_shard(ctx, clip, 64, 48, GOPs, TRLs, rank, world, block_size=16, update_factor=0.25,
                        analyze_fn=fn, boundary_relay=shard.BoundaryRelay(rank, ranges))
    assert ctx.fn is None
    q.put((rank, ctx.log))
    dist.barrier()
    dist.destroy_process_group()


def test_boundary_relay_passes_planes_right_and_the_frame_back():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_boundary_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    for t, has_right, got, first, back in res[0]:
        assert has_right and not got and first == 0 and back == 110 + t  # rank 1's finished frame came back
    for t, has_right, got, first, back in res[1]:
        assert not has_right and got and first == t and back == 0        # rank 0's planes arrived
