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
        shard.check_exact(64, 40, 16, 0.0, world=2)
    shard.check_exact(64, 40, 16, 0.25, world=1)
