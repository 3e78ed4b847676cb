"""Multi-GPU check (not collected by pytest): one sequence whose height is not a multiple
of the block size, GOP-sharded over the ranks of a torchrun job (one GPU per rank, NCCL
point-to-point for the prediction tail state and, with update_factor != 0, for the boundary
frame), analysis and synthesis, compared byte for byte on rank 0 with the CPU oracle.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tests/run_sharded_parity.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from qsvc_b200 import shard, yuv  # noqa: E402
from qsvc_b200.mctf import Context  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world = dist.get_world_size()
    X, Y, TRLs, bs, sr, a = 128, 72, 4, 16, 4, 2
    GOPs = max(2, world)
    clip = yuv.synthetic_clip(X, Y, GOPs * 2 ** (TRLs - 1) + 1, 31, max_shift=12)
    ok = True
    with Context(local) as ctx:
        # update_factor 0: tail hand-over only; 0.25: also the boundary-frame hand-over
        for uf in (0.0, 0.25):
            kw = dict(block_size=bs, search_range=sr, subpixel_accuracy=a, update_factor=uf)
            got = shard.analyze_distributed(ctx, clip, X, Y, GOPs, TRLs, block_size_min=bs, **kw)
            ref = orc.analyze(clip, X, Y, TRLs, bs, sr, a, uf, block_size_min=bs)
            sub = {f"low_{TRLs-1}": ref[f"low_{TRLs-1}"]}
            for t in range(1, TRLs):
                sub[f"high_{t}"], sub[f"motion_{t}"] = ref[f"high_{t}"], ref[f"motion_filtered_{t}"]
                sub[f"frame_types_{t}"] = ref[f"frame_types_{t}"]
            rec = shard.synthesize_distributed(ctx, sub, X, Y, GOPs, TRLs, **kw)
            if rank == 0:
                good = True
                for k, v in got.items():
                    same = np.array_equal(ref[k], v) if not isinstance(v, bytes) else ref[k] == v
                    good &= bool(same)
                    if not same:
                        print("MISMATCH", k)
                whole = ctx.synthesize(sub, X, Y, GOPs, TRLs, bs, sr, a, uf)
                good &= bool(np.array_equal(rec, whole))
                ok &= good
                print(f"sharded parity over {world} GPUs, update_factor {uf}: analysis+synthesis "
                      f"{'OK' if good else 'FAILED'}")
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
