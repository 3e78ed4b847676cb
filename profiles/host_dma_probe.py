"""Aggregate host <-> device copy bandwidth of one box with N ranks copying at once (torchrun).

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 profiles/host_dma_probe.py

Every rank moves the byte counts of one bench.py step (401 MB up, 417 MB down) between pinned host
memory and its GPU, first alone (rank 0 only), then all ranks together: upload only, download only,
both directions at once on two streams (what the end-to-end leg of bench.py does).  The ratio of the
last line to N x the first is the ceiling the platform puts on end-to-end scaling, independent of
this library.
"""
import os
import time

import torch
import torch.distributed as dist

UP, DOWN, REPS = 401_241_600, 417_193_084, 6


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        dist.init_process_group("nccl")
    h_up = torch.empty(UP, dtype=torch.uint8).pin_memory()
    h_dn = torch.empty(DOWN, dtype=torch.uint8).pin_memory()
    h_up.fill_(1)
    h_dn.fill_(0)
    d_up = torch.empty(UP, dtype=torch.uint8, device="cuda")
    d_dn = torch.ones(DOWN, dtype=torch.uint8, device="cuda")
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def run(up, down, active):
        barrier()
        t0 = time.perf_counter()
        if active:
            for _ in range(REPS):
                if up:
                    with torch.cuda.stream(s_up):
                        d_up.copy_(h_up, non_blocking=True)
                if down:
                    with torch.cuda.stream(s_dn):
                        h_dn.copy_(d_dn, non_blocking=True)
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt if active else 0.0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for label, up, down in (("upload", True, False), ("download", False, True), ("both", True, True)):
        nbytes = REPS * ((UP if up else 0) + (DOWN if down else 0))
        run(up, down, True)  # warm
        solo = run(up, down, rank == 0)
        together = run(up, down, True)
        if rank == 0:
            print(f"{label:9s} rank 0 alone {nbytes / solo / 1e9:7.1f} GB/s | {world} ranks together "
                  f"{world * nbytes / together / 1e9:7.1f} GB/s aggregate = {nbytes / together / 1e9:6.1f} per rank "
                  f"(x{world * solo / together / world:.2f} of alone)", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
