import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
import bench
from qsvc_b200 import yuv
from qsvc_b200.mctf import Context
w = bench.WORKLOADS["cfg3"]
clip = yuv.synthetic_clip(w["X"], w["Y"], bench.n_frames(w), 2, max_shift=48)
with Context(0) as ctx:
    pin = ctx.pinned_like(clip) if hasattr(ctx, "pinned_like") else clip
    try:
        import torch
        t = torch.from_numpy(clip).pin_memory(); pin = t.numpy()
    except Exception as e:
        print("no pin", e)
    kw = dict(block_size=16, search_range=16, subpixel_accuracy=2, update_factor=0.0, always_B=1, block_size_min=16)
    for i in range(3):
        ctx.analyze(pin, w["X"], w["Y"], w["GOPs"], w["TRLs"], reuse_buffers=True, **kw)
    ws, gs = [], []
    for i in range(6):
        t0 = time.perf_counter()
        ctx.analyze(pin, w["X"], w["Y"], w["GOPs"], w["TRLs"], reuse_buffers=True, **kw)
        ws.append((time.perf_counter() - t0) * 1e3)
        gs.append(ctx.resident_stats()["total_ms"])
    print("wall ms", [round(x, 2) for x in ws])
    print("gpu span ms (ev2..ev3)", [round(x, 2) for x in gs])
    ctx.resident_load(pin, w["X"], w["Y"])
    for i in range(3):
        ctx.resident_analyze(TRLs=w["TRLs"], **kw)
        print("resident span", round(ctx.resident_stats()["total_ms"], 2))
