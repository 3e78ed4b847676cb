"""One warm-up and one measured analysis step of the bench workload (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from qsvc_b200 import yuv
from qsvc_b200.mctf import Context

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
uf = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0  # update_factor (0.25 = the codec's default)
w = bench.WORKLOADS[name]
clip = yuv.synthetic_clip(w["X"], w["Y"], bench.n_frames(w), bench.SEEDS[name], max_shift=min(48, 3 * w["sr"]))
with Context(0) as ctx:
    ctx.resident_load(clip, w["X"], w["Y"])
    for _ in range(steps):
        l0 = ctx.launches
        ctx.resident_analyze(TRLs=w["TRLs"], block_size=w["bs"], search_range=w["sr"],
                             subpixel_accuracy=w["a"], update_factor=uf, always_B=w["always_B"],
                             block_size_min=w["bs"])
        print("launches per step:", ctx.launches - l0, "total_ms", ctx.resident_stats()["total_ms"])
