"""CPU: host-side logic of the drop-in boundary (schedules, file formats, CLI)."""
import os
import subprocess
import sys

import numpy as np

from qsvc_b200 import shard, yuv
from qsvc_b200.mctf import gop_size, level_schedule, merge, split

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MCTF = os.path.join(ROOT, "bin", "mctf")


def test_level_schedule_follows_analyze_py():
    # cfg3: --TRLs=6 --GOPs=4 --block_size=16 --search_range=16
    s = level_schedule(4, 6, 16, 16)
    assert [x["pictures"] for x in s] == [129, 65, 33, 17, 9]
    assert [x["pairs"] for x in s] == [64, 32, 16, 8, 4]
    assert [x["search_range"] for x in s] == [16, 32, 64, 128, 128]
    assert all(x["block_size"] == 16 for x in s)  # block_size_min collapses to block_size
    # block size halves down to block_size_min (analyze.py:149-151)
    assert [x["block_size"] for x in level_schedule(1, 4, 64, 4, 32)] == [64, 32, 32]
    assert gop_size(5) == 16


def test_split_merge_round_trip():
    clip = yuv.synthetic_clip(32, 16, 9, 0)
    even, odd = split(clip)
    assert even.shape[0] == 5 and odd.shape[0] == 4
    assert np.array_equal(merge(even, odd), clip)


def test_motion_file_layout(tmp_path):
    mv = np.arange(2 * 4 * 3 * 5, dtype=np.int16).reshape(2, 4, 3, 5) - 40
    p = str(tmp_path / "motion_1")
    yuv.write_motion(p, mv)
    raw = np.fromfile(p, "<i2")
    assert raw.size == 2 * 4 * 3 * 5 and raw[0] == -40  # PREV.X plane first, row-major
    assert np.array_equal(yuv.read_motion(p, 5 * 16, 3 * 16, 16, 2), mv)


def test_gop_partition_and_gather():
    assert shard.partition(8, 4) == [(0, 2), (2, 4), (4, 6), (6, 8)]
    assert shard.partition(4, 8)[:5] == [(0, 1), (1, 2), (2, 3), (3, 4), (4, 4)]
    assert shard.partition(5, 2) == [(0, 3), (3, 5)]
    clip = yuv.synthetic_clip(32, 16, 4 * 4 + 1, 0)
    a = shard.shard_frames(clip, 3, 0, 2)
    b = shard.shard_frames(clip, 3, 2, 4)
    assert a.shape[0] == 9 and b.shape[0] == 9 and np.array_equal(a[-1], b[0])


def _run(args, cwd):
    env = dict(os.environ, PYTHONPATH=ROOT)
    return subprocess.run([MCTF] + args, cwd=cwd, env=env, capture_output=True)


def test_cli_split_merge_and_exit_codes(tmp_path):
    d = str(tmp_path)
    clip = yuv.synthetic_clip(64, 48, 5, 1)
    yuv.write_frames(os.path.join(d, "low_0"), clip)
    geo = ["--pictures=5", "--pixels_in_x=64", "--pixels_in_y=48"]
    assert _run(["split", "--even_fn=even_1", "--low_fn=low_0", "--odd_fn=odd_1"] + geo, d).returncode == 0
    assert os.path.getsize(os.path.join(d, "even_1")) == 3 * 4608
    assert os.path.getsize(os.path.join(d, "odd_1")) == 2 * 4608
    # abbreviated flags as used by synthesize_step.py:134-141
    assert _run(["merge", "--even=even_1", "--low=low_back", "--odd=odd_1"] + geo, d).returncode == 0
    assert open(os.path.join(d, "low_back"), "rb").read() == clip.tobytes()
    # every call is appended to ./trace (mctf.sh:35)
    assert len(open(os.path.join(d, "trace")).read().splitlines()) == 2
    # motion_estimate refuses to overwrite an existing motion file: exit 1, no work
    open(os.path.join(d, "motion_1"), "wb").write(b"x")
    r = _run(["motion_estimate", "--even_fn=even_1", "--odd_fn=odd_1", "--motion_fn=motion_1"] + geo, d)
    assert r.returncode == 1 and open(os.path.join(d, "motion_1"), "rb").read() == b"x"
    # --help exits 1 like the reference tools
    assert _run(["decorrelate", "--help"], d).returncode == 1
    # unreadable input: abort
    r = _run(["split", "--low_fn=missing"] + geo, d)
    assert r.returncode == 134 and b"aborting" in r.stderr


def test_thread_boundary_relay_routes_between_neighbours():
    """shard.ThreadBoundaryRelay: planes travel right (phase 1 -> phase 0), the finished frame travels
    back left (phase 2 -> phase 3); the first shard has no left neighbour, the last no right one."""
    import threading

    from qsvc_b200 import shard
    n = 3
    relays = shard.ThreadBoundaryRelay.make(n)
    log = [None] * n

    def work(r):
        planes = np.full(8, 10 + r, np.uint8)
        has_right = relays[r](1, 0, 1, planes)          # last frame after the NEXT pass
        first = np.zeros(8, np.uint8)
        got = relays[r](1, 0, 0, first)                 # first frame: the left neighbour's planes
        relays[r](1, 0, 2, np.full(4, 100 + r, np.uint8))
        back = np.zeros(4, np.uint8)
        if has_right:
            assert relays[r](1, 0, 3, back)
        log[r] = (bool(has_right), bool(got), int(first[0]), int(back[0]))

    th = [threading.Thread(target=work, args=(r,)) for r in range(n)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=30)
    assert log[0] == (True, False, 0, 101)
    assert log[1] == (True, True, 10, 102)
    assert log[2] == (False, True, 11, 0)
