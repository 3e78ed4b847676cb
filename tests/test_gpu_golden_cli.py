"""GPU: the CUDA path against the committed reference fixtures, through the
Python API, through the flag-compatible command-line tools, GOP-sharded, and
through size-independent properties at full 1080p size."""
import os
import subprocess

import numpy as np
import pytest

from golden_util import NAMES, load, schedule
from qsvc_b200 import shard, yuv
from qsvc_b200.mctf import level_schedule

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MCTF = os.path.join(ROOT, "bin", "mctf")


@pytest.mark.parametrize("name", NAMES)
def test_api_matches_golden(ctx, name):
    g = load(name)
    X, Y, bs, a, uf, T = g["X"], g["Y"], g["bs"], g["a"], g["uf"], g["TRLs"]
    got = ctx.analyze(g["low_0"], X, Y, g["GOPs"], T, bs, g["sr"], a, uf, g["always_B"],
                      block_overlaping=g["ov"], block_size_min=bs)
    sub = {f"low_{T-1}": got[f"low_{T-1}"]}
    for t in range(1, T):
        for n in ("motion", "motion_filtered", "high", "low"):
            assert np.array_equal(got[f"{n}_{t}"], g[f"{n}_{t}"]), f"{n}_{t}"
        assert got[f"frame_types_{t}"] == bytes(g[f"frame_types_{t}"])
        sub[f"high_{t}"], sub[f"motion_{t}"] = got[f"high_{t}"], got[f"motion_filtered_{t}"]
        sub[f"frame_types_{t}"] = got[f"frame_types_{t}"]
    rec = ctx.synthesize(sub, X, Y, g["GOPs"], T, bs, g["sr"], a, uf, g["ov"])
    assert np.array_equal(rec, g["syn_low_0"])


def _mctf(args, cwd):
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([MCTF] + args, cwd=cwd, env=env, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()[-2000:]


@pytest.mark.parametrize("name", ["ib_types_a0", "quarter_pel", "obmc_half_pel"])
def test_cli_tool_chain_matches_golden(tmp_path, name):
    """bin/mctf analyze (fused) and the per-step tools (analyze_step) write the same
    files the reference chain writes; bin/mctf synthesize reconstructs its low_0."""
    g = load(name)
    X, Y, bs, a, uf, T, GOPs = g["X"], g["Y"], g["bs"], g["a"], g["uf"], g["TRLs"], g["GOPs"]
    common = [f"--pixels_in_x={X}", f"--pixels_in_y={Y}", f"--block_size={bs}",
              f"--subpixel_accuracy={a}", f"--update_factor={uf}"]
    ovf = [f"--block_overlaping={g['ov']}"] if g["ov"] else []
    d1, d2 = tmp_path / "fused", tmp_path / "steps"
    for d in (d1, d2):
        d.mkdir()
        yuv.write_frames(str(d / "low_0"), g["low_0"])
    _mctf(["analyze", f"--GOPs={GOPs}", f"--TRLs={T}", f"--search_range={g['sr']}",
           f"--block_size_min={bs}", f"--always_B={g['always_B']}"] + ovf + common, str(d1))
    pictures = GOPs * 2 ** (T - 1) + 1
    for t, sr in schedule(g):
        _mctf(["analyze_step", f"--pictures={pictures}", f"--search_range={sr}",
               f"--temporal_subband={t}", f"--always_B={g['always_B']}"] + ovf + common, str(d2))
        pictures = (pictures + 1) // 2
    for d in (d1, d2):
        for t in range(1, T):
            n = g[f"high_{t}"].shape[0]
            assert np.array_equal(yuv.read_motion(str(d / f"motion_{t}"), X, Y, bs, n), g[f"motion_{t}"])
            assert np.array_equal(yuv.read_motion(str(d / f"motion_filtered_{t}"), X, Y, bs, n), g[f"motion_filtered_{t}"])
            assert np.array_equal(yuv.read_frames(str(d / f"high_{t}"), X, Y), g[f"high_{t}"])
            assert np.array_equal(yuv.read_frames(str(d / f"low_{t}"), X, Y), g[f"low_{t}"])
            assert (d / f"frame_types_{t}").read_bytes() == bytes(g[f"frame_types_{t}"])
    for t in range(1, T):  # decoder side: motion_filtered_t is delivered as motion_t
        os.replace(str(d2 / f"motion_filtered_{t}"), str(d2 / f"motion_{t}"))
        assert np.array_equal(yuv.read_frames(str(d2 / f"prediction_even_{t}"), X, Y), g[f"prediction_even_{t}"])
    os.remove(str(d2 / "low_0"))
    lists = ",".join([str(bs)] * T)
    _mctf(["synthesize", f"--GOPs={GOPs}", f"--TRLs={T}", f"--search_range={g['sr']}",
           f"--block_size={lists}", f"--pixels_in_x={','.join([str(X)] * (T + 1))}",
           f"--pixels_in_y={','.join([str(Y)] * (T + 1))}",
           f"--subpixel_accuracy={','.join([str(a)] * (T + 1))}", f"--update_factor={uf}"] + ovf, str(d2))
    assert np.array_equal(yuv.read_frames(str(d2 / "low_0"), X, Y), g["syn_low_0"])


def test_gop_sharded_analysis_on_gpu(ctx):
    """Two GOP shards analysed independently and gathered == whole sequence
    (the shard that does not start the sequence carries reference[0])."""
    from oracle import oracle as orc
    X, Y, GOPs, TRLs, bs, sr, a = 128, 96, 4, 3, 16, 4, 1
    clip = yuv.synthetic_clip(X, Y, GOPs * 2 ** (TRLs - 1) + 1, 8, max_shift=12)
    kw = dict(block_size=bs, search_range=sr, subpixel_accuracy=a, update_factor=0.0, block_size_min=bs)
    parts = [shard.analyze_shard(ctx, clip, X, Y, GOPs, TRLs, r, 2, **kw) for r in range(2)]
    out = shard.gather(parts, TRLs)
    full = orc.analyze(clip, X, Y, TRLs, bs, sr, a, 0.0, block_size_min=bs)
    for k, v in out.items():
        assert (v == full[k]) if isinstance(v, bytes) else np.array_equal(v, full[k]), k


def test_properties_at_1080p(ctx):
    """Size-independent properties at the bench's full picture size (one GOP of 8)."""
    X, Y, GOPs, TRLs, bs, sr, a = 1920, 1080, 1, 4, 16, 16, 2
    clip = yuv.synthetic_clip(X, Y, GOPs * 2 ** (TRLs - 1) + 1, 2, max_shift=48)
    got = ctx.analyze(clip, X, Y, GOPs, TRLs, bs, sr, a, 0.0, always_B=1, block_size_min=bs)
    low = clip
    sub = {f"low_{TRLs-1}": got[f"low_{TRLs-1}"]}
    for s in level_schedule(GOPs, TRLs, bs, sr, bs):
        t = s["t"]
        assert got[f"frame_types_{t}"] == b"B" * s["pairs"]                    # always_B
        assert np.array_equal(got[f"motion_filtered_{t}"], got[f"motion_{t}"])  # B frames keep their vectors
        assert np.array_equal(got[f"low_{t}"], low[0::2])                       # update_factor 0: low_t == even_t
        lim = (s["search_range"] << a) + (1 << a) - 1                             # reach of the +-1 descent
        assert np.abs(got[f"motion_{t}"]).max() <= lim
        low = got[f"low_{t}"]
        sub[f"high_{t}"], sub[f"motion_{t}"] = got[f"high_{t}"], got[f"motion_filtered_{t}"]
        sub[f"frame_types_{t}"] = got[f"frame_types_{t}"]
    rec = ctx.synthesize(sub, X, Y, GOPs, TRLs, bs, sr, a, 0.0)
    # even frames of the coarsest level come back untouched; the odd frame of the
    # coarsest level (input frame 4) is exact wherever its residue was not clamped
    assert np.array_equal(rec[0::8], clip[0::8])
    h = got[f"high_{TRLs-1}"][0]
    unclamped = (h > 0) & (h < 255)
    assert unclamped.mean() > 0.9
    assert np.array_equal(rec[4][unclamped], clip[4][unclamped])


def test_cli_synthesize_with_per_level_lists(tmp_path):
    """synthesize.py:127-133 indexes block_size / pixels / subpixel_accuracy per temporal level
    (spatially scalable decoding): lists that vary run level by level with
    each level's own values, like the reference's chain of synthesize_step calls."""
    from oracle import oracle as orc
    g = load("quarter_pel")
    X, Y, bs, T, GOPs, uf = g["X"], g["Y"], g["bs"], g["TRLs"], g["GOPs"], g["uf"]
    acc = {3: 2, 2: 1, 1: 2}  # per temporal level; the list is indexed [TRLs - t]
    d = tmp_path
    for t in range(1, T):
        yuv.write_frames(str(d / f"high_{t}"), g[f"high_{t}"])
        yuv.write_motion(str(d / f"motion_{t}"), g[f"motion_filtered_{t}"])
        (d / f"frame_types_{t}").write_bytes(bytes(g[f"frame_types_{t}"]))
    yuv.write_frames(str(d / f"low_{T-1}"), g[f"low_{T-1}"])
    acc_list = ["0"] * (T + 1)
    for t, v in acc.items():
        acc_list[T - t] = str(v)
    _mctf(["synthesize", f"--GOPs={GOPs}", f"--TRLs={T}", f"--search_range={g['sr']}",
           f"--block_size={','.join([str(bs)] * T)}", f"--pixels_in_x={','.join([str(X)] * (T + 1))}",
           f"--pixels_in_y={','.join([str(Y)] * (T + 1))}", f"--subpixel_accuracy={','.join(acc_list)}",
           f"--update_factor={uf}"], str(d))
    low = g[f"low_{T-1}"]
    for t, sr in reversed(schedule(g)):
        types, mv = bytes(g[f"frame_types_{t}"]), g[f"motion_filtered_{t}"]
        even = orc.update(low, g[f"high_{t}"], mv, types, X, Y, bs, uf, inverse=True)
        odd, _ = orc.correlate(even, g[f"high_{t}"], mv, types, X, Y, bs, sr, acc[t])
        low = np.empty((2 * odd.shape[0] + 1, even.shape[1]), np.uint8)
        low[0::2], low[1::2] = even, odd
    assert np.array_equal(yuv.read_frames(str(d / "low_0"), X, Y), low)


def test_cli_synthesize_with_per_level_picture_sizes_matches_reference(tmp_path):
    """Lists that vary in pixels_in_x / pixels_in_y AND subpixel_accuracy per temporal level
    (synthesize.py:127-133 indexes block_size [(TRLs-1)-t], the others
    [TRLs-t]) against the low_0 the unmodified reference tools produced for the same files
    (tests/golden/level_lists.npz)."""
    from golden_util import load_level_lists
    g = load_level_lists()
    T, d = g["TRLs"], tmp_path
    for t in range(1, T):
        yuv.write_frames(str(d / f"high_{t}"), g[f"high_{t}"])
        yuv.write_motion(str(d / f"motion_{t}"), g[f"motion_{t}"])
        (d / f"frame_types_{t}").write_bytes(bytes(g[f"frame_types_{t}"]))
    yuv.write_frames(str(d / f"low_{T-1}"), g[f"low_{T-1}"])
    xs, ys, acc = ["0"] * (T + 1), ["0"] * (T + 1), ["0"] * (T + 1)
    for t, (X, Y, a) in g["geo"].items():
        xs[T - t], ys[T - t], acc[T - t] = str(X), str(Y), str(a)
    xs[0], ys[0] = xs[1], ys[1]
    _mctf(["synthesize", f"--GOPs={g['GOPs']}", f"--TRLs={T}", f"--search_range={g['sr']}",
           f"--block_size={','.join([str(g['bs'])] * T)}", f"--pixels_in_x={','.join(xs)}",
           f"--pixels_in_y={','.join(ys)}", f"--subpixel_accuracy={','.join(acc)}",
           f"--update_factor={g['uf']}"], str(d))
    X1, Y1, _ = g["geo"][1]
    assert np.array_equal(yuv.read_frames(str(d / "low_0"), X1, Y1), g["syn_low_0"])


def test_cli_analyze_stops_at_an_existing_motion_file(tmp_path):
    """motion_estimate exits 1 without computing when its motion file exists
    (motion_estimate.cpp:659-682); analyze.py turns that into exit -1 after the levels
    before it (and that level's split) have run."""
    g = load("quarter_pel")
    X, Y, bs, T, GOPs = g["X"], g["Y"], g["bs"], g["TRLs"], g["GOPs"]
    d = tmp_path
    yuv.write_frames(str(d / "low_0"), g["low_0"])
    (d / "motion_2").write_bytes(b"stale")
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([MCTF, "analyze", f"--GOPs={GOPs}", f"--TRLs={T}", f"--block_size={bs}",
                        f"--block_size_min={bs}", f"--search_range={g['sr']}",
                        f"--subpixel_accuracy={g['a']}", f"--pixels_in_x={X}", f"--pixels_in_y={Y}",
                        f"--update_factor={g['uf']}"], cwd=str(d), env=env, capture_output=True)
    assert r.returncode == 255
    assert np.array_equal(yuv.read_frames(str(d / "high_1"), X, Y), g["high_1"])   # level 1 ran
    assert (d / "even_2").exists() and (d / "odd_2").exists()                      # level 2's split ran
    assert not (d / "high_2").exists() and (d / "motion_2").read_bytes() == b"stale"
