"""Byte-exact parity of the CUDA path against the CPU oracle AT THE BENCHMARKED SIZES.

The small geometries of test_gpu_parity.py never reach the size-dependent branches of the
fused kernels (64-wide TMA boxes on 7680-byte planes, the 1080-line tail chain, strips and
segments of the decorrelate pass, chunking by the HBM budget).  Here the kernels bench.py
times are forced (me_mode 2 / mc_mode 2 where the fused path applies) and compared with the
oracle on:
  (i)   1920x1080, block 16, quarter-pel (cfg3): a 5-frame clip through levels sr 16 and 32,
        single pairs at sr 64 and 128 (non-invertible pyramid: carried reference[0]);
  (ii)  704x576 half-pel sr 8 (cfg2): one GOP of 16, analysis and synthesis round trip;
  (iii) 3840x2160 integer-pel (cfg4): sr 16 and sr 128;
  (iv)  the 1080p clip cut into two GOP shards with the prediction-tail relay;
  (v)   qsvc_analyze on a pinned clip uploaded GOP by GOP behind the motion estimation
        (the upload / decorrelate ordering).
The oracle jobs (about 25 s of one core per 1080p pair) run once per session in a process pool.
"""
import concurrent.futures as cf
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

from qsvc_b200 import shard, yuv
from qsvc_b200.mctf import level_schedule

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

HD = dict(X=1920, Y=1080, bs=16, a=2)
UHD = dict(X=3840, Y=2160, bs=16, a=0)
CIF4 = dict(X=704, Y=576, bs=16, a=1)


def hd_clip():
    return yuv.synthetic_clip(1920, 1080, 5, 2, max_shift=48)


def hd_pair(sr):
    # frames 7 apart on the pan trajectory: motion of 0.66 / 0.33 * (sr - 8) pixels, enough to
    # drive vectors beyond +-127 quarter-pels and search windows across the picture borders
    return yuv.synthetic_clip(1920, 1080, 15, 40 + sr, max_shift=sr - 8)[[0, 7, 14]]


def uhd_clip():
    return yuv.synthetic_clip(3840, 2160, 3, 13, max_shift=48)


def cif4_clip():
    return yuv.synthetic_clip(704, 576, 17, 11, max_shift=24)


from oracle_jobs import _job_chain, _job_pair, _job_synth  # noqa: E402  (run in worker processes)


@pytest.fixture(scope="module")
def ref():
    """Every oracle result of this module, computed concurrently (one core each)."""
    from oracle import oracle as orc
    orc.build()
    workers = max(1, min(7, (os.cpu_count() or 2) - 1))
    # spawn: the session's CUDA context must not be forked
    with cf.ProcessPoolExecutor(workers, mp_context=mp.get_context("spawn")) as ex:
        fut = {
            "hd_chain": ex.submit(_job_chain, hd_clip(), 1920, 1080, 16, 16, 2, 3, 0.0, 1, True),
            "hd_64": ex.submit(_job_pair, hd_pair(64), 1920, 1080, 16, 64, 2),
            "hd_128": ex.submit(_job_pair, hd_pair(128), 1920, 1080, 16, 128, 2),
            "uhd_16": ex.submit(_job_pair, uhd_clip(), 3840, 2160, 16, 16, 0),
            "uhd_128": ex.submit(_job_pair, uhd_clip(), 3840, 2160, 16, 128, 0),
            "cif4": ex.submit(_job_chain, cif4_clip(), 704, 576, 16, 8, 1, 5, 0.0, 0, False),
        }
        out = {k: f.result() for k, f in fut.items()}
        c4 = out["cif4"]
        sub = {"low_4": c4["low_4"]}
        for t in range(1, 5):
            sub[f"high_{t}"], sub[f"motion_{t}"] = c4[f"high_{t}"], c4[f"motion_filtered_{t}"]
            sub[f"frame_types_{t}"] = c4[f"frame_types_{t}"]
        out["cif4_sub"] = sub
        out["cif4_rec"] = ex.submit(_job_synth, sub, 704, 576, 1, 5, 16, 8, 1, 0.0).result()
    return out


@pytest.fixture()
def fused(ctx):
    """Forces the kernels bench.py times: fused ME and byte-plane decorrelate or fail."""
    ctx.set_me_mode(2)
    ctx.set_mc_mode(2)
    yield ctx
    ctx.set_me_mode(0)
    ctx.set_mc_mode(0)


def _same(name, got, want):
    if isinstance(want, (bytes, bytearray)):
        assert bytes(got) == bytes(want), name
        return
    bad = np.argwhere(np.asarray(got) != np.asarray(want))
    assert bad.size == 0, f"{name}: {len(bad)} values differ, first {bad[:4].tolist()}"


def test_1080p_quarter_pel_tools_match_oracle(fused, ref):
    """(i) levels sr 16 and sr 32 of a 5-frame 1080p clip, tool by tool, incl. prediction."""
    ctx, want, low = fused, ref["hd_chain"], hd_clip()
    X, Y, bs, a = HD["X"], HD["Y"], HD["bs"], HD["a"]
    for s in level_schedule(1, 3, bs, 16, block_size_min=bs):
        t, r = s["t"], s["search_range"]
        even, odd = low[0::2], low[1::2]
        mv = ctx.motion_estimate(even, odd, X, Y, bs, r, a)
        _same(f"motion_{t}", mv, want[f"motion_{t}"])
        high, types, mvf, pred = ctx.decorrelate(even, odd, want[f"motion_{t}"], X, Y, bs, r, a, always_B=1,
                                                 want_prediction=True)
        _same(f"prediction_{t}", pred, want[f"prediction_{t}"])
        _same(f"high_{t}", high, want[f"high_{t}"])
        _same(f"frame_types_{t}", types, want[f"frame_types_{t}"])
        _same(f"motion_filtered_{t}", mvf, want[f"motion_filtered_{t}"])
        back, _ = ctx.correlate(even, high, mvf, types, X, Y, bs, r, a)
        unclamped = (high > 0) & (high < 255)
        assert np.array_equal(back[unclamped], odd[unclamped])  # exact wherever the residue was not clamped
        low = want[f"low_{t}"]


def test_1080p_resident_and_pinned_analysis_match_oracle(fused, ref):
    """(i) the same clip through the whole-sequence paths bench.py times: resident analysis
    (device-resident `value`) and qsvc_analyze on a pinned clip (`e2e`)."""
    ctx, want, clip = fused, ref["hd_chain"], hd_clip()
    X, Y, bs, a = HD["X"], HD["Y"], HD["bs"], HD["a"]
    ctx.resident_load(clip, X, Y)
    ctx.resident_analyze(3, bs, 16, a, 0.0, always_B=1, block_size_min=bs)
    for s in level_schedule(1, 3, bs, 16, block_size_min=bs):
        got = ctx.resident_fetch(s["t"], s["pairs"], bs)
        for name in ("motion", "motion_filtered", "high", "low", "frame_types"):
            _same(f"resident {name}_{s['t']}", got[name], want[f"{name}_{s['t']}"])
    pinned = ctx.host_alloc(clip.shape)
    pinned[:] = clip
    got = ctx.analyze(pinned, X, Y, 1, 3, bs, 16, a, 0.0, always_B=1, block_size_min=bs)
    for t in (1, 2):
        for name in ("motion", "motion_filtered", "high", "low", "frame_types"):
            _same(f"pinned {name}_{t}", got[f"{name}_{t}"], want[f"{name}_{t}"])


@pytest.mark.parametrize("sr", [64, 128])
def test_1080p_large_search_ranges_match_oracle(fused, ref, sr):
    """(i) sr 64 / 128 at quarter-pel: six / seven pyramid levels over 1080 lines (odd 135 at
    depth 3: the descent is not the inverse of the analysis), vectors up to +-511, decorrelate
    borders of 1024 / 2048 up-sampled samples."""
    ctx, want, clip = fused, ref[f"hd_{sr}"], hd_pair(sr)
    X, Y, bs, a = HD["X"], HD["Y"], HD["bs"], HD["a"]
    even, odd = clip[0::2], clip[1::2]
    mv = ctx.motion_estimate(even, odd, X, Y, bs, sr, a)
    _same("motion", mv, want["motion"])
    assert np.abs(want["motion"]).max() > 127  # the case really leaves the int8 range
    high, types, mvf, pred = ctx.decorrelate(even, odd, want["motion"], X, Y, bs, sr, a, always_B=1,
                                             want_prediction=True)
    _same("prediction", pred, want["prediction"])
    _same("high", high, want["high"])
    _same("motion_filtered", mvf, want["motion_filtered"])
    rec, _ = ctx.correlate(even, want["high"], want["motion_filtered"], want["types"], X, Y, bs, sr, a)
    _same("odd", rec, want["odd"])


def test_cfg2_gop_round_trip_matches_oracle(fused, ref):
    """(ii) 704x576 half-pel sr 8: one GOP of 16 through qsvc_analyze, then synthesis."""
    ctx, want, clip = fused, ref["cif4"], cif4_clip()
    X, Y, bs, a = CIF4["X"], CIF4["Y"], CIF4["bs"], CIF4["a"]
    got = ctx.analyze(clip, X, Y, 1, 5, bs, 8, a, 0.0, always_B=0, block_size_min=bs)
    for t in range(1, 5):
        for name in ("motion", "motion_filtered", "high", "low", "frame_types"):
            _same(f"{name}_{t}", got[f"{name}_{t}"], want[f"{name}_{t}"])
    rec = ctx.synthesize(ref["cif4_sub"], X, Y, 1, 5, bs, 8, a, 0.0)
    _same("low_0", rec, ref["cif4_rec"])


@pytest.mark.parametrize("sr", [16, 128])
def test_2160p_integer_pel_matches_oracle(ctx, ref, sr):
    """(iii) cfg4 geometry (a = 0: literal ME, byte-plane decorrelate); sr 128 reaches the
    non-invertible pyramid of 2160 lines (odd 135 at depth 4)."""
    want, clip = ref[f"uhd_{sr}"], uhd_clip()
    X, Y, bs, a = UHD["X"], UHD["Y"], UHD["bs"], UHD["a"]
    even, odd = clip[0::2], clip[1::2]
    _same("motion", ctx.motion_estimate(even, odd, X, Y, bs, sr, a), want["motion"])
    ctx.set_mc_mode(2)
    try:
        high, types, mvf, pred = ctx.decorrelate(even, odd, want["motion"], X, Y, bs, sr, a, always_B=1,
                                                 want_prediction=True)
        rec, _ = ctx.correlate(even, want["high"], want["motion_filtered"], want["types"], X, Y, bs, sr, a)
    finally:
        ctx.set_mc_mode(0)
    _same("prediction", pred, want["prediction"])
    _same("high", high, want["high"])
    _same("odd", rec, want["odd"])


def test_1080p_two_gop_shards_with_tail_relay_match_the_whole_sequence(fused, ref):
    """(iv) 1080 lines, block 16: the 8 uncovered rows chain through the pairs of a level.  The
    5-frame clip as two GOPs of 2 (TRLs 2), one shard each, tail state handed left to right,
    equals the single-process level 1 of the oracle; without the hand-over it does not."""
    ctx, want, clip = fused, ref["hd_chain"], hd_clip()
    X, Y, bs, a = HD["X"], HD["Y"], HD["bs"], HD["a"]
    kw = dict(block_size=bs, search_range=16, subpixel_accuracy=a, update_factor=0.0, always_B=1,
              block_size_min=bs)

    def run(relay):
        parts = []
        for r in range(2):
            parts.append(shard.analyze_shard(ctx, clip, X, Y, 2, 2, r, 2, relay=relay, **kw))
            if hasattr(relay, "next_shard"):
                relay.next_shard()
        return shard.gather(parts, 2)

    got = run(shard.LocalTailRelay())
    for name in ("motion", "motion_filtered", "high", "low", "frame_types"):
        _same(f"sharded {name}_1", got[f"{name}_1"], want[f"{name}_1"])
    naive = run(lambda level, synthesis, phase, state: False)
    assert not np.array_equal(naive["high_1"], want["high_1"])
    assert np.array_equal(naive["high_1"][0], want["high_1"][0])  # the first shard needs nothing


def test_pinned_upload_overlap_does_not_race_the_decorrelate(fused, ref):
    """(v) qsvc_analyze uploads a pinned clip GOP by GOP on the copy stream while level 1's
    motion estimation starts on the ME lane; the decorrelate lane must not read the clip before
    it has landed.  The context is first poisoned with a different clip of the same geometry so
    that a premature read sees wrong pixels."""
    from oracle import oracle as orc
    ctx = fused
    X, Y, bs, a, GOPs, TRLs = 704, 576, 16, 1, 4, 3
    frames = GOPs * 2 ** (TRLs - 1) + 1
    clip = yuv.synthetic_clip(X, Y, frames, 5, max_shift=16)
    poison = yuv.synthetic_clip(X, Y, frames, 6, max_shift=16)
    want = orc.analyze(clip, X, Y, TRLs, bs, 8, a, 0.0, block_size_min=bs)
    pinned = ctx.host_alloc(clip.shape)
    for _ in range(3):
        pinned[:] = poison
        ctx.analyze(pinned, X, Y, GOPs, TRLs, bs, 8, a, 0.0, block_size_min=bs)
        pinned[:] = clip
        got = ctx.analyze(pinned, X, Y, GOPs, TRLs, bs, 8, a, 0.0, block_size_min=bs)
        for t in range(1, TRLs):
            for name in ("motion", "motion_filtered", "high", "low", "frame_types"):
                _same(f"{name}_{t}", got[f"{name}_{t}"], want[f"{name}_{t}"])


def _random_update_case(X, Y, bs, n, reach, seed, types):
    """Random frames, residues and motion fields with vectors of up to `reach` "pixels" (update applies
    the quarter-pel vectors of the upper levels as whole pixels, A.4): most blocks near the picture
    borders fold onto the edge rows / columns, whose targets then take long ordered chains."""
    rng = np.random.default_rng(seed)
    fb = X * Y * 3 // 2
    frames = rng.integers(0, 256, (n + 1, fb), dtype=np.uint8)
    # residues around 128 with a heavy tail, so that chains really saturate at both ends
    high = np.clip(rng.normal(128, 40, (n, fb)), 0, 255).astype(np.uint8)
    mv = rng.integers(-reach, reach + 1, (n, 4, Y // bs, X // bs)).astype(np.int16)
    mv[0, :, ::2] //= 8  # one field with moderate vectors in every other block row: short lists and long ones
    return frames, high, mv, types.encode()


@pytest.mark.parametrize("uf", [0.25, 0.5, 1.0, 0.3])
@pytest.mark.parametrize("X,Y,reach", [(640, 352, 511), (320, 176, 127), (1920, 1080, 511)])
def test_update_with_long_vectors_matches_oracle(ctx, X, Y, reach, uf):
    """Tile lists that overflow into the ordered scan, targets on the picture edge that collect
    thousands of contributions (the composed saturating adds of the dyadic kernel), 'I' pairs."""
    from oracle import oracle as orc
    orc.build()
    n = 2 if X > 1000 else 4
    frames, high, mv, types = _random_update_case(X, Y, 16, n, reach, 1000 + X + reach, "BIBB"[:n] if n > 2 else "BB")
    want = orc.update(frames, high, mv, types, X, Y, 16, uf)
    _same(f"update uf={uf}", ctx.update(frames, high, mv, types, X, Y, 16, uf), want)
    back = orc.update(want, high, mv, types, X, Y, 16, uf, inverse=True)
    _same(f"un_update uf={uf}", ctx.un_update(want, high, mv, types, X, Y, 16, uf), back)


@pytest.mark.parametrize("uf", [0.25, 0.3])
@pytest.mark.parametrize("X,Y", [(640, 352), (1920, 1080)])
def test_update_with_converging_vectors_matches_oracle(ctx, X, Y, uf):
    """Every block displaced onto (almost) the same spot inside the picture: tiles far from the edges whose
    block lists overflow, hundreds of ordered contributions per target that is NOT on the picture edge."""
    from oracle import oracle as orc
    orc.build()
    bs, n = 16, 2
    rng = np.random.default_rng(X + int(uf * 100))
    fb = X * Y * 3 // 2
    frames = rng.integers(0, 256, (n + 1, fb), dtype=np.uint8)
    high = np.clip(rng.normal(128, 12, (n, fb)), 0, 255).astype(np.uint8)
    by, bx = np.mgrid[0:Y // bs, 0:X // bs]
    mv = np.zeros((n, 4, Y // bs, X // bs), np.int16)
    for i in range(n):
        for d, (cy, cx) in enumerate([(Y // 2 - 40, X // 2 + 24), (Y // 3, X // 4)]):
            jy, jx = rng.integers(-6, 7, by.shape), rng.integers(-6, 7, bx.shape)
            mv[i, 2 * d] = np.clip(cx - bx * bs + jx, -511, 511)      # x component
            mv[i, 2 * d + 1] = np.clip(cy - by * bs + jy, -511, 511)  # y component
    want = orc.update(frames, high, mv, b"BB", X, Y, bs, uf)
    _same(f"update uf={uf}", ctx.update(frames, high, mv, b"BB", X, Y, bs, uf), want)
    back = orc.update(want, high, mv, b"BB", X, Y, bs, uf, inverse=True)
    _same(f"un_update uf={uf}", ctx.un_update(want, high, mv, b"BB", X, Y, bs, uf), back)
