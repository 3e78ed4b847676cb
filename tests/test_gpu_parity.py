"""Parity of the CUDA path (through the C ABI) against the CPU oracle.

Bit-exact: motion fields, high/low/odd/even frames, frame types, predictions.
Geometries follow SURVEY.md Appendix C item 10: each one triggers a distinct
quirk of the reference (heap aliasing, border pollution at sub-pixel levels, odd
pyramid sizes, heights that are not a multiple of the block size, I/B decision,
update scatter order).
"""
import numpy as np
import pytest

from oracle import oracle as orc
from qsvc_b200 import yuv
from qsvc_b200.mctf import level_schedule

pytestmark = pytest.mark.gpu

#        X    Y   GOPs TRLs bs  sr  a  uf    flat
CASES = [
    (352, 288, 1, 5, 16, 4, 0, 0.0, 0),    # cfg1 geometry; sr reaches 32 (heap alias at a=0)
    (128, 96, 1, 4, 16, 8, 2, 0.25, 0),    # quarter-pel: size-field read + border pollution
    (128, 96, 1, 3, 16, 16, 1, 0.3, 0),    # half-pel, non-dyadic update factor
    (128, 120, 1, 4, 8, 32, 0, 0.0, 0),    # odd pyramid sizes, sr 32/64/128: carried reference[0]
    (128, 72, 2, 5, 16, 4, 2, 0.25, 0),    # height % block_size != 0: chained prediction tail
    (128, 96, 2, 3, 16, 4, 0, 0.25, 2),    # IBIB frame types
    (96, 64, 1, 3, 16, 6, 1, 0.25, 0),     # search range that is not a power of two
    (80, 48, 1, 3, 16, 16, 2, 0.25, 0),    # border larger than the picture
    (176, 144, 2, 3, 32, 4, 1, 0.5, 3),    # X % bs != 0 (uncovered columns)
    (352, 288, 1, 3, 16, 8, 2, 0.0, 0),    # CIF quarter-pel: mostly fast-path blocks in the fused ME
    (320, 192, 1, 4, 16, 16, 1, 0.0, 0),   # half-pel, sr up to 64
]


def clip_for(X, Y, GOPs, TRLs, sr, flat, seed=7):
    frames = GOPs * 2 ** (TRLs - 1) + 1
    return yuv.synthetic_clip(X, Y, frames, seed, max_shift=min(48, 3 * sr), flat_every=flat)


@pytest.mark.parametrize("X,Y,GOPs,TRLs,bs,sr,a,uf,flat", CASES)
def test_tools_match_oracle(ctx, X, Y, GOPs, TRLs, bs, sr, a, uf, flat):
    """Each tool on its own, levels chained through the oracle's outputs."""
    low = clip_for(X, Y, GOPs, TRLs, sr, flat)
    for s in level_schedule(GOPs, TRLs, bs, sr, block_size_min=bs):
        even, odd = low[0::2], low[1::2]
        b, r = s["block_size"], s["search_range"]
        mv_o = orc.motion_estimate(even, odd, X, Y, b, r, a)
        mv_g = ctx.motion_estimate(even, odd, X, Y, b, r, a)
        assert np.array_equal(mv_g, mv_o), f"motion_{s['t']}: {(mv_g != mv_o).sum()} components differ"
        high_o, types_o, mvf_o, pred_o, rc = orc.decorrelate(even, odd, mv_o, X, Y, b, r, a)
        assert rc == 0
        high_g, types_g, mvf_g, pred_g = ctx.decorrelate(even, odd, mv_o, X, Y, b, r, a,
                                                         want_prediction=True)
        assert types_g == types_o
        assert np.array_equal(pred_g, pred_o), f"prediction_{s['t']}: {(pred_g != pred_o).sum()} samples differ"
        assert np.array_equal(high_g, high_o), f"high_{s['t']}: {(high_g != high_o).sum()} samples differ"
        assert np.array_equal(mvf_g, mvf_o)
        low_o = orc.update(even, high_o, mvf_o, types_o, X, Y, b, uf)
        low_g = ctx.update(even, high_o, mvf_o, types_o, X, Y, b, uf)
        assert np.array_equal(low_g, low_o), f"low_{s['t']}: {(low_g != low_o).sum()} samples differ"
        # inverse tools on the same level
        even_o = orc.update(low_o, high_o, mvf_o, types_o, X, Y, b, uf, inverse=True)
        even_g = ctx.un_update(low_o, high_o, mvf_o, types_o, X, Y, b, uf)
        assert np.array_equal(even_g, even_o)
        odd_o, _ = orc.correlate(even_o, high_o, mvf_o, types_o, X, Y, b, r, a)
        odd_g, _ = ctx.correlate(even_o, high_o, mvf_o, types_o, X, Y, b, r, a)
        assert np.array_equal(odd_g, odd_o)
        low = low_o


@pytest.mark.parametrize("X,Y,GOPs,TRLs,bs,sr,a,uf,flat", CASES[:6])
def test_resident_analyze_synthesize_match_oracle(ctx, X, Y, GOPs, TRLs, bs, sr, a, uf, flat):
    """Whole-sequence analysis and synthesis with frames resident in HBM."""
    clip = clip_for(X, Y, GOPs, TRLs, sr, flat, seed=11)
    ref = orc.analyze(clip, X, Y, TRLs, bs, sr, a, uf, block_size_min=bs)
    got = ctx.analyze(clip, X, Y, GOPs, TRLs, bs, sr, a, uf, block_size_min=bs)
    for t in range(1, TRLs):
        for name in ("motion", "motion_filtered", "high", "low"):
            assert np.array_equal(got[f"{name}_{t}"], ref[f"{name}_{t}"]), f"{name}_{t}"
        assert got[f"frame_types_{t}"] == ref[f"frame_types_{t}"]
    # synthesis from the analysis outputs (motion_filtered_t plays motion_t)
    sub = {f"low_{TRLs-1}": ref[f"low_{TRLs-1}"]}
    low = ref[f"low_{TRLs-1}"]
    sched = level_schedule(GOPs, TRLs, bs, sr, block_size_min=bs)
    for s in reversed(sched):
        t = s["t"]
        sub[f"high_{t}"], sub[f"motion_{t}"] = ref[f"high_{t}"], ref[f"motion_filtered_{t}"]
        sub[f"frame_types_{t}"] = ref[f"frame_types_{t}"]
        even = orc.update(low, sub[f"high_{t}"], sub[f"motion_{t}"], sub[f"frame_types_{t}"], X, Y,
                          bs, uf, inverse=True)
        odd, _ = orc.correlate(even, sub[f"high_{t}"], sub[f"motion_{t}"], sub[f"frame_types_{t}"],
                               X, Y, bs, s["search_range"], a)
        low = np.empty((2 * odd.shape[0] + 1, even.shape[1]), np.uint8)
        low[0::2], low[1::2] = even, odd
    rec = ctx.synthesize(sub, X, Y, GOPs, TRLs, bs, sr, a, uf)
    assert np.array_equal(rec, low)


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("X,Y,GOPs,TRLs,bs,sr,a,uf,flat", [c for c in CASES if c[6] in (1, 2) and c[4] << c[6] <= 64])
def test_me_literal_and_fused_paths_agree_with_oracle(ctx, mode, X, Y, GOPs, TRLs, bs, sr, a, uf, flat):
    """mode 1: literal path (materialised up-sampled images); mode 2: fused sub-pixel
    path (u8 planes + packed-byte SAD + exact generator for polluted/edge blocks)."""
    from qsvc_b200._lib import QsvcError
    low = clip_for(X, Y, GOPs, TRLs, sr, flat, seed=13)
    ctx.set_me_mode(mode)
    try:
        for s in level_schedule(GOPs, TRLs, bs, sr, block_size_min=bs):
            even, odd = low[0::2], low[1::2]
            mv_o = orc.motion_estimate(even, odd, X, Y, bs, s["search_range"], a)
            try:
                mv_g = ctx.motion_estimate(even, odd, X, Y, bs, s["search_range"], a)
            except QsvcError:
                assert mode == 2  # geometry outside the fused path's domain: refused, not wrong
                continue
            bad = np.argwhere(mv_g != mv_o)
            assert bad.size == 0, f"motion_{s['t']}: {len(bad)} components differ, first {bad[:5].tolist()}"
            low = even
    finally:
        ctx.set_me_mode(0)


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("X,Y,GOPs,TRLs,bs,sr,a,uf,flat", CASES)
def test_mc_literal_and_fused_paths_agree_with_oracle(ctx, mode, X, Y, GOPs, TRLs, bs, sr, a, uf, flat):
    """decorrelate / correlate: mode 1 materialises int16 planes like the reference,
    mode 2 runs on byte planes (k_mc_march + chained tail rows)."""
    from qsvc_b200._lib import QsvcError
    low = clip_for(X, Y, GOPs, TRLs, sr, flat, seed=17)
    ctx.set_mc_mode(mode)
    try:
        for s in level_schedule(GOPs, TRLs, bs, sr, block_size_min=bs):
            even, odd = low[0::2], low[1::2]
            r = s["search_range"]
            mv = orc.motion_estimate(even, odd, X, Y, bs, r, a)
            high_o, types_o, mvf_o, pred_o, rc = orc.decorrelate(even, odd, mv, X, Y, bs, r, a)
            try:
                high_g, types_g, mvf_g, pred_g = ctx.decorrelate(even, odd, mv, X, Y, bs, r, a,
                                                                 want_prediction=True)
            except QsvcError:
                assert mode == 2
                continue
            assert types_g == types_o
            bad = np.argwhere(pred_g != pred_o)
            assert bad.size == 0, f"prediction_{s['t']}: {len(bad)} differ, first {bad[:4].tolist()}"
            assert np.array_equal(high_g, high_o) and np.array_equal(mvf_g, mvf_o)
            odd_o, _ = orc.correlate(even, high_o, mvf_o, types_o, X, Y, bs, r, a)
            odd_g, _ = ctx.correlate(even, high_o, mvf_o, types_o, X, Y, bs, r, a)
            assert np.array_equal(odd_g, odd_o)
            low = even
    finally:
        ctx.set_mc_mode(0)


#             X    Y  bs sr  a  ov
OBMC_CASES = [
    (96, 64, 16, 4, 0, 2),     # one transform level on 20 x 20 blocks
    (64, 48, 16, 4, 1, 2),     # half-pel: two levels on 40 x 40 blocks
    (64, 64, 16, 4, 2, 4),     # quarter-pel, overlap 4: four levels on 96 x 96 blocks
    (64, 48, 16, 8, 1, 3),     # overlap that is not a power of two (odd sub-band sizes)
    (128, 64, 32, 4, 1, 2),    # 32 x 32 blocks
]


@pytest.mark.parametrize("X,Y,bs,sr,a,ov", OBMC_CASES)
def test_overlapped_block_prediction_matches_oracle(ctx, X, Y, bs, sr, a, ov):
    """decorrelate / correlate with --block_overlaping > 0 (decorrelate.cpp:84-88, 99-172):
    per-block 5/3 analysis of the extended block, sub-band scatter, picture synthesis."""
    clip = yuv.synthetic_clip(X, Y, 5, 29, max_shift=min(24, 3 * sr))
    even, odd = clip[0::2], clip[1::2]
    mv = orc.motion_estimate(even, odd, X, Y, bs, sr, a)
    high_o, types_o, mvf_o, pred_o, rc = orc.decorrelate(even, odd, mv, X, Y, bs, sr, a, ov)
    assert rc == 0
    high_g, types_g, mvf_g, pred_g = ctx.decorrelate(even, odd, mv, X, Y, bs, sr, a, block_overlaping=ov,
                                                     want_prediction=True)
    assert types_g == types_o
    bad = np.argwhere(pred_g != pred_o)
    assert bad.size == 0, f"prediction: {len(bad)} differ, first {bad[:4].tolist()}"
    assert np.array_equal(high_g, high_o) and np.array_equal(mvf_g, mvf_o)
    odd_o, _ = orc.correlate(even, high_o, mvf_o, types_o, X, Y, bs, sr, a, ov)
    odd_g, _ = ctx.correlate(even, high_o, mvf_o, types_o, X, Y, bs, sr, a, block_overlaping=ov)
    assert np.array_equal(odd_g, odd_o)


#                   X    Y  bs sr  a  ov
OBMC_RAGGED_CASES = [
    (96, 72, 16, 4, 0, 2),     # Y % bs != 0: uncovered rows in every sub-band go through the picture synthesis
    (88, 64, 16, 4, 1, 2),     # X % bs != 0
    (88, 72, 16, 4, 2, 4),     # both, quarter-pel, four levels
    (104, 56, 16, 4, 1, 3),    # overlap that is not a power of two
]


@pytest.mark.parametrize("X,Y,bs,sr,a,ov", OBMC_RAGGED_CASES)
def test_overlapped_prediction_on_ragged_pictures_matches_oracle(ctx, X, Y, bs, sr, a, ov):
    """Areas no block covers carry the previous pair's analysed leftovers through the picture
    synthesis (SURVEY.md A.2.6 with A.2.2): four pairs in one call, so that the carried buffer
    matters, analysis and synthesis."""
    clip = yuv.synthetic_clip(X, Y, 9, 37, max_shift=min(24, 3 * sr))
    even, odd = clip[0::2], clip[1::2]
    mv = orc.motion_estimate(even, odd, X, Y, bs, sr, a)
    high_o, types_o, mvf_o, pred_o, rc = orc.decorrelate(even, odd, mv, X, Y, bs, sr, a, ov)
    assert rc == 0
    high_g, types_g, mvf_g, pred_g = ctx.decorrelate(even, odd, mv, X, Y, bs, sr, a, block_overlaping=ov,
                                                     want_prediction=True)
    assert types_g == types_o
    bad = np.argwhere(pred_g != pred_o)
    assert bad.size == 0, f"prediction: {len(bad)} differ, first {bad[:4].tolist()}"
    assert np.array_equal(high_g, high_o) and np.array_equal(mvf_g, mvf_o)
    odd_o, _ = orc.correlate(even, high_o, mvf_o, types_o, X, Y, bs, sr, a, ov)
    odd_g, _ = ctx.correlate(even, high_o, mvf_o, types_o, X, Y, bs, sr, a, block_overlaping=ov)
    assert np.array_equal(odd_g, odd_o)


def test_overlapped_ragged_analysis_chain_matches_oracle(ctx):
    """The same through the whole resident analysis and synthesis (update included)."""
    X, Y, GOPs, TRLs, bs, sr, a, uf, ov = 88, 72, 1, 4, 16, 4, 1, 0.25, 2
    clip = yuv.synthetic_clip(X, Y, GOPs * 2 ** (TRLs - 1) + 1, 41, max_shift=12)
    ref = orc.analyze(clip, X, Y, TRLs, bs, sr, a, uf, block_overlaping=ov, block_size_min=bs)
    got = ctx.analyze(clip, X, Y, GOPs, TRLs, bs, sr, a, uf, block_overlaping=ov, block_size_min=bs)
    assert any(k.startswith("high_") for k in got)
    for k, g in got.items():
        v = ref[k]
        if isinstance(v, (bytes, bytearray)):
            assert bytes(g) == bytes(v), k
        else:
            assert np.array_equal(np.asarray(g), v), k


def test_gop_shards_with_tail_exchange_match_the_whole_sequence(ctx):
    """Height % block_size != 0: GOP shards that hand the
    prediction tail state from left to right reproduce the single-process result, analysis
    and synthesis; without the hand-over the later shard's bottom rows differ."""
    from qsvc_b200 import shard
    X, Y, GOPs, TRLs, bs, sr, a = 128, 72, 2, 4, 16, 4, 2
    clip = yuv.synthetic_clip(X, Y, GOPs * 2 ** (TRLs - 1) + 1, 23, max_shift=12)
    ref = orc.analyze(clip, X, Y, TRLs, bs, sr, a, 0.0, block_size_min=bs)
    kw = dict(block_size=bs, search_range=sr, subpixel_accuracy=a, update_factor=0.0, block_size_min=bs)

    def run(relay):
        parts = []
        for r in range(2):
            parts.append(shard.analyze_shard(ctx, clip, X, Y, GOPs, TRLs, r, 2, relay=relay, **kw))
            if hasattr(relay, "next_shard"):
                relay.next_shard()
        return shard.gather(parts, TRLs)

    got = run(shard.LocalTailRelay())
    for k, v in got.items():
        assert (np.array_equal(ref[k], v) if not isinstance(v, bytes) else ref[k] == v), k
    naive = run(lambda level, synthesis, phase, state: False)
    assert any(not np.array_equal(naive[f"high_{t}"], ref[f"high_{t}"]) for t in range(1, TRLs))

    sub = {f"low_{TRLs-1}": ref[f"low_{TRLs-1}"]}
    for t in range(1, TRLs):
        sub[f"high_{t}"], sub[f"motion_{t}"] = ref[f"high_{t}"], ref[f"motion_filtered_{t}"]
        sub[f"frame_types_{t}"] = ref[f"frame_types_{t}"]
    whole = ctx.synthesize(sub, X, Y, GOPs, TRLs, bs, sr, a, 0.0)
    relay = shard.LocalTailRelay()
    parts = []
    for r in range(2):
        parts.append(shard.synthesize_shard(ctx, sub, X, Y, GOPs, TRLs, r, 2, block_size=bs, search_range=sr,
                                            subpixel_accuracy=a, update_factor=0.0, relay=relay))
        relay.next_shard()
    assert np.array_equal(shard.gather_frames(parts), whole)


def test_first_pair_flag_for_gop_shards(ctx):
    """A later GOP shard must start from the carried (non-restored) reference[0]
    when the pyramid is not perfectly reconstructing."""
    X, Y, bs, sr = 128, 120, 8, 64
    clip = yuv.synthetic_clip(X, Y, 9, 5, max_shift=40)
    even, odd = clip[0::2], clip[1::2]
    full = orc.motion_estimate(even, odd, X, Y, bs, sr, 0)
    part = ctx.motion_estimate(even[2:], odd[2:], X, Y, bs, sr, 0, first_pair_is_global_first=False)
    assert np.array_equal(part, full[2:])
    part_fresh = ctx.motion_estimate(even[2:], odd[2:], X, Y, bs, sr, 0)
    assert np.array_equal(part_fresh, orc.motion_estimate(even[2:], odd[2:], X, Y, bs, sr, 0))


def test_domain_error_on_large_vectors(ctx):
    """always_B=0 with |mv| > 127 is undefined behaviour in the reference; the
    library must refuse loudly instead of inventing a result."""
    from qsvc_b200._lib import QsvcError, QSVC_EDOMAIN
    X, Y, bs = 64, 48, 16
    clip = yuv.synthetic_clip(X, Y, 3, 1)
    mv = np.zeros((1, 4, Y // bs, X // bs), np.int16)
    mv[0, 0, 0, 0] = 200
    with pytest.raises(QsvcError) as e:
        ctx.decorrelate(clip[0::2], clip[1::2], mv, X, Y, bs, 64, 0, always_B=0)
    assert e.value.code == QSVC_EDOMAIN
    high, types, mvo, _ = ctx.decorrelate(clip[0::2], clip[1::2], mv, X, Y, bs, 64, 0, always_B=1)
    ref = orc.decorrelate(clip[0::2], clip[1::2], mv, X, Y, bs, 64, 0, always_B=1)
    assert np.array_equal(high, ref[0]) and types == ref[1] == b"B"


@pytest.mark.parametrize("GOPs,shards", [(2, 2), (3, 3)])
def test_gop_shards_with_boundary_exchange_match_the_whole_sequence(GOPs, shards):
    """update_factor != 0: the frame two shards share takes the left
    shard's NEXT update, then the right shard's PREV update.  Shards run concurrently (one
    context per thread on this GPU) and hand the int16 planes right / the finished frame left."""
    import threading
    from qsvc_b200 import shard
    from qsvc_b200.mctf import Context
    X, Y, TRLs, bs, sr, a, uf = 128, 96, 4, 16, 4, 1, 0.25
    clip = yuv.synthetic_clip(X, Y, GOPs * 2 ** (TRLs - 1) + 1, 37, max_shift=12)
    ref = orc.analyze(clip, X, Y, TRLs, bs, sr, a, uf, block_size_min=bs)
    assert any(b"B" in ref[f"frame_types_{t}"] for t in range(1, TRLs))
    kw = dict(block_size=bs, search_range=sr, subpixel_accuracy=a, update_factor=uf)

    def run_threads(fn):
        out, err = [None] * shards, []

        def work(r):
            try:
                with Context(0) as c:
                    out[r] = fn(c, r)
            except Exception as e:  # noqa: BLE001
                err.append(e)

        th = [threading.Thread(target=work, args=(r,)) for r in range(shards)]
        for t in th:
            t.start()
        for t in th:
            t.join(timeout=300)
        assert not err, err
        return out

    relays = shard.ThreadBoundaryRelay.make(shards)
    parts = run_threads(lambda c, r: shard.analyze_shard(c, clip, X, Y, GOPs, TRLs, r, shards, block_size_min=bs,
                                                         boundary_relay=relays[r], **kw))
    got = shard.gather(parts, TRLs)
    for k, v in got.items():
        assert (np.array_equal(ref[k], v) if not isinstance(v, bytes) else ref[k] == v), k
    naive = run_threads(lambda c, r: shard.analyze_shard(c, clip, X, Y, GOPs, TRLs, r, shards, block_size_min=bs,
                                                         allow_inexact=True, **kw))
    naive = shard.gather(naive, TRLs)
    assert any(not np.array_equal(naive[f"low_{t}"], ref[f"low_{t}"]) for t in range(1, TRLs))

    sub = {f"low_{TRLs-1}": ref[f"low_{TRLs-1}"]}
    for t in range(1, TRLs):
        sub[f"high_{t}"], sub[f"motion_{t}"] = ref[f"high_{t}"], ref[f"motion_filtered_{t}"]
        sub[f"frame_types_{t}"] = ref[f"frame_types_{t}"]
    with Context(0) as c:
        whole = c.synthesize(sub, X, Y, GOPs, TRLs, bs, sr, a, uf)
    relays = shard.ThreadBoundaryRelay.make(shards)
    parts = run_threads(lambda c, r: shard.synthesize_shard(c, sub, X, Y, GOPs, TRLs, r, shards,
                                                            boundary_relay=relays[r], **kw))
    assert np.array_equal(shard.gather_frames(parts), whole)
