"""CPU: the oracle (oracle/mctf_oracle.c) against the committed golden vectors
produced by the unmodified reference tools (tests/golden, oracle/make_golden.py)."""
import numpy as np
import pytest

from golden_util import NAMES, load, schedule
from oracle import oracle as orc


@pytest.mark.parametrize("name", NAMES)
def test_oracle_analysis_matches_reference(name):
    g = load(name)
    X, Y, bs, a, uf = g["X"], g["Y"], g["bs"], g["a"], g["uf"]
    low = g["low_0"]
    for t, sr in schedule(g):
        even, odd = low[0::2], low[1::2]
        mv = orc.motion_estimate(even, odd, X, Y, bs, sr, a)
        assert np.array_equal(mv, g[f"motion_{t}"]), f"motion_{t}"
        high, types, mvf, pred, rc = orc.decorrelate(even, odd, mv, X, Y, bs, sr, a, g["ov"], g["always_B"])
        assert rc == 0
        assert types == bytes(g[f"frame_types_{t}"])
        assert np.array_equal(pred, g[f"prediction_even_{t}"]), f"prediction_even_{t}"
        assert np.array_equal(high, g[f"high_{t}"]), f"high_{t}"
        assert np.array_equal(mvf, g[f"motion_filtered_{t}"])
        low = orc.update(even, high, mvf, types, X, Y, bs, uf)
        assert np.array_equal(low, g[f"low_{t}"]), f"low_{t}"


@pytest.mark.parametrize("name", NAMES)
def test_oracle_synthesis_matches_reference(name):
    g = load(name)
    X, Y, bs, a, uf, T = g["X"], g["Y"], g["bs"], g["a"], g["uf"], g["TRLs"]
    low = g[f"low_{T-1}"]
    for t, sr in reversed(schedule(g)):
        types = bytes(g[f"frame_types_{t}"])
        mv = g[f"motion_filtered_{t}"]
        even = orc.update(low, g[f"high_{t}"], mv, types, X, Y, bs, uf, inverse=True)
        assert np.array_equal(even, g[f"syn_even_{t}"]), f"even_{t}"
        odd, _ = orc.correlate(even, g[f"high_{t}"], mv, types, X, Y, bs, sr, a, g["ov"])
        assert np.array_equal(odd, g[f"syn_odd_{t}"]), f"odd_{t}"
        low = np.empty((2 * odd.shape[0] + 1, even.shape[1]), np.uint8)
        low[0::2], low[1::2] = even, odd
    assert np.array_equal(low, g["syn_low_0"])


def test_entropy_matches_reference_expression():
    # uniform over 4 symbols -> exactly 2 bits; empty bins are skipped
    count = np.zeros(256, np.int32)
    count[[3, 9, 77, 200]] = 5
    assert orc.entropy(count) == 2.0
    count[:] = 0
    count[0] = 10
    assert orc.entropy(count) == 0.0


def test_dwt53_round_trip_and_layout():
    rng = np.random.default_rng(0)
    for (y, x) in [(16, 24), (15, 24), (16, 23), (7, 9), (2, 2), (3, 5)]:
        img = rng.integers(-300, 600, size=(y, x)).astype(np.int16)
        a = orc.dwt53(img, 1)
        assert np.array_equal(orc.dwt53(a, 1, synth=True), img)  # lifting is exactly invertible
    # constant image: all high bands vanish, low band keeps the constant
    img = np.full((8, 8), 37, np.int16)
    a = orc.dwt53(img, 2)
    assert (a[:2, :2] == 37).all() and (a[4:, :] == 0).all() and (a[:, 4:] == 0).all()


def test_oracle_synthesis_with_per_level_geometry_matches_reference():
    """synthesize.py:127-133 with lists that vary per level: each level's
    un_update / correlate run with that level's own picture size and sub-pixel accuracy, the
    merged frames are re-read with the next level's geometry (fixture made by the unmodified
    reference tools, oracle/make_golden.py make_level_lists)."""
    from golden_util import load_level_lists
    g = load_level_lists()
    low, sr = g["low_3"], {1: g["sr"], 2: 2 * g["sr"], 3: 4 * g["sr"]}
    for t in (3, 2, 1):
        X, Y, a = g["geo"][t]
        types = bytes(g[f"frame_types_{t}"])
        low = low.reshape(-1, X * Y * 3 // 2)
        even = orc.update(low, g[f"high_{t}"], g[f"motion_{t}"], types, X, Y, g["bs"], g["uf"], inverse=True)
        odd, _ = orc.correlate(even, g[f"high_{t}"], g[f"motion_{t}"], types, X, Y, g["bs"], sr[t], a)
        low = np.empty((2 * odd.shape[0] + 1, even.shape[1]), np.uint8)
        low[0::2], low[1::2] = even, odd
    assert np.array_equal(low, g["syn_low_0"])
