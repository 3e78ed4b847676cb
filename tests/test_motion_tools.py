"""Motion-field (de)correlation tools: the oracle against the reference
binaries run live (CPU), the CUDA path against the oracle through the C ABI, the flag-compatible
command-line tools, and the device-resident motion_residue of a whole analysis."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc, run_ref
from qsvc_b200 import yuv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MCTF = os.path.join(ROOT, "bin", "mctf")


def fields(n, by, bx, seed, big=False):
    r = np.random.default_rng(seed)
    lim = 32000 if big else 130
    return r.integers(-lim, lim + 1, size=(n, 4, by, bx), dtype=np.int64).astype(np.int16)


# n_fields, n_reference (None: no reference file), fields_in_predicted, by, bx
INTERLEVEL = [(8, 4, 8, 3, 4), (7, 4, 7, 2, 5), (8, 2, 8, 3, 3), (8, None, 8, 2, 2), (8, 4, 3, 2, 2),
              (1, 1, 1, 67, 120)]


@pytest.mark.skipif(not run_ref.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("n,nref,fip,by,bx", INTERLEVEL)
def test_oracle_interlevel_matches_live_reference(tmp_path, n, nref, fip, by, bx):
    pred = fields(n, by, bx, 1)
    ref = None if nref is None else fields(nref, by, bx, 2, big=True)
    yuv.write_motion(str(tmp_path / "pred"), pred)
    if ref is not None:
        yuv.write_motion(str(tmp_path / "ref"), ref)
    flags = dict(blocks_in_x=bx, blocks_in_y=by, fields_in_predicted=fip, predicted_fn="pred",
                 reference_fn="ref", residue_fn="res")
    assert run_ref.tool("interlevel_motion_decorrelate", str(tmp_path), **flags) == 0
    res = np.fromfile(str(tmp_path / "res"), "<i2").reshape(-1, 4, by, bx)
    want = orc.interlevel_motion(pred, ref, fip)
    assert np.array_equal(res, want)
    flags["predicted_fn"] = "back"
    assert run_ref.tool("interlevel_motion_correlate", str(tmp_path), **flags) == 0
    back = np.fromfile(str(tmp_path / "back"), "<i2").reshape(-1, 4, by, bx)
    assert np.array_equal(back, orc.interlevel_motion(res, ref, fip, inverse=True))
    assert np.array_equal(back, pred[: len(back)])


@pytest.mark.skipif(not run_ref.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("n,by,bx", [(4, 3, 4), (1, 67, 120), (5, 1, 1)])
def test_oracle_bidirectional_matches_live_reference(tmp_path, n, by, bx):
    f = fields(n, by, bx, 3, big=True)
    yuv.write_motion(str(tmp_path / "in"), f)
    flags = dict(blocks_in_x=bx, blocks_in_y=by, fields=n, input_fn="in", output_fn="out")
    assert run_ref.tool("bidirectional_motion_decorrelate", str(tmp_path), **flags) == 0
    out = np.fromfile(str(tmp_path / "out"), "<i2").reshape(n, 4, by, bx)
    assert np.array_equal(out, orc.bidirectional_motion(f))
    flags.update(input_fn="out", output_fn="back")
    assert run_ref.tool("bidirectional_motion_correlate", str(tmp_path), **flags) == 0
    back = np.fromfile(str(tmp_path / "back"), "<i2").reshape(n, 4, by, bx)
    assert np.array_equal(back, orc.bidirectional_motion(out, inverse=True))
    assert np.array_equal(back, f)


@pytest.mark.gpu
@pytest.mark.parametrize("n,nref,fip,by,bx", INTERLEVEL)
def test_gpu_interlevel_matches_oracle(ctx, n, nref, fip, by, bx):
    pred = fields(n, by, bx, 5, big=True)
    ref = None if nref is None else fields(nref, by, bx, 6, big=True)
    want = orc.interlevel_motion(pred, ref)
    got = ctx.interlevel_motion_decorrelate(pred, ref)
    assert np.array_equal(got, want)
    assert np.array_equal(ctx.interlevel_motion_decorrelate(got, ref, inverse=True), pred)


@pytest.mark.gpu
@pytest.mark.parametrize("n,by,bx", [(4, 3, 4), (16, 67, 120), (5, 1, 1), (0, 3, 4)])
def test_gpu_bidirectional_matches_oracle(ctx, n, by, bx):
    f = fields(n, by, bx, 7, big=True)
    got = ctx.bidirectional_motion_decorrelate(f)
    assert np.array_equal(got, orc.bidirectional_motion(f))
    assert np.array_equal(ctx.bidirectional_motion_decorrelate(got, inverse=True), f)


@pytest.mark.gpu
def test_cli_motion_tools_follow_the_reference_contract(tmp_path):
    """motion_compress.py:141-182 / motion_expand.py:147-179 chained through bin/mctf."""
    by, bx = 4, 6
    lv1, lv2 = fields(4, by, bx, 8), fields(2, by, bx, 9)
    yuv.write_motion(str(tmp_path / "motion_filtered_1"), lv1)
    yuv.write_motion(str(tmp_path / "motion_filtered_2"), lv2)
    env = dict(os.environ, PYTHONPATH=ROOT)

    def mctf(*args):
        r = subprocess.run([MCTF] + list(args), cwd=str(tmp_path), env=env, capture_output=True)
        assert r.returncode == 0, r.stderr.decode()[-2000:]

    geo = [f"--blocks_in_x={bx}", f"--blocks_in_y={by}"]
    mctf("interlevel_motion_decorrelate", *geo, "--fields_in_predicted=4", "--predicted=motion_filtered_1",
         "--reference=motion_filtered_2", "--residue=motion_residue_1")
    mctf("bidirectional_motion_decorrelate", *geo, "--fields=2", "--input=motion_filtered_2",
         "--output=motion_residue_2")
    r1 = np.fromfile(str(tmp_path / "motion_residue_1"), "<i2").reshape(-1, 4, by, bx)
    r2 = np.fromfile(str(tmp_path / "motion_residue_2"), "<i2").reshape(-1, 4, by, bx)
    assert np.array_equal(r1, orc.interlevel_motion(lv1, lv2)) and np.array_equal(r2, orc.bidirectional_motion(lv2))
    mctf("bidirectional_motion_correlate", *geo, "--fields=2", "--input=motion_residue_2", "--output=motion_2")
    mctf("interlevel_motion_correlate", *geo, "--fields_in_predicted=4", "--predicted=motion_1",
         "--reference=motion_2", "--residue=motion_residue_1")
    assert np.array_equal(np.fromfile(str(tmp_path / "motion_2"), "<i2").reshape(-1, 4, by, bx), lv2)
    assert np.array_equal(np.fromfile(str(tmp_path / "motion_1"), "<i2").reshape(-1, 4, by, bx), lv1)
    # a reference file that does not exist reads as zeros (interlevel_motion_decorrelate.cpp:229-238)
    mctf("interlevel_motion_decorrelate", *geo, "--fields_in_predicted=4", "--predicted=motion_filtered_1",
         "--reference=missing", "--residue=plain")
    assert np.array_equal(np.fromfile(str(tmp_path / "plain"), "<i2").reshape(-1, 4, by, bx), lv1)


@pytest.mark.gpu
def test_resident_motion_residue_matches_the_tool_chain(ctx):
    X, Y, GOPs, TRLs, bs, sr, a = 128, 96, 2, 4, 16, 4, 1
    clip = yuv.synthetic_clip(X, Y, GOPs * 2 ** (TRLs - 1) + 1, 41, max_shift=12)
    ref = orc.analyze(clip, X, Y, TRLs, bs, sr, a, 0.0, block_size_min=bs)
    ctx.resident_load(clip, X, Y)
    ctx.resident_analyze(TRLs=TRLs, block_size=bs, search_range=sr, subpixel_accuracy=a, update_factor=0.0,
                         block_size_min=bs)
    by, bx = Y // bs, X // bs
    for t in range(1, TRLs):
        mvf = ref[f"motion_filtered_{t}"]
        want = (orc.bidirectional_motion(mvf) if t == TRLs - 1
                else orc.interlevel_motion(mvf, ref[f"motion_filtered_{t + 1}"]))
        got = ctx.resident_motion_residue(t, mvf.shape[0], by, bx)
        assert np.array_equal(got, want), f"motion_residue_{t}"
