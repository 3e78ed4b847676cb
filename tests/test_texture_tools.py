"""The step after the hot path on the texture side: component demux for
the per-component coders and the distortion meter behind psnr.py.  Both are external programs
in the reference (sources not in its tree), so the checks are against numpy restatements of the
contract their call sites spell out (texture_compress_fb_j2k.py:155-163,255-257; psnr.py:78-90)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from qsvc_b200 import yuv  # noqa: E402


def _run(tool, args, cwd, stdin=None):
    env = dict(os.environ, PYTHONPATH=ROOT)
    return subprocess.run([os.path.join(ROOT, "bin", tool)] + args, cwd=cwd, env=env, input=stdin,
                          capture_output=True)


def test_demux_pulls_one_component_out_of_every_picture(tmp_path):
    X, Y, n = 64, 48, 5
    clip = yuv.synthetic_clip(X, Y, n, 3)
    Ysz, Csz = X * Y, X * Y // 4
    for off, ln in ((0, Ysz), (Ysz, Csz), (Ysz + Csz, Csz)):
        r = _run("demux", [str(Ysz + 2 * Csz), str(off), str(ln)], str(tmp_path), stdin=clip.tobytes())
        assert r.returncode == 0, r.stderr
        assert r.stdout == clip[:, off:off + ln].tobytes()
    assert _run("demux", ["1"], str(tmp_path), stdin=b"").returncode == 1


@pytest.mark.gpu
def test_sse_and_psnr_match_numpy(ctx):
    rng = np.random.default_rng(5)
    fb = 1920 * 1080 * 3 // 2
    a = rng.integers(0, 256, size=3 * fb + 1000, dtype=np.uint8)
    b = np.clip(a.astype(np.int32) + rng.integers(-9, 10, size=a.size), 0, 255).astype(np.uint8)
    b[fb:2 * fb] = a[fb:2 * fb]  # one identical picture
    for block in (fb, 4097, 16):
        n = a.size // block
        d = a[: n * block].astype(np.int64) - b[: n * block]
        want = (d * d).reshape(n, block).sum(axis=1).astype(np.uint64)
        assert np.array_equal(ctx.sse(a, b, block), want)
    per, tot = ctx.psnr(a, b, fb)
    assert np.isinf(per[1]) and abs(tot - 10 * np.log10(255.0 ** 2 * 3 * fb / float(((a[:3 * fb].astype(np.int64) - b[:3 * fb]) ** 2).sum()))) < 1e-9
    # unaligned streams take the byte path
    assert np.array_equal(ctx.sse(a[1:], b[1:], 4096), ((a[1:].astype(np.int64) - b[1:])[: (a.size - 1) // 4096 * 4096] ** 2)
                          .reshape(-1, 4096).sum(axis=1).astype(np.uint64))


@pytest.mark.gpu
def test_snr_tool_prints_the_line_psnr_py_parses(tmp_path):
    X, Y, n = 64, 48, 4
    a = yuv.synthetic_clip(X, Y, n, 8)
    b = a.copy()
    b[:, ::7] ^= 3
    yuv.write_frames(str(tmp_path / "orig"), a)
    yuv.write_frames(str(tmp_path / "low_0"), b)
    fb = a.shape[1]
    r = _run("snr", ["--type=uchar", "--peak=255", "--file_A=orig", "--file_B=low_0", f"--block_size={fb}"],
             str(tmp_path))
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.decode().splitlines() if "PSNR" in ln and "dB" in ln]  # | grep PSNR | grep dB
    got = float("\n".join(lines).split(" ")[2])                                         # psnr.py:88
    mse = ((a.astype(np.float64) - b) ** 2).mean()
    assert abs(got - 10 * np.log10(255.0 ** 2 / mse)) < 1e-4
