"""File formats of the MCTF hot path and the seeded synthetic clip generator.

Formats (reference trunk/src):
  * frame files (``low_t``, ``even_t``, ``odd_t``, ``high_t``): headerless I420,
    frames concatenated, each frame Y (Y*X) then U, V ((Y/2)*(X/2)) bytes
    (decorrelate.cpp:583-585, split.cpp:229-253);
  * motion files (``motion_t``, ``motion_filtered_t``): per frame pair four
    planes PREV.X, PREV.Y, NEXT.X, NEXT.Y of ``by*bx`` little-endian int16
    (motion.cpp:9-15,93-101), ``by = Y // block_size`` (motion_estimate.cpp:745);
  * frame-type files: one ASCII byte per pair, ``'I'`` or ``'B'``
    (decorrelate.cpp:982,1005).
"""
from __future__ import annotations

import numpy as np


def frame_bytes(X: int, Y: int) -> int:
    return X * Y + 2 * (X // 2) * (Y // 2)


def field_shorts(X: int, Y: int, block_size: int) -> int:
    return 4 * (Y // block_size) * (X // block_size)


def read_frames(path: str, X: int, Y: int, count: int | None = None) -> np.ndarray:
    """Returns a (frames, frame_bytes) uint8 array."""
    fb = frame_bytes(X, Y)
    data = np.fromfile(path, dtype=np.uint8)
    n = data.size // fb if count is None else count
    if data.size < n * fb:
        raise IOError(f"{path}: {data.size} bytes, expected {n} frames of {fb}")
    return data[: n * fb].reshape(n, fb)


def write_frames(path: str, frames: np.ndarray) -> None:
    np.ascontiguousarray(frames, dtype=np.uint8).tofile(path)


def read_motion(path: str, X: int, Y: int, block_size: int, count: int) -> np.ndarray:
    """Returns a (count, 4, by, bx) int16 array: PREV.X, PREV.Y, NEXT.X, NEXT.Y."""
    by, bx = Y // block_size, X // block_size
    data = np.fromfile(path, dtype="<i2")
    if data.size < count * 4 * by * bx:
        raise IOError(f"{path}: {data.size} shorts, expected {count * 4 * by * bx}")
    return data[: count * 4 * by * bx].reshape(count, 4, by, bx)


def write_motion(path: str, mv: np.ndarray) -> None:
    np.ascontiguousarray(mv, dtype="<i2").tofile(path)


def planes(frame: np.ndarray, X: int, Y: int):
    """Splits one I420 frame (flat uint8) into Y, U, V 2-D views."""
    cx, cy = X // 2, Y // 2
    y = frame[: X * Y].reshape(Y, X)
    u = frame[X * Y : X * Y + cx * cy].reshape(cy, cx)
    v = frame[X * Y + cx * cy : X * Y + 2 * cx * cy].reshape(cy, cx)
    return y, u, v


def synthetic_clip(X: int, Y: int, frames: int, seed: int, max_shift: int = 48,
                   noise: int = 2, flat_every: int = 0) -> np.ndarray:
    """Seeded synthetic I420 clip: a blurred blocky random canvas panned along a
    smooth closed trajectory, plus small integer noise; chroma are two other
    crops of the same canvas.  ``flat_every=k`` makes every k-th odd-indexed
    frame nearly flat (low entropy) so that the I/B decision produces 'I'
    frames.  Returns a (frames, frame_bytes) uint8 array.
    """
    rng = np.random.default_rng(seed)
    pad = max_shift + 8
    H, W = Y + 2 * pad, X + 2 * pad
    coarse = rng.integers(0, 256, size=((H + 7) // 8 + 1, (W + 7) // 8 + 1)).astype(np.float32)
    canvas = np.kron(coarse, np.ones((8, 8), np.float32))[:H, :W]
    for _ in range(3):
        canvas = (canvas + np.roll(canvas, 1, 0) + np.roll(canvas, -1, 0)) / 3.0
        canvas = (canvas + np.roll(canvas, 1, 1) + np.roll(canvas, -1, 1)) / 3.0
    # stretch contrast back to most of [0,255]
    canvas = (canvas - canvas.min()) / max(1e-6, float(canvas.max() - canvas.min())) * 235.0 + 10.0
    out = np.empty((frames, frame_bytes(X, Y)), np.uint8)
    cx, cy = X // 2, Y // 2
    for t in range(frames):
        ph = 2.0 * np.pi * t / 61.0
        dx = int(round(max_shift * np.sin(ph)))
        dy = int(round(max_shift * 0.6 * np.sin(2.0 * ph + 0.7)))
        y0, x0 = pad + dy, pad + dx
        luma = canvas[y0 : y0 + Y, x0 : x0 + X]
        if noise:
            luma = luma + rng.integers(-noise, noise + 1, size=(Y, X))
        if flat_every and (t % 2 == 1) and ((t // 2) % flat_every == 0):
            luma = np.full((Y, X), 100.0) + (rng.integers(0, 2, size=(Y, X)))
        u = canvas[y0 // 2 : y0 // 2 + cy, x0 // 2 : x0 // 2 + cx] * 0.5 + 64.0
        v = canvas[y0 // 2 + 3 : y0 // 2 + 3 + cy, x0 // 2 + 5 : x0 // 2 + 5 + cx] * 0.4 + 80.0
        fr = out[t]
        fr[: X * Y] = np.clip(luma, 0, 255).astype(np.uint8).ravel()
        fr[X * Y : X * Y + cx * cy] = np.clip(u, 0, 255).astype(np.uint8).ravel()
        fr[X * Y + cx * cy :] = np.clip(v, 0, 255).astype(np.uint8).ravel()
    return out
