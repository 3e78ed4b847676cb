"""Python-3 restatement of the reference's four MCTF driver scripts, issuing the
same argv to the UNMODIFIED reference tools compiled into oracle/_ref/.

TEST INFRASTRUCTURE ONLY (see oracle/mctf_oracle.c).  Follows
  trunk/src/analyze.py:107-153, analyze_step.py:115-232,
  synthesize.py:95-153,  synthesize_step.py:84-143
(Python-2 integer `/` rewritten as `//`).  All files live in `workdir`, exactly
like the reference's tools-communicate-through-CWD contract.
"""
from __future__ import annotations

import os
import subprocess
import time

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(_HERE, "_ref")
SEARCH_RANGE_MAX = 128

TOOLS = ("split", "merge", "motion_estimate", "decorrelate", "correlate", "update", "un_update",
         "bidirectional_motion_decorrelate", "bidirectional_motion_correlate",
         "interlevel_motion_decorrelate", "interlevel_motion_correlate")


def available() -> bool:
    return all(os.access(os.path.join(REF_BIN, t), os.X_OK) for t in TOOLS)


def build() -> bool:
    """Compiles oracle/_ref from /root/reference when that tree is present."""
    if available():
        return True
    if not os.path.isdir("/root/reference/trunk/src"):
        return False
    subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return available()


def tool(name: str, workdir: str, timings: dict | None = None, **flags) -> int:
    argv = [os.path.join(REF_BIN, name)] + [f"--{k}={v}" for k, v in flags.items()]
    t0 = time.perf_counter()
    rc = subprocess.call(argv, cwd=workdir, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    if timings is not None:
        timings[name] = timings.get(name, 0.0) + time.perf_counter() - t0
    return rc


def analyze_step(workdir, t, pictures, X, Y, block_size, search_range, subpixel_accuracy,
                 update_factor, always_B=0, block_overlaping=0, border_size=0, timings=None):
    def run(name, **kw):
        rc = tool(name, workdir, timings, **kw)
        if rc != 0:
            raise RuntimeError(f"reference {name} exited {rc} at temporal_subband {t}")

    for stale in (f"motion_{t}",):
        p = os.path.join(workdir, stale)
        if os.path.exists(p):
            os.remove(p)  # motion_estimate exits 1 if it exists (motion_estimate.cpp:659-682)
    run("split", even_fn=f"even_{t}", low_fn=f"low_{t-1}", odd_fn=f"odd_{t}", pictures=pictures,
        pixels_in_x=X, pixels_in_y=Y)
    run("motion_estimate", block_size=block_size, border_size=border_size, even_fn=f"even_{t}",
        imotion_fn=f"imotion_{t}", motion_fn=f"motion_{t}", odd_fn=f"odd_{t}", pictures=pictures,
        pixels_in_x=X, pixels_in_y=Y, search_range=search_range,
        subpixel_accuracy=subpixel_accuracy)
    run("decorrelate", block_overlaping=block_overlaping, block_size=block_size,
        even_fn=f"even_{t}", frame_types_fn=f"frame_types_{t}", high_fn=f"high_{t}",
        motion_in_fn=f"motion_{t}", motion_out_fn=f"motion_filtered_{t}", odd_fn=f"odd_{t}",
        pictures=pictures, pixels_in_x=X, pixels_in_y=Y, search_range=search_range,
        subpixel_accuracy=subpixel_accuracy, always_B=always_B)
    run("update", block_size=block_size, even_fn=f"even_{t}", frame_types_fn=f"frame_types_{t}",
        high_fn=f"high_{t}", low_fn=f"low_{t}", motion_fn=f"motion_filtered_{t}",
        pictures=pictures, pixels_in_x=X, pixels_in_y=Y, subpixel_accuracy=subpixel_accuracy,
        update_factor=update_factor)


def analyze(workdir, X, Y, GOPs, TRLs, block_size=32, search_range=4, subpixel_accuracy=0,
            update_factor=0.0, always_B=0, block_overlaping=0, border_size=0, block_size_min=32,
            timings=None):
    """`low_0` must exist in workdir.  Returns the per-level schedule used."""
    pictures = GOPs * 2 ** (TRLs - 1) + 1
    if block_size < block_size_min:
        block_size_min = block_size
    sched = []
    for t in range(1, TRLs):
        sched.append(dict(t=t, pictures=pictures, search_range=search_range, block_size=block_size))
        analyze_step(workdir, t, pictures, X, Y, block_size, search_range, subpixel_accuracy,
                     update_factor, always_B, block_overlaping, border_size, timings)
        pictures = (pictures + 1) // 2
        search_range = min(search_range * 2, SEARCH_RANGE_MAX)
        block_size = max(block_size // 2, block_size_min)
    return sched


def synthesize_step(workdir, t, pictures, X, Y, block_size, search_range, subpixel_accuracy,
                    update_factor, block_overlaping=0, timings=None):
    def run(name, **kw):
        rc = tool(name, workdir, timings, **kw)
        if rc != 0:
            raise RuntimeError(f"reference {name} exited {rc} at temporal_subband {t}")

    run("un_update", block_size=block_size, even_fn=f"even_{t}", frame_types_fn=f"frame_types_{t}",
        high_fn=f"high_{t}", low_fn=f"low_{t}", motion_fn=f"motion_{t}", pictures=pictures,
        pixels_in_x=X, pixels_in_y=Y, subpixel_accuracy=subpixel_accuracy,
        update_factor=update_factor)
    run("correlate", block_overlaping=block_overlaping, block_size=block_size, even_fn=f"even_{t}",
        frame_types_fn=f"frame_types_{t}", high_fn=f"high_{t}", motion_in_fn=f"motion_{t}",
        odd_fn=f"odd_{t}", pictures=pictures, pixels_in_x=X, pixels_in_y=Y,
        search_range=search_range, subpixel_accuracy=subpixel_accuracy)
    run("merge", even=f"even_{t}", low=f"low_{t-1}", odd=f"odd_{t}", pictures=pictures,
        pixels_in_x=X, pixels_in_y=Y)


def synthesize(workdir, X, Y, GOPs, TRLs, block_size=16, search_range=4, subpixel_accuracy=0,
               update_factor=0.25, block_overlaping=0, timings=None):
    """Scalar geometry for every level (the reference takes comma lists; the
    spatially-scalable per-level variant is SURVEY §8f).  `low_{TRLs-1}`,
    `high_t`, `motion_t`, `frame_types_t` must exist in workdir."""
    all_pictures = GOPs * 2 ** (TRLs - 1) + 1
    for t in range(TRLs - 1, 0, -1):
        pictures, sr = all_pictures, search_range
        for _ in range(1, t):
            sr = min(sr * 2, SEARCH_RANGE_MAX)
            pictures = (pictures + 1) // 2
        synthesize_step(workdir, t, pictures, X, Y, block_size, sr, subpixel_accuracy,
                        update_factor, block_overlaping, timings)
