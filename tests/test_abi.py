"""CPU: the C-ABI library loads and exports every symbol include/qsvc_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "qsvc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qsvc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from qsvc_b200 import _lib
    so = _lib.build()
    lib = ctypes.CDLL(so)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/qsvc_b200.h but not exported"
    # the Python binding covers the same set
    assert set(_lib.SIGNATURES) == set(names)


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, not compute on the CPU."""
    from qsvc_b200 import _lib
    L = _lib.lib()
    if L.qsvc_device_count() > 0:
        pytest.skip("a GPU is present")
    assert not L.qsvc_create(0)
    assert b"no CUDA device" in L.qsvc_last_error()
    from qsvc_b200.mctf import Context
    with pytest.raises(_lib.QsvcError):
        Context(0)


def test_product_path_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "qsvc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("(validated against the oracle)", ""), f
