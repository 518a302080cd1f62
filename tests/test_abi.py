"""The drop-in boundary is a C-ABI shared library: every entry point declared in include/axisem3d_b200.h must be exported by
axisem3d_b200/libaxisem3d_b200.so (no compute calls here: this runs without a GPU), the ctypes binding must cover the same set,
and without a CUDA device the library must fail loudly instead of falling back to anything."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "axisem3d_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ax3d_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    from axisem3d_b200 import capi
    names = _declared()
    assert len(names) >= 40
    lib = capi.load(build_if_missing=False)
    for n in names:
        assert hasattr(lib, n), "declared in the header but not exported: " + n
    missing = sorted(set(names) - set(capi.SYMBOLS))
    assert not missing, "declared in the header but not bound in capi.SYMBOLS: %s" % missing
    extra = sorted(set(capi.SYMBOLS) - set(names))
    assert not extra, "bound but not declared in the header: %s" % extra


def test_header_cites_the_reference_for_every_entry_point():
    """each declaration sits under a comment that names the reference interface it replaces (file:line)"""
    text = open(HEADER).read()
    assert len(re.findall(r"[A-Za-z0-9_]+\.(?:cpp|h):\d+", text)) >= 40


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from axisem3d_b200 import capi
    lib = capi.load(build_if_missing=False)
    h = ctypes.c_void_p()
    rc = lib.ax3d_create(0, ctypes.byref(h))
    assert rc != 0
    assert b"no CUDA device" in lib.ax3d_last_error()
