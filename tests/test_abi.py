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


def test_drop_in_binary_fails_loudly_without_a_device(tmp_path):
    """oracle/_ref/axisem3d_gpu (the reference's own main + preloop + recorders over the binding, oracle/Makefile.dropin) runs
    the reference's whole preloop on the template inputs and then -- on a box without a GPU -- stops in Domain::Domain with the
    library's error, printed by the reference's own exception handler.  There is no CPU path to fall back to."""
    import shutil
    import subprocess
    import sys
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oracle", "_ref", "axisem3d_gpu")
    if not os.path.exists(exe) or torch.cuda.is_available():
        pytest.skip("needs the drop-in binary (built where /root/reference exists) and a box without a CUDA device")
    sys.path.insert(0, os.path.join(root, "oracle"))
    import main_case as MC
    from nc_flatten import flatten
    run = os.path.join(str(tmp_path), "run")
    os.makedirs(run)
    inp = MC.input_dir("cfg1_template", run)
    flatten(os.path.join(inp, MC.MESH))
    link = os.path.join(run, "axisem3d_gpu")
    os.symlink(exe, link)
    r = subprocess.run([link], cwd=run, capture_output=True, text=True, timeout=300)
    assert "AXISEM3D ABORTED UPON RUNTIME EXCEPTION" in r.stdout and "FROM: Domain::Domain" in r.stdout
    assert "no CUDA device" in r.stdout
    assert "Attenuation Builder" in r.stdout            # the reference's preloop ran up to the point where the Domain is created
    assert not os.path.exists(os.path.join(run, "output", "stations", "axisem3d_synthetics.nc.ncflat"))
    shutil.rmtree(run, ignore_errors=True)
