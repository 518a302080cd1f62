"""The C++ host facade (axisem3d_b200/host/ax3d_host.hpp): the reference's Domain / Point / Element / Newmark class
interface over the C-ABI.  tests/cpp/host_driver.cpp replays a serialised Mesh::release through those classes."""
import os
import subprocess

import numpy as np
import pytest

from dump_domain import DumpDomain, build_driver, read_displacement
from helpers import build_oracle
from axisem3d_b200.mesh_synth import SynthMesh

MESH = dict(n_theta=5, n_r=8, nu=7, law="aniso", model3d=True, attenuation="cg4", fluid3d=True, perturb_rho=True, prt=True, ocean=True)


def _dump(tmp_path, nstep=40):
    m = SynthMesh(**MESH)
    dt = m.estimate_dt()
    d = DumpDomain()
    rel = m.release(d, dt)
    d.addSourceTerm(m.make_source(rel["elements"], rel["dec"], amp=1e18))
    stf = np.exp(-((np.arange(nstep) - 12) / 4.0) ** 2)
    path = os.path.join(str(tmp_path), "domain.bin")
    d.write(path, dt, stf)
    return m, dt, stf, d, path


def test_facade_compiles_links_and_fails_loudly_without_gpu(tmp_path):
    """g++-compiles the facade + driver against the library; on a box without a GPU the Domain constructor must throw
    the library's 'no CUDA device' error (exit code 3) -- never fall back to a CPU path."""
    import torch
    exe = build_driver()
    _, _, _, _, path = _dump(tmp_path, nstep=2)
    r = subprocess.run([exe, path, os.path.join(str(tmp_path), "out.bin")], capture_output=True, text=True, timeout=300)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 3, (r.returncode, r.stderr)
        assert "no CUDA device" in r.stderr and "Domain::Domain" in r.stderr


@pytest.mark.gpu
def test_cpp_driver_matches_oracle(tmp_path):
    exe = build_driver()
    m, dt, stf, d, path = _dump(tmp_path)
    out = os.path.join(str(tmp_path), "out.bin")
    r = subprocess.run([exe, path, out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    ora, _ = build_oracle(m, dt, np.float64)
    for s in stf:
        ora.step(dt, float(s))
    got = read_displacement(out, d.points)
    sscale, fscale = np.abs(ora.S["displ"]).max(), np.abs(ora.F["displ"]).max()
    assert sscale > 0 and fscale > 0
    for t, (s, f) in got.items():
        if s is not None:
            assert np.abs(s - ora.get_solid(t, "displ")).max() <= 1e-4 * sscale, t
        if f is not None:
            assert np.abs(f - ora.get_fluid(t, "displ")).max() <= 1e-4 * fscale, t


@pytest.mark.gpu
def test_cpp_driver_replays_the_reference_preloop(tmp_path):
    """INTEGRATION.md's arrangement end to end: the REFERENCE's own preloop (tests/golden/main_bubbles_3d_domain.bin.xz = what
    its Mesh / Source / STF release put into its Domain on the template inputs with three 3-D bubble models, dumped by
    oracle/ref_main_dump.cpp) replayed through the C++ facade classes onto the CUDA path -- no Python preloop in between --
    against the repo's own preloop + Python binding on the same input directory."""
    import lzma
    import struct
    import main_case as MC
    from axisem3d_b200.domain import Domain
    name, nstep = "bubbles_3d", 300
    raw = lzma.open(os.path.join(MC.GOLDEN, "main_%s_domain.bin.xz" % name), "rb").read()
    end = raw.rfind(b"RECV")
    nref = MC.golden(name)["steps"]
    start = end - (4 + 8 + 4 * nref)
    n, dt = struct.unpack("<id", raw[start:start + 12])
    assert n == nref
    stf = np.frombuffer(raw, dtype="<f4", count=n, offset=start + 12)[:nstep]
    path, out = os.path.join(str(tmp_path), "ref_domain.bin"), os.path.join(str(tmp_path), "out.bin")
    with open(path, "wb") as f:
        f.write(raw[:start] + struct.pack("<id", nstep, dt) + stf.tobytes())
    r = subprocess.run([build_driver(), path, out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    case = MC.get_case(name)
    assert abs(case.dt - dt) <= 1e-12 * dt
    g = Domain(0)
    rel = case.release(g)
    g.finalize()
    g.runSteps(case.dt, case.stf[:nstep])
    got = read_displacement(out, rel["points"])
    sscale = max(np.abs(g.get_bulk("displ", False)).max(), 1e-300)
    fscale = max(np.abs(g.get_bulk("displ", True)).max(), 1e-300)
    assert sscale > 1e-12
    for t, (s, f) in got.items():
        if s is not None:
            assert np.abs(s - g.get_solid(t, "displ")).max() <= 1e-4 * sscale, t
        if f is not None:
            assert np.abs(f - g.get_fluid(t, "displ")).max() <= 1e-4 * fscale, t
