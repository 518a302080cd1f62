"""The shipped Exodus mesh (template/input/AxiSEM_prem_ani_one_crust_50.e, kept as a fixture under tests/golden/) read
with the dependency-free HDF5 reader and turned into solver descriptors by the preloop restatement
(axisem3d_b200/h5lite.py, exodus_mesh.py): file-level facts, geometric invariants of the three element mappings and the
integral factor, material / attenuation sanity, and a stable run of the CPU oracle on it."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MESH = os.path.join(ROOT, "tests", "golden", "AxiSEM_prem_ani_one_crust_50.e")


@pytest.fixture(scope="module")
def mesh():
    from axisem3d_b200.exodus_mesh import ExodusMesh
    return ExodusMesh(MESH, nu=2)


def test_h5lite_reads_the_exodus_file():
    from axisem3d_b200 import h5lite
    f = h5lite.File(MESH)
    assert f.root_attrs()["title"] == "AxiSEM_prem_ani_one_crust_50"
    con = f["connect1"].read()
    assert con.shape == (2016, 4) and con.min() == 1 and con.max() == 2115
    x, y = f["coordx"].read(), f["coordy"].read()
    assert x.shape == (2115,) and abs(np.hypot(x, y).max() - 6371e3) < 1e-3
    names = [b"".join(r).split(b"\0")[0].decode() for r in f["name_glo_var"].read()]
    vals = dict(zip(names, f["vals_glo_var"].read().reshape(-1)))
    assert vals["nr_lin_solids"] == 3 and vals["radius"] == 6371e3 and abs(vals["dt"] - 2.5241379) < 1e-6
    ss = [b"".join(r).split(b"\0")[0].decode() for r in f["ss_names"].read()]
    assert ss == ["r1", "solid_fluid_boundary", "t0"]
    assert f["vals_elem_var1eb1"].read().shape == (1, 2016)          # chunked + shuffle + deflate
    assert "no_such_variable" not in f


def test_geometry_invariants(mesh):
    assert mesh.nelem == 2016 and mesh.ngll == 32649
    assert sorted(np.bincount(mesh.map_kind).tolist()) == [128, 192, 1696]      # semi-spherical, linear, spherical elements
    assert all(g["det"].min() > 0 for g in mesh.geo)
    # the integral factors (w_xi w_eta s |J|, axial form on the axis) integrate the volume of the sphere
    vol = 2 * np.pi * sum(f.sum() for f in mesh.ifact)
    assert abs(vol / (4.0 / 3.0 * np.pi * mesh.r_outer ** 3) - 1.0) < 1e-12
    # axial elements have side 3 on the axis after the rotation of ExodusModel::formAuxiliary
    for iq in np.nonzero(mesh.axial)[0]:
        assert np.all(np.abs(mesh.nodes[iq, 0, [0, 3]]) < 1e-6 * mesh.r_outer)
        assert np.all(mesh.geo[iq]["s"][0] == 0.0)
    # shared GLL points have the same coordinates from every element that holds them
    s, z = np.zeros(mesh.ngll), np.zeros(mesh.ngll)
    seen = np.zeros(mesh.ngll, dtype=bool)
    for iq in range(mesh.nelem):
        t = mesh.e2g[iq]
        g = mesh.geo[iq]
        old = seen[t]
        assert np.all(np.abs(s[t][old] - g["s"][old]) < 1e-5) and np.all(np.abs(z[t][old] - g["z"][old]) < 1e-5)
        s[t], z[t], seen[t] = g["s"], g["z"], True
    assert seen.all()


def test_material_and_points(mesh):
    assert int(mesh.is_fluid.sum()) == 256                                       # the outer core
    # mass of the Earth from the solid masses (rho x integral factor) and the file's density
    m_solid = 2 * np.pi * sum(float(m[0]) for m in mesh.mass_s)
    rho_f = 2 * np.pi * sum(float((mesh.ifact[iq] * mesh.mat[iq]["rho"]).sum()) for iq in np.nonzero(mesh.is_fluid)[0])
    assert 5.9e24 < m_solid + rho_f < 6.05e24
    # Nu = 2: Nr = 5 everywhere except where the circumference caps it at 3 on the axis
    assert set(np.unique(mesh.p_nr).tolist()) <= {3, 5}
    assert all((mesh.sf_n[t] is not None) == (mesh.mass_s[t].any() and mesh.mass_f[t].any()) for t in range(mesh.ngll))
    dt = mesh.estimate_dt()
    assert 0.2 < dt < 0.8                                                        # 50 s mesh at nPol = 4


@pytest.mark.timeout(600)
def test_oracle_runs_stably_on_the_real_mesh(mesh):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from axisem_oracle import OracleDomain
    from c_oracle import COracle
    dt = mesh.estimate_dt()
    d = OracleDomain(np.float32)
    rel = mesh.release(d, dt)
    d.addSourceTerm(mesh.make_source(rel["elements"], rel["dec"], amp=1e18))
    d.finalize()
    assert {(g.kind, getattr(g, "law", None)) for g in d.groups} == {("solid", "ti"), ("solid", "iso"), ("fluid", None)}
    co = COracle(d)
    stf = np.exp(-((np.arange(250) - 40) / 12.0) ** 2)
    peak = 0.0
    for sft in stf:
        co.step(dt, float(sft))
        peak = max(peak, float(np.abs(d.S["displ"]).max()))
    assert np.isfinite(d.S["displ"]).all() and np.isfinite(d.F["displ"]).all()
    assert 0 < np.abs(d.S["displ"]).max() <= peak < 1e6 and np.abs(d.F["displ"]).max() > 0     # the wave has reached the core, no blow-up
