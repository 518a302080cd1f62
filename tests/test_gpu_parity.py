"""Parity of the CUDA path (through the C-ABI) against the oracle.  BASELINE.json north_star:
per-step stiffness forces rel. L2 <= 1e-5; seismograms rel. L2 <= 1e-4 after 2000 steps."""
import numpy as np
import pytest

from helpers import build_oracle, build_gpu, randomize_displ, push_fields, compare_field, rel_l2
from axisem3d_b200.mesh_synth import SynthMesh

pytestmark = pytest.mark.gpu

TOL_FORCE = 1e-5     # north_star tolerance on per-step stiffness forces


CASES = {
    # cfg1-like: 1D TI PREM + CG4 attenuation, Nu = 2 (Nr = 5), fluid core, SF coupling
    "cfg1_ti1d_cg4": dict(n_theta=8, n_r=8, nu=2, law="ti", model3d=False, attenuation="cg4"),
    "iso1d_full": dict(n_theta=6, n_r=6, nu=5, law="iso", model3d=False, attenuation="full"),
    "aniso1d": dict(n_theta=6, n_r=6, nu=4, law="aniso", model3d=False, attenuation=None),
    # cfg2-like: 3D isotropic, no attenuation
    "cfg2_iso3d": dict(n_theta=8, n_r=8, nu=12, law="iso", model3d=True, attenuation=None),
    "cfg2_iso3d_nu100": dict(n_theta=4, n_r=6, nu=100, law="iso", model3d=True, attenuation=None),
    "ti3d_cg4": dict(n_theta=6, n_r=8, nu=9, law="ti", model3d=True, attenuation="cg4"),
    # cfg3-like: 3D anisotropic + SLS (CG4 and Full), 3D fluid, 3D mass
    "cfg3_aniso3d_cg4": dict(n_theta=6, n_r=8, nu=20, law="aniso", model3d=True, attenuation="cg4", fluid3d=True),
    "aniso3d_full_mass3d": dict(n_theta=5, n_r=8, nu=7, law="aniso", model3d=True, attenuation="full",
                                fluid3d=True, perturb_rho=True),
}


def ragged_nu(s, z):
    """cfg4-like: per-point Nu growing with distance from the axis."""
    return int(3 + 40 * s / 6371e3)


# fused-kernel corner cases: gather tile smaller than the spectrum (Nr = 288), and the split pipeline (Nr = 416)
CASES["iso3d_nu140_tiled"] = dict(n_theta=3, n_r=4, nu=140, law="iso", model3d=True, attenuation=None, fluid3d=True)
CASES["ti3d_nu200_split"] = dict(n_theta=3, n_r=4, nu=200, law="ti", model3d=True, attenuation="cg4", fluid3d=True)
# split pipeline with one point per CTA (Nr = 2016 needs > 224 KB for five points)
def equator_nu1000(s, z):
    """two equatorial elements of a thin shell at Nu = 1000 (Nr = 2016: the circumference cap needs ~20 km GLL spacing)."""
    return 1000 if s > 0.9895 * 6371e3 and abs(z) < 40e3 else 6


CASES["iso3d_nu1000_split_np1"] = dict(n_theta=320, n_r=1, r_in=6371e3 - 80e3, nu_fn=equator_nu1000, law="ti", model3d=True,
                                       attenuation="cg4", fluid_layers=())
def equator_nu500(s, z):
    """equatorial elements at Nu = 500 (Nr = 1008) and a neighbouring band at Nu = 290: the 5-CTA cluster kernel's range."""
    if s > 0.9895 * 6371e3 and abs(z) < 40e3:
        return 500
    return 290 if s > 0.97 * 6371e3 and abs(z) < 120e3 else 6


CASES["ti3d_nu500_cluster"] = dict(n_theta=320, n_r=1, r_in=6371e3 - 80e3, nu_fn=equator_nu500, law="ti", model3d=True,
                                   attenuation="cg4", fluid_layers=())
# particle relabelling: 9-component path (computeGrad9 / computeQuad9, 9-component rotation, PRT_1D in Fourier space, PRT_3D in
# physical space through the split pipeline with 5 Z-form pairs per point), solid and fluid elements
CASES["prt1d_aniso1d_full"] = dict(n_theta=6, n_r=6, nu=5, law="aniso", model3d=False, attenuation="full", prt=True)
CASES["prt3d_ti3d_cg4_nu60"] = dict(n_theta=4, n_r=6, nu=60, law="ti", model3d=True, attenuation="cg4", fluid3d=True, prt=True)
CASES["prt3d_ragged"] = dict(n_theta=8, n_r=6, nu_fn=ragged_nu, law="iso", model3d=True, attenuation=None, fluid3d=True, prt=True)
CASES["cfg4_ragged"] = dict(n_theta=10, n_r=6, nu_fn=ragged_nu, law="iso", model3d=True, attenuation=None)


@pytest.mark.parametrize("name", sorted(CASES))
def test_stiffness_forces_match_oracle(name):
    m = SynthMesh(**CASES[name])
    dt = m.estimate_dt()
    d, _ = build_oracle(m, dt, np.float64)
    g, _ = build_gpu(m, dt)
    randomize_displ(d, seed=7)
    push_fields(d, g, ("displ",))
    for it in range(2):                       # second pass exercises the memory variables
        d.computeStiff()
        d.coupleSolidFluid()
        g.computeStiff()
        g.coupleSolidFluid()
        err = compare_field(d, g, "stiff")
        for k, v in err.items():
            assert v <= TOL_FORCE, (name, it, k, v)
        if it == 0:
            d.S["stiff"][:] = 0
            d.F["stiff"][:] = 0
            s, f = g.get_bulk("stiff", False), g.get_bulk("stiff", True)
            if s.size: g.set_bulk("stiff", False, np.zeros_like(s))
            if f.size: g.set_bulk("stiff", True, np.zeros_like(f))


@pytest.mark.parametrize("name", ["cfg1_ti1d_cg4", "cfg2_iso3d", "aniso3d_full_mass3d"])
def test_time_loop_matches_oracle(name):
    m = SynthMesh(**CASES[name])
    dt = m.estimate_dt()
    d, rel = build_oracle(m, dt, np.float64)
    g, _ = build_gpu(m, dt)
    nstep = 60
    stf = np.exp(-((np.arange(nstep) - 15) / 5.0) ** 2)
    for i in range(nstep):
        d.step(dt, stf[i])
    g.runSteps(dt, stf)
    assert g.checkStability()
    err = compare_field(d, g, "displ")
    for k, v in err.items():
        assert v <= 1e-4, (name, k, v)


def test_seismograms_2000_steps():
    """north_star: station seismograms within 1e-4 (rel. L2) of the reference after 2000 steps."""
    m = SynthMesh(n_theta=6, n_r=6, nu=2, law="ti", model3d=False, attenuation="cg4")
    dt = m.estimate_dt()
    d, rel = build_oracle(m, dt, np.float64)
    g, _ = build_gpu(m, dt)
    nstep = 2000
    stf = np.exp(-((np.arange(nstep) - 30) / 8.0) ** 2).astype(np.float32)
    # 6 receivers in surface elements (top layer), arbitrary interpolation weights
    rng = np.random.default_rng(3)
    etags = [e.domain_tag for e in rel["elements"] if e.kind == "solid"][-6:]
    phi = rng.uniform(0, 2 * np.pi, len(etags))
    w = rng.uniform(0, 1, (len(etags), 25))
    w /= w.sum(axis=1, keepdims=True)
    so, sg = [], []
    for i in range(nstep):
        d.step(dt, float(stf[i]))
        g.step(dt, float(stf[i]))
        if i % 10 == 0:
            so.append([d.ground_motion(t, p, ww) for t, p, ww in zip(etags, phi, w)])
            sg.append(g.ground_motion(etags, phi, w))
    so, sg = np.array(so), np.array(sg)
    assert np.abs(so).max() > 0
    assert rel_l2(so, sg) <= 1e-4


def test_error_behaviour_matches_reference():
    from axisem3d_b200 import model as M
    from axisem3d_b200.domain import Domain
    g = Domain(0)
    with pytest.raises(RuntimeError, match="setGMat|finalize"):
        g.updateNewmark(0.1)
    sp = M.SolidPoint(5, False, [1.0, 2.0], M.Mass1D(1.0))
    g.addPoint(sp)
    with pytest.raises(RuntimeError, match="Incompatible size"):
        M.SolidPoint(6, False, [1.0, 2.0], M.Mass3D(np.ones(5)))


@pytest.mark.parametrize("kw", [
    dict(n_theta=24, n_r=10, nu=24, law="iso", model3d=True, attenuation=None),                      # 240 quads > 148 SMs
    dict(n_theta=10, n_r=6, nu_fn=ragged_nu, law="ti", model3d=True, attenuation="cg4", fluid3d=True),
])
def test_in_kernel_newmark_matches_verbwise_stepping(kw):
    """ax3d_run_steps advances the plain solid points inside the fused element kernel (fused.cuh: Newmark warps fed by
    TMA bulk loads, arrival counters, ready queue); the verb-wise calls use the stand-alone Newmark kernel.  Both must
    give the same wavefield (the only difference is the fp32 summation order of the scatter atomics), and both must
    match the oracle."""
    m = SynthMesh(**kw)
    dt = m.estimate_dt()
    d, _ = build_oracle(m, dt, np.float64)
    g1, _ = build_gpu(m, dt)
    g2, _ = build_gpu(m, dt)
    nstep = 40
    stf = np.exp(-((np.arange(nstep) - 12) / 4.0) ** 2)
    for i in range(nstep):
        d.step(dt, stf[i])
        g2.step(dt, float(stf[i]))
    g1.runSteps(dt, stf[:25])          # three calls: first / middle / last graph variants and the hand-over between calls
    g1.runSteps(dt, stf[25:26])
    g1.runSteps(dt, stf[26:])
    assert g1.checkStability() and g2.checkStability()
    for which in ("displ", "veloc", "accel", "stiff"):
        a, b = g1.get_bulk(which, False), g2.get_bulk(which, False)
        assert rel_l2(b.astype(np.complex128), a.astype(np.complex128)) <= 2e-5, which
        err = compare_field(d, g1, which)
        for k, v in err.items():
            assert v <= 1e-4, (which, k, v)


def test_device_side_recorder_matches_per_step_record():
    """ax3d_run_steps_record (samples buffered on the device, in-kernel Newmark active) == step(); record() per step."""
    m = SynthMesh(n_theta=16, n_r=10, nu=16, law="iso", model3d=True, attenuation=None)
    dt = m.estimate_dt()
    g1, rel = build_gpu(m, dt)
    g2, _ = build_gpu(m, dt)
    rng = np.random.default_rng(11)
    solid = [e.domain_tag for e in rel["elements"] if e.kind == "solid"]
    etags = [solid[i] for i in rng.integers(0, len(solid), 9)]
    phi = rng.uniform(0, 2 * np.pi, len(etags))
    w = rng.uniform(0, 1, (len(etags), 25))
    w /= w.sum(axis=1, keepdims=True)
    g1.setReceivers(etags, phi, w)
    g2.setReceivers(etags, phi, w)
    nstep = 45
    stf = np.exp(-((np.arange(nstep) - 10) / 4.0) ** 2).astype(np.float32)
    a = np.concatenate([g1.runStepsRecord(dt, stf[:30]), g1.runStepsRecord(dt, stf[30:])])
    b = []
    for i in range(nstep):
        g2.updateNewmark(dt)
        b.append(g2.record())
        g2.applySource(float(stf[i]))
        g2.computeStiff()
        g2.coupleSolidFluid()
    b = np.array(b)
    assert a.shape == b.shape and np.abs(b).max() > 0
    assert rel_l2(b, a) <= 2e-5


@pytest.mark.parametrize("name", ["cfg3_aniso3d_cg4", "cfg4_ragged", "ti3d_nu500_cluster"])
def test_cluster_kernel_matches_oracle(name, monkeypatch):
    """k_elem3d_cluster (csrc/cluster.cuh: one 5-CTA thread-block cluster per element, xi-derivative through distributed
    shared memory) is opt-in (AX3D_CLUSTER, read at finalize); every 3D element of these cases goes through it."""
    monkeypatch.setenv("AX3D_CLUSTER", "2")
    m = SynthMesh(**CASES[name])
    dt = m.estimate_dt()
    d, _ = build_oracle(m, dt, np.float64)
    g, _ = build_gpu(m, dt)
    randomize_displ(d, seed=11)
    push_fields(d, g, ("displ",))
    d.computeStiff()
    d.coupleSolidFluid()
    g.computeStiff()
    g.coupleSolidFluid()
    for k, v in compare_field(d, g, "stiff").items():
        assert v <= TOL_FORCE, (name, k, v)


def test_check_stability_reports_non_finite_displacement():
    """Domain::checkStability (Domain.cpp:237-275): Point::stable() = mDispl.allFinite() (SolidPoint.h:21, FluidPoint.h:21).
    A NaN or an infinity anywhere in the solid or the fluid displacement must come back as `*stable == 0`; a clean field as 1."""
    m = SynthMesh(n_theta=6, n_r=8, nu=8, law="iso", model3d=True, attenuation=None)
    dt = m.estimate_dt()
    g, _ = build_gpu(m, dt)
    g.runSteps(dt, np.exp(-((np.arange(8) - 3) / 2.0) ** 2).astype(np.float32))
    assert g.checkStability()
    for fluid, bad in ((False, np.nan), (True, np.inf), (False, -np.inf), (True, np.nan)):
        u = g.get_bulk("displ", fluid).copy()
        keep = u.copy()
        k = u.size - 3 if fluid else u.size // 2        # anywhere in the array, including the tail the last CTA covers
        u[k] = complex(bad, 0.0) if fluid else complex(0.0, bad)
        g.set_bulk("displ", fluid, u)
        assert not g.checkStability(), (fluid, bad)
        g.set_bulk("displ", fluid, keep)
        assert g.checkStability()
    # an unstable time step must be caught by the loop itself: dt far above the CFL limit blows up within a few hundred steps
    g2, _ = build_gpu(m, dt)
    g2.runSteps(50.0 * dt, np.ones(400, np.float32))
    assert not g2.checkStability()


# ---------------------------------------------------------------------------------------------------------------
# The shipped Exodus mesh itself (template/input/AxiSEM_prem_ani_one_crust_50.e; axisem3d_b200/exodus_mesh.py): all three
# element mappings, PREM's TI upper mantle + isotropic rest + fluid outer core, CG4 attenuation from the file's Q.
def _real_mesh(**kw):
    import os
    from axisem3d_b200.exodus_mesh import ExodusMesh
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return ExodusMesh(os.path.join(root, "tests", "golden", "AxiSEM_prem_ani_one_crust_50.e"), **kw)


def _ramp_nu(s, z):
    return int(3 + 30 * s / 6371e3)


@pytest.mark.timeout(1200)
@pytest.mark.parametrize("kw", [
    dict(nu=2, attenuation="cg4"),                                   # configs[0]: the template's own inparam (1D PREM, Nu = 2)
    dict(nu_fn=_ramp_nu, attenuation="cg4", model3d=True),           # 3D (phi-dependent) material, ragged Nu, on the real geometry
], ids=["cfg1_real_mesh", "ragged3d_real_mesh"])
def test_real_mesh_matches_oracle(kw):
    m = _real_mesh(**kw)
    dt = m.estimate_dt()
    d, _ = build_oracle(m, dt, np.float64)
    g, _ = build_gpu(m, dt)
    randomize_displ(d, seed=5)
    push_fields(d, g, ("displ",))
    _two_evaluations(d, g, kw)
    # and a time loop with the source, from rest
    d2, _ = build_oracle(m, dt, np.float64)
    g2, _ = build_gpu(m, dt)
    nstep = 40
    stf = np.exp(-((np.arange(nstep) - 12) / 4.0) ** 2)
    for i in range(nstep):
        d2.step(dt, stf[i])
    g2.runSteps(dt, stf)
    assert g2.checkStability()
    for k, v in compare_field(d2, g2, "displ").items():
        if k == "solid":          # the wave has not reached the fluid core in 40 steps
            assert v <= 1e-4, (kw, k, v)


def test_fluid_strain_and_curl_receivers():
    """FluidElement::computeStrain (FluidElement.cpp:219-306; Acoustic1D elements: k_strain_fluid1d) against the numpy oracle, and
    FluidElement::computeCurl == 0 (FluidElement.cpp:308-311); a 3D fluid is rejected with the reference-style message."""
    m = SynthMesh(n_theta=6, n_r=8, nu=9, law="iso", model3d=True, attenuation=None)       # 3D solid, 1D fluid
    dt = m.estimate_dt()
    d, rel = build_oracle(m, dt, np.float64)
    g, _ = build_gpu(m, dt)
    randomize_displ(d, seed=4)
    push_fields(d, g, ("displ",))
    fl = [e.domain_tag for e in rel["elements"] if e.kind == "fluid"]
    rng = np.random.default_rng(2)
    tags = [fl[i] for i in rng.integers(0, len(fl), 7)] + [fl[0], fl[-1]]      # includes axial fluid elements
    phi = rng.uniform(0, 2 * np.pi, len(tags))
    w = rng.uniform(0, 1, (len(tags), 25))
    w /= w.sum(axis=1, keepdims=True)
    ref = np.array([d.strain(t, p, ww) for t, p, ww in zip(tags, phi, w)])
    got = g.strain(tags, phi, w)
    assert np.abs(ref).max() > 0 and rel_l2(ref, got) <= 1e-4
    assert np.all(g.curl(tags, phi, w) == 0.0)
    m3 = SynthMesh(n_theta=5, n_r=8, nu=7, law="iso", model3d=True, attenuation=None, fluid3d=True)
    g3, rel3 = build_gpu(m3, m3.estimate_dt())
    f3 = [e.domain_tag for e in rel3["elements"] if e.kind == "fluid" and e.acoustic.K.shape[0] > 1][:1]
    with pytest.raises(RuntimeError, match="FluidElement::computeStrain"):
        g3.strain(f3, phi[:1], w[:1])


@pytest.mark.parametrize("name", ["iso3d_nu1000_split_np1", "ti3d_nu200_split", "cfg4_ragged"])
def test_split_pipeline_still_matches_oracle(name, monkeypatch):
    """The fused kernel now takes every element up to Nr ~ 2700 (partial-row passes); the split pipeline (k_grad3d ->
    k_fft3d_v2 -> k_quad3d, one point per CTA for Nr = 2016) remains for particle relabelling and beyond.  AX3D_NO_FUSED=1
    (read at finalize) sends every 3D element through it, so that it stays under test."""
    monkeypatch.setenv("AX3D_NO_FUSED", "1")
    m = SynthMesh(**CASES[name])
    dt = m.estimate_dt()
    d, _ = build_oracle(m, dt, np.float64)
    g, _ = build_gpu(m, dt)
    randomize_displ(d, seed=13)
    push_fields(d, g, ("displ",))
    _two_evaluations(d, g, name)


def _two_evaluations(d, g, tag):
    """two stiffness evaluations (the second one exercises the memory variables), forces zeroed in between as the Newmark update does"""
    for it in range(2):
        d.computeStiff()
        d.coupleSolidFluid()
        g.computeStiff()
        g.coupleSolidFluid()
        for k, v in compare_field(d, g, "stiff").items():
            assert v <= TOL_FORCE, (tag, it, k, v)
        d.S["stiff"][:] = 0
        d.F["stiff"][:] = 0
        for fluid in (False, True):
            a = g.get_bulk("stiff", fluid)
            if a.size:
                g.set_bulk("stiff", fluid, np.zeros_like(a))


def test_partial_row_passes_cover_mixed_sizes(monkeypatch):
    """Elements of Nr = 2016 (10 partial-row passes; opt-in, AX3D_PARTIAL_ROWS=1 read at finalize), Nr = 1008 (5 passes) and
    small ones side by side."""
    monkeypatch.setenv("AX3D_PARTIAL_ROWS", "1")
    def nu(s, z):
        if s > 0.9895 * 6371e3 and abs(z) < 40e3:
            return 1000
        return 500 if s > 0.97 * 6371e3 and abs(z) < 120e3 else 8
    m = SynthMesh(n_theta=320, n_r=1, r_in=6371e3 - 80e3, nu_fn=nu, law="aniso", model3d=True, attenuation="cg4", fluid_layers=())
    dt = m.estimate_dt()
    d, _ = build_oracle(m, dt, np.float64)
    g, _ = build_gpu(m, dt)
    assert m.e_nr.max() == 2016
    randomize_displ(d, seed=21)
    push_fields(d, g, ("displ",))
    _two_evaluations(d, g, "partial rows")
