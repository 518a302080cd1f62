"""Size-independent properties of the stiffness operator at BASELINE.json's full sizes (the oracle only finishes small
meshes in seconds): on the bench's own cfg2, cfg3 and cfg4 domains (72 x 28 = 2016 quads; Nu = 100, Nu = 200 anisotropic, Nu = 20..500) the CUDA
path must be
  * linear:        K(u1 + c u2) = K u1 + c K u2,
  * self-adjoint:  <u2, K u1> = <u1, K u2> in the azimuthally integrated inner product the reference's own
                   SolidElement::test / FluidElement::test use (SolidElement.cpp:96-187: mode 0 once, modes >= 1 twice),
  * positive:      <u, K u> > 0,
and deterministic up to the order of the scatter atomics between two evaluations.
The test fields are made admissible (axial and Nyquist masks) by the library's own Newmark update of a random force."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _build(cfg):
    import bench
    from axisem3d_b200.domain import Domain
    bench.CFG = cfg
    m = bench.make_mesh(bench.N_THETA)
    dt = m.estimate_dt()
    g = Domain(0)
    rel = m.release(g, dt)
    g.finalize()
    return g, rel, dt


def _mode_weights(points):
    """weight of every entry of the solid / fluid bulk arrays in the inner product: 1 for mode 0, 2 for modes >= 1"""
    ws, wf = [], []
    for p in points:
        w = np.full(p.nu + 1, 2.0)
        w[0] = 1.0
        if p.kind != "fluid":
            ws.append(np.tile(w, 3))
        if p.kind != "solid":
            wf.append(w)
    cat = lambda l: np.concatenate(l) if l else np.zeros(0)
    return cat(ws), cat(wf)


def _admissible_fields(g, dt, seed):
    """random force -> one Newmark update: u = dt^2 M^-1 mask(f), masked by SolidPoint/FluidPoint::updateNewmark"""
    rng = np.random.default_rng(seed)
    g.resetZero()
    out = []
    for fluid in (False, True):
        n = g.field_size(fluid)
        if n:
            g.set_bulk("stiff", fluid, ((rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64))
    g.updateNewmark(dt)
    for fluid in (False, True):
        u = g.get_bulk("displ", fluid) if g.field_size(fluid) else np.zeros(0, np.complex64)
        s = np.abs(u).max()
        out.append((u / s * 1e-6).astype(np.complex64) if s > 0 else u)
    return out


def _apply(g, us, uf):
    """f = -stiff after Domain::computeStiff on displacement (us, uf)"""
    g.resetZero()
    if us.size:
        g.set_bulk("displ", False, us)
    if uf.size:
        g.set_bulk("displ", True, uf)
    g.computeStiff()
    fs = -g.get_bulk("stiff", False).astype(np.complex128) if us.size else np.zeros(0, np.complex128)
    ff = -g.get_bulk("stiff", True).astype(np.complex128) if uf.size else np.zeros(0, np.complex128)
    return fs, ff


def _dot(w, u, f):
    return float(np.sum(w * (u.real * f.real + u.imag * f.imag)))


# cfg3 carries SLS attenuation, but every evaluation below starts from Domain::resetZero: with zero memory variables one
# computeStiff is the (unrelaxed) elastic operator, still linear and self-adjoint
@pytest.mark.parametrize("cfg", ["cfg2", "cfg3", "cfg4"])
def test_stiffness_operator_properties_at_full_size(cfg):
    g, rel, dt = _build(cfg)
    assert len(rel["elements"]) == 2016
    ws, wf = _mode_weights(rel["points"])
    u1s, u1f = _admissible_fields(g, dt, 1)
    u2s, u2f = _admissible_fields(g, dt, 2)
    assert np.abs(u1s).max() > 0 and np.abs(u1f).max() > 0
    f1s, f1f = _apply(g, u1s, u1f)
    f2s, f2f = _apply(g, u2s, u2f)
    # deterministic up to the order of the scatter atomics
    r1s, r1f = _apply(g, u1s, u1f)
    assert np.linalg.norm(r1s - f1s) <= 2e-6 * np.linalg.norm(f1s)
    assert np.linalg.norm(r1f - f1f) <= 2e-6 * np.linalg.norm(f1f)
    # linearity
    c = 0.37
    f3s, f3f = _apply(g, (u1s + c * u2s).astype(np.complex64), (u1f + c * u2f).astype(np.complex64))
    assert np.linalg.norm(f3s - (f1s + c * f2s)) <= 1e-5 * np.linalg.norm(f3s)
    assert np.linalg.norm(f3f - (f1f + c * f2f)) <= 1e-5 * np.linalg.norm(f3f)
    # self-adjoint and positive, solid and fluid operators separately (computeStiff does not couple them)
    for w, ua, ub, fa, fb in ((ws, u1s, u2s, f1s, f2s), (wf, u1f, u2f, f1f, f2f)):
        ua, ub = ua.astype(np.complex128), ub.astype(np.complex128)
        ab, ba = _dot(w, ub, fa), _dot(w, ua, fb)
        scale = np.sqrt(_dot(w, ua, fa) * _dot(w, ub, fb))
        assert _dot(w, ua, fa) > 0 and _dot(w, ub, fb) > 0
        assert abs(ab - ba) <= 1e-5 * scale, (cfg, ab, ba, scale)


# ---------------------------------------------------------------------------------------------------------------
# The bench's own domains against the CPU oracle.  oracle/oracle.c (C/OpenMP, fp32; pinned to the reference's golden
# vectors by tests/test_golden_reference.py::test_c_oracle_matches_reference_sources) steps the whole 2016-quad mesh in
# about a second, so the full-size CUDA path -- the exact kernel instances, shared-memory plans, work queues and chunking
# bench.py times -- is compared entry by entry: per-step stiffness forces of each field family to rel. L2 <= 1e-5
# (BASELINE.json's force tolerance), for three steps from a random admissible displacement, once verb by verb and once
# through ax3d_run_steps (CUDA graph replay + in-kernel Newmark).
@pytest.mark.timeout(1500)
@pytest.mark.parametrize("cfg", ["cfg2", "cfg3", "cfg4"])
def test_full_size_forces_match_c_oracle(cfg):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bench
    from helpers import build_oracle, build_gpu, randomize_displ, push_fields, compare_field
    from c_oracle import COracle
    bench.CFG = cfg
    m = bench.make_mesh(bench.N_THETA)
    assert m.nelem == 2016
    dt = m.estimate_dt()
    ora, _ = build_oracle(m, dt, np.float32, source=False)
    co = COracle(ora)
    randomize_displ(ora, seed=11)
    g1, _ = build_gpu(m, dt, source=False)
    g2, _ = build_gpu(m, dt, source=False)
    push_fields(ora, g1)
    push_fields(ora, g2)
    nstep = 3
    for i in range(nstep):
        co.step(dt, 0.0)
        g1.step(dt, 0.0)
        for which in ("stiff", "displ"):
            err = compare_field(ora, g1, which)
            assert err and all(v <= 1e-5 for v in err.values()), (cfg, "verbs", i, which, err)
    g2.runSteps(dt, np.zeros(nstep, np.float32))
    assert g2.checkStability()
    for which in ("stiff", "displ", "veloc"):
        err = compare_field(ora, g2, which)
        assert err and all(v <= 1e-5 for v in err.values()), (cfg, "run_steps", which, err)
