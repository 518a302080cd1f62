"""Domain::assembleStiff (Domain.cpp:111-163) on ONE GPU: the mesh is cut into 2 strips, and into 2 x 2 blocks whose
central corner point is shared by all four ranks (three neighbours per rank, non-zero peer_slot / peer_begin), every rank
is its own `ax3d_domain` on device 0, and the ranks are connected through `ax3d_halo_connect(ptrs = ...)` -- the
same-process form of the peer-memory halo (k_halo_put into the neighbour's window, k_halo_wait_add in neighbour-rank
order, both inside the step graph).  The ranks advance in lock step from one host thread: every call enqueues a few
steps on the rank's own stream and returns, the arrival counters order the streams on the device.

Checked: every point a rank owns against the single-domain fp64 oracle (1e-4 of the field's magnitude after 30 steps =
BASELINE.json's seismogram tolerance) and against a single-domain CUDA run (1e-5: same arithmetic, other summation order
at the shared points), and that the copies of a shared point held by different ranks agree to rounding."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

NSTEP = 30
CHUNK = 5


def _stf():
    return np.exp(-((np.arange(NSTEP) - 8) / 3.0) ** 2).astype(np.float32)


def _partition(mesh, kind):
    a, b = mesh.ab[:, 0], mesh.ab[:, 1]
    if kind == "strips2":
        return (a >= mesh.nth // 2).astype(np.int64), 2
    if kind == "blocks4":       # 2 x 2 blocks: the point in the middle belongs to all four ranks
        return (2 * (a >= mesh.nth // 2) + (b >= mesh.nr_ // 2)).astype(np.int64), 4
    if kind == "strips3":       # the middle rank has two neighbours, the outer ones one each
        return np.minimum(a * 3 // mesh.nth, 2).astype(np.int64), 3
    raise ValueError(kind)


@pytest.mark.timeout(900)
@pytest.mark.parametrize("kind,mesh_kw", [
    ("strips2", dict(n_theta=8, n_r=6, nu=12, law="ti", model3d=True, attenuation="cg4")),
    ("blocks4", dict(n_theta=8, n_r=6, nu=12, law="iso", model3d=True, attenuation=None)),
    ("blocks4", dict(n_theta=6, n_r=6, nu=3, law="ti", model3d=False, attenuation="cg4")),
    ("strips3", dict(n_theta=9, n_r=5, nu=10, law="aniso", model3d=True, attenuation="full")),
])
@pytest.mark.parametrize("inkernel", [False, True], ids=["put_kernels", "inkernel_put"])
def test_same_process_peer_halo_matches_oracle(kind, mesh_kw, inkernel, monkeypatch):
    # AX3D_INKERNEL_PUT (read at ax3d_halo_connect): the solid element kernel sends the boundary forces itself (fused.cuh:
    # halo_put_cta) instead of k_halo_put kernels behind the coupling
    monkeypatch.setenv("AX3D_INKERNEL_PUT", "1" if inkernel else "0")
    from helpers import build_oracle, build_gpu
    from axisem3d_b200.domain import Domain
    from axisem3d_b200.mesh_synth import SynthMesh

    m = SynthMesh(**mesh_kw)
    dt = m.estimate_dt()
    e2p, world = _partition(m, kind)
    assert len(set(e2p.tolist())) == world
    doms, rels = [], []
    for r in range(world):
        d = Domain(0)
        rel = m.release(d, dt, rank=r, elem_to_proc=e2p)
        st = m.make_source(rel["elements"], rel["dec"], amp=1e18)
        if st is not None:
            d.addSourceTerm(st)
        d.setMessaging(rel["msg"], r, world, None)
        d.finalize()
        doms.append(d)
        rels.append(rel)
    if kind == "blocks4":
        assert all(len(rel["msg"].mIProcComm) == 3 for rel in rels)      # corner contact: everyone neighbours everyone
    wins = [d.haloExport([int(x) for x in rel["msg"].mIProcComm]) for d, rel in zip(doms, rels)]
    for r, (d, rel) in enumerate(zip(doms, rels)):
        neigh = [int(x) for x in rel["msg"].mIProcComm]
        d.haloConnect(r, neigh, {q: wins[q] for q in neigh}, same_process=True)
    stf = _stf()
    for s0 in range(0, NSTEP, CHUNK):
        for d in doms:
            d.runSteps(dt, stf[s0:s0 + CHUNK])
    for d in doms:
        assert d.checkStability()

    ref, _ = build_oracle(m, dt, np.float64)
    for s in stf:
        ref.step(dt, float(s))
    one, _ = build_gpu(m, dt)
    one.runSteps(dt, stf)
    scale = float(np.abs(ref.S["displ"]).max())
    fscale = float(np.abs(ref.F["displ"]).max()) if ref.F["displ"].size else 0.0
    assert scale > 0
    seen_s, seen_f = {}, {}
    for r, (d, rel) in enumerate(zip(doms, rels)):
        l2g = rel["dec"].local_to_global_gll
        for t, p in enumerate(d.points):
            g = int(l2g[t])
            if p.kind != "fluid":
                u = d.get_solid(t, "displ")
                assert np.abs(u - ref.get_solid(g, "displ")).max() <= 1e-4 * scale, (kind, r, g)
                assert np.abs(u - one.get_solid(g, "displ")).max() <= 1e-5 * scale, (kind, "vs one domain", r, g)
                if g in seen_s:      # copies of a shared point: same sum in the same neighbour order up to the scatter atomics
                    assert np.abs(u - seen_s[g]).max() <= 2e-6 * scale, (kind, "copies", r, g)
                seen_s[g] = u
            if p.kind != "solid":
                u = d.get_fluid(t, "displ")
                assert np.abs(u - ref.get_fluid(g, "displ")).max() <= 1e-4 * fscale, (kind, r, g)
                assert np.abs(u - one.get_fluid(g, "displ")).max() <= 1e-5 * fscale, (kind, "vs one domain", r, g)
                if g in seen_f:
                    assert np.abs(u - seen_f[g]).max() <= 2e-6 * fscale, (kind, "copies", r, g)
                seen_f[g] = u
    assert set(seen_s) | set(seen_f) == set(range(m.ngll))
