"""Shared helpers of the parity tests: build the same synthetic domain for the oracle and the
CUDA library, move fields between the two layouts, compare."""
import numpy as np

from axisem3d_b200.mesh_synth import SynthMesh
from axisem_oracle import OracleDomain


def build_oracle(mesh, dt, dtype=np.float32, source=True, rank=0, elem_to_proc=None):
    d = OracleDomain(dtype)
    rel = mesh.release(d, dt, rank=rank, elem_to_proc=elem_to_proc)
    if source:
        st = mesh.make_source(rel["elements"], rel["dec"], amp=1e18)
        if st is not None:
            d.addSourceTerm(st)
    d.finalize()
    return d, rel


def build_gpu(mesh, dt, source=True, device=0):
    from axisem3d_b200.domain import Domain
    g = Domain(device)
    rel = mesh.release(g, dt)
    if source:
        st = mesh.make_source(rel["elements"], rel["dec"], amp=1e18)
        if st is not None:
            g.addSourceTerm(st)
    g.finalize()
    return g, rel


def oracle_to_bulk(d, which):
    """Oracle padded arrays -> (solid_bulk, fluid_bulk) complex64 in the C-ABI bulk layout
    (point-tag order; solid block = [3][Nu+1], fluid block = [Nu+1])."""
    s, f = [], []
    for t, p in enumerate(d.points):
        n = p.nu + 1
        if d.s_idx[t] >= 0:
            s.append(d.S[which][d.s_idx[t], :, :n].reshape(-1))
        if d.f_idx[t] >= 0:
            f.append(d.F[which][d.f_idx[t], :n])
    cat = lambda l: np.concatenate(l).astype(np.complex64) if l else np.zeros(0, np.complex64)
    return cat(s), cat(f)


def randomize_displ(d, seed=1, scale=1e-6):
    """complex normal * scale, masked like Point::randomDispl (SolidPoint.cpp:47-58)."""
    rng = np.random.default_rng(seed)
    for fld, rows in ((d.S, d.s_rows[:, None, :]), (d.F, d.f_rows)):
        a = fld["displ"]
        r = (rng.standard_normal(a.shape) + 1j * rng.standard_normal(a.shape)) * scale
        a[:] = np.where(np.broadcast_to(rows, a.shape), r, 0).astype(a.dtype)
    d.maskDispl()


def push_fields(d, g, names=("displ",)):
    for w in names:
        s, f = oracle_to_bulk(d, w)
        if s.size:
            g.set_bulk(w, False, s)
        if f.size:
            g.set_bulk(w, True, f)


def rel_l2(a, b):
    na = np.linalg.norm(a)
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / na) if na > 0 else float(np.linalg.norm(b))


def compare_field(d, g, which):
    s, f = oracle_to_bulk(d, which)
    out = {}
    if s.size:
        out["solid"] = rel_l2(s.astype(np.complex128), g.get_bulk(which, False).astype(np.complex128))
    if f.size:
        out["fluid"] = rel_l2(f.astype(np.complex128), g.get_bulk(which, True).astype(np.complex128))
    return out
