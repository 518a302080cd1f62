"""Index maps and partition halos (integer work; BASELINE.json north_star: "gather/scatter index maps and
partition halos must be bit-exact").  The reference pins none of this with fixtures, so the restated
algorithm (axisem3d_b200/connectivity.py <- S/preloop/graph/Connectivity.cpp:42-221) is checked against
its defining properties on structured meshes whose numbering can be derived by hand."""
import numpy as np
import pytest

from axisem3d_b200 import connectivity as CN
from axisem3d_b200.mesh_synth import SynthMesh


def strip_mesh(n):
    """n quads in a row: nodes 0..n on the bottom, n+1..2n+1 on the top, counter-clockwise connectivity."""
    conn = [[i, i + 1, n + 2 + i, n + 1 + i] for i in range(n)]
    return np.array(conn)


def test_elem_to_gll_strip_by_hand():
    """Connectivity::formElemToGLL (42-94): element 0 gets tags 0..24 row-major (ipol, jpol); element 1 shares its
    ipol = 0 edge with element 0's ipol = 4 edge and numbers the remaining 20 points row-major."""
    ngll, e2g = CN.form_elem_to_gll(strip_mesh(3))
    assert ngll == 25 + 20 + 20
    assert np.array_equal(e2g[0].reshape(-1), np.arange(25))
    assert np.array_equal(e2g[1, 0, :], e2g[0, 4, :])
    assert np.array_equal(e2g[1, 1:, :].reshape(-1), np.arange(25, 45))
    assert np.array_equal(e2g[2, 0, :], e2g[1, 4, :])


def test_global_numbering_is_a_bijection_on_shared_points():
    m = SynthMesh(n_theta=6, n_r=5, nu=3)
    e2g = m.e2g
    assert e2g.min() == 0 and e2g.max() == m.ngll - 1
    assert len(np.unique(e2g)) == m.ngll
    # geometric check: two (element, ipol, jpol) with the same tag sit at the same (s, z)
    crd = {}
    for e in range(m.nelem):
        g = m.geo[e]
        for i in range(5):
            for j in range(5):
                t = int(e2g[e, i, j])
                sz = np.array([g["s"][i, j], g["z"][i, j]])
                if t in crd:
                    assert np.allclose(crd[t], sz, rtol=0, atol=1e-3)      # metres, Earth-sized mesh
                else:
                    crd[t] = sz
    # and distinct tags are distinct points
    pts = np.array([crd[t] for t in range(m.ngll)])
    assert len(np.unique(np.round(pts, 1), axis=0)) == m.ngll


@pytest.mark.parametrize("nproc", [2, 3, 4])
def test_decompose_halos_are_symmetric_and_ordered(nproc):
    """Connectivity::decompose (96-221): both sides of a rank pair list the SAME global GLL tags in the SAME
    (ascending global tag) order -- that is what makes the packed halo buffers line up without exchanging indices
    (Domain.cpp:111-163) -- and local numbering is the global numbering restricted to the rank, first-seen order."""
    m = SynthMesh(n_theta=8, n_r=6, nu=3)
    e2p = CN.partition_contiguous(m.e_nr.astype(np.float64), nproc)
    assert set(np.unique(e2p)) == set(range(nproc))
    decs = [CN.decompose(m.conn, e2p, r, m.e2g, m.neighbours) for r in range(nproc)]
    owned = np.zeros(m.nelem, dtype=int)
    for r, d in enumerate(decs):
        owned[d.local_elems] += 1
        assert np.all(np.diff(d.local_elems) > 0)                       # global-id order within the rank
        assert d.iProcComm == sorted(d.iProcComm)
        # local tags: element by element, row-major, new tag unless seen before
        seen, nxt = {}, 0
        for il, e in enumerate(d.local_elems):
            for i in range(5):
                for j in range(5):
                    g = int(m.e2g[e, i, j])
                    if g not in seen:
                        seen[g] = nxt
                        nxt += 1
                    assert d.elemToGllLocal[il, i, j] == seen[g]
        assert d.nGllLocal == nxt
        for other, loc, glb in zip(d.iProcComm, d.iLocalPoints, d.iGlobalPoints):
            assert glb == sorted(glb) and len(set(glb)) == len(glb)
            assert [int(d.local_to_global_gll[t]) for t in loc] == glb
            k = decs[other].iProcComm.index(r)
            assert decs[other].iGlobalPoints[k] == glb                  # bit-exact halo lists on both sides
            # the halo is exactly the set of GLL points both ranks touch
            mine = set(int(t) for t in d.local_to_global_gll)
            theirs = set(int(t) for t in decs[other].local_to_global_gll)
            assert set(glb) == (mine & theirs)
    assert np.all(owned == 1)


def test_partition_contiguous_balances_weight():
    w = np.random.default_rng(0).uniform(1, 10, 500)
    for n in (2, 4, 8):
        p = CN.partition_contiguous(w, n)
        assert np.all(np.diff(p) >= 0) and p[0] == 0 and p[-1] == n - 1
        loads = np.array([w[p == r].sum() for r in range(n)])
        assert loads.max() / loads.mean() < 1.05
