"""DualGraph::decompose / formNeighbourhood (S/preloop/graph/DualGraph.cpp:12-94) over the METIS inside the CUDA toolkit
(axisem3d_b200/host/dual_graph.cpp, partition.py): index width self-test, the dual graph against the direct node-incidence
neighbourhood of connectivity.py, and the properties the reference asks of the partition (contiguous parts, 1 % imbalance of
the weighted load, small edge cut), plus the measured-cost weights of bench.py."""
import numpy as np
import pytest

from axisem3d_b200 import connectivity as CN
from axisem3d_b200 import partition as PT
from axisem3d_b200.mesh_synth import SynthMesh


def test_metis_selftest_and_dual_graph():
    lib = PT.load()
    assert lib.ax3d_metis_selftest() == 0
    m = SynthMesh(n_theta=9, n_r=7, nu=3)
    nb1 = PT.dual_graph(m.conn, 1)
    ref = CN.form_neighbourhood(m.conn)
    assert all(sorted(int(x) for x in a) == sorted(b) for a, b in zip(nb1, ref))          # ncommon = 1: corner contacts included
    nb2 = PT.dual_graph(m.conn, 2)
    assert all(set(int(x) for x in a) <= set(b) for a, b in zip(nb2, ref))
    assert max(len(a) for a in nb2) == 4 and max(len(a) for a in nb1) == 8


@pytest.mark.parametrize("nproc", [2, 3, 4, 8])
def test_kway_partition_properties(nproc):
    m = SynthMesh(n_theta=48, n_r=24, nu_fn=lambda s, z: int(4 + 60 * s / 6371e3))
    w = m.e_nr.astype(np.float64) * np.log2(np.maximum(m.e_nr, 2)) + 16.0
    e2p, info = PT.partition_kway(m.conn, w, nproc, imbalance=0.01, ntrials=4)
    assert e2p.shape == (m.nelem,) and set(e2p.tolist()) == set(range(nproc))
    assert info["contiguous"]
    loads = np.array([w[e2p == r].sum() for r in range(nproc)])
    assert loads.max() / loads.mean() <= 1.03 and abs(info["imbalance"] - loads.max() / loads.mean()) < 1e-9
    # the cut is of the order of a few mesh lines, not of the element count
    assert 0 < info["edgecut"] < 6 * (m.nth + m.nr_)
    again, info2 = PT.partition_kway(m.conn, w, nproc, imbalance=0.01, ntrials=4)
    assert np.array_equal(e2p, again) and info2["edgecut"] == info["edgecut"]                 # deterministic for given seeds
    # every rank's halo lists are consistent with its neighbours' (Connectivity::decompose on this elemToProc)
    decs = [CN.decompose(m.conn, e2p, r, m.e2g, m.neighbours) for r in range(nproc)]
    for r, d in enumerate(decs):
        for q, pts in zip(d.iProcComm, d.iLocalPoints):
            other = decs[q]
            back = other.iLocalPoints[list(other.iProcComm).index(r)]
            assert [int(d.local_to_global_gll[t]) for t in pts] == [int(other.local_to_global_gll[t]) for t in back]


def test_single_part_and_errors():
    m = SynthMesh(n_theta=6, n_r=5, nu=2)
    e2p, info = PT.partition_kway(m.conn, None, 1)
    assert not e2p.any() and info["edgecut"] == 0
    with pytest.raises(RuntimeError, match="weights must be positive"):
        PT.partition_kway(m.conn, np.zeros(m.nelem), 2)


def test_bench_weights_use_the_measured_cost_model():
    import bench
    bench.CFG = "cfg4"
    m = bench.make_mesh(24, n_r=12)
    w = bench.element_weights(m)
    assert w.shape == (m.nelem,) and (w > 0).all()
    sol = ~m.is_fluid
    big, small = sol & (m.e_nr >= np.percentile(m.e_nr[sol], 90)), sol & (m.e_nr <= np.percentile(m.e_nr[sol], 10))
    assert w[big].mean() > 3 * w[small].mean()                      # cost grows with Nr
    assert w[m.is_fluid].mean() < 0.5 * w[sol].mean()               # 1D fluid elements are cheap
