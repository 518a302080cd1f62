"""N > 1 host logic on CPU: two `torch.distributed` (gloo) ranks each own one part of the mesh (Connectivity::decompose),
run the reference's step order with the halo sum of Domain::assembleStiff (Domain.cpp:111-163) over send/recv, and
must reproduce the single-rank run on every point they own -- invariant (7) of SURVEY.md §4.  The ranks here are
oracle domains (the GPU library is exercised the same way by bench.py --gpus N on the box); what is under test is the
host side every backend shares: partition -> local numbering -> per-neighbour point lists -> pack order -> unpack-add."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MESH = dict(n_theta=8, n_r=6, nu=4, law="ti", model3d=True, attenuation="cg4")
NSTEP = 25


def _stf():
    return np.exp(-((np.arange(NSTEP) - 8) / 3.0) ** 2)


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from axisem3d_b200 import connectivity as CN
    from axisem3d_b200.mesh_synth import SynthMesh
    from axisem_oracle import OracleDomain

    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = SynthMesh(**MESH)
    dt = m.estimate_dt()
    e2p = CN.partition_contiguous(m.e_nr.astype(np.float64), world)
    d = OracleDomain(np.float64)
    rel = m.release(d, dt, rank=rank, elem_to_proc=e2p)
    st = m.make_source(rel["elements"], rel["dec"], amp=1e18)
    if st is not None:
        d.addSourceTerm(st)
    info = rel["msg"]

    def exchange(send):
        recv = [torch.zeros(2 * len(b), dtype=torch.float64) for b in send]
        reqs = []
        for peer, b, r in zip(info.mIProcComm, send, recv):
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(b, dtype=np.complex128)).view(np.float64).copy())
            reqs.append(dist.isend(t, peer))
            reqs.append(dist.irecv(r, peer))
        for q in reqs:
            q.wait()
        return [r.numpy().view(np.complex128) for r in recv]

    d.setMessaging(info, exchange)
    d.finalize()
    for s in _stf():
        d.step(dt, float(s))
    l2g = rel["dec"].local_to_global_gll
    sol = {int(l2g[t]): d.get_solid(t, "displ") for t in range(len(d.points)) if d.s_idx[t] >= 0}
    flu = {int(l2g[t]): d.get_fluid(t, "displ") for t in range(len(d.points)) if d.f_idx[t] >= 0}
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.array([sol, flu], dtype=object), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_halo_sum_matches_single_rank(tmp_path):
    import torch.multiprocessing as mp
    from helpers import build_oracle
    from axisem3d_b200.mesh_synth import SynthMesh

    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)

    m = SynthMesh(**MESH)
    dt = m.estimate_dt()
    ref, _ = build_oracle(m, dt, np.float64)
    for s in _stf():
        ref.step(dt, float(s))
    scale = float(np.abs(ref.S["displ"]).max())
    fscale = float(np.abs(ref.F["displ"]).max())       # the fluid potential has its own magnitude
    assert scale > 0 and fscale > 0
    seen = set()
    for r in range(world):
        sol, flu = np.load(os.path.join(str(tmp_path), "rank%d.npy" % r), allow_pickle=True)
        for g, u in sol.items():
            assert np.abs(u - ref.get_solid(g, "displ")).max() <= 1e-10 * scale, (r, g)
            seen.add(g)
        for g, u in flu.items():
            assert np.abs(u - ref.get_fluid(g, "displ")).max() <= 1e-10 * fscale, (r, g)
            seen.add(g)
    assert seen == set(range(m.ngll))          # every global GLL point is owned by some rank
