"""N > 1 on real GPUs: two ranks, one GPU each, run the partitioned mesh through the C-ABI with the halo sum of
Domain::assembleStiff (Domain.cpp:111-163) done (a) by the peer-memory kernels k_halo_put / k_halo_wait_add over
NVLink (steps replayed as CUDA graphs) and (b) by the NCCL send/recv group (eager steps).  Every point a rank owns must
match the single-domain fp64 oracle within the seismogram tolerance of BASELINE.json (1e-4, relative to the field's
magnitude), and the two halo transports must agree to rounding (same pack and sum order; the element scatter uses
floating-point atomics whose order differs between two runs, so the last bits may differ).
With 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`) each rank has its own device and both
transports are compared.  On a 1-GPU box the two rank processes share device 0: the windows are still exchanged as CUDA IPC
handles between processes and the peer-memory kernels still order the ranks through the arrival counters (the processes are
time-sliced on the device), but NCCL refuses two ranks on one GPU, so only the peer transport runs there."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

MESH = dict(n_theta=8, n_r=6, nu=12, law="ti", model3d=True, attenuation="cg4")
NSTEP = 30
TOL = 1e-4


def _stf():
    return np.exp(-((np.arange(NSTEP) - 8) / 3.0) ** 2).astype(np.float32)


def _worker(rank, world, port, out_dir, halo):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from axisem3d_b200 import connectivity as CN
    from axisem3d_b200.domain import Domain, nccl_unique_id
    from axisem3d_b200.mesh_synth import SynthMesh

    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo", rank=rank, world_size=world)   # control plane only (handles, unique id)
    m = SynthMesh(**MESH)
    dt = m.estimate_dt()
    e2p = CN.partition_contiguous(m.e_nr.astype(np.float64), world)
    d = Domain(dev)
    rel = m.release(d, dt, rank=rank, elem_to_proc=e2p)
    st = m.make_source(rel["elements"], rel["dec"], amp=1e18)
    if st is not None:
        d.addSourceTerm(st)
    uid = [nccl_unique_id() if (rank == 0 and halo == "nccl") else None]
    dist.broadcast_object_list(uid, 0)
    d.setMessaging(rel["msg"], rank, world, uid[0])
    d.finalize()
    if halo == "peer":
        d.connectHalo(rel["msg"], rank, dist)
    d.runSteps(dt, _stf())
    assert d.checkStability()
    l2g = rel["dec"].local_to_global_gll
    sol, flu = {}, {}
    for t, p in enumerate(d.points):
        if p.kind != "fluid":
            sol[int(l2g[t])] = d.get_solid(t, "displ")
        if p.kind != "solid":
            flu[int(l2g[t])] = d.get_fluid(t, "displ")
    np.save(os.path.join(out_dir, "%s_rank%d.npy" % (halo, rank)), np.array([sol, flu], dtype=object), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_two_gpu_halo_peer_and_nccl_match_oracle(tmp_path):
    import torch
    two = torch.cuda.device_count() >= 2
    import torch.multiprocessing as mp
    from helpers import build_oracle
    from axisem3d_b200.mesh_synth import SynthMesh

    world = 2
    for k, halo in enumerate(("peer", "nccl") if two else ("peer",)):
        port = 29600 + (os.getpid() % 2000) + k
        mp.spawn(_worker, args=(world, port, str(tmp_path), halo), nprocs=world, join=True)

    m = SynthMesh(**MESH)
    dt = m.estimate_dt()
    ref, _ = build_oracle(m, dt, np.float64)
    for s in _stf():
        ref.step(dt, float(s))
    scale = float(np.abs(ref.S["displ"]).max())
    fscale = float(np.abs(ref.F["displ"]).max())
    assert scale > 0 and fscale > 0
    seen = set()
    for r in range(world):
        sol, flu = np.load(os.path.join(str(tmp_path), "peer_rank%d.npy" % r), allow_pickle=True)
        sol2, flu2 = np.load(os.path.join(str(tmp_path), "%s_rank%d.npy" % ("nccl" if two else "peer", r)), allow_pickle=True)
        for g, u in sol.items():
            assert np.abs(u - ref.get_solid(g, "displ")).max() <= TOL * scale, (r, g)
            assert np.abs(u - sol2[g]).max() <= 1e-5 * scale, ("peer vs nccl", r, g)
            seen.add(g)
        for g, u in flu.items():
            assert np.abs(u - ref.get_fluid(g, "displ")).max() <= TOL * fscale, (r, g)
            assert np.abs(u - flu2[g]).max() <= 1e-5 * fscale, ("peer vs nccl", r, g)
            seen.add(g)
    assert seen == set(range(m.ngll))
