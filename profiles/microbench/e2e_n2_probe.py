"""Probe of the e2e (ax3d_run_steps_record) path at N > 1: wall time of successive calls on each rank.
torchrun --nproc-per-node 2 profiles/microbench/e2e_n2_probe.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import bench
from axisem3d_b200 import connectivity as CN
from axisem3d_b200.domain import Domain, nccl_unique_id

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mesh = bench.make_mesh(bench.N_THETA * world)
dt = mesh.estimate_dt()
dom = Domain(local)
e2p = CN.partition_contiguous(mesh.e_nr.astype(np.float64), world) if world > 1 else None
rel = mesh.release(dom, dt, rank=rank, elem_to_proc=e2p)
if world > 1:
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    dom.setMessaging(rel["msg"], rank, world, bytes(uid.cpu().numpy().tolist()))
dom.finalize()
surf = [e.domain_tag for e in rel["elements"] if e.kind == "solid"]
rng = np.random.default_rng(5)
etags = [surf[i] for i in rng.integers(0, len(surf), 128)]
w = rng.uniform(0, 1, (128, 25)); w /= w.sum(axis=1, keepdims=True)
dom.setReceivers(etags, rng.uniform(0, 2 * np.pi, 128), w)
stf = bench.stf_series(400)
def sync():
    if world > 1: dist.barrier()
    dom.synchronize(); torch.cuda.synchronize()
dom.runSteps(dt, stf[:5]); sync()
for name, fn in (("runSteps 25", lambda: dom.runSteps(dt, stf[:25])), ("runStepsRecord 25", lambda: dom.runStepsRecord(dt, stf[:25]))):
    for rep in range(4):
        sync(); t0 = time.perf_counter(); fn(); t1 = time.perf_counter(); dom.synchronize(); t2 = time.perf_counter()
        print("rank %d %s rep %d: call %.3f ms, +sync %.3f ms -> %.3f ms/step" % (rank, name, rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t2 - t0) / 25), flush=True)
if world > 1: dist.destroy_process_group()
