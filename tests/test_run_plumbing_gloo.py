"""Host logic of `python -m axisem3d_b200.run` on one and on two ranks (gloo, CPU): partition, per-rank release of the
template mesh, station ownership, the gather of the traces on rank 0 and the two station writers -- with a stand-in for the
CUDA domain that records nothing but which rank a station was recorded on.  (The same entry point on real devices:
tests/test_gpu_run_dir.py.)"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NSTEP = 30


class _StandInDomain:
    """what run.main needs from axisem3d_b200.domain.Domain, minus the device"""

    def __init__(self, dev):
        self.n, self.pts, self.els, self.nsrc = 0, [], [], 0

    def setGMat(self, *a):
        pass

    def addPoint(self, p):
        p.domain_tag = len(self.pts)
        self.pts.append(p)
        return p.domain_tag

    def addElement(self, e):
        e.domain_tag = len(self.els)
        self.els.append(e)
        return e.domain_tag

    def addSourceTerm(self, s):
        self.nsrc += 1

    def setMessaging(self, info, rank, world, uid):
        self.neigh = list(info.mIProcComm)

    def finalize(self):
        pass

    def connectHalo(self, info, rank, dist):
        every = [None] * dist.get_world_size()
        dist.all_gather_object(every, sorted(int(r) for r in info.mIProcComm))
        for r, nb in enumerate(every):                           # neighbourhood must be symmetric
            for q in nb:
                assert r in every[q]

    def setReceivers(self, tags, phi, w):
        self.n = len(tags)

    def runStepsRecord(self, dt, stf):
        return np.full((len(stf), self.n, 3), float(os.environ.get("RANK", "0")) + 1.0, np.float32)

    def runSteps(self, dt, stf):
        pass

    def checkStability(self):
        return True

    def synchronize(self):
        pass


def _run_dir(tmp):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import main_case as MC
    run = os.path.join(tmp, "run")
    os.makedirs(run)
    inp = MC.input_dir("cfg1_template", run)
    for name, key, val in (("inparam.advanced", "DEVELOP_MAX_TIME_STEPS", str(NSTEP)), ("inparam.time_src_recv", "OUT_STATIONS_FORMAT", "ascii netcdf"),
                           ("inparam.time_src_recv", "OUT_STATIONS_COMPONENTS", "SPZ")):
        path = os.path.join(inp, name)
        lines = [("%s %s" % (key, val)) if ln.split()[:1] == [key] else ln for ln in open(path).read().split("\n")]
        open(path, "w").write("\n".join(lines))
    return run


def _worker(rank, world, port, run):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch
    torch.cuda.set_device = lambda d: None
    import axisem3d_b200.domain as D
    D.Domain = _StandInDomain
    from axisem3d_b200 import run as R
    assert R.main([run]) == 0
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2])
def test_run_main_plumbing(world, tmp_path):
    run = _run_dir(str(tmp_path))
    if world == 1:
        import multiprocessing as mp
        p = mp.get_context("spawn").Process(target=_worker, args=(0, 1, 0, run))      # a fresh interpreter: the stand-in must not leak
        p.start()
        p.join()
        assert p.exitcode == 0
    else:
        import torch.multiprocessing as tmp_mp
        tmp_mp.spawn(_worker, args=(world, 29850 + (os.getpid() % 1000), run), nprocs=world, join=True)
    st = os.path.join(run, "output", "stations")
    files = sorted(f for f in os.listdir(st) if f.endswith(".ascii"))
    assert len(files) == 129
    vals = set()
    for f in files:
        a = np.loadtxt(os.path.join(st, f))
        assert a.shape == (NSTEP, 4) and abs(a[1, 0] - a[0, 0] - 0.435851) < 1e-5
        assert np.ptp(a[:, 1:]) == 0.0
        vals.add(float(a[0, 1]))
    assert vals == ({1.0} if world == 1 else {1.0, 2.0})                              # every station recorded on exactly one rank
    from scipy.io import netcdf_file
    with netcdf_file(os.path.join(st, "axisem3d_synthetics.nc"), "r", mmap=False) as nc:
        assert len(nc.variables) == 130 and nc.variables["IU.SSPA.SPZ"][:].shape == (NSTEP, 3)
