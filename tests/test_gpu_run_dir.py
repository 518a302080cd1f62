"""`python -m axisem3d_b200.run <run_dir>`: the reference's template run directory, unchanged, end to end on the CUDA path --
on one rank and on two ranks (METIS partition, halo sum over peer-memory windows; on a one-GPU box the two rank processes share
device 0 like tests/test_gpu_multirank.py) -- against the station seismograms of the REFERENCE's whole program
(tests/golden/main_cfg1_template.npz, oracle/make_golden_main.py)."""
import os
import shutil
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NSTEP = 600
TOL = 1e-4          # the ascii station files carry 6 significant digits


def _run_dir(tmp_path):
    import main_case as MC
    run = os.path.join(str(tmp_path), "run")
    os.makedirs(run)
    inp = MC.input_dir("cfg1_template", run)
    assert inp == os.path.join(run, "input")
    path = os.path.join(inp, "inparam.advanced")
    lines = [("DEVELOP_MAX_TIME_STEPS %d" % NSTEP) if ln.split()[:1] == ["DEVELOP_MAX_TIME_STEPS"] else ln for ln in open(path).read().split("\n")]
    open(path, "w").write("\n".join(lines))
    _set(inp, "inparam.time_src_recv", "OUT_STATIONS_FORMAT", "ascii netcdf")     # both writers of run.py
    return run


def _set(inp, name, key, value):
    path = os.path.join(inp, name)
    lines = [("%s %s" % (key, value)) if ln.split()[:1] == [key] else ln for ln in open(path).read().split("\n")]
    open(path, "w").write("\n".join(lines))


def _check(run):
    import main_case as MC
    gold = MC.golden("cfg1_template")
    st = os.path.join(run, "output", "stations")
    got = np.stack([np.loadtxt(os.path.join(st, k + ".ascii")) for k in gold["keys"]])         # [nrec][nstep][1 + 3]
    assert got.shape[1] == NSTEP
    n = NSTEP // gold["stride"]
    assert np.abs(got[0, ::gold["stride"], 0][:n] - gold["time"][:n]).max() < 1e-3
    a, b = got[:, ::gold["stride"], 1:][:, :n], gold["seis"].astype(np.float64)[:, :n]
    mis = float(np.linalg.norm(a - b) / np.linalg.norm(b))
    assert mis <= TOL, mis
    # the NetCDF writer carries the same traces in full fp32
    from scipy.io import netcdf_file
    with netcdf_file(os.path.join(st, "axisem3d_synthetics.nc"), "r", mmap=False) as nc:
        assert np.abs(np.array(nc.variables["time_points"][:]) - got[0, :, 0]).max() < 1e-3
        k = gold["keys"][len(gold["keys"]) // 2]
        tr = np.array(nc.variables[k][:], dtype=np.float64)
        assert tr.shape == (NSTEP, 3)
        asc = got[gold["keys"].index(k), :, 1:]
        assert np.abs(tr - asc).max() <= 2e-6 * max(np.abs(asc).max(), 1e-30) + 1e-30


def _worker(rank, world, port, run):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from axisem3d_b200 import run as R
    assert R.main([run]) == 0
    import torch.distributed as dist
    dist.destroy_process_group()


@pytest.mark.gpu
def test_run_dir_single_rank(tmp_path):
    from axisem3d_b200 import run as R
    run = _run_dir(tmp_path)
    assert R.main([run]) == 0
    _check(run)


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_run_dir_two_ranks(tmp_path):
    import torch.multiprocessing as mp
    run = _run_dir(tmp_path)
    mp.spawn(_worker, args=(2, 29700 + (os.getpid() % 2000), run), nprocs=2, join=True)
    _check(run)
    shutil.rmtree(run, ignore_errors=True)


@pytest.mark.gpu
def test_run_dir_learns_and_reuses_wisdom(tmp_path):
    """The reference's two-run workflow: a learning run (NU_WISDOM_LEARN true) leaves output/<name>.nu_wisdom.nc behind
    (Domain::dumpWisdom), a second run reads it as its Nu field (NU_TYPE wisdom, WisdomNrField).  The learnt values are
    compared with the reference program's own wisdom file (tests/golden/main_wisdom_learn.npz)."""
    import main_case as MC
    from axisem3d_b200 import preloop as PL
    from axisem3d_b200 import run as R
    import test_wisdom_reference as TW
    run = os.path.join(str(tmp_path), "learn")
    os.makedirs(run)
    inp = MC.input_dir("wisdom_learn", run)
    assert R.main([run]) == 0
    wis = PL.NuWisdom.read(os.path.join(run, "output", "learn.nu_wisdom.nc"))
    sz, learn, orign = TW._gold()
    assert np.array_equal(wis.nu_orign, orign) and np.abs(wis.sz - sz).max() < 1.0
    near = np.hypot(sz[:, 0], sz[:, 1] - (6371e3 - 12e3)) < 2500e3
    assert (wis.nu_learn[near] == learn[near]).all() and float((wis.nu_learn == learn).mean()) >= 0.95
    # second run: the learnt field as NU_TYPE wisdom
    run2 = os.path.join(str(tmp_path), "reuse")
    os.makedirs(run2)
    inp2 = MC.input_dir("wisdom_learn", run2)
    shutil.copy(os.path.join(run, "output", "learn.nu_wisdom.nc"), os.path.join(inp2, "learn.nu_wisdom.nc"))
    _set(inp2, "inparam.nu", "NU_TYPE", "wisdom")
    _set(inp2, "inparam.nu", "NU_WISDOM_LEARN", "false")
    _set(inp2, "inparam.nu", "NU_WISDOM_REUSE_INPUT", "learn.nu_wisdom.nc")
    _set(inp2, "inparam.advanced", "DEVELOP_MAX_TIME_STEPS", "100")
    _set(inp2, "inparam.time_src_recv", "OUT_STATIONS_FORMAT", "ascii")
    assert R.main([run2]) == 0
    a = np.loadtxt(os.path.join(run2, "output", "stations", "IU.SSPA.RTZ.ascii"))
    assert a.shape == (100, 4) and np.isfinite(a).all()
