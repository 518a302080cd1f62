"""The drop-in of INTEGRATION.md section 1, executed: oracle/_ref/axisem3d_gpu is the REFERENCE's own main program, preloop and
station recorder (axisem.cpp, preloop/**, 3d_model/**, core/output/**, compiled unmodified from /root/reference by
oracle/Makefile.dropin) linked against axisem3d_b200/host/ax3d_reference_binding.hpp + libaxisem3d_b200.so in place of its
core/{domain,element,point,newmark,fftw,source}.  It is run in a run directory like `./axisem3d`; the station file it writes
through the reference's PointwiseRecorder is compared with the one the pure reference program wrote
(tests/golden/main_cfg1_template.npz).  The binary is built in the container that has /root/reference (it travels with the
snapshot); where it is absent the test has nothing to run."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "axisem3d_gpu")
# (case, steps): the template itself (first 600 of its 2000 steps) and every other reference-main case at full length
CASES = (("cfg1_template", 600), ("emp_full_enz", 400), ("bubbles_3d", 600), ("ellipticity_prt", 300), ("pointforce_spz", 400),
         ("wisdom_learn", 300))


@pytest.mark.gpu
@pytest.mark.parametrize("name,nstep", CASES)
def test_reference_main_program_runs_on_the_cuda_path(name, nstep, tmp_path):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/axisem3d_gpu is built from /root/reference (make -C oracle -f Makefile.dropin), which this box does not have")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import main_case as MC
    from nc_flatten import flatten, read_flat
    run = os.path.join(str(tmp_path), "run")
    os.makedirs(run)
    inp = MC.input_dir(name, run)                                # template inputs + the case's overrides (NetCDF station format)
    path = os.path.join(inp, "inparam.advanced")
    lines = [("DEVELOP_MAX_TIME_STEPS %d" % nstep) if ln.split()[:1] == ["DEVELOP_MAX_TIME_STEPS"] else ln for ln in open(path).read().split("\n")]
    open(path, "w").write("\n".join(lines))
    flatten(os.path.join(inp, MC.MESH))                          # the NetCDF stand-in of this build reads <mesh>.ncflat
    exe = os.path.join(run, "axisem3d_gpu")
    os.symlink(EXE, exe)
    r = subprocess.run([exe], cwd=run, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ABORTED" not in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
    nc = read_flat(os.path.join(run, "output", "stations", "axisem3d_synthetics.nc.ncflat"))
    gold = MC.golden(name)
    n = nstep // gold["stride"]
    got = np.stack([nc[k] for k in gold["keys"]]).astype(np.float64)[:, ::gold["stride"]][:, :n]
    ref = gold["seis"].astype(np.float64)[:, :n]
    assert np.abs(nc["time_points"][::gold["stride"]][:n] - gold["time"][:n]).max() < 1e-9
    mis = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
    log = os.environ.get("AX3D_MISFIT_LOG")
    if log:
        with open(log, "a") as f:
            f.write("%-16s %-12s %5d steps   rel. L2 over all stations %.3e   (reference main + preloop + recorder, CUDA core)\n" % (name, "drop-in", nstep, mis))
    assert mis <= 1e-4, mis
    if name == "wisdom_learn":                                   # Domain::dumpWisdom through the reference's own NuWisdom writer
        import test_wisdom_reference as TW
        w = read_flat(os.path.join(run, "output", "learn.nu_wisdom.nc.ncflat"))["axisem3d_wisdom"]
        sz, learn, orign = TW._gold()
        near = np.hypot(sz[:, 0], sz[:, 1] - (6371e3 - 12e3)) < 2500e3
        got_nu = np.round(w[:, 2]).astype(np.int64)
        assert np.array_equal(np.round(w[:, 3]).astype(np.int64), orign)
        assert (got_nu[near] == learn[near]).all() and float((got_nu == learn).mean()) >= 0.95
