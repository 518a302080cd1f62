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
NSTEP = 600


@pytest.mark.gpu
def test_reference_main_program_runs_on_the_cuda_path(tmp_path):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/axisem3d_gpu is built from /root/reference (make -C oracle -f Makefile.dropin), which this box does not have")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import main_case as MC
    from nc_flatten import flatten, read_flat
    run = os.path.join(str(tmp_path), "run")
    os.makedirs(run)
    inp = MC.input_dir("cfg1_template", run)                     # template inputs; the golden case already asks for NetCDF stations
    path = os.path.join(inp, "inparam.advanced")
    lines = [("DEVELOP_MAX_TIME_STEPS %d" % NSTEP) if ln.split()[:1] == ["DEVELOP_MAX_TIME_STEPS"] else ln for ln in open(path).read().split("\n")]
    open(path, "w").write("\n".join(lines))
    flatten(os.path.join(inp, MC.MESH))                          # the NetCDF stand-in of this build reads <mesh>.ncflat
    exe = os.path.join(run, "axisem3d_gpu")
    os.symlink(EXE, exe)
    r = subprocess.run([exe], cwd=run, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ABORTED" not in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
    nc = read_flat(os.path.join(run, "output", "stations", "axisem3d_synthetics.nc.ncflat"))
    gold = MC.golden("cfg1_template")
    n = NSTEP // gold["stride"]
    got = np.stack([nc[k] for k in gold["keys"]]).astype(np.float64)[:, ::gold["stride"]][:, :n]
    ref = gold["seis"].astype(np.float64)[:, :n]
    assert np.abs(nc["time_points"][::gold["stride"]][:n] - gold["time"][:n]).max() < 1e-9
    mis = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
    assert mis <= 1e-4, mis
    log = os.environ.get("AX3D_MISFIT_LOG")
    if log:
        with open(log, "a") as f:
            f.write("%-16s %-12s %5d steps   rel. L2 over all stations %.3e   (reference main + preloop + recorder, CUDA core)\n" % ("cfg1_template", "drop-in", NSTEP, mis))
