"""End to end on the reference's own inputs: the unchanged template input directory (Exodus mesh, inparam.*, CMTSOLUTION,
STATIONS) through the repo's preloop and time loop, against the station seismograms written by the REFERENCE's whole program
(oracle/_ref/axisem3d_ref = axisem.cpp's main over the stand-ins of oracle/shim; tests/golden/main_<case>.npz, written by
oracle/make_golden_main.py).  BASELINE.json's third check: relative L2 misfit <= 1e-4 after 2000 steps.

CPU: the oracle's time loop (oracle.c through oracle/c_oracle.py, fp32) -- this pins the preloop restatement + oracle as a whole.
GPU: ax3d_run_steps_record on the CUDA domain (`-m gpu`)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import main_case as MC  # noqa: E402

TOL = 1e-4
# the CPU oracle follows only the first steps of every case (long enough for the near stations to carry the body and surface
# waves); the CUDA test runs all of them (2000 for cfg1_template)
CPU_STEPS = {"cfg1_template": 300, "emp_full_enz": 120, "bubbles_3d": 100, "ellipticity_prt": 120, "pointforce_spz": 200, "wisdom_learn": 100,
             "ellipticity_pole": 60, "cylinder_3d": 80, "deep_stations": 200, "ocean_on_ellipsoid": 40}


def _log(name, impl, steps, tot, worst):
    """AX3D_MISFIT_LOG=<file>: keep the measured misfits (profiles/ quotes them)"""
    path = os.environ.get("AX3D_MISFIT_LOG")
    if path:
        with open(path, "a") as f:
            f.write("%-16s %-12s %5d steps   rel. L2 over all stations %.3e   worst live trace %.3e\n" % (name, impl, steps, tot, worst))


def _misfit(got, ref):
    """relative L2 misfit over all stations and components, and the worst single trace among those that carry signal"""
    tot = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
    amp = np.linalg.norm(ref, axis=1)                                     # [nrec][3]
    live = amp > 1e-3 * amp.max()
    per = np.linalg.norm(got - ref, axis=1)[live] / amp[live]
    return tot, float(per.max())


@pytest.mark.parametrize("name", MC.CASES + MC.CPU_ONLY_CASES)
def test_oracle_seismograms_match_reference_main(name):
    from axisem_oracle import OracleDomain
    from c_oracle import COracle
    if name in ("ellipticity_prt", "ocean_on_ellipsoid") and not os.environ.get("AX3D_SLOW_TESTS"):
        # oracle.c has no particle-relabelling path, the numpy oracle needs ~1.2 s per step on this case (2.5 min for the 120
        # steps the near stations need); run with AX3D_SLOW_TESTS=1 (passes: 1e-6).  The case's preloop arrays are compared
        # in test_preloop_reference.py and its seismograms, all 300 steps, by the CUDA test below.
        pytest.skip("slow (numpy oracle, particle relabelling): set AX3D_SLOW_TESTS=1")
    gold = MC.golden(name)
    case = MC.get_case(name)
    if True:
        d = OracleDomain(np.float32)
        rel = case.release(d)
        d.finalize()
        try:
            co = COracle(d)
        except NotImplementedError:           # oracle.c has no particle-relabelling path: the numpy oracle steps those cases
            co = d
        rc = case.receivers
        tags = [rel["elements"][int(q)].domain_tag for q in rc.quad]
        assert len(case.stf) == gold["steps"]
        nstep = min(gold["steps"], CPU_STEPS[name])
        got = []
        for i in range(nstep):
            co.updateNewmark(case.dt)
            if i % gold["stride"] == 0:                      # Domain::record follows the update of the step (Newmark.cpp:49-70)
                got.append([d.ground_motion(t, float(p), w) for t, p, w in zip(tags, rc.phi, rc.weights)])
            d.applySource(float(case.stf[i]))
            co.computeStiff()
            d.coupleSolidFluid()
        got = rc.rotate(np.array(got)).transpose(1, 0, 2).astype(np.float64)          # [nrec][nt][3]
        assert rc.keys == [k.rsplit(".", 1)[0] for k in gold["keys"]]
        tot, worst = _misfit(got, gold["seis"].astype(np.float64)[:, :got.shape[1]])
        _log(name, "oracle (CPU)", nstep, tot, worst)
        assert tot <= TOL and worst <= 10 * TOL, (tot, worst)


@pytest.mark.gpu
@pytest.mark.parametrize("name", MC.CASES)
def test_cuda_seismograms_match_reference_main(name):
    from axisem3d_b200.domain import Domain
    gold = MC.golden(name)
    case = MC.get_case(name)
    if True:
        g = Domain(0)
        rel = case.release(g)
        g.finalize()
        rc = case.receivers
        mine = rc.release(g, rel["elements"])
        assert len(mine) == len(rc.keys)
        series = np.asarray(g.runStepsRecord(case.dt, case.stf))                    # [step][nrec][3], SPZ
        assert g.checkStability()
        got = rc.rotate(series)[::gold["stride"]].transpose(1, 0, 2).astype(np.float64)
        tot, worst = _misfit(got, gold["seis"].astype(np.float64))
        _log(name, "CUDA", gold["steps"], tot, worst)
        assert tot <= TOL and worst <= 10 * TOL, (tot, worst)
