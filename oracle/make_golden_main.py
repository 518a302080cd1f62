"""oracle/make_golden_main.py -- TEST INFRASTRUCTURE.  Golden vectors from the REFERENCE's whole program on its template inputs.

For every case below: copy /root/reference/template/input (the unchanged Exodus mesh, inparam.*, CMTSOLUTION, STATIONS) into
a scratch run directory, apply the case's inparam overrides (always: DEVELOP_MAX_TIME_STEPS and the NetCDF station format, so
that the traces come back in full float32 instead of the 6 digits of the ascii writer), flatten the mesh for the NetCDF
stand-in (oracle/nc_flatten.py), and run

  * oracle/_ref/axisem3d_ref  -- the reference's own main() (axisem.cpp) -> station seismograms,
  * oracle/_ref/axisem3d_dump -- the same preloop behind oracle/ref_main_dump.cpp -> what Mesh / Source / STF / Receiver
    release put into the reference's Domain (AX3D serialisation, tests/dump_domain.py: parse_dump), and, with `solve`, the
    same time loop again (its traces must equal the first program's bit for bit: checked here).

Both are built by oracle/Makefile.main from the reference's sources, unmodified, over the stand-ins of oracle/shim.
Written: tests/golden/main_<case>.npz (traces, decimated in time, + the case's overrides) and
tests/golden/main_<case>_domain.bin.xz (the dump).  /root/reference does not exist on the GPU box: this runs in the build
container only:   make -C oracle -f Makefile.main && python oracle/make_golden_main.py [case ...]
"""
import lzma
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]

from nc_flatten import flatten, read_flat  # noqa: E402

REF = os.environ.get("AX3D_REFERENCE", "/root/reference")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
BIN = os.path.join(ROOT, "oracle", "_ref")

# overrides on top of template/input; "steps" = DEVELOP_MAX_TIME_STEPS, "stride" = decimation of the stored traces
CASES = {
    # BASELINE.json configs[0]: the template run itself (1-D PREM, Nu = 2, CG4 attenuation, CMT source, 128 stations, RTZ)
    "cfg1_template": dict(steps=2000, stride=8, par={}),
    # the other branches of the same callers: empirical Nu(s, z) with lucky numbers and the odd-Nr rule near the axis, full
    # (non-CG4) attenuation, no Q_kappa, Ricker wavelet with a given half duration, dt factor, ENZ components, and a constant
    # ocean load (MassOcean1D on the surface points)
    "emp_full_enz": dict(steps=400, stride=2, par={
        "NU_TYPE": "empirical", "NU_EMP_REF": "6", "NU_EMP_MIN": "2", "ATTENUATION_CG4": "false", "ATTENUATION_QKAPPA": "false",
        "SOURCE_TIME_FUNCTION": "ricker", "SOURCE_STF_HALF_DURATION": "20.0", "TIME_DELTA_T_FACTOR": "0.8",
        "OUT_STATIONS_COMPONENTS": "ENZ", "MODEL_3D_ELLIPTICITY_MODE": "off", "MODEL_3D_OCEAN_LOAD": "constant$2.5"}),
    # 3-D volumetric models (Volumetric3D_bubble, no data files): a slow S anomaly around the source, a density anomaly across
    # the core-mantle boundary that also acts on the fluid (Mass3D, Acoustic3D, SFCoupling3D) and a P anomaly in the outer
    # core, Nu = 8: the reference's own Material 3-D path, Isotropic3D / TransverselyIsotropic3D + Attenuation3D_CG4 elements
    "bubbles_3d": dict(steps=600, stride=2, par={
        "NU_CONST": "8", "MODEL_3D_VOLUMETRIC_NUM": "3",
        "MODEL_3D_VOLUMETRIC_LIST": "bubble$VS$Ref1D$-0.08$300$100$37$-75 bubble$RHO$Ref1D$0.05$500$2891$45$-100$false$true "
                                    "bubble$VP$Ref3D$0.06$300$4000$40$-80$false$true"}),
    # a 3-D geometric model without data files: the full ellipticity of the Earth (MODEL_3D_ELLIPTICITY_MODE full) -> particle
    # relabelling of every element (PRT_3D + Jacobian-scaled 3-D moduli, Mass3D, tilted solid-fluid normals, a source and
    # receivers placed in the undeformed mesh): the reference's own Relabelling / Geometric3D / Ellipticity classes.  Every
    # element is 3-D, so the dump keeps the element arrays of every 6th element only ("thin").
    "ellipticity_prt": dict(steps=300, stride=2, thin=6, par={"MODEL_3D_ELLIPTICITY_MODE": "full"}),
    # the remaining branches of the source / receiver callers: a point force at the north pole (the pole singularity of
    # Source::Source), Gaussian source-time function, stations given in source-centred coordinates, SPZ components, no attenuation
    "pointforce_spz": dict(steps=400, stride=2, thin=100000, par={
        "SOURCE_TYPE": "point_force", "SOURCE_FILE": "POINTFORCE", "SOURCE_TIME_FUNCTION": "gauss", "SOURCE_STF_HALF_DURATION": "15.0",
        "OUT_STATIONS_FILE": "STATIONS_SC", "OUT_STATIONS_SYSTEM": "source-centered", "OUT_STATIONS_COMPONENTS": "SPZ",
        "ATTENUATION": "false"},
        files={"POINTFORCE": "latitude:   90.0\nlongitude:  10.0\ndepth:      20.0\nFt:         1.0e18\nFp:        -0.5e18\nFr:         2.0e18\n",
               "STATIONS_SC": "".join("S%02d  SC  %7.3f  %7.3f  0.0  %5.1f\n" % (i, 2.0 + 3.1 * i, (37.0 * i) % 360.0, 0.0 if i % 3 else 25.0 * i)
                                      for i in range(24))}),
    # full ellipticity seen from a source at the pole: the undulation is axisymmetric, so the relabelling is 1-D (PRT_1D + 1-D
    # moduli scaled by the Jacobian, Mass1D, SFCoupling1D) -- the other branch of Relabelling::createPRT / Quad::release
    "ellipticity_pole": dict(steps=200, stride=2, thin=6, par={
        "MODEL_3D_ELLIPTICITY_MODE": "full", "SOURCE_TYPE": "point_force", "SOURCE_FILE": "POINTFORCE", "ATTENUATION": "false"},
        files={"POINTFORCE": "latitude:   90.0\nlongitude:  0.0\ndepth:      30.0\nFt:         1.0e18\nFp:         0.0\nFr:         1.0e18\n"}),
    # the second data-free volumetric model: a slanted fast cylinder through the upper mantle next to the source, given in
    # source-centred coordinates, absolute density in a second, vertical one (Absolute reference type: no decay), Nu = 6
    "cylinder_3d": dict(steps=120, stride=2, thin=8, par={
        "NU_CONST": "6", "MODEL_3D_VOLUMETRIC_NUM": "2",
        "MODEL_3D_VOLUMETRIC_LIST": "cylinder$VP$Ref1D$0.05$150$50$3$20$600$8$60$true cylinder$RHO$Abs$3.1$100$0$35$-70$24$35$-70"}),
    # stations inside the Earth, down into the fluid outer core (Receiver::computeInterpFact divides by the integral factor there,
    # FluidElement::computeGroundMotion differentiates the potential), geographic coordinates with depths, ellipticity mode
    # geographic (depth-dependent flattening in Geodesy::lat2Theta_d)
    "deep_stations": dict(steps=400, stride=2, thin=100000, par={"OUT_STATIONS_FILE": "STATIONS_DEEP"},
        files={"STATIONS_DEEP": "".join("D%02d  DP  %8.3f  %8.3f  0.0  %9.1f\n" % (i, 37.9 - 1.7 * (i % 9), -77.9 + 4.3 * (i % 7), d)
                                        for i, d in enumerate([0.0, 15e3, 35e3, 3.0e5, 6.6e5, 1.5e6, 2.8e6, 2.95e6, 3.5e6, 4.2e6, 5.0e6, 5.2e6, 5.9e6, 1.0e5, 2.891e6]))}),
    # an ocean on the ellipsoid: full ellipticity tilts the surface normal with azimuth, so the surface points get MassOcean3D
    "ocean_on_ellipsoid": dict(steps=200, stride=2, thin=12, par={"MODEL_3D_ELLIPTICITY_MODE": "full", "MODEL_3D_OCEAN_LOAD": "constant$4.0"}),
    # wisdom learning (Domain::learnWisdom / dumpWisdom, Point::learnWisdom): empirical Nu, learn every 5th step with cutoff
    # 1e-3; the wisdom file the reference writes (s, z, learnt Nu, original Nu per point) is kept next to the traces
    "wisdom_learn": dict(steps=300, stride=2, thin=100000, par={
        "NU_TYPE": "empirical", "NU_EMP_REF": "6", "NU_EMP_MIN": "2", "NU_WISDOM_LEARN": "true", "NU_WISDOM_LEARN_EPSILON": "1e-3",
        "NU_WISDOM_LEARN_INTERVAL": "5", "NU_WISDOM_LEARN_OUTPUT": "learn.nu_wisdom.nc"}),
}


def write_overrides(input_dir, par):
    for name in ("inparam.model", "inparam.nu", "inparam.time_src_recv", "inparam.advanced"):
        path = os.path.join(input_dir, name)
        lines = open(path).read().split("\n")
        for i, line in enumerate(lines):
            w = line.split()
            if w and w[0] in par:
                lines[i] = "%-43s %s" % (w[0], par[w[0]])
        open(path, "w").write("\n".join(lines))


def prepare(case, run_dir):
    cfg = CASES[case]
    inp = os.path.join(run_dir, "input")
    shutil.copytree(os.path.join(REF, "template", "input"), inp)
    for f in os.listdir(inp):
        os.chmod(os.path.join(inp, f), 0o644)
    par = dict(cfg["par"])
    par["DEVELOP_MAX_TIME_STEPS"] = str(cfg["steps"])
    par["OUT_STATIONS_FORMAT"] = "netcdf"
    par["OPTION_VERBOSE_LEVEL"] = "essential"
    write_overrides(inp, par)
    for name, text in cfg.get("files", {}).items():
        open(os.path.join(inp, name), "w").write(text)
    mesh = [f for f in os.listdir(inp) if f.endswith(".e")][0]
    flatten(os.path.join(inp, mesh))
    return par


def run(exe, run_dir, *args):
    link = os.path.join(run_dir, exe)
    if not os.path.exists(link):
        os.symlink(os.path.join(BIN, exe), link)
    shutil.rmtree(os.path.join(run_dir, "output"), ignore_errors=True)
    res = subprocess.run([link] + list(args), cwd=run_dir, capture_output=True, text=True)
    if res.returncode != 0 or "ABORTED" in res.stdout:
        raise RuntimeError("%s failed:\n%s\n%s" % (exe, res.stdout[-3000:], res.stderr[-3000:]))
    return res.stdout


def traces(run_dir):
    nc = read_flat(os.path.join(run_dir, "output", "stations", "axisem3d_synthetics.nc.ncflat"))
    keys = [k for k in nc if "@" not in k and k != "time_points"]
    return nc["time_points"].copy(), keys, np.stack([nc[k] for k in keys])          # [nrec][nstep][3]


def make(case, keep=None):
    cfg = CASES[case]
    run_dir = keep or tempfile.mkdtemp(prefix="ax3d_main_")
    if keep:
        shutil.rmtree(run_dir, ignore_errors=True)
        os.makedirs(run_dir)
    par = prepare(case, run_dir)
    log = run("axisem3d_ref", run_dir)
    t, keys, seis = traces(run_dir)
    dump_path = os.path.join(run_dir, "domain.bin")
    run("axisem3d_dump", run_dir, dump_path, "solve", *(["thin", str(cfg["thin"])] if cfg.get("thin") else []))
    t2, keys2, seis2 = traces(run_dir)
    assert keys == keys2 and np.array_equal(t, t2) and seis.tobytes() == seis2.tobytes(), \
        "ref_main_dump.cpp does not mirror axisem_main: the two programs' traces differ"
    stride = cfg["stride"]
    extra = {}
    wis = os.path.join(run_dir, "output", par.get("NU_WISDOM_LEARN_OUTPUT", "-") + ".ncflat")
    if os.path.exists(wis):                                   # NuWisdom::writeToFile: rows (s, z, nu_learn, nu_orign), domain point order
        w = read_flat(wis)["axisem3d_wisdom"]
        extra = dict(wisdom_sz=w[:, :2].astype(np.float32), wisdom_nu_learn=np.round(w[:, 2]).astype(np.int16),
                     wisdom_nu_orign=np.round(w[:, 3]).astype(np.int16))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "main_%s.npz" % case), time=t[::stride], keys=np.array(keys),
                        seis=seis[:, ::stride].astype(np.float32), stride=stride, steps=cfg["steps"],
                        par_keys=np.array(sorted(par)), par_vals=np.array([par[k] for k in sorted(par)]),
                        file_names=np.array(sorted(cfg.get("files", {}))), file_texts=np.array([cfg["files"][k] for k in sorted(cfg.get("files", {}))]),
                        **extra)
    raw = open(dump_path, "rb").read()
    with lzma.open(os.path.join(GOLDEN_DIR, "main_%s_domain.bin.xz" % case), "wb", preset=9) as f:
        f.write(raw)
    print("%-16s %d stations x %d steps (stored every %d), max |u| %.3e, dump %d bytes" %
          (case, len(keys), seis.shape[1], stride, float(np.abs(seis).max()), len(raw)))
    if not keep:
        shutil.rmtree(run_dir, ignore_errors=True)
    return log


if __name__ == "__main__":
    for c in (sys.argv[1:] or list(CASES)):
        make(c)
