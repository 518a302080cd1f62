"""Shared set-up of the "reference main()" cases (tests/golden/main_<case>.npz + main_<case>_domain.bin.xz, written by
oracle/make_golden_main.py from the reference's whole program): the unchanged template inputs (tests/golden/template_input.json +
the shipped Exodus mesh) with the case's inparam overrides, run through the repo's preloop (axisem3d_b200/exodus_mesh.py,
preloop.py) into any domain that offers the Domain verbs (numpy oracle, DumpDomain, CUDA)."""
import atexit
import json
import lzma
import os
import shutil
import tempfile

import numpy as np

from axisem3d_b200 import preloop as PL
from dump_domain import parse_dump

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
MESH = "AxiSEM_prem_ani_one_crust_50.e"
CASES = ("cfg1_template", "emp_full_enz", "bubbles_3d", "ellipticity_prt", "pointforce_spz", "wisdom_learn")
# cases whose preloop arrays are compared (tests/test_preloop_reference.py) and that the CPU oracle steps, but that have no CUDA test
CPU_ONLY_CASES = ("ellipticity_pole", "cylinder_3d", "deep_stations", "ocean_on_ellipsoid")


def golden(case):
    z = np.load(os.path.join(GOLDEN, "main_%s.npz" % case))
    files = dict(zip([str(k) for k in z["file_names"]], [str(v) for v in z["file_texts"]])) if "file_names" in z.files else {}
    return dict(time=z["time"], keys=[str(k) for k in z["keys"]], seis=z["seis"], stride=int(z["stride"]), steps=int(z["steps"]),
                par=dict(zip([str(k) for k in z["par_keys"]], [str(v) for v in z["par_vals"]])), files=files)


def reference_domain(case):
    with lzma.open(os.path.join(GOLDEN, "main_%s_domain.bin.xz" % case), "rb") as f:
        return parse_dump(f.read())


def input_dir(case, tmp=None):
    """template_input + mesh + the case's overrides in a fresh directory"""
    tmp = tmp or tempfile.mkdtemp(prefix="ax3d_in_")
    inp = os.path.join(tmp, "input")
    os.makedirs(inp)
    for fname, text in json.load(open(os.path.join(GOLDEN, "template_input.json")))["files"].items():
        open(os.path.join(inp, fname), "w").write(text)
    shutil.copy(os.path.join(GOLDEN, MESH), os.path.join(inp, MESH))
    gold = golden(case)
    par = gold["par"]
    for name, text in gold["files"].items():           # input files the case adds (a point-force file, a station list)
        open(os.path.join(inp, name), "w").write(text)
    for name in PL.Parameters.FILES:
        path = os.path.join(inp, name)
        lines = open(path).read().split("\n")
        for i, line in enumerate(lines):
            w = line.split()
            if w and w[0] in par:
                lines[i] = "%s %s" % (w[0], par[w[0]])
        open(path, "w").write("\n".join(lines))
    return inp


class Case:
    """The case's input directory through the public entry point of the repo (axisem3d_b200.run.Simulation = everything
    axisem_main builds before the time loop)."""

    def __init__(self, case):
        from axisem3d_b200.run import Simulation
        self.name = case
        self.inp = input_dir(case)
        sim = self.sim = Simulation(self.inp)
        self.par, self.mesh, self.dt, self.geodesy = sim.par, sim.mesh, sim.dt, sim.geodesy
        self.source, self.stf, self.shift, self.receivers = sim.source, sim.stf, sim.shift, sim.receivers

    def release(self, domain):
        self.rel = self.sim.release(domain)
        return self.rel

    def cleanup(self):
        shutil.rmtree(os.path.dirname(self.inp), ignore_errors=True)


_CACHE = {}


def get_case(name):
    """One Simulation per case and test process (building the mesh is the expensive part; release() makes fresh objects)."""
    if name not in _CACHE:
        _CACHE[name] = Case(name)
    return _CACHE[name]


@atexit.register
def _cleanup():
    for c in _CACHE.values():
        c.cleanup()
