import sys, os, numpy as np
sys.path.insert(0, '/root/repo')
from bench import make_mesh, stf_series
from axisem3d_b200.domain import Domain
mesh = make_mesh(72); dt = mesh.estimate_dt(); dom = Domain(0)
rel = mesh.release(dom, dt); src = mesh.make_source(rel["elements"], rel["dec"], amp=1e18); dom.addSourceTerm(src); dom.finalize()
stf = stf_series(64)
dom.runSteps(dt, stf[:20]); dom.checkStability()
for n in (10, 10):
    ms = dom.runStepsTimed(dt, stf[:n]); print("ms/step", ms / n); dom.checkStability()
dom.runSteps(dt, stf[:1]); dom.checkStability()
