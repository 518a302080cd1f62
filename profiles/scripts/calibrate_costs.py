#!/usr/bin/env python
"""Calibrates the element cost model of the partitioner from costs measured on the device (ax3d_measure_costs =
the reference's cost-measure pass, Mesh.cpp:412-588): for every element kind (solid / fluid, 1D / 3D material, elastic
law, attenuation) fit  cost_us = a Nr log2 Nr + b Nr + c  over the elements of the cfg1-cfg4 bench domains and write
gpurun_out/cost_model.json (committed as axisem3d_b200/cost_model.json)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from axisem3d_b200.domain import Domain  # noqa: E402

rows = {}
for cfg in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5-shell"):
    if cfg == "cfg5-shell":      # cfg5's element kind over its whole Nr range: a thin outer shell with the cfg5 order field
        from axisem3d_b200.mesh_synth import SynthMesh, R_EARTH
        bench.CFG = "cfg5"
        m = SynthMesh(n_theta=240, n_r=8, r_in=R_EARTH - 400e3, nu_fn=bench.cfg5_nu, law="aniso", model3d=True, attenuation="cg4",
                      fluid_layers=(), dtype_coef=np.float32)
    else:
        bench.CFG = cfg
        m = bench.make_mesh(bench.N_THETA)
    dt = m.estimate_dt()
    g = Domain(0)
    rel = m.release(g, dt)
    g.finalize()
    cost = g.measure_costs(5)
    for e, c in zip(rel["elements"], cost):
        nr = max(p.nr for p in e.points)
        if e.kind == "solid":
            el = e.elastic
            key = "solid|%s|%s|%s" % ("3d" if el.coef.shape[1] > 1 else "1d", el.law, "none" if el.att is None else ("cg4" if el.att.cg4 else "full"))
        else:
            key = "fluid|%s" % ("3d" if e.acoustic.K.shape[0] > 1 else "1d")
        rows.setdefault(key, []).append((nr, c))
    print(cfg, "total SM-us", cost.sum(), "= %.3f ms on 148 SMs" % (cost.sum() / 148e3), flush=True)
    del g

model = {}
for key, lst in rows.items():
    a = np.array(lst, dtype=np.float64)
    nr, c = a[:, 0], a[:, 1]
    A = np.stack([nr * np.log2(np.maximum(nr, 2)), nr, np.ones_like(nr)], 1)
    coef, *_ = np.linalg.lstsq(A, c, rcond=None)
    fit = A @ coef
    un = np.unique(nr)
    model[key] = {"a_nr_log2nr": coef[0], "b_nr": coef[1], "c": coef[2], "n": int(len(nr)), "nr_min": int(nr.min()), "nr_max": int(nr.max()),
                  "rel_rms": float(np.sqrt(np.mean((fit - c) ** 2)) / max(np.mean(c), 1e-12)),
                  # the measured table itself: mean cost of the elements of every Nr (the row-group count changes the cost in steps)
                  "table_nr": [int(x) for x in un], "table_us": [float(np.mean(c[nr == x])) for x in un]}
    print(key, {k: v for k, v in model[key].items() if not k.startswith('table')}, flush=True)
model["_point_us_per_mode"] = 192.0 / (6456.2e9 / 148) * 1e6          # Newmark bytes per solid point-mode at one SM's share of HBM
model["_about"] = "cost_us = a Nr log2 Nr + b Nr + c per element kind, least squares over device-measured costs (profiles/scripts/calibrate_costs.py)"
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(model, open(os.path.join(ROOT, "gpurun_out", "cost_model.json"), "w"), indent=1)
