#!/usr/bin/env python
"""Load balance of an N-way partition measured on ONE GPU: every rank's part of the weak-scaling mesh is released as a
domain of its own (no halo exchange) and its step is timed -- the compute time each rank would need, against the cost
model's prediction.  usage: rank_balance.py N [config] [metis|contiguous]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from axisem3d_b200 import connectivity as CN, partition as PT  # noqa: E402
from axisem3d_b200.domain import Domain  # noqa: E402

N = int(sys.argv[1])
bench.CFG = sys.argv[2] if len(sys.argv) > 2 else "cfg4"
how = sys.argv[3] if len(sys.argv) > 3 else "metis"
n_theta, n_r = bench.mesh_shape(N, "weak")
mesh = bench.make_mesh(n_theta, n_r)
dt = mesh.estimate_dt()
w = bench.element_weights(mesh)
if how == "metis":
    e2p, info = PT.partition_kway(mesh.conn, w, N, imbalance=0.01, ntrials=4)
else:
    e2p, info = CN.partition_contiguous(w, N), {}
stf = bench.stf_series(200)
rows = []
for r in range(N):
    d = Domain(0)
    rel = mesh.release(d, dt, rank=r, elem_to_proc=e2p)
    d.finalize()
    bench.apply_kick(d, rel)
    d.runSteps(dt, stf[:10])
    ms = min(d.runStepsTimed(dt, stf[10:60]) for _ in range(3)) / 50
    d.enable_timers(True)
    d.get_timers(reset=True)
    d.kernel_stats(reset=True)
    d.runSteps(dt, stf[:8])
    fam = d.get_timers(reset=True) / 8
    ks = {k: round(v[0] / 8, 4) for k, v in d.kernel_stats().items()}
    d.enable_timers(False)
    rows.append(dict(rank=r, elements=int((e2p == r).sum()), model_us=float(w[e2p == r].sum()), ms_per_step=ms, family_ms=[round(float(x), 4) for x in fam],
                     kernels_ms=ks, point_modes=int(d.work_per_step())))
    print(rows[-1], flush=True)
    del d
t = np.array([x["ms_per_step"] for x in rows])
m = np.array([x["model_us"] for x in rows])
print("N", N, how, info.get("edgecut"), "measured max/mean %.3f" % (t.max() / t.mean()), "model max/mean %.3f" % (m.max() / m.mean()),
      "ms per model-us-per-SM:", np.round(t / (m / 148e3), 3))
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "rank_balance_%s_n%d_%s.json" % (bench.CFG, N, how)), "w"), indent=1)
