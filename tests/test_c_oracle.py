"""The C restatement (oracle/oracle.c, the timed CPU baseline) against the numpy oracle."""
import copy

import numpy as np
import pytest

from helpers import build_oracle, randomize_displ, rel_l2
from axisem3d_b200.mesh_synth import SynthMesh
from c_oracle import COracle

CASES = [
    dict(n_theta=6, n_r=6, nu=2, law="ti", model3d=False, attenuation="cg4"),
    dict(n_theta=5, n_r=6, nu=5, law="aniso", model3d=False, attenuation="full"),
    dict(n_theta=6, n_r=8, nu=12, law="iso", model3d=True, attenuation=None),
    dict(n_theta=5, n_r=8, nu=9, law="aniso", model3d=True, attenuation="cg4", fluid3d=True),
    dict(n_theta=5, n_r=8, nu=7, law="ti", model3d=True, attenuation="full"),
]


@pytest.mark.parametrize("kw", CASES)
def test_c_oracle_matches_numpy_oracle(kw):
    m = SynthMesh(**kw)
    dt = m.estimate_dt()
    a, _ = build_oracle(m, dt, np.float32)
    b, _ = build_oracle(m, dt, np.float32)
    cb = COracle(b)
    assert cb.threads() >= 1
    stf = np.exp(-((np.arange(12) - 4) / 2.0) ** 2)
    randomize_displ(a, seed=3)
    randomize_displ(b, seed=3)
    for i in range(len(stf)):
        a.step(dt, stf[i])
        cb.step(dt, stf[i])
    for k in ("S", "F"):
        x, y = getattr(a, k)["stiff"], getattr(b, k)["stiff"]
        if x.size:
            assert rel_l2(x, y) < 5e-5
        x, y = getattr(a, k)["displ"], getattr(b, k)["displ"]
        if x.size:
            assert rel_l2(x, y) < 5e-5
