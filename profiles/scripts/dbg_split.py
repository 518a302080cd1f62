import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
from helpers import build_oracle, build_gpu, randomize_displ, push_fields, compare_field
from axisem3d_b200.mesh_synth import SynthMesh
import test_gpu_parity as T
for nofused in ("0", "1"):
    os.environ["AX3D_NO_FUSED"] = nofused
    for zero in (True, False):
        for name in ("ti3d_nu200_split", "cfg3_aniso3d_cg4"):
            m = SynthMesh(**T.CASES[name]); dt = m.estimate_dt()
            d, _ = build_oracle(m, dt, np.float64); g, _ = build_gpu(m, dt)
            randomize_displ(d, seed=13); push_fields(d, g, ("displ",))
            out = []
            for it in range(3):
                d.computeStiff(); d.coupleSolidFluid(); g.computeStiff(); g.coupleSolidFluid()
                out.append({k: float('%.2e' % v) for k, v in compare_field(d, g, "stiff").items()})
                if zero:
                    d.S["stiff"][:] = 0; d.F["stiff"][:] = 0
                    s, f = g.get_bulk("stiff", False), g.get_bulk("stiff", True)
                    if s.size: g.set_bulk("stiff", False, np.zeros_like(s))
                    if f.size: g.set_bulk("stiff", True, np.zeros_like(f))
            print("nofused", nofused, "zero", zero, name, out, flush=True)
