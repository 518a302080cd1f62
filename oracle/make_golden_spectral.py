#!/usr/bin/env python
"""Golden fixture for axisem3d_b200/spectral.py: the nPol = 4 constants the reference hard-codes
(/root/reference/SOLVER/src/preloop/spectral/SpectralConstants.cpp:43-49: pgll_4, pglj_4, wgll_4, wglj_4, Ggll_4, Gglj_4),
parsed from the source where it lies and written to tests/golden/spectral_npol4.json.  TEST INFRASTRUCTURE ONLY."""
import json
import os
import re

SRC = "/root/reference/SOLVER/src/preloop/spectral/SpectralConstants.cpp"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "spectral_npol4.json")

txt = open(SRC).read()
out = {"source": "SOLVER/src/preloop/spectral/SpectralConstants.cpp:43-49 (12 decimals)"}
for name in ("pgll_4", "pglj_4", "wgll_4", "wglj_4", "Ggll_4", "Gglj_4"):
    m = re.search(r"double\s+%s\[\]\s*=\s*\{([^}]*)\}" % name, txt)
    out[name] = [float(x) for x in m.group(1).split(",")]
assert len(out["Ggll_4"]) == 25 and len(out["pgll_4"]) == 5
json.dump(out, open(OUT, "w"), indent=1)
print("wrote", os.path.normpath(OUT))
