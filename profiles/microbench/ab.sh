#!/bin/bash
# ab.sh NAME... : bench.py (cfg2, no CPU leg) once per variant library, prints ms/step and the fused kernel's ms/launch
cd "$(dirname "$0")/../.."
for v in "$@"; do
  AX3D_LIB=profiles/microbench/variants/$v.so python bench.py --no-cpu --steps 60 --warmup 5 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/ab_$v.json"))
    print("%-12s step %.4f ms  e2e %.4f  fused %.4f  families %s" % ("$v", d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], {k: round(x, 4) for k, x in d["roofline"]["family_ms"].items()}))
except Exception as e:
    print("$v failed", e); print(open("gpurun_out/ab_$v.err").read()[-600:])
P
done
