mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" > gpurun_out/ab3_$name.json 2> gpurun_out/ab3_$name.err
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/ab3_$name.json").read().strip().splitlines()[-1]); r = d["roofline"]
    print("%-18s step %.4f ms  %s %.4f" % ("$name", d["ms_per_step"], r["kernel"][:24], r["kernel_ms_per_launch"]))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/ab3_$name.err").read()[-600:])
P
}
for c in cfg2 cfg4 cfg3; do
  run base_$c X=1 python bench.py --config $c --no-cpu --steps 50 --min-seconds 0.5
  run noinl_$c AX3D_LIB=profiles/microbench/variants/noinl.so python bench.py --config $c --no-cpu --steps 50 --min-seconds 0.5
done
