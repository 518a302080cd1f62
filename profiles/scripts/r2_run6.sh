mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" > gpurun_out/r2m_$name.json 2> gpurun_out/r2m_$name.err
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/r2m_$name.json").read().strip().splitlines()[-1]); r = d["roofline"]
    print("%-22s step %.4f ms  %s %.4f  families %s" % ("$name", d["ms_per_step"], r["kernel"][:24], r["kernel_ms_per_launch"], {k: round(x, 4) for k, x in r["family_ms"].items()}))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r2m_$name.err").read()[-600:])
P
}
for c in cfg4 cfg3 cfg2; do
  run auto_$c X=1 python bench.py --config $c --no-cpu --steps 50 --min-seconds 0.5
  run pass_$c AX3D_PASS_ITEMS=1 python bench.py --config $c --no-cpu --steps 50 --min-seconds 0.5
  run elem_$c AX3D_PASS_ITEMS=0 python bench.py --config $c --no-cpu --steps 50 --min-seconds 0.5
done
timeout 900 python profiles/scripts/rank_balance.py 8 cfg4 metis 2>&1 | grep -v Warning | tail -2
