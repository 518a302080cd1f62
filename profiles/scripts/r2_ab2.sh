# A/B: compute threads of the fused kernel (512 / 640 / 768 / 1024) and radix-8 plans; cfg2 source-level profile
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/ab_$name.json").read().strip().splitlines()[-1]); r = d["roofline"]
    print("%-18s step %.4f ms  %s %.4f  families %s" % ("$name", d["ms_per_step"], r["kernel"][:24], r["kernel_ms_per_launch"], {k: round(x, 4) for k, x in r["family_ms"].items()}))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/ab_$name.err").read()[-600:])
P
}
for c in cfg2 cfg4 cfg3; do
  run base_$c X=1 python bench.py --config $c --no-cpu --steps 50 --min-seconds 0.5
  for v in nt640 nt768 nt768r8 nt1024r8; do
    run ${v}_$c AX3D_LIB=profiles/microbench/variants/$v.so python bench.py --config $c --no-cpu --steps 50 --min-seconds 0.5
  done
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_elem3d_fused -s 4 -c 1 -o gpurun_out/r2c_fused_cfg2 python bench.py --config cfg2 --no-cpu --steps 3 --warmup 3 --min-seconds 0 > gpurun_out/r2c_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
