# final measurements of round 1: bench lines for cfg1-cfg4 + the reference arm, and the ncu launch list of the cfg2 step
mkdir -p gpurun_out
for c in cfg2 cfg1 cfg3 cfg4; do python bench.py --config $c > gpurun_out/r1f_bench_$c.json 2> gpurun_out/r1f_bench_$c.err; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1f_bench_reference.json 2> gpurun_out/r1f_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1f_launches_cfg2.csv python bench.py --no-cpu --steps 3 --warmup 3 > gpurun_out/r1f_ncu.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r1f_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],4), d['value'], d.get('roofline',{}).get('whole_step'))
    except Exception as e: print(f,'ERR',e)
PY
