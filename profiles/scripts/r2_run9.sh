mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log
tail -6 gpurun_out/r2t_pytest.log
for c in cfg4 cfg2; do timeout 600 python bench.py --config $c --no-cpu --min-seconds 0.5 > gpurun_out/r2t_bench_$c.json 2> gpurun_out/r2t_bench_$c.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2t_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']; print(f, round(d['ms_per_step'],4), {k:round(v,4) for k,v in r['family_ms'].items()}, round(r['whole_step']['frac'],3), r['kernel'][:30], round(r['kernel_ms_per_launch'],4))
    except Exception as e: print(f,'ERR',e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
