mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -6 gpurun_out/r2h_pytest.log
timeout 900 python bench.py > gpurun_out/r2h_bench_default.json 2> gpurun_out/r2h_bench_default.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2h_bench_reference.json 2> gpurun_out/r2h_bench_reference.err
for c in cfg1 cfg2 cfg4; do timeout 600 python bench.py --mesh exodus --config $c --no-cpu --min-seconds 0.5 > gpurun_out/r2h_bench_exodus_$c.json 2> gpurun_out/r2h_bench_exodus_$c.err; done
timeout 600 python bench.py --config cfg1 --no-cpu --min-seconds 0.5 > gpurun_out/r2h_bench_cfg1.json 2> gpurun_out/r2h_bench_cfg1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2h_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if 'roofline' in d:
            r=d['roofline']; print(f, round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k:round(v,4) for k,v in r['family_ms'].items()}, round(r['whole_step']['frac'],3), r['kernel'][:30], round(r['frac'],3), d.get('cpu_baseline',{}).get('value'), d['data'][:30])
        else: print(f, d.get('value'), d.get('ms_per_step'), d.get('cpu_baseline'))
    except Exception as e: print(f,'ERR',e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
