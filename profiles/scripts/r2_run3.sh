mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -5 gpurun_out/r2d_pytest.log
for c in cfg2 cfg4 cfg3; do timeout 600 python bench.py --config $c --no-cpu --min-seconds 0.5 > gpurun_out/r2d_bench_$c.json 2> gpurun_out/r2d_bench_$c.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2d_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']; print(f, round(d['ms_per_step'],4), {k:round(v,4) for k,v in r['family_ms'].items()}, round(r['whole_step']['frac'],3), r['kernel'], round(r['kernel_ms_per_launch'],4), round(r['frac'],3))
    except Exception as e: print(f,'ERR',e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_elem3d_fused -s 4 -c 1 -o gpurun_out/r2d_fused_cfg2 python bench.py --config cfg2 --no-cpu --steps 3 --warmup 3 --min-seconds 0 > gpurun_out/r2d_ncu2.log 2>&1
