# round 2, call 2: row-group fused kernel -- parity, then cfg4 / cfg3 / cfg2 bench, launch list and ncu of the fused kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
for c in cfg4 cfg3 cfg2; do timeout 600 python bench.py --config $c --no-cpu > gpurun_out/r2b_bench_$c.json 2> gpurun_out/r2b_bench_$c.err; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 60 --csv --log-file gpurun_out/r2b_launches_cfg4.csv python bench.py --config cfg4 --no-cpu --steps 3 --warmup 3 --min-seconds 0 > gpurun_out/r2b_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_elem3d_fused' -s 8 -c 2 -o gpurun_out/r2b_fused_cfg4 python bench.py --config cfg4 --no-cpu --steps 3 --warmup 3 --min-seconds 0 > gpurun_out/r2b_ncu2.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2b_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],4), d['roofline']['family_ms'], d['roofline']['whole_step'], d['roofline']['kernel'], round(d['roofline']['frac'],3))
    except Exception as e: print(f,'ERR',e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
