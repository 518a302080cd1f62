# round 2, call 1: "before" evidence for the split pipeline (cfg4 / cfg3) on the round-1 kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -3 gpurun_out/r2a_pytest.log
for c in cfg4 cfg3; do python bench.py --config $c --no-cpu > gpurun_out/r2a_bench_$c.json 2> gpurun_out/r2a_bench_$c.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/r2a_launches_cfg4.csv python bench.py --config cfg4 --no-cpu --steps 3 --warmup 3 > gpurun_out/r2a_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_fft3d_v2|k_grad3d|k_quad3d' -s 24 -c 6 -o gpurun_out/r2a_split_cfg4 python bench.py --config cfg4 --no-cpu --steps 3 --warmup 3 > gpurun_out/r2a_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_fft3d_v2' -s 4 -c 2 -o gpurun_out/r2a_fft_cfg3 python bench.py --config cfg3 --no-cpu --steps 3 --warmup 3 > gpurun_out/r2a_ncu3.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2a_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],4), d['roofline']['family_ms'], d['roofline']['whole_step'])
    except Exception as e: print(f,'ERR',e)
PY
