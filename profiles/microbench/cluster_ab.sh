export AX3D_CLUSTER=2
python -m pytest tests/test_gpu_parity.py tests/test_golden_reference.py -m gpu -x -q > gpurun_out/s15_pytest_cl2.log 2>&1; tail -4 gpurun_out/s15_pytest_cl2.log
for nt in 128 192 256; do AX3D_CL_NT=$nt python bench.py --steps 30 --warmup 5 > gpurun_out/s15_cfg2_cl2_nt$nt.json 2> gpurun_out/s15_cfg2_cl2_nt$nt.err; done
unset AX3D_CLUSTER
for m in 0 1; do AX3D_CLUSTER=$m python bench.py --config cfg4 --steps 20 --warmup 5 > gpurun_out/s15_cfg4_cl$m.json 2> gpurun_out/s15_cfg4_cl$m.err; done
AX3D_CLUSTER=1 python bench.py --config cfg3 --steps 20 --warmup 5 > gpurun_out/s15_cfg3_cl1.json 2> gpurun_out/s15_cfg3_cl1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s15_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['roofline']['family_ms'], d['roofline']['whole_step']['frac'])
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
