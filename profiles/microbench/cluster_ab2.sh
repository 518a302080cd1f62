# A/B of the cluster kernel for elements too large for the single-CTA kernel (AX3D_CLUSTER=1) against the split pipeline (0)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k cluster 2>&1 | tail -2
for cfg in cfg4 cfg3; do for m in 0 1; do
  AX3D_CLUSTER=$m python bench.py --config $cfg --steps 20 --warmup 5 > gpurun_out/s19_${cfg}_cl$m.json 2> gpurun_out/s19_${cfg}_cl$m.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/s19_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],4), d['roofline']['family_ms'], round(d['roofline']['whole_step']['frac'],4))
    except Exception as e: print(f, 'ERR', e, open(f.replace('.json','.err')).read()[-600:])
PY
