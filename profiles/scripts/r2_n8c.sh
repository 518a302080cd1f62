mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "partial_row or split_pipeline or nu1000" 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
timeout 900 $TR bench.py --gpus 8 --config cfg5 --steps 20 --warmup 3 --min-seconds 0.3 > gpurun_out/r2w_bench_cfg5_n8.json 2> gpurun_out/r2w_bench_cfg5_n8.err
echo "cfg5 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2w_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']; print(f, 'value %.4g'%d['value'], round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k:round(v,4) for k,v in r['family_ms'].items()}, round(r['whole_step']['frac'],3), d['config']['elements'], d['config']['stable'], d['config']['partition'])
        for pr in d.get('per_rank',[]): print('    ', pr['rank'], pr['elements'], pr['point_modes'], pr['family_ms'], pr['neighbours'], pr['kernels_ms'])
    except Exception as e: print(f,'ERR',e); print(open(f.replace('.json','.err')).read()[-3000:])
PY
