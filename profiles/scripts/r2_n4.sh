mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fluid_strain" 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521"
timeout 900 $TR bench.py --gpus 4 --steps 50 --warmup 5 --min-seconds 1 > gpurun_out/r2k_bench_n4.json 2> gpurun_out/r2k_bench_n4.err
AX3D_NO_INKERNEL_PUT=1 timeout 900 $TR bench.py --gpus 4 --steps 50 --warmup 5 --min-seconds 1 --no-parity > gpurun_out/r2k_bench_n4_noput.json 2> gpurun_out/r2k_bench_n4_noput.err
AX3D_HALO=nccl timeout 900 $TR bench.py --gpus 4 --steps 50 --warmup 5 --min-seconds 1 --no-parity > gpurun_out/r2k_bench_n4_nccl.json 2> gpurun_out/r2k_bench_n4_nccl.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2k_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']; print(f, 'value %.4g'%d['value'], round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), {k:round(v,4) for k,v in r['family_ms'].items()}, (d.get('parity') or {}).get('rel_l2'), d['config']['timing'])
        for pr in d.get('per_rank',[]): print('    ', pr['rank'], pr['elements'], pr['family_ms'], pr['neighbours'])
    except Exception as e: print(f,'ERR',e); print(open(f.replace('.json','.err')).read()[-2500:])
PY
