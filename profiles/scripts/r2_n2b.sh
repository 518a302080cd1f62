mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 50 --warmup 5 --min-seconds 1 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err
AX3D_NO_INKERNEL_PUT=1 timeout 900 $TR bench.py --gpus 2 --steps 50 --warmup 5 --min-seconds 1 --no-parity > gpurun_out/r2g_bench_n2_noput.json 2> gpurun_out/r2g_bench_n2_noput.err
timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 --min-seconds 1 --no-cpu > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err
timeout 900 $TR bench.py --gpus 2 --steps 50 --warmup 5 --min-seconds 1 --scaling strong > gpurun_out/r2g_bench_n2_strong.json 2> gpurun_out/r2g_bench_n2_strong.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2g_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']; print(f, 'value %.4g'%d['value'], round(d['ms_per_step'],4), {k:round(v,4) for k,v in r['family_ms'].items()}, round(r['whole_step']['frac'],3), d.get('parity'), (d['config'].get('partition') or {}).get('edgecut'))
    except Exception as e: print(f,'ERR',e); print(open(f.replace('.json','.err')).read()[-2500:])
PY
