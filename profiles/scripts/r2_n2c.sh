mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 50 --warmup 5 --min-seconds 1 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err
AX3D_NO_INKERNEL_PUT=1 timeout 900 $TR bench.py --gpus 2 --steps 50 --warmup 5 --min-seconds 1 --no-parity > gpurun_out/r2j_bench_n2_noput.json 2> gpurun_out/r2j_bench_n2_noput.err
timeout 900 $TR bench.py --gpus 2 --steps 50 --warmup 5 --min-seconds 1 --no-parity --partition contiguous > gpurun_out/r2j_bench_n2_contig.json 2> gpurun_out/r2j_bench_n2_contig.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2j_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']; print(f, 'value %.4g'%d['value'], round(d['ms_per_step'],4), {k:round(v,4) for k,v in r['family_ms'].items()}, (d.get('parity') or {}).get('rel_l2'))
        for pr in d.get('per_rank',[]): print('    ', pr)
    except Exception as e: print(f,'ERR',e); print(open(f.replace('.json','.err')).read()[-2500:])
PY
