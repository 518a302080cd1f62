# 8-GPU box: weak scaling cfg4 at N = 8 (= the strong-scaling mesh), cfg5 at N = 8
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
timeout 1200 $TR bench.py --gpus 8 --steps 50 --warmup 5 --min-seconds 1 > gpurun_out/r2i_bench_n8.json 2> gpurun_out/r2i_bench_n8.err
echo "n8 rc=$?"
timeout 1500 $TR bench.py --gpus 8 --config cfg5 --steps 20 --warmup 3 --min-seconds 0.5 > gpurun_out/r2i_bench_cfg5_n8.json 2> gpurun_out/r2i_bench_cfg5_n8.err
echo "cfg5 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2i_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']; print(f, 'value %.4g'%d['value'], round(d['ms_per_step'],4), {k:round(v,4) for k,v in r['family_ms'].items()}, round(r['whole_step']['frac'],3), d.get('parity'), d['config'].get('partition'), d['config']['elements'], d['config']['stable'])
    except Exception as e: print(f,'ERR',e); print(open(f.replace('.json','.err')).read()[-3000:])
PY
