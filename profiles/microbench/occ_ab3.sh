# after adopting AX_FFT_MIN_CTAS = 4 and chunks cut by residency class: cfg2 / cfg3 / cfg4 + the full gpu test suite
mkdir -p gpurun_out
for c in cfg4 cfg3 cfg2; do python bench.py --config $c --no-cpu --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$c', round(d['ms_per_step'],4), d['roofline']['family_ms'], d['gpu_launches'])"; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
