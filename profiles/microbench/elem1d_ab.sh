# A/B: register cap of k_elem1d<solid> (1 vs 2 resident 400-thread CTAs per SM) on cfg1 (all elements 1D) and a 1D aniso + Full case
mkdir -p gpurun_out
for v in e1d3 e1d4; do
  AX3D_LIB=profiles/microbench/variants/$v.so python bench.py --config cfg1 --no-cpu --steps 60 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v cfg1', round(d['ms_per_step'],4), d['roofline']['family_ms'])"
done
