# A/B on cfg4 (Nr 42..1008: both k_fft3d_v2 instances) and cfg3: resident CTAs of the 256-thread instance, threads of the big-tile one
mkdir -p gpurun_out
run() { AX3D_LIB=$1 python bench.py --config $2 --no-cpu --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$3 $2', round(d['ms_per_step'],4), d['roofline']['family_ms']['elements'])"; }
for c in cfg4 cfg3; do
  run axisem3d_b200/libaxisem3d_b200.so $c base
  for v in fft3 fft4 fft4b; do run profiles/microbench/variants/$v.so $c $v; done
done
AX3D_LIB=profiles/microbench/variants/fft4b.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "split or cluster or ragged or cfg4 or cfg3" 2>&1 | tail -1 | sed "s/^/fft4b parity: /"
