# A/B: more threads / resident CTAs at lower register caps (with small spills) for the fused element kernel (cfg2) and k_fft3d_v2 (cfg3)
mkdir -p gpurun_out
run() { AX3D_LIB=$1 python bench.py --config $2 --no-cpu --steps 40 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$3 $2', round(d['ms_per_step'],4), round(d['roofline']['kernel_ms_per_launch'],4), d['roofline']['family_ms'])"; }
run axisem3d_b200/libaxisem3d_b200.so cfg2 base
for v in fus704 fus960; do run profiles/microbench/variants/$v.so cfg2 $v; done
run axisem3d_b200/libaxisem3d_b200.so cfg3 base
for v in fft3 fft4; do run profiles/microbench/variants/$v.so cfg3 $v; done
for v in fus704 fus960 fft3 fft4; do AX3D_LIB=profiles/microbench/variants/$v.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cfg2_iso3d or ti3d_nu200_split or in_kernel" 2>&1 | tail -1 | sed "s/^/$v parity: /"; done
