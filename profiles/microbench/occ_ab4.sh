# A/B: k_grad3d / k_quad3d compiled for 4 resident 400-thread CTAs per SM (32 registers) against 3 (43-47 registers)
run() { AX3D_LIB=$1 python bench.py --config $2 --no-cpu --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$3 $2', round(d['ms_per_step'],4), d['roofline']['family_ms']['elements'])"; }
for c in cfg4 cfg3; do run axisem3d_b200/libaxisem3d_b200.so $c base; run profiles/microbench/variants/gq4.so $c gq4; done
