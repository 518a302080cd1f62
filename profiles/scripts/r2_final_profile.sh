# final evidence of round 2: launch list of the cfg4 step, ncu --set full of the solid fused kernel (cfg4) and of the point-update
# kernel, per-kernel sections for the split pipeline kernels on a case that uses them (Nr = 2016)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 64 --csv --log-file gpurun_out/r2_launches_cfg4.csv python bench.py --config cfg4 --no-cpu --steps 3 --warmup 3 --min-seconds 0 > gpurun_out/r2n_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_elem3d_fused|k_newmark_solid|k_elem1d' -s 12 -c 4 -o gpurun_out/r2_step_cfg4 python bench.py --config cfg4 --no-cpu --steps 3 --warmup 3 --min-seconds 0 > gpurun_out/r2n_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fft3d_v2|k_grad3d|k_quad3d' -c 3 -o gpurun_out/r2_split_nu1000 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "iso3d_nu1000_split_np1" > gpurun_out/r2n_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
# strong scaling reference: the N = 8 mesh (576 x 28 = 16128 quads) on ONE GPU
timeout 900 python bench.py --gpus 1 --scaling strong --steps 20 --warmup 3 --min-seconds 0.5 --no-cpu > gpurun_out/r2_bench_strong_n1.json 2> gpurun_out/r2_bench_strong_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_strong_n1.json').read().strip().splitlines()[-1]); print('strong N=1', d['value'], d['ms_per_step'], d['config']['elements'])
PY
