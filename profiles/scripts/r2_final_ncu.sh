# final build of round 2: launch list of the cfg4 step and ncu --set full of the dominant kernels (refreshes profiles/traffic.json)
mkdir -p gpurun_out
timeout 110 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 64 --csv --log-file gpurun_out/r2f_launches_cfg4.csv python bench.py --config cfg4 --no-cpu --steps 3 --warmup 3 --min-seconds 0 > gpurun_out/r2f_ncu1.log 2>&1
echo "ncu1 rc=$?"
timeout 110 ncu --set full --clock-control none --import-source on -k regex:'k_elem3d_fused|k_newmark_solid|k_elem1d' -s 12 -c 4 -o gpurun_out/r2f_step_cfg4 python bench.py --config cfg4 --no-cpu --steps 3 --warmup 3 --min-seconds 0 > gpurun_out/r2f_ncu2.log 2>&1
echo "ncu2 rc=$?"
ls -la gpurun_out/r2f_*
