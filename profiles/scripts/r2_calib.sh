mkdir -p gpurun_out
timeout 1200 python profiles/scripts/calibrate_costs.py > gpurun_out/r2f_calib.log 2>&1; tail -6 gpurun_out/r2f_calib.log
