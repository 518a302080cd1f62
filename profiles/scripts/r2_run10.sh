mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log
tail -4 gpurun_out/r2u_pytest.log
timeout 1200 python profiles/scripts/calibrate_costs.py > gpurun_out/r2f_calib.log 2>&1; tail -3 gpurun_out/r2f_calib.log
