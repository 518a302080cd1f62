mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_run_dir.py tests/test_cpp_facade.py -m gpu -q --durations=5 -p no:cacheprovider > gpurun_out/r2z_pytest_new.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2z_pytest_new.log
tail -30 gpurun_out/r2z_pytest_new.log
