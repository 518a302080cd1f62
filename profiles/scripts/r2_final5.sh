mkdir -p gpurun_out
timeout 28 python -m pytest tests/test_gpu_run_dir.py::test_run_dir_single_rank -m gpu -q -p no:cacheprovider > gpurun_out/r2u_pytest_run_dir_formats.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2u_pytest_run_dir_formats.log
tail -8 gpurun_out/r2u_pytest_run_dir_formats.log
