mkdir -p gpurun_out
export AX3D_MISFIT_LOG=$PWD/gpurun_out/r2z_misfit_cuda.log; rm -f $AX3D_MISFIT_LOG
timeout 300 python -m pytest tests/test_gpu_run_dir.py tests/test_cpp_facade.py tests/test_main_seismograms.py -m gpu -q --durations=5 -p no:cacheprovider > gpurun_out/r2z_pytest_new.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2z_pytest_new.log
tail -30 gpurun_out/r2z_pytest_new.log
cat gpurun_out/r2z_misfit_cuda.log
