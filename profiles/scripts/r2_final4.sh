mkdir -p gpurun_out
export AX3D_MISFIT_LOG=$PWD/gpurun_out/r2v_misfit_dropin.log; true
timeout 100 python -m pytest tests/test_gpu_dropin.py --durations=6 -m gpu -q -p no:cacheprovider > gpurun_out/r2v_pytest_dropin.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2v_pytest_dropin.log
tail -40 gpurun_out/r2v_pytest_dropin.log
cat gpurun_out/r2v_misfit_dropin.log
