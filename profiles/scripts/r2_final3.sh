mkdir -p gpurun_out
export AX3D_MISFIT_LOG=$PWD/gpurun_out/r2w_misfit_cuda.log; rm -f $AX3D_MISFIT_LOG
timeout 175 python -m pytest tests/test_wisdom_reference.py "tests/test_main_seismograms.py::test_cuda_seismograms_match_reference_main[wisdom_learn]" tests/test_gpu_run_dir.py::test_run_dir_learns_and_reuses_wisdom -m gpu -q --durations=5 -p no:cacheprovider > gpurun_out/r2w_pytest_wisdom.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2w_pytest_wisdom.log
tail -40 gpurun_out/r2w_pytest_wisdom.log
cat gpurun_out/r2w_misfit_cuda.log
