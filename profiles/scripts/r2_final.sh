# round 2, final GPU call: the whole -m gpu suite (incl. the reference-main seismogram tests), the template run directory end
# to end through `python -m axisem3d_b200.run`, and one default bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/r2x_gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q --durations=12 -p no:cacheprovider > gpurun_out/r2x_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log
tail -25 gpurun_out/r2x_pytest.log
RUN=/tmp/ax3d_template_run
rm -rf $RUN && mkdir -p $RUN && python -c "import json,os,sys; d=json.load(open('tests/golden/template_input.json'))['files']; os.makedirs(sys.argv[1]); [open(os.path.join(sys.argv[1],k),'w').write(v) for k,v in d.items()]" $RUN/input && cp tests/golden/AxiSEM_prem_ani_one_crust_50.e $RUN/input/
( time timeout 300 python -m axisem3d_b200.run $RUN ) > gpurun_out/r2x_template_run.log 2>&1
echo "run rc=$?" >> gpurun_out/r2x_template_run.log
ls $RUN/output/stations | wc -l >> gpurun_out/r2x_template_run.log
head -3 $RUN/output/stations/II.AAK.RTZ.ascii >> gpurun_out/r2x_template_run.log
tail -2 $RUN/output/stations/IU.SSPA.RTZ.ascii >> gpurun_out/r2x_template_run.log
cat gpurun_out/r2x_template_run.log
timeout 200 python bench.py > gpurun_out/r2x_bench_default.json 2> gpurun_out/r2x_bench_default.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r2x_bench_default.json
