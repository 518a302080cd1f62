# the unchanged template run directory (3600 s record = 8274 steps, 128 stations) end to end on the GPU
mkdir -p gpurun_out
RUN=/tmp/ax3d_template_run
rm -rf $RUN && mkdir -p $RUN && python -c "import json,os,sys; d=json.load(open('tests/golden/template_input.json'))['files']; os.makedirs(sys.argv[1]); [open(os.path.join(sys.argv[1],k),'w').write(v) for k,v in d.items()]" $RUN/input && cp tests/golden/AxiSEM_prem_ani_one_crust_50.e $RUN/input/
( time timeout 200 python -m axisem3d_b200.run $RUN ) > gpurun_out/r2y_template_run.log 2>&1
echo "run rc=$?" >> gpurun_out/r2y_template_run.log
ls $RUN/output/stations | wc -l >> gpurun_out/r2y_template_run.log
head -2 $RUN/output/stations/II.AAK.RTZ.ascii >> gpurun_out/r2y_template_run.log
sed -n '2000,2002p' $RUN/output/stations/IU.SSPA.RTZ.ascii >> gpurun_out/r2y_template_run.log
tail -1 $RUN/output/stations/IU.SSPA.RTZ.ascii >> gpurun_out/r2y_template_run.log
cp $RUN/output/stations/IU.SSPA.RTZ.ascii gpurun_out/r2y_IU.SSPA.RTZ.ascii
cat gpurun_out/r2y_template_run.log
