# A/B: round-1 tree vs the row-group kernel on cfg2 (same box), Newmark-warp chunk size, instruction counts
mkdir -p gpurun_out
R1=profiles/microbench/variants/r1tree
run() { # name, env..., cmd
  name=$1; shift
  env "$@" > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/ab_$name.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%-16s step %.4f ms  fused %.4f  families %s" % ("$name", d["ms_per_step"], r["kernel_ms_per_launch"], {k: round(x, 4) for k, x in r["family_ms"].items()}))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/ab_$name.err").read()[-600:])
P
}
run r1_cfg2 X=1 python $R1/bench.py --config cfg2 --no-cpu --steps 50
run r1_cfg2_nonw AX3D_NO_NW=1 python $R1/bench.py --config cfg2 --no-cpu --steps 50
run new_cfg2 X=1 python bench.py --config cfg2 --no-cpu --steps 50 --min-seconds 0.5
run new_cfg2_nonw AX3D_NO_NW=1 python bench.py --config cfg2 --no-cpu --steps 50 --min-seconds 0.5
run new_cfg2_chr160 AX3D_LIB=profiles/microbench/variants/chr160.so python bench.py --config cfg2 --no-cpu --steps 50 --min-seconds 0.5
run new_cfg4_chr160 AX3D_LIB=profiles/microbench/variants/chr160.so python bench.py --config cfg4 --no-cpu --steps 50 --min-seconds 0.5
run new_cfg4_nonw AX3D_NO_NW=1 python bench.py --config cfg4 --no-cpu --steps 50 --min-seconds 0.5
M=smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_pipe_lsu.sum
AX3D_NO_NW=1 ncu --metrics $M --clock-control none -k regex:k_elem3d_fused -s 6 -c 2 --csv --log-file gpurun_out/ab_r1_inst.csv python $R1/bench.py --config cfg2 --no-cpu --steps 3 --warmup 3 > /dev/null 2>&1
AX3D_NO_NW=1 ncu --metrics $M --clock-control none -k regex:k_elem3d_fused -s 6 -c 2 --csv --log-file gpurun_out/ab_new_inst.csv python bench.py --config cfg2 --no-cpu --steps 3 --warmup 3 --min-seconds 0 > /dev/null 2>&1
grep -h "k_elem3d_fused" gpurun_out/ab_r1_inst.csv gpurun_out/ab_new_inst.csv | awk -F'","' '{print substr($5,1,40), $(NF-2), $NF}'
