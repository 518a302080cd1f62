mkdir -p gpurun_out
timeout 900 python profiles/scripts/rank_balance.py 8 cfg4 metis 2>&1 | grep -v Warning | tail -12
timeout 900 python profiles/scripts/rank_balance.py 8 cfg4 contiguous 2>&1 | grep -v Warning | tail -3
