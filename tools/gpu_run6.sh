timeout 400 python -m pytest tests/test_gpu_workloads.py -x -q --timeout 60 > gpurun_out/t_wl.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_wl.log
tools/sweep.sh where_agg 50000000 "0 0 0 0 -1 0" "0 0 0 0 -1 3" "0 0 0 0 2048 0" > gpurun_out/sweep_w7.log 2>&1
tools/sweep.sh nogrp_agg 100000000 "0 0 0 0 -1 0" >> gpurun_out/sweep_w7.log 2>&1
