timeout 300 python -m pytest tests/test_gpu_workloads.py -x -q --timeout 60 > gpurun_out/t_wl.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_wl.log
tools/sweep.sh where_agg 50000000 "0 0 0 0 -1 2" "0 0 0 0 -1 0" > gpurun_out/sweep_w3.log 2>&1
tools/sweep.sh high_cardinality 50000000 "0 0 0 0 -1 0" >> gpurun_out/sweep_w3.log 2>&1
