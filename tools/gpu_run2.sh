timeout 300 python -m pytest tests/test_gpu_workloads.py -x -q --timeout 60 > gpurun_out/t_wl.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_wl.log
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_nogrp.json 2> gpurun_out/b_nogrp.err
timeout 120 python bench.py --workload where_agg --rows 100000000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_where.json 2> gpurun_out/b_where.err
timeout 200 python bench.py --workload high_cardinality --rows 50000000 --steps 3 --warmup 3 --no-cpu-baseline --no-check > gpurun_out/b_hc.json 2> gpurun_out/b_hc.err
