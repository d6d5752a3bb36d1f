timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gpu.log
timeout 300 python bench.py --workload where_agg --no-cpu-baseline > gpurun_out/bench_where.json 2> gpurun_out/bench_where.err; echo "rc=$?" >> gpurun_out/bench_where.err
tools/sweep.sh where_agg 125000000 "0 0 8 0 -1" "0 0 8 0 2048" "4096 2 8 0 -1" > gpurun_out/sweep_where.log 2>&1
timeout 300 python bench.py --workload high_cardinality --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/plain_hc.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_hc.csv python bench.py --workload high_cardinality --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-check > gpurun_out/ncu_launches_hc.log 2>&1
timeout 300 python bench.py --workload nogrp_agg --format row --rows 40000000 --chunk-rows 10000000 --e2e-chunk-rows 5000000 --no-cpu-baseline > gpurun_out/bench_nogrp_row.json 2> gpurun_out/bench_nogrp_row.err; echo "rc=$?" >> gpurun_out/bench_nogrp_row.err
