# baseline re-measurement: gpu tests + the three workloads (no ncu)
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gpu.log
timeout 600 python bench.py > gpurun_out/bench_nogrp.json 2> gpurun_out/bench_nogrp.err; echo "rc=$?" >> gpurun_out/bench_nogrp.err
timeout 300 python bench.py --workload where_agg --no-cpu-baseline > gpurun_out/bench_where.json 2> gpurun_out/bench_where.err; echo "rc=$?" >> gpurun_out/bench_where.err
timeout 300 python bench.py --workload high_cardinality --no-cpu-baseline > gpurun_out/bench_hc.json 2> gpurun_out/bench_hc.err; echo "rc=$?" >> gpurun_out/bench_hc.err
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.log
