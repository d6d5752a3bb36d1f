# full round-end style check on one B200: tests, bench, launch list, ncu captures
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gpu.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "rc=$?" >> gpurun_out/bench_r01.err
timeout 300 python bench.py --workload where_agg --no-cpu-baseline > gpurun_out/bench_r01_where.json 2> gpurun_out/bench_r01_where.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_launches.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
bash tools/gpu_ncu.sh nogrp_r01 --workload nogrp_agg
bash tools/gpu_ncu.sh where_r01 --workload where_agg
