# round-end evidence on one B200: default bench (with CPU baseline), launch list, ncu captures
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "rc=$?" >> gpurun_out/bench_r01.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_reference.json 2> gpurun_out/bench_r01_reference.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_launches.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
bash tools/gpu_ncu.sh nogrp_r01 --workload nogrp_agg
bash tools/gpu_ncu.sh where_r01 --workload where_agg
timeout 200 python bench.py --workload high_cardinality --rows 50000000 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-check > gpurun_out/plain_hc_r01.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gpupreagg_main|gpupreagg_partagg" -s 6 -c 2 -f -o gpurun_out/prof_hc_r01 python bench.py --workload high_cardinality --rows 50000000 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-check > gpurun_out/ncu_hc_r01.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_hc_r01.log
