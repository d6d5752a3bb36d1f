timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gpu.log
timeout 300 python bench.py --workload where_agg --no-cpu-baseline > gpurun_out/bench_where.json 2> gpurun_out/bench_where.err; echo "rc=$?" >> gpurun_out/bench_where.err
timeout 300 python bench.py --workload high_cardinality --no-cpu-baseline > gpurun_out/bench_hc.json 2> gpurun_out/bench_hc.err; echo "rc=$?" >> gpurun_out/bench_hc.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gpupreagg_main -s 3 -c 1 -f -o gpurun_out/prof_where_c python bench.py --workload where_agg --rows 50000000 --chunk-rows 50000000 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-check > gpurun_out/ncu_where_c.log 2>&1
