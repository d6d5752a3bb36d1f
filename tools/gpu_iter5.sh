timeout 1500 python -m pytest tests -m gpu -x -q --timeout 400 > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gpu.log
timeout 300 python bench.py --workload where_agg --no-cpu-baseline > gpurun_out/bench_where.json 2> gpurun_out/bench_where.err; echo "rc=$?" >> gpurun_out/bench_where.err
timeout 300 python bench.py --workload high_cardinality --no-cpu-baseline > gpurun_out/bench_hc.json 2> gpurun_out/bench_hc.err; echo "rc=$?" >> gpurun_out/bench_hc.err
