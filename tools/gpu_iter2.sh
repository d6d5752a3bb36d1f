timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gpu.log
timeout 300 python bench.py --workload where_agg --no-cpu-baseline > gpurun_out/bench_where.json 2> gpurun_out/bench_where.err; echo "rc=$?" >> gpurun_out/bench_where.err
PGSTROM_DEBUG_LEVEL=4 timeout 200 python tools/dbg_counters.py where_agg 50000000 > gpurun_out/dbg_where.log 2>&1
tools/sweep.sh where_agg 125000000 "0 0 0 0 1536" "0 0 0 0 1632" > gpurun_out/sweep_where.log 2>&1
