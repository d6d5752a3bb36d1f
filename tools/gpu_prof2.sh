# source-level profile of the grouped kernels
PGSTROM_DEBUG_LEVEL=4 timeout 200 python tools/dbg_counters.py where_agg 50000000 > gpurun_out/dbg_where.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gpupreagg_main -s 3 -c 1 -f -o gpurun_out/prof_where_b python bench.py --workload where_agg --rows 50000000 --chunk-rows 50000000 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-check > gpurun_out/ncu_where_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gpupreagg_main -s 3 -c 1 -f -o gpurun_out/prof_hc_b python bench.py --workload high_cardinality --rows 50000000 --chunk-rows 50000000 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-check > gpurun_out/ncu_hc_b.log 2>&1
ls -la gpurun_out > gpurun_out/ls.log
