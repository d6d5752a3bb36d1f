# Round 2, GPU call H (one B200): where does the gather scan spend its time?
mkdir -p gpurun_out
PGSTROM_DEBUG_LEVEL=4 timeout 200 python tools/dbg_counters.py where_agg 125000000 10 > gpurun_out/h_dbg_sel10.txt 2>&1
PGSTROM_DEBUG_LEVEL=4 timeout 200 python tools/dbg_counters.py where_agg 125000000 1 > gpurun_out/h_dbg_sel1.txt 2>&1
X="--steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-check"
timeout 200 python bench.py $X --workload where_agg --rows 2000000 --chunk-rows 2000000 > gpurun_out/h_where_2M.json 2> gpurun_out/h_where_2M.err
timeout 200 python bench.py $X --workload where_agg --rows 2000000 --chunk-rows 2000000 --selectivity 1 > gpurun_out/h_where_2M_sel1.json 2> gpurun_out/h_where_2M_sel1.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'^gpupreagg_main$' -s 3 -c 1 -f \
    -o gpurun_out/prof_h_where_sel1 python bench.py --workload where_agg --selectivity 1 --steps 2 --warmup 3 \
    --no-cpu-baseline --e2e-steps 1 --no-check > gpurun_out/ncu_h_where_sel1.log 2>&1
ls -la gpurun_out > gpurun_out/h_ls.txt
