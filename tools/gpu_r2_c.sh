# Round 2, GPU call C (one B200): GPU suite, default bench line (four workloads),
# many-groups experiments (cursor spreading, warps), ncu of the where / deal kernels.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -rf --timeout 300 --durations=8 ) > gpurun_out/c_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c_tests.log
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "rc=$?" >> gpurun_out/c_bench.err
H="--workload high_cardinality --rows 50000000 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
PGSTROM_CURSOR_SHIFT=0 timeout 200 python bench.py $H > gpurun_out/c_hc_shift0.json 2> gpurun_out/c_hc_shift0.err
PGSTROM_CURSOR_SHIFT=3 timeout 200 python bench.py $H > gpurun_out/c_hc_shift3.json 2> gpurun_out/c_hc_shift3.err
PGSTROM_CURSOR_SHIFT=5 timeout 200 python bench.py $H > gpurun_out/c_hc_shift5.json 2> gpurun_out/c_hc_shift5.err
PGSTROM_CONSUMER_WARPS=24 timeout 300 python bench.py $H > gpurun_out/c_hc_warps24.json 2> gpurun_out/c_hc_warps24.err
PGSTROM_NUM_STAGES=2 PGSTROM_TILE_ROWS=2048 timeout 200 python bench.py $H > gpurun_out/c_hc_ring2x2048.json 2> gpurun_out/c_hc_ring2x2048.err
W="--workload where_agg --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 200 python bench.py $W --rows 50000000 > gpurun_out/c_where_50M.json 2> gpurun_out/c_where_50M.err
timeout 200 python bench.py $W --rows 50000000 --selectivity 1 > gpurun_out/c_where_50M_sel1.json 2> gpurun_out/c_where_50M_sel1.err
timeout 200 python bench.py $W --rows 50000000 --selectivity 50 > gpurun_out/c_where_50M_sel50.json 2> gpurun_out/c_where_50M_sel50.err
PGSTROM_CONSUMER_WARPS=24 timeout 300 python bench.py $W --rows 50000000 > gpurun_out/c_where_50M_warps24.json 2> gpurun_out/c_where_50M_warps24.err
bash tools/gpu_ncu.sh c_where --workload where_agg
bash tools/gpu_ncu.sh c_hc --workload high_cardinality --rows 50000000
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/c_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 \
    > gpurun_out/c_ncu_launches.log 2>&1; echo "rc=$?" >> gpurun_out/c_ncu_launches.log
ls -la gpurun_out > gpurun_out/c_ls.txt
