# Round 2, GPU call D (one B200): parity of the staged heap kernel and the 256-bit record
# stores, their bench numbers, L2 fetch granularity and fixed-overhead experiments.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -rf --timeout 300 --durations=5 ) > gpurun_out/d_tests.log 2>&1; echo "rc=$?" >> gpurun_out/d_tests.log
PGSTROM_HEAP_STAGED=0 timeout 300 python -m pytest tests/test_gpu_regression.py -m gpu -q -k "heap or flat" > gpurun_out/d_tests_heap_unstaged.log 2>&1; echo "rc=$?" >> gpurun_out/d_tests_heap_unstaged.log
X="--steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 300 python bench.py $X --workload nogrp_agg_heap > gpurun_out/d_heap.json 2> gpurun_out/d_heap.err
PGSTROM_HEAP_STAGED=0 timeout 300 python bench.py $X --workload nogrp_agg_heap > gpurun_out/d_heap_unstaged.json 2> gpurun_out/d_heap_unstaged.err
timeout 300 python bench.py $X --workload high_cardinality > gpurun_out/d_hc.json 2> gpurun_out/d_hc.err
timeout 300 python bench.py $X --workload where_agg > gpurun_out/d_where.json 2> gpurun_out/d_where.err
PGSTROM_L2_FETCH_GRANULARITY=32 timeout 300 python bench.py $X --workload where_agg > gpurun_out/d_where_l2g32.json 2> gpurun_out/d_where_l2g32.err
PGSTROM_GATHER_PAYLOAD=0 timeout 300 python bench.py $X --workload where_agg > gpurun_out/d_where_nogather.json 2> gpurun_out/d_where_nogather.err
timeout 300 python bench.py $X --workload where_agg --rows 2000000 --chunk-rows 2000000 > gpurun_out/d_where_2M.json 2> gpurun_out/d_where_2M.err
timeout 300 python bench.py $X --workload where_agg --rows 2000000 --chunk-rows 2000000 --selectivity 1 > gpurun_out/d_where_2M_sel1.json 2> gpurun_out/d_where_2M_sel1.err
bash tools/gpu_ncu.sh d_heap --workload nogrp_agg_heap
ls -la gpurun_out > gpurun_out/d_ls.txt
