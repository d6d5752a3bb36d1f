# Round 2, GPU call G (one B200): whole GPU suite on the changed kernels (two-phase gather
# scan, partagg record double-buffer, unrolled heap walk), then their bench numbers.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -rf --timeout 300 ) > gpurun_out/g_tests.log 2>&1; echo "rc=$?" >> gpurun_out/g_tests.log
X="--steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1"
run() { tag=$1; shift; timeout 300 python bench.py $X "$@" > gpurun_out/g_$tag.json 2> gpurun_out/g_$tag.err; echo "rc=$?" >> gpurun_out/g_$tag.err; }
run where --workload where_agg
run where_sel1 --workload where_agg --selectivity 1 --no-check
run where_sel50 --workload where_agg --selectivity 50 --no-check
PGSTROM_CONSUMER_WARPS=20 PGSTROM_TILE_ROWS=5120 run where_w20 --workload where_agg --no-check
PGSTROM_CONSUMER_WARPS=24 PGSTROM_TILE_ROWS=6144 run where_w24 --workload where_agg --no-check
PGSTROM_NUM_STAGES=3 run where_st3 --workload where_agg --no-check
run hc --workload high_cardinality
run heap --workload nogrp_agg_heap
run nogrp --workload nogrp_agg
ls -la gpurun_out > gpurun_out/g_ls.txt
