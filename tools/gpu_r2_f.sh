# Round 2, GPU call F (one B200): parity of the changed paths (scan fast path of the gather
# variant, 8-row deal pass, push-and-go finish of the peer merge), then the knob sweeps.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_zz_state_merge_gpu.py tests/test_gpu_workloads.py tests/test_gpu_regression.py -m gpu -q -rf --timeout 300 -x ) > gpurun_out/f_tests.log 2>&1; echo "rc=$?" >> gpurun_out/f_tests.log
X="--steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1"
run() { tag=$1; shift; timeout 300 python bench.py $X "$@" > gpurun_out/f_$tag.json 2> gpurun_out/f_$tag.err; echo "rc=$?" >> gpurun_out/f_$tag.err; }
run where --workload where_agg
PGSTROM_CONSUMER_WARPS=20 run where_w20 --workload where_agg --no-check
PGSTROM_CONSUMER_WARPS=24 run where_w24 --workload where_agg --no-check
PGSTROM_CONSUMER_WARPS=12 run where_w12 --workload where_agg --no-check
run where_sel1 --workload where_agg --selectivity 1 --no-check
run where_sel50 --workload where_agg --selectivity 50 --no-check
run hc --workload high_cardinality
PGSTROM_DEAL_STEPS=1 run hc_d1 --workload high_cardinality --no-check
PGSTROM_CONSUMER_WARPS=20 run hc_w20 --workload high_cardinality --no-check
PGSTROM_NUM_STAGES=3 PGSTROM_TILE_ROWS=2048 run hc_3x2048 --workload high_cardinality --no-check
run hc_zipf --workload high_cardinality --zipf
ls -la gpurun_out > gpurun_out/f_ls.txt
