# Round 2, GPU call I (one B200): whole GPU suite on the segment-mode deal pass, the whole-tile
# push of the gather scan and the producer back-off; their bench numbers with A/B switches.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -rf --timeout 300 ) > gpurun_out/i_tests.log 2>&1; echo "rc=$?" >> gpurun_out/i_tests.log
X="--steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1"
run() { tag=$1; shift; timeout 300 python bench.py $X "$@" > gpurun_out/i_$tag.json 2> gpurun_out/i_$tag.err; echo "rc=$?" >> gpurun_out/i_$tag.err; }
run where --workload where_agg
run where_sel1 --workload where_agg --selectivity 1 --no-check
run where_sel50 --workload where_agg --selectivity 50 --no-check
run hc --workload high_cardinality
PGSTROM_NO_SEGMENTS=1 run hc_noseg --workload high_cardinality --no-check
run hc_zipf --workload high_cardinality --zipf
run nogrp --workload nogrp_agg
ls -la gpurun_out > gpurun_out/i_ls.txt
