# Round 2, GPU call L (one B200, the last seconds of the budget): CTA-local table size of the gather scan
mkdir -p gpurun_out
X="--steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-check --workload where_agg"
run() { tag=$1; shift; timeout 60 python bench.py $X "$@" > gpurun_out/l_$tag.json 2> gpurun_out/l_$tag.err; echo "rc=$?" >> gpurun_out/l_$tag.err; }
PGSTROM_SH_SLOTS=1408 run s1408
PGSTROM_SH_SLOTS=1536 run s1536
PGSTROM_SH_SLOTS=2016 run s2016
