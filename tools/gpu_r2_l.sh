# Round 2, GPU call L (one B200): CTA-local table size / ring depth of the gather scan (where_agg)
mkdir -p gpurun_out
X="--steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-check --workload where_agg"
run() { tag=$1; shift; timeout 200 python bench.py $X "$@" > gpurun_out/l_$tag.json 2> gpurun_out/l_$tag.err; echo "rc=$?" >> gpurun_out/l_$tag.err; }
run base
PGSTROM_SH_SLOTS=1408 run s1408
PGSTROM_SH_SLOTS=1536 run s1536
PGSTROM_SH_SLOTS=1536 PGSTROM_NUM_STAGES=4 run s1536_st4
PGSTROM_SH_SLOTS=2016 run s2016
PGSTROM_NUM_STAGES=2 run st2
ls -la gpurun_out > gpurun_out/l_ls.txt
