# First GPU call of the next round (one B200): the experimental where_agg variant that
# stages only the qual's column - parity first, then the
# bench at 1 / 10 / 50 % selectivity with and without it, then an ncu capture.
# usage: gpurun --timeout 900 -- 'bash tools/gpu_round2_first.sh'
mkdir -p gpurun_out
# whole GPU suite first (no -x): several tests written at the end of round 1 have never run on a
# GPU and are xfail(strict=False) until they have (test_zz_*, the experimental variant)
timeout 1500 python -m pytest tests -m gpu -q -rxXf --timeout 300 > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
PGSTROM_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_workloads.py -x -q \
    --timeout 120 -k gather > gpurun_out/t_gather.log 2>&1; echo "rc=$?" >> gpurun_out/t_gather.log
for sel in 1 10 50; do
  for g in 0 1; do
    PGSTROM_GATHER_PAYLOAD=$g timeout 120 python bench.py --workload where_agg --rows 50000000 \
        --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --selectivity $sel \
        > gpurun_out/bench_where_sel${sel}_gather${g}.json 2> gpurun_out/bench_where_sel${sel}_gather${g}.err
  done
done
PGSTROM_GATHER_PAYLOAD=1 bash tools/gpu_ncu.sh where_gather --workload where_agg
# bytes in flight vs table size on the default path: a smaller CTA-local table leaves room
# for a third stage (two in flight); 1000 groups need >= 1344 slots at the 75 % fill limit
for cfg in "1728 2 2048" "1408 3 2048" "1344 3 2048" "1408 4 1024" "1728 3 1024" "1408 2 3072"; do
  set -- $cfg
  PGSTROM_SH_SLOTS=$1 PGSTROM_NUM_STAGES=$2 PGSTROM_TILE_ROWS=$3 timeout 120 python bench.py \
      --workload where_agg --rows 50000000 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 \
      > gpurun_out/bench_where_slots$1_st$2_tile$3.json 2> gpurun_out/bench_where_slots$1_st$2_tile$3.err
done
# the many-groups path has no ncu capture yet (roofline.traffic is null for it): deal kernel first,
# then gpupreagg_partagg (-k picks by name; make_profiles.py takes the .ncu-rep files)
bash tools/gpu_ncu.sh hc --workload high_cardinality --rows 50000000
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gpupreagg_partagg -s 3 -c 1 -f \
    -o gpurun_out/prof_hc_partagg python bench.py --workload high_cardinality --rows 50000000 --steps 2 \
    --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-check > gpurun_out/ncu_hc_partagg.log 2>&1
# not yet measured in round 1: Zipf(1.0) keys, where_agg at 50 % (covered by the loop above)
timeout 200 python bench.py --workload high_cardinality --zipf --rows 50000000 --steps 3 --warmup 3 \
    --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_hc_zipf.json 2> gpurun_out/bench_hc_zipf.err
