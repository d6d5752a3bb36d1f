# Round 2, GPU call A (one B200): whole GPU suite without -x (what fails, what is slow),
# the gather-payload variant of where_agg (parity, then bench), ring/table sweep of the
# default where_agg path, ncu of the two many-groups kernels.
# usage: gpurun --timeout 1700 -- 'bash tools/gpu_r2_a.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/a_gpu.txt 2>&1
nproc >> gpurun_out/a_gpu.txt; numactl -H >> gpurun_out/a_gpu.txt 2>&1
( time timeout 1100 python -m pytest tests -m gpu -q -rf --timeout 300 --durations=40 ) > gpurun_out/a_t_all.log 2>&1; echo "rc=$?" >> gpurun_out/a_t_all.log
PGSTROM_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_workloads.py -x -q \
    --timeout 200 -k gather > gpurun_out/a_t_gather.log 2>&1; echo "rc=$?" >> gpurun_out/a_t_gather.log
B="--workload where_agg --rows 50000000 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1"
for sel in 10 1 50; do
  for g in 1 0; do
    PGSTROM_GATHER_PAYLOAD=$g timeout 150 python bench.py $B --selectivity $sel \
        > gpurun_out/a_where_sel${sel}_gather${g}.json 2> gpurun_out/a_where_sel${sel}_gather${g}.err
  done
done
for cfg in "1408 3 2048" "1344 3 2048" "1408 4 1024" "1728 3 1024" "1344 5 1024"; do
  set -- $cfg
  PGSTROM_SH_SLOTS=$1 PGSTROM_NUM_STAGES=$2 PGSTROM_TILE_ROWS=$3 timeout 150 python bench.py $B --no-check \
      > gpurun_out/a_where_slots$1_st$2_tile$3.json 2> gpurun_out/a_where_slots$1_st$2_tile$3.err
done
PGSTROM_GATHER_PAYLOAD=1 bash tools/gpu_ncu.sh a_where_gather --workload where_agg --rows 50000000
bash tools/gpu_ncu.sh a_hc --workload high_cardinality --rows 50000000
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gpupreagg_partagg -s 3 -c 1 -f \
    -o gpurun_out/prof_a_hc_partagg python bench.py --workload high_cardinality --rows 50000000 --steps 2 \
    --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-check > gpurun_out/ncu_a_hc_partagg.log 2>&1
ls -la gpurun_out > gpurun_out/a_ls.txt
