# Round 2, GPU call Z (final) (one B200): the evidence run.  GPU suite, smoke, default bench line
# (four workloads), reference arm, launch list, ncu --set full of every hot kernel.
# usage: gpurun --timeout 2400 -- 'bash tools/gpu_r2_e.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/z_gpu.txt 2>&1
nproc >> gpurun_out/z_gpu.txt; numactl -H >> gpurun_out/z_gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -q -rf --timeout 300 --durations=10 ) > gpurun_out/z_tests.log 2>&1; echo "rc=$?" >> gpurun_out/z_tests.log
( time timeout 200 python __graft_entry__.py smoke ) > gpurun_out/z_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/z_smoke.log
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo "rc=$?" >> gpurun_out/z_bench.err
( time timeout 400 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/z_ref.json 2> gpurun_out/z_ref.err; echo "rc=$?" >> gpurun_out/z_ref.err
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
    --log-file gpurun_out/z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 \
    > gpurun_out/z_ncu_launches.log 2>&1; echo "rc=$?" >> gpurun_out/z_ncu_launches.log
X="--steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-check"
cap() {  # tag, kernel regex, skip, count, bench args...
  tag=$1; k=$2; s=$3; c=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -f \
      -o gpurun_out/prof_z_$tag python bench.py "$@" $X > gpurun_out/ncu_z_$tag.log 2>&1
  echo "rc=$?" >> gpurun_out/ncu_z_$tag.log
}
cap nogrp '^gpupreagg_main$' 3 1 --workload nogrp_agg
cap where '^gpupreagg_main$' 3 1 --workload where_agg
cap hc 'gpupreagg_main$|gpupreagg_partagg' 6 2 --workload high_cardinality
cap heap 'gpupreagg_heap_index|gpupreagg_main_heap' 6 2 --workload nogrp_agg_heap
ls -la gpurun_out > gpurun_out/z_ls.txt
