# Round 2, GPU call B (one B200): the whole GPU suite (no -x), the default bench
# line (all four workloads), the reference arm, a launch list of the default bench.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_r2_b.sh'
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -rf --timeout 300 --durations=15 ) > gpurun_out/b_tests.log 2>&1; echo "rc=$?" >> gpurun_out/b_tests.log
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "rc=$?" >> gpurun_out/b_bench.err
( time timeout 400 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/b_ref.json 2> gpurun_out/b_ref.err; echo "rc=$?" >> gpurun_out/b_ref.err
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 \
    > gpurun_out/b_ncu_launches.log 2>&1; echo "rc=$?" >> gpurun_out/b_ncu_launches.log
ls -la gpurun_out > gpurun_out/b_ls.txt
