python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_gpu.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?" >> gpurun_out/bench_default.err
python bench.py --workload where_agg --rows 100000000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_where.json 2> gpurun_out/bench_where.err
python bench.py --workload high_cardinality --rows 100000000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_hc.json 2> gpurun_out/bench_hc.err
nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current --format=csv > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt; lscpu | head -20 >> gpurun_out/gpu.txt
