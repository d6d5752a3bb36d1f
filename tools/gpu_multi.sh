# usage: gpu_multi.sh N [workloads...]
N=$1; shift
for wl in "$@"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/bench_${wl}_${N}gpu.json 2> gpurun_out/bench_${wl}_${N}gpu.err
  echo "rc=$?" >> gpurun_out/bench_${wl}_${N}gpu.err
done
