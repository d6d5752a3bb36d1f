# usage: gpurun --gpus N --timeout T -- 'bash tools/gpu_r2_multi2.sh N tag [workload ...|all]'
# one torchrun per workload with a tight timeout (a hang costs one timeout x N GPUs)
N=$1; TAG=$2; shift 2
mkdir -p gpurun_out
for W in "$@"; do
  if [ "$W" = "all" ]; then A="--steps 10 --warmup 3"; TO=400; else A="--workload $W --steps 5 --warmup 3 --no-cpu-baseline"; TO=170; fi
  ( time timeout -k 10 $TO python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N $A ) \
      > gpurun_out/${TAG}${N}_$W.json 2> gpurun_out/${TAG}${N}_$W.err
  echo "rc=$?" >> gpurun_out/${TAG}${N}_$W.err
done
ls -la gpurun_out > gpurun_out/${TAG}${N}_ls.txt
