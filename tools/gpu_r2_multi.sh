# usage: gpurun --gpus N --timeout T -- 'bash tools/gpu_r2_multi.sh N [each|all]'
# the merge paths on N GPUs, one workload at a time (a hang costs one timeout),
# then the whole default line.
N=$1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/m${N}_topo.txt 2>&1
run() {  # tag, timeout, args...
  tag=$1; to=$2; shift 2
  ( time timeout -k 10 $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
      --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" ) \
      > gpurun_out/m${N}_$tag.json 2> gpurun_out/m${N}_$tag.err
  echo "rc=$?" >> gpurun_out/m${N}_$tag.err
}
if [ "$2" != "all" ]; then
run nogrp 300 --workload nogrp_agg --steps 10 --warmup 3 --no-cpu-baseline
run where 300 --workload where_agg --steps 5 --warmup 3 --no-cpu-baseline
run hc 400 --workload high_cardinality --steps 3 --warmup 3 --no-cpu-baseline
run heap 300 --workload nogrp_agg_heap --steps 5 --warmup 3 --no-cpu-baseline
fi
if [ "$2" != "each" ]; then
run all 900 --steps 10 --warmup 3
fi
ls -la gpurun_out > gpurun_out/m${N}_ls.txt
