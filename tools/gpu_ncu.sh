# usage: gpu_ncu.sh <tag> <bench args...>   (plain run first, then ncu --set full of gpupreagg_main)
tag=$1; shift
timeout 200 python bench.py "$@" --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-check > gpurun_out/plain_$tag.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gpupreagg_main -s 3 -c 1 -f -o gpurun_out/prof_$tag python bench.py "$@" --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-check > gpurun_out/ncu_$tag.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_$tag.log
