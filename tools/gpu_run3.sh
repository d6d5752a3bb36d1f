timeout 300 python -m pytest tests/test_gpu_workloads.py -x -q --timeout 60 > gpurun_out/t_wl.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_wl.log
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_nogrp.json 2> gpurun_out/b_nogrp.err
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chunk-rows 100000000 > gpurun_out/b_nogrp1.json 2> gpurun_out/b_nogrp1.err
