# Round 2, last GPU call (two B200): the PostgreSQL glue executing a rewritten plan on the
# device, then the no-group merge over NVLink with the cheaper fences at N = 2.
mkdir -p gpurun_out
( time timeout 110 python -m pytest tests/test_pg_glue_plan.py -m gpu -q -rf --timeout 100 ) > gpurun_out/y_glue_test.log 2>&1; echo "rc=$?" >> gpurun_out/y_glue_test.log
( time timeout -k 10 130 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29517 bench.py --gpus 2 --workload nogrp_agg --steps 10 --warmup 3 --no-cpu-baseline ) \
    > gpurun_out/y2_nogrp_agg.json 2> gpurun_out/y2_nogrp_agg.err; echo "rc=$?" >> gpurun_out/y2_nogrp_agg.err
