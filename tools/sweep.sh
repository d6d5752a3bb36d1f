#!/bin/bash
# usage: sweep.sh <workload> <rows> "<cfg>" ... ; cfg = "tile stages warps ctas shslots" (0 = default, shslots -1 = default)
wl=${1:-nogrp_agg}; rows=${2:-50000000}; shift; shift
for cfg in "$@" ; do
  set -- $cfg
  unset PGSTROM_TILE_ROWS PGSTROM_NUM_STAGES PGSTROM_CONSUMER_WARPS PGSTROM_MIN_CTAS PGSTROM_SH_SLOTS
  [ "$1" != "0" ] && export PGSTROM_TILE_ROWS=$1
  [ "$2" != "0" ] && export PGSTROM_NUM_STAGES=$2
  [ "$3" != "0" ] && export PGSTROM_CONSUMER_WARPS=$3
  [ "$4" != "0" ] && export PGSTROM_MIN_CTAS=$4
  [ -n "$5" ] && [ "$5" != "-1" ] && export PGSTROM_SH_SLOTS=$5
  unset PGSTROM_DEBUG_LEVEL; [ -n "$6" ] && export PGSTROM_DEBUG_LEVEL=$6
  timeout 150 python bench.py --workload $wl --rows $rows --chunk-rows $rows --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-check 2>gpurun_out/sweep_err.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$wl cfg=[$cfg]', 'launch_ms=%.4f'%r['launch_ms'], 'GB/s=%.0f'%r['achieved'], 'frac=%.3f'%r['frac'], 'rows/s=%.3e'%d['value'], 'ms/step=%.3f'%d['ms_per_step'])
"
done
