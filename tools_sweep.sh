#!/bin/bash
# usage: tools_sweep.sh <workload> <rows> "<cfg>" ... ; cfg = "tile stages warps ctas"
wl=${1:-nogrp_agg}; rows=${2:-50000000}; shift; shift
for cfg in "$@" ; do
  set -- $cfg
  unset PGSTROM_TILE_ROWS PGSTROM_NUM_STAGES PGSTROM_CONSUMER_WARPS PGSTROM_MIN_CTAS
  [ "$1" != "0" ] && export PGSTROM_TILE_ROWS=$1
  [ "$2" != "0" ] && export PGSTROM_NUM_STAGES=$2
  [ "$3" != "0" ] && export PGSTROM_CONSUMER_WARPS=$3
  [ "$4" != "0" ] && export PGSTROM_MIN_CTAS=$4
  python bench.py --workload $wl --rows $rows --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>gpurun_out/sweep_err.log | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('$wl tile=$1 stages=$2 warps=$3 ctas=$4', 'launch_ms=%.4f'%r['launch_ms'], 'GB/s=%.0f'%r['achieved'], 'frac=%.3f'%r['frac'], 'rows/s=%.3e'%d['value'], 'e2e=%.3e'%d['e2e']['value'])
"
done
