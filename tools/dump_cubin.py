#!/usr/bin/env python3
"""dump_cubin.py <workload> <out.cubin>: the NVRTC cubin of a bench workload's program
(honours the PGSTROM_* build knobs in the environment)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pg_strom_b200 import _capi, gpupreagg as gp, workloads as W
GUCS = {"pg_strom.enabled": "on", "pg_strom.debug_force_gpupreagg": "on", "pg_strom.perfmon": "on"}
lib = _capi.load()
mk = {"nogrp_agg": W.nogrp_plan, "where_agg": W.where_plan, "high_cardinality": W.hc_plan}[sys.argv[1]]
plan = gp.Plan(mk(), gucs=GUCS)
prog = plan.build_program()
n = C.c_size_t()
p = lib.pgs_program_cubin(prog, C.byref(n))
open(sys.argv[2], "wb").write(C.string_at(p, n.value))
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write(plan.kernel_source())
print("wrote", sys.argv[2], n.value)
