#!/usr/bin/env python3
"""PGSTROM_DEBUG_LEVEL=4 python tools/dbg_counters.py [workload] [rows]:
runs the resident scan a few times and prints the consumer-warp cycle
counters (wait for tile / qual+queue / hash chain / number of chains)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PGSTROM_DEBUG_LEVEL", "4")
from pg_strom_b200 import _capi, gpupreagg as gp, workloads as W
wl = sys.argv[1] if len(sys.argv) > 1 else "where_agg"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000_000
sel = int(sys.argv[3]) if len(sys.argv) > 3 else 10
lib = _capi.load(); gp.cuda_init([0])
w = W.WORKLOADS[wl]
plan = gp.Plan(w["plan"](**({"selectivity_pct": sel} if wl == "where_agg" and sel != 10 else {})), gucs={"pg_strom.enabled": "on", "pg_strom.debug_force_gpupreagg": "on", "pg_strom.perfmon": "on"})
cols = w["columns"](0, rows)
ds = gp.DataStore([t for _, t in w["table"].columns], cols, nrows=rows)
dptr = lib.pgs_device_alloc(0, ds.length); _capi.check(lib.pgs_device_upload(0, dptr, ds.ptr, ds.length))
sess = gp.Session(plan, max_async_chunks=2, max_chunk_rows=rows, max_chunk_bytes=ds.length)
N = 5
for i in range(N):
    sess.submit_device(dptr, ds.length, rows); sess.finish_raw()
pm = sess.perfmon()
d = pm["debug_counters"]; nw = 148 * 16
print("kernel ms/launch %.4f" % (pm["time_kern_main_ms"] / pm["num_kern_main"]))
tot = sum(d[:3])
print("per warp per launch (cycles): wait %.0f  scan %.0f  chain %.0f  chains %.1f -> cycles/chain %.0f" % (
    d[0] / nw / N, d[1] / nw / N, d[2] / nw / N, d[3] / nw / N, d[2] / max(d[3], 1)))
print("shares: wait %.2f scan %.2f chain %.2f" % (d[0] / tot, d[1] / tot, d[2] / tot))
print({k: pm[k] for k in ("tile_rows", "num_stages", "sh_nslots", "smem_main")})
