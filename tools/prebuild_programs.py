#!/usr/bin/env python3
"""Builds, ahead of time and in parallel, the device programs the GPU tests
and the bench will ask for, into the in-tree cubin cache
(pg_strom_b200/_cubin_cache/, keyed by SHA-256 of source + flags + runtime
headers).  NVRTC needs no GPU, so this runs in the build container; the cache
travels with the snapshot and the GPU box does not spend its minutes compiling
(round 1: ~250 s of a 317 s GPU test run were NVRTC builds of the regression
statements' programs, one at a time).

    python tools/prebuild_programs.py [--jobs N] [--force]

Sources: every statement of tests/golden/{nogrp,group,where,zero,overflow}_agg.json
the planner offloads (180 distinct programs for 437 statements), the bench
workloads and their test variants, and tests/golden/extra_programs.jsonl.gz when
present (sources a GPU run dumped with PGSTROM_PROGRAM_DUMP).
"""
import argparse
import ctypes as C
import gzip
import hashlib
import json
import multiprocessing
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
SUITES = ["nogrp_agg", "group_agg", "where_agg", "zero_agg", "overflow_agg"]
GUCS = {"pg_strom.enabled": "on", "pg_strom.debug_force_gpupreagg": "on"}
EXTRA = os.path.join(GOLDEN, "extra_programs.jsonl.gz")


def regression_sources():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness
    from pg_strom_b200 import gpupreagg as gp
    from pg_strom_b200 import pgplan as P
    out = []
    for name in SUITES:
        path = os.path.join(GOLDEN, name + ".json")
        if not os.path.exists(path):
            continue
        with open(path) as f:
            stmts = json.load(f)
        for s in stmts:
            q = P.parse_regression_sql(s["sql"])
            table, _rows = harness.fixture_table(q["table"])
            plan = gp.Plan(P.plan_regression_sql(s["sql"], table), gucs=GUCS)
            try:
                for i in range(plan.num_gpupreagg):
                    out.append((plan.kernel_source(i), plan.extra_flags(i)))
            finally:
                plan.free()
    return out


def workload_sources():
    from pg_strom_b200 import gpupreagg as gp
    from pg_strom_b200 import workloads as W
    out = []
    variants = [("nogrp_agg", {}), ("where_agg", {}), ("where_agg", {"selectivity_pct": 1}),
                ("where_agg", {"selectivity_pct": 50}), ("where_agg", {"selectivity_pct": 100}),
                ("where_agg", {"num_groups": 120_000}), ("high_cardinality", {}),
                ("high_cardinality", {"num_groups": 200_000}),
                ("high_cardinality", {"num_groups": 100_000}),
                ("high_cardinality", {"num_groups": 150_000})]
    for name, kw in variants:
        # where_agg: both flavours of the scan (see test_where_gather_payload_variant)
        for gather in ((None, "0") if name == "where_agg" else (None,)):
            old = os.environ.pop("PGSTROM_GATHER_PAYLOAD", None)
            if gather is not None:
                os.environ["PGSTROM_GATHER_PAYLOAD"] = gather
            try:
                plan = gp.Plan(W.WORKLOADS[name]["plan"](**kw), gucs=GUCS)
                try:
                    out.append((plan.kernel_source(), plan.extra_flags()))
                finally:
                    plan.free()
            finally:
                os.environ.pop("PGSTROM_GATHER_PAYLOAD", None)
                if old is not None:
                    os.environ["PGSTROM_GATHER_PAYLOAD"] = old
    return out


def textkey_sources():
    """GROUP BY (text, character(5)): tests/test_typelib_gpu.py and
    tools/gpu_keyheap_check.py (long keys go through the key heap)."""
    from pg_strom_b200 import gpupreagg as gp
    from pg_strom_b200 import pgplan as P
    t = P.Table("cats", [("cat", "text"), ("code", "bpchar"), ("f", "int4"), ("v", "int8")],
                typmods={"code": 4 + 5})
    out = []
    for num_groups in (200, 3000):
        for with_qual in (False, True):
            tree = P.make_agg_plan(
                t, [(t.col("cat"), "cat"), (t.col("code"), "code"),
                    (P.Agg("count", star=True), "count"), (P.Agg("sum", [t.col("f")]), "sum"),
                    (P.Agg("min", [t.col("v")]), "min")],
                group_by=["cat", "code"], num_groups=num_groups,
                where=[P.Op("<", t.col("f"), P.Const("int4", 50))] if with_qual else [])
            plan = gp.Plan(tree, gucs=GUCS)
            try:
                out.append((plan.kernel_source(), plan.extra_flags()))
            finally:
                plan.free()
    return out


def extra_sources():
    out = []
    if os.path.exists(EXTRA):
        with gzip.open(EXTRA, "rt") as f:
            for ln in f:
                d = json.loads(ln)
                out.append((d["source"], d["extra_flags"]))
    return out


def _build_one(item):
    src, flags = item
    from pg_strom_b200 import _capi
    lib = _capi.load()
    prog = C.c_void_p()
    log = C.c_char_p()
    t0 = time.time()
    rc = lib.pgs_program_build(src.encode(), flags, C.byref(prog), C.byref(log))
    if rc != 0:
        return (rc, (log.value or b"").decode(errors="replace")[:2000])
    lib.pgs_program_release(prog)
    return (0, time.time() - t0)


def main(jobs=None, force=False, quiet=False):
    from pg_strom_b200 import build as B
    B.build_library()
    cache = os.path.join(ROOT, "pg_strom_b200", "_cubin_cache")
    os.makedirs(cache, mode=0o700, exist_ok=True)
    # nothing to do when the library, the goldens and this script are unchanged
    h = hashlib.sha256()
    for p in [B.LIB, os.path.abspath(__file__), EXTRA,
              os.path.join(ROOT, "pg_strom_b200", "workloads.py"),
              os.path.join(ROOT, "pg_strom_b200", "pgplan.py")] + \
            [os.path.join(GOLDEN, s + ".json") for s in SUITES]:
        if os.path.exists(p):
            with open(p, "rb") as f:
                h.update(f.read())
    for k in sorted(os.environ):
        if k.startswith("PGSTROM_"):
            h.update(("%s=%s;" % (k, os.environ[k])).encode())
    stamp = h.hexdigest()
    stamp_file = os.path.join(cache, "prebuild.stamp")
    if not force and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return 0
    t0 = time.time()
    # binaries of earlier library builds can never be hit again (the key
    # covers the runtime headers); they only bloat what travels to the GPU box
    for fn in os.listdir(cache):
        if fn.endswith(".cubin") or ".cubin.tmp" in fn:
            try:
                os.unlink(os.path.join(cache, fn))
            except OSError:
                pass
    items = regression_sources() + workload_sources() + textkey_sources() + extra_sources()
    uniq = list({hashlib.sha1((s + "|%d" % f).encode()).hexdigest(): (s, f)
                 for s, f in items}.values())
    jobs = jobs or max(1, (os.cpu_count() or 2))
    with multiprocessing.get_context("spawn").Pool(jobs) as pool:
        res = pool.map(_build_one, uniq, chunksize=1)
    bad = [r for r in res if r[0] != 0]
    if bad:
        raise RuntimeError("device program build failure: %s" % (bad[0][1],))
    with open(stamp_file, "w") as f:
        f.write(stamp)
    if not quiet:
        built = sum(1 for r in res if r[1] > 0.2)
        print("prebuild: %d statements -> %d programs, %d compiled, %.1fs with %d jobs"
              % (len(items), len(uniq), built, time.time() - t0, jobs))
    return len(uniq)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=0)
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    main(a.jobs or None, a.force)
