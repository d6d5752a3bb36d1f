#!/usr/bin/env python
"""Longer runs of the host-simulation fuzzers of tests/test_codegen_hostsim.py
(CPU only): random typed expression trees and aggregates through the
generated device code compiled with g++, against the oracle.

  python tools/fuzz_hostsim.py [--seeds 8] [--queries 40] [--first-seed 1000]

Per seed: `queries` random queries over 200-300 random rows, each checked
three ways - qual + projection row by row (check_query), every flavour of the
merge rules + flush (check_aggregation), and the original query on
PostgreSQL's own aggregates against the rewritten plan (the end-to-end test,
FILTER clauses included).  Prints one line per seed; exits 1 on a mismatch.
"""
import argparse
import os
import random
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=8)
    ap.add_argument("--queries", type=int, default=40)
    ap.add_argument("--first-seed", type=int, default=1000)
    args = ap.parse_args()

    import __graft_entry__ as ge
    ge.build()
    import test_codegen_hostsim as H
    from pg_strom_b200 import _capi
    from pg_strom_b200 import gpupreagg as gp
    from pg_strom_b200 import pgplan as P
    lib = _capi.load()
    simdir = H.simdir.__wrapped__(lib)          # the fixture's body
    H.install_device_numeric_range(setattr)     # what the autouse fixture does under pytest
    src = open(H.__file__).read()
    failures = 0
    for seed in range(args.first_seed, args.first_seed + args.seeds):
        rng = random.Random(seed)
        rows = H.rand_rows(rng.choice([200, 300]), rng)
        nq = 0
        H._nbuilt[0] = seed * 10000
        for _ in range(args.queries):
            quals = [H.gen("bool", 4, rng) for _ in range(rng.choice([0, 1, 1, 2]))]
            aggs = []
            for _ in range(rng.choice([1, 2, 3, 4])):
                typ = rng.choice(H.NUM)
                fns = ["min", "max", "avg", "count"] + ([] if typ == "int8" else ["sum"]) + \
                    (["stddev", "variance"] if typ in ("float4", "float8") else [])
                aggs.append((P.Agg(rng.choice(fns), [H.gen(typ, 3, rng)]), "agg"))
            keyed = rng.random() < 0.6
            keycol = rng.choice(["k", "f8", "d", "b", "tx", "f4", "i8", "ts", "s2"])
            targets = ([(H.TBL.col(keycol), keycol)] if keyed else []) + \
                [(P.Agg("count", star=True), "count")] + aggs
            tree = P.make_agg_plan(H.TBL, targets, group_by=[keycol] if keyed else [],
                                   where=quals, num_groups=8)
            plan = gp.Plan(tree, gucs=H.GUCS)
            ok = plan.num_gpupreagg == 1 and \
                ("#define GPUPREAGG_HAS_QUAL 1" in plan.kernel_source()) == bool(quals)
            plan.free()
            if not ok:
                continue
            nq += 1
            try:
                H.check_query(H.TBL, tree, rows, simdir)
                H.check_aggregation(H.TBL, tree, rows, simdir, lib)
            except AssertionError as e:
                failures += 1
                print("seed %d: MISMATCH\n%s" % (seed, str(e)[:2000]))
        # the end-to-end test with this seed (its generator is inside the test)
        code = src.replace("rng = random.Random(4242)", "rng = random.Random(%d)" % seed) \
                  .replace("assert nqueries >= 25", "pass") \
                  .replace("_nbuilt = [0]", "_nbuilt = [%d]" % (seed * 10000 + 5000))
        mod = types.ModuleType("hostsim_seed%d" % seed)
        mod.__file__ = H.__file__
        exec(compile(code, mod.__name__, "exec"), mod.__dict__)
        try:
            mod.test_end_to_end_against_postgres_own_aggregates(simdir, lib)
            e2e = "ok"
        except AssertionError as e:
            failures += 1
            e2e = "MISMATCH " + str(e)[:1500]
        print("seed %d: %d fuzzed queries ok, end-to-end %s" % (seed, nq, e2e))
        for f in os.listdir(simdir):
            if f.startswith("q") and f.endswith((".so", ".cpp")):
                os.remove(os.path.join(simdir, f))
    print("failures: %d" % failures)
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
