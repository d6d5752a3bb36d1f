"""CPU: the SQL-side half of the path - the partial
placeholders and the accumulators of the `pgstrom.*` final aggregates
(gpupreagg.c:4251-4773, pg_strom--1.0.sql:99-401) - through the C ABI
(include/pgstrom_cuda.h section 7) against the oracle's restatement
(oracle/pg_agg.FinalAgg, pinned on the reference's goldens by
test_oracle_golden.py).

The regression statements are cut into chunks so that every group arrives as
several partial rows; the partial rows come from the oracle's restatement of
the device path (oracle/partial.py) - this test is about the merge on
PostgreSQL's side, the device path has its own parity tests.
"""
import ctypes as C
import json
import math
import os
import random
from decimal import Decimal

import pytest

import harness
from oracle import partial, pg_agg
from pg_strom_b200 import _capi
from pg_strom_b200 import finalfn
from pg_strom_b200 import gpupreagg as gp
from pg_strom_b200 import pgplan as P

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SUITES = ["nogrp_agg", "group_agg", "where_agg", "zero_agg", "overflow_agg"]


@pytest.fixture(scope="module")
def lib():
    return _capi.load()


def _oracle_state(fa):
    """What the same accumulator holds in the oracle."""
    if fa.agg == "sum":
        return [fa.nn, fa.v]
    if fa.agg == "avg":
        return [fa.N, fa.S]
    return list(fa.s)


def _same(a, b):
    if isinstance(a, float) and isinstance(b, float):
        return a == b or (math.isnan(a) and math.isnan(b))
    return a == b


@pytest.mark.parametrize("suite", SUITES)
def test_accumulators_on_regression_partials(suite, lib):
    with open(os.path.join(GOLDEN, suite + ".json")) as f:
        stmts = json.load(f)
    checked = 0
    seen = set()
    for s in stmts:
        if s["error"]:
            continue
        q = P.parse_regression_sql(s["sql"])
        table, rows = harness.fixture_table(q["table"])
        # a slice of every block of the fixture (positive / negative / mixed /
        # all-NULL, 10000 rows each) keeps the CPU suite short
        tuples = [t for i, t in enumerate(harness.rows_as_tuples(table, rows)) if i % 10000 < 1500]
        plan = gp.Plan(P.plan_regression_sql(s["sql"], table), gucs=harness.GUCS)
        if plan.num_gpupreagg != 1:
            plan.free()
            continue
        desc = plan.describe()
        shape = (q["table"], q["where_key"] is not None, q["group"], tuple(
            (t["expr"]["orig_aggname"], tuple(t["expr"].get("orig_aggargtypes") or []))
            for t in desc["agg_targetlist"] if t["expr"]["node"] == "Aggref"))
        if shape in seen:                   # same aggregates over the same kind of query
            plan.free()
            continue
        seen.add(shape)
        node = harness.find_gpreagg_node(plan.tree())
        key_idx = [i for i, c in enumerate(desc["columns"]) if c["role"] == 1]
        groups = {}
        for row0 in range(0, len(tuples), 1301):
            g, order = partial.partial_rows(node, tuples[row0:row0 + 1301], len(table.columns))
            for k in order:
                pr = tuple(g[k])
                groups.setdefault(tuple(pr[i] for i in key_idx), []).append(pr)
        for tle in desc["agg_targetlist"]:
            e = tle["expr"]
            if e["node"] != "Aggref":
                continue
            agg, argtypes = e["orig_aggname"], e.get("orig_aggargtypes") or []
            try:
                finalfn.FinalAccum(agg, argtypes)
            except KeyError:
                continue                    # merged by PostgreSQL's own sum / min / max
            argcols = [a["varattno"] - 1 for a in e["args"]]
            for k, prs in groups.items():
                mine = finalfn.FinalAccum(agg, argtypes)
                ref = pg_agg.FinalAgg(agg, argtypes)
                for pr in prs:
                    mine.accum([pr[c] for c in argcols])
                    ref.accum([pr[c] for c in argcols])
                got, exp = mine.state(), _oracle_state(ref)
                if mine.kind == "sum_float8":
                    got = got[:2]
                assert len(got) == len(exp) and all(_same(a, b) for a, b in zip(got, exp)), \
                    (s["sql"], agg, argtypes, k, got, exp)
                if mine.kind == "sum_int8":
                    assert mine.sum_int8_final() == ref.final()
                checked += 1
        plan.free()
    assert checked > 0 or suite == "zero_agg"


def test_placeholders():
    assert finalfn.partial_nrows() == 1                       # count(*)
    assert finalfn.partial_nrows(True, True) == 1
    assert finalfn.partial_nrows(True, False) == 0
    assert finalfn.partial_nrows(True, None, True) == 0
    assert finalfn.psum_x2(None) is None
    assert finalfn.psum_x2(-1.5) == 2.25
    assert math.isinf(finalfn.psum_x2(1e200))                 # float8mul would raise in PG; the
    # device flags such a row for the host instead (kern_mathlib.cuh)
    x, y = 3.0, -0.5
    assert [finalfn.pcov(k, True, x, y) for k in ("x", "y", "x2", "y2", "xy")] == \
        [3.0, -0.5, 9.0, 0.25, -1.5]
    for filt, a, b in ((False, x, y), (None, x, y), (True, None, y), (True, x, None)):
        assert all(finalfn.pcov(k, filt, a, b) is None for k in ("x", "y", "x2", "y2", "xy"))


def test_float8_overflow_and_infinity(lib):
    big = 1.7e308
    acc = finalfn.FinalAccum("avg", ["float8"])
    acc.accum([1, big])
    with pytest.raises(finalfn.FinalFnError) as ei:
        acc.accum([1, big])                                   # finite + finite -> inf
    assert ei.value.code == 1
    assert acc.state() == [1.0, big, 0.0]                     # state untouched by the failed call
    acc.accum([2, float("inf")])                              # an infinite input is not an error
    assert acc.state()[0] == 3.0 and math.isinf(acc.state()[1])
    var = finalfn.FinalAccum("variance", ["float8"])
    var.accum([1, 1.0, 1.6e308])
    with pytest.raises(finalfn.FinalFnError):
        var.accum([1, 1.0, 1.6e308])
    cov = finalfn.FinalAccum("corr", ["float8", "float8"])
    cov.accum([1, 1.0, 1.0, big, 1.0, 1.0])
    with pytest.raises(finalfn.FinalFnError):                 # the sum of Y is checked
        cov.accum([1, 1.0, 1.0, big, 1.0, 1.0])               # (gpupreagg.c:4729 checks X twice)
    assert cov.state() == [1.0, 1.0, 1.0, big, 1.0, 1.0]


def test_strict_and_null_group(lib):
    s = finalfn.FinalAccum("sum", ["int4"])
    s.accum([None])
    assert s.sum_int8_final() is None                         # PostgreSQL: NULL, reference: 0
    s.accum([-5])
    s.accum([2 ** 40])
    assert s.sum_int8_final() == 2 ** 40 - 5
    a = finalfn.FinalAccum("avg", ["int2"])
    a.accum([None, 3])
    a.accum([4, None])
    assert a.state() == [0, 0]
    a.accum([2 ** 31 - 1, -7])
    a.accum([2 ** 31 - 1, -7])                                # N beyond int4: several partial rows
    assert a.state() == [2 ** 32 - 2, -14]


def test_numeric_avg_state(lib):
    rng = random.Random(99)
    for _ in range(30):
        acc = finalfn.FinalAccum("avg", ["numeric"])
        n, total = 0, Decimal(0)
        for _ in range(rng.randrange(1, 40)):
            scale = rng.choice([0, 0, 3, 16, 20])
            v = Decimal(rng.randrange(-10 ** rng.randrange(1, 40), 10 ** rng.randrange(1, 40))) \
                .scaleb(-scale)
            nrows = rng.randrange(0, 5000)
            if rng.random() < 0.15:
                acc.accum([nrows, None])                      # NULL psum: state untouched
                continue
            acc.accum([nrows, v])
            n += nrows
            total += v
        cnt, sm = acc.state()
        assert cnt == n and sm == total, (cnt, n, sm, total)
        ref_scale = max(0, -sm.as_tuple().exponent)
        assert ref_scale == max(0, -total.as_tuple().exponent) or sm == 0
    # the N rule: exactly nrows, also for a partial row with nrows = 0
    # (the reference adds one too many, gpupreagg.c:4556-4561)
    acc = finalfn.FinalAccum("avg", ["int8"])
    acc.accum([0, 0])
    acc.accum([3, 2 ** 63 - 1])
    acc.accum([2, 2 ** 63 - 1])
    assert acc.state() == [5, Decimal(2 ** 64 - 2)]
    # device numerics arrive as "<mantissa>e<exp>" (pgstrom_fixup_kernel_numeric)
    st = C.c_void_p(lib.pgs_numeric_avg_init())
    assert lib.pgs_numeric_avg_accum(st, 1, 0, b"-12345e-3") == 0
    assert lib.pgs_numeric_avg_accum(st, 1, 0, b"5e2") == 0
    assert lib.pgs_numeric_avg_accum(st, 1, 0, b"0.0000") == 0
    buf = C.create_string_buffer(64)
    assert lib.pgs_numeric_avg_sum_text(st, buf, 64) and buf.value == b"487.6550"
    assert lib.pgs_numeric_avg_accum(st, -1, 0, b"1") == 2
    assert lib.pgs_numeric_avg_accum(st, 1, 1, b"1") == 2
    assert lib.pgs_numeric_avg_accum(st, 1, 0, b"12x") == 3
    assert lib.pgs_numeric_avg_count(st) == 3
    assert lib.pgs_numeric_avg_sum_text(st, buf, 4) == 0      # buffer too small
    lib.pgs_numeric_avg_free(st)
