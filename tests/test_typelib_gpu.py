"""-m gpu: date / time / timestamp arithmetic, text / bpchar comparison and
float -> numeric casts inside the fused WHERE qualifier and the projections
of GpuPreAgg (kern_timelib.cuh, kern_textlib.cuh, kern_numeric.cuh; the
catalogue of /root/reference/codegen.c:519-629), through the executor half of
the C ABI, on column chunks and on heap-page chunks.  The checker is the
oracle's per-row restatement of PostgreSQL (oracle/pg_expr.py +
oracle/pg_typelib.py + oracle/partial.py); integer, count and numeric results
must be identical, float8 sums are exact here (dyadic grid)."""
import random
from decimal import Decimal

import numpy as np
import pytest

from oracle import bench_oracle, partial, pg_typelib as T
from pg_strom_b200 import gpupreagg as gp
from pg_strom_b200 import pgplan as P
from test_typelib_device_code import EVENTS, GUCS, typelib_queries

pytestmark = pytest.mark.gpu

WORDS = [b"aaa", b"bbb", b"ccc", b"ab", b"ab ", b"ab   ", b"abc", b"", b" ", b"xx", b"xy",
         b"x" * 40, b"caf\xc3\xa9", b"\xe3\x81\x82", b"zzz", b"m" * 200]


def make_events(n, seed, wild_dates=False):
    """n rows of `events` as python tuples (None = NULL)."""
    rng = random.Random(seed)
    rows = []
    for i in range(n):
        d = rng.randrange(7000, 8000)
        if wild_dates and rng.random() < 0.02:
            d = rng.randrange(106751992, 2 ** 31 - 2)       # beyond timestamp's range
        ts = d * T.USECS_PER_DAY + rng.choice([-1, 0, 1, 3600000000, 86399999999,
                                               -40000000000, 90000000000])
        if rng.random() < 0.3:
            ts = rng.randrange(7000, 8000) * T.USECS_PER_DAY + rng.randrange(0, T.USECS_PER_DAY)
        if wild_dates:
            ts = rng.randrange(7000, 8000) * T.USECS_PER_DAY
        tm = rng.choice([0, 1, 43200000000, T.USECS_PER_DAY - 1, rng.randrange(0, T.USECS_PER_DAY)])
        s = rng.choice(WORDS)
        c = rng.choice(WORDS[:8]).ljust(5)[:5] if rng.random() < 0.7 else rng.choice(WORDS)
        k = rng.randrange(0, 12)
        v = rng.randrange(-10 ** 12, 10 ** 12)
        # multiples of 2^-10 below 2^30: float8 sums of 6000 rows are exact
        f = rng.choice([rng.randrange(-10 ** 6, 10 ** 6) / 1024.0, rng.randrange(0, 1000) / 8.0,
                        0.0, 536870912.0, 123456.5])
        row = [d, ts, tm, s, c, k, v, f]
        for j in (0, 1, 2, 3, 4, 6, 7):
            if rng.random() < 0.05:
                row[j] = None
        rows.append(tuple(row))
    return rows


def make_chunk(rows, fmt):
    coltypes = [t for _, t in EVENTS.columns]
    columns = []
    for c, typ in enumerate(coltypes):
        raw = [r[c] for r in rows]
        attlen = gp.PGTYPES[typ][0]
        if attlen > 0:
            mask = np.array([v is None for v in raw], dtype=np.uint8)
            arr = np.array([0 if v is None else v for v in raw], dtype=gp.PGTYPES[typ][3])
            columns.append((arr, mask if mask.any() else None))
        else:
            # every other datum with a 4-byte header, like a datum that was
            # never stored in a heap tuple
            vals = [None if v is None else T.varlena(v, short=None if i % 2 else False)
                    for i, v in enumerate(raw)]
            columns.append((vals, None))
    if fmt == "column":
        return gp.DataStore(coltypes, columns, nrows=len(rows))
    return gp.HeapDataStore(coltypes, columns, nrows=len(rows), flat=(fmt == "flat"))


def find_node(tree):
    n = tree
    while n is not None:
        if n.get("node") == "CustomPlan" and n.get("custom_name") == "GpuPreAgg":
            return n
        n = n.get("lefttree")
    raise AssertionError("no GpuPreAgg node")


def run(plan_tree, rows, fmt="column", chunk_rows=None):
    plan = gp.Plan(plan_tree, gucs=GUCS)
    try:
        assert plan.num_gpupreagg == 1, plan.reject_reason
        desc = plan.describe()
        node = find_node(plan.tree())
        chunk_rows = chunk_rows or len(rows)
        chunks = [make_chunk(rows[lo:lo + chunk_rows], fmt)
                  for lo in range(0, len(rows), chunk_rows)]
        st = gp.GpuPreAggState(plan, chunks)
        try:
            device_rows = st.fetch_all()
            recheck = [s * chunk_rows + r for s, r in st.recheck_rows()]
        finally:
            st.end()
        for ds in chunks:
            ds.free()
        return desc, node, device_rows, sorted(recheck)
    finally:
        plan.free()


def check(desc, node, device_rows, rows, skip=()):
    """device partial rows, merged per group == oracle partial rows."""
    skip = set(skip)
    exp, _order = partial.partial_rows(node, [r for i, r in enumerate(rows) if i not in skip],
                                       len(EVENTS.columns))
    got = bench_oracle.combine_device_rows(desc, device_rows)
    assert set(got) == set(exp), (sorted(got)[:5], sorted(exp)[:5])
    for key, erow in exp.items():
        drow = got[key]
        for i, c in enumerate(desc["columns"]):
            if c["role"] == 0:
                assert drow[i] is None
            elif isinstance(erow[i], Decimal) or isinstance(drow[i], Decimal):
                assert drow[i] is not None and Decimal(drow[i]) == Decimal(erow[i]), \
                    (key, c["text"], drow[i], erow[i])
            else:
                assert drow[i] == erow[i], (key, c["text"], drow[i], erow[i])
    return len(exp)


@pytest.mark.parametrize("fmt", ["column", "row"])
@pytest.mark.parametrize("name", [n for n, _ in typelib_queries()])
def test_typelib_query(cuda, name, fmt):
    rows = make_events(6000, seed=11)
    desc, node, device_rows, recheck = run(dict(typelib_queries())[name], rows, fmt=fmt,
                                           chunk_rows=3500)
    assert recheck == []        # nothing in this table is out of the device's ranges
    ngroups = check(desc, node, device_rows, rows)
    assert ngroups >= 1


def test_qual_selects_something(cuda):
    """Guards the tests above against vacuous passes: the text and date quals
    keep a non-trivial share of the rows."""
    rows = make_events(6000, seed=11)
    for name in ("text_eq", "date_arith", "bpchar_eq", "text_order"):
        desc, node, device_rows, _ = run(dict(typelib_queries())[name], rows)
        idx = [i for i, c in enumerate(desc["columns"]) if c["text"].startswith("pgstrom.nrows")]
        total = sum(r[idx[0]] for r in device_rows)
        assert 50 < total < 5900, (name, total)


def test_date_beyond_timestamp_range_is_rechecked(cuda):
    """A date that does not fit a timestamp makes date -> timestamp raise in
    PostgreSQL; the device flags exactly those rows CpuReCheck (here the first
    qual already rejects them, so the host's re-evaluation drops them) and
    aggregates all the others."""
    t = EVENTS
    rows = make_events(5000, seed=12, wild_dates=True)
    tree = P.make_agg_plan(
        t, [(t.col("k"), "k"), (P.Agg("count", star=True), "count"),
            (P.Agg("min", [t.col("v")]), "min")],
        group_by=["k"], num_groups=16,
        where=[P.Op("<", t.col("d"), P.Const("date", 8000)),
               P.Op(">=", P.Cast(t.col("d"), "timestamp"), t.col("ts"))])
    desc, node, device_rows, recheck = run(tree, rows, chunk_rows=2048)
    wild = [i for i, r in enumerate(rows) if r[0] is not None and r[0] > 106751991]
    assert len(wild) > 20
    assert recheck == wild
    check(desc, node, device_rows, rows)


def test_math_domain_errors_are_rechecked(cuda):
    """sqrt() of a negative number / ln() of zero raise in PostgreSQL: the
    generated wrappers flag those rows CpuReCheck, all other rows are
    filtered and aggregated on the device (sqrt is correctly rounded on both
    sides; ln is only compared far away from the threshold)."""
    from test_mathlib_device_code import MATHT, _fn
    rng = random.Random(21)
    rows = []
    for i in range(5000):
        x = rng.choice([rng.randrange(-50, 4000) / 8.0, 0.0, 16.0, 9.0, 2.0 ** 40])
        rows.append((None if rng.random() < 0.05 else x, rng.randrange(0, 7)))
    x, k = MATHT.col("x"), MATHT.col("k")
    tree = P.make_agg_plan(
        MATHT, [(k, "k"), (P.Agg("count", star=True), "count"), (P.Agg("sum", [x]), "sum")],
        group_by=["k"], num_groups=8,
        where=[P.Op(">=", _fn("sqrt", x), P.Const("float8", "3")),
               P.Op("<", _fn("ln", x), P.Const("float8", "20.5"))])
    plan = gp.Plan(tree, gucs=GUCS)
    try:
        assert plan.num_gpupreagg == 1, plan.reject_reason
        desc = plan.describe()
        node = find_node(plan.tree())
        coltypes = [t for _, t in MATHT.columns]
        xs = np.array([0.0 if r[0] is None else r[0] for r in rows])
        xm = np.array([r[0] is None for r in rows], dtype=np.uint8)
        ds = gp.DataStore(coltypes, [(xs, xm), (np.array([r[1] for r in rows], np.int32), None)],
                          nrows=len(rows))
        clean = gp.DataStore(coltypes, [(np.full(64, 16.0), None), (np.zeros(64, np.int32), None)],
                             nrows=64)
        st = gp.GpuPreAggState(plan, [ds, clean])
        try:
            device_rows = st.fetch_all()
            recheck = sorted(r for _s, r in st.recheck_rows())
            # the chunk with re-check rows (sequence number 0) stays with the
            # node until the host has walked it; the clean one went back at once
            assert {s for s, _r in st.recheck_rows()} == {0}
            assert st.released == [clean.ptr]
            assert st.recheck_chunk(0) == ds.ptr and st.recheck_chunk(1) == 0
            st.recheck_done(0)
            assert st.released == [clean.ptr, ds.ptr] and st.recheck_chunk(0) == 0
            with pytest.raises(gp._capi.StromError):
                st.recheck_done(0)
        finally:
            st.end()
        ds.free()
        clean.free()
        rows = rows + [(16.0, 0)] * 64
    finally:
        plan.free()
    bad = [i for i, r in enumerate(rows) if r[0] is not None and r[0] <= 0.0]
    assert len(bad) > 100 and recheck == bad
    exp, _ = partial.partial_rows(node, [r for i, r in enumerate(rows) if i not in set(bad)], 2)
    got = bench_oracle.combine_device_rows(desc, device_rows)
    assert set(got) == set(exp)
    for key, erow in exp.items():
        assert list(got[key]) == list(erow), (key, got[key], erow)
