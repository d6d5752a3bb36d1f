"""The benchmark workloads of BASELINE.json: synthetic
tables, the SQL they stand for, and the plan trees handed to the planner
half.  Data is generated with a counter based RNG (Philox) so that row i does
not depend on how the table is cut into chunks.

  C2 nogrp_agg : bench_nogrp(x int4, y float8), 100M rows
       SELECT count(*), count(x), sum(x), avg(x), min(x), max(x),
              sum(y), avg(y), min(y), max(y) FROM bench_nogrp
  C3 where_agg : bench_where(f int4, key int4, v float8, w int4), 1B rows / 8 GPUs
       SELECT key, count(*), sum(w), avg(v), min(v), max(v)
         FROM bench_where WHERE f < 10 GROUP BY key
  C4 high-card : bench_hc(key int8, v int8, y float8), ~10M distinct keys
       SELECT key, count(*), avg(v), avg(y), variance(y) FROM bench_hc GROUP BY key
       (sum(int8) is not in the reference's aggfunc_catalog, gpupreagg.c:183-189;
        avg(int8) exercises the same 128-bit psum)
"""
import numpy as np

from . import pgplan as P

SEED = 42
NULL_FRACTION = 0.05


def _raw(seed, stream, row0, n):
    """n uint64 values for rows [row0, row0+n); row0 must be a multiple of 4
    (Philox4x64 yields 4 words per counter step)."""
    assert row0 % 4 == 0
    bg = np.random.Philox(key=np.uint64(seed * 1000 + stream), counter=np.uint64(row0 // 4))
    return bg.random_raw(n)


def _nullmask(seed, stream, row0, n, with_nulls):
    if not with_nulls:
        return None
    u = _raw(seed, stream, row0, n)
    return (u < np.uint64(int(NULL_FRACTION * 2.0 ** 64))).astype(np.uint8)


# ------------------------------------------------------------------ C2
NOGRP_TABLE = P.Table("bench_nogrp", [("x", "int4"), ("y", "float8")])


def nogrp_plan():
    t = NOGRP_TABLE
    x, y = t.col("x"), t.col("y")
    targets = [(P.Agg("count", star=True), "count"), (P.Agg("count", [x]), "count"),
               (P.Agg("sum", [x]), "sum"), (P.Agg("avg", [x]), "avg"),
               (P.Agg("min", [x]), "min"), (P.Agg("max", [x]), "max"),
               (P.Agg("sum", [y]), "sum"), (P.Agg("avg", [y]), "avg"),
               (P.Agg("min", [y]), "min"), (P.Agg("max", [y]), "max")]
    return P.make_agg_plan(t, targets)


def nogrp_columns(row0, n, with_nulls=True, seed=SEED):
    """x ~ U{-10^6..10^6}; y = k/1024, k ~ U{0..102399} (dyadic: any summation
    order gives the same float8 sum); 5% NULLs in each column independently."""
    x = (_raw(seed, 1, row0, n) % np.uint64(2000001)).astype(np.int64) - 1000000
    y = (_raw(seed, 2, row0, n) % np.uint64(102400)).astype(np.float64) / 1024.0
    return [(x.astype(np.int32), _nullmask(seed, 3, row0, n, with_nulls)),
            (y, _nullmask(seed, 4, row0, n, with_nulls))]


# ------------------------------------------------------------------ C3
WHERE_TABLE = P.Table("bench_where", [("f", "int4"), ("key", "int4"),
                                      ("v", "float8"), ("w", "int4")])


def where_plan(selectivity_pct=10, num_groups=1000):
    t = WHERE_TABLE
    f, key, v, w = t.col("f"), t.col("key"), t.col("v"), t.col("w")
    targets = [(key, "key"), (P.Agg("count", star=True), "count"),
               (P.Agg("sum", [w]), "sum"), (P.Agg("avg", [v]), "avg"),
               (P.Agg("min", [v]), "min"), (P.Agg("max", [v]), "max")]
    where = [P.Op("<", f, P.Const("int4", selectivity_pct))]
    return P.make_agg_plan(t, targets, group_by=["key"], where=where,
                           num_groups=num_groups)


def where_columns(row0, n, num_groups=1000, with_nulls=False, seed=SEED):
    f = (_raw(seed, 11, row0, n) % np.uint64(100)).astype(np.int32)
    key = (_raw(seed, 12, row0, n) % np.uint64(num_groups)).astype(np.int32)
    v = (_raw(seed, 13, row0, n) % np.uint64(102400)).astype(np.float64) / 1024.0
    w = ((_raw(seed, 14, row0, n) % np.uint64(2000001)).astype(np.int64) - 1000000).astype(np.int32)
    return [(f, None), (key, None), (v, _nullmask(seed, 15, row0, n, with_nulls)),
            (w, _nullmask(seed, 16, row0, n, with_nulls))]


# ------------------------------------------------------------------ C4
HC_TABLE = P.Table("bench_hc", [("key", "int8"), ("v", "int8"), ("y", "float8")])


def _mix64(u):
    u = u.copy()
    u ^= u >> np.uint64(33)
    u *= np.uint64(0xff51afd7ed558ccd)
    u ^= u >> np.uint64(33)
    u *= np.uint64(0xc4ceb9fe1a85ec53)
    u ^= u >> np.uint64(33)
    return u


def hc_plan(num_groups=10_000_000):
    t = HC_TABLE
    key, v, y = t.col("key"), t.col("v"), t.col("y")
    targets = [(key, "key"), (P.Agg("count", star=True), "count"),
               (P.Agg("avg", [v]), "avg"), (P.Agg("avg", [y]), "avg"),
               (P.Agg("variance", [y]), "variance")]
    return P.make_agg_plan(t, targets, group_by=["key"], num_groups=num_groups)


def hc_columns(row0, n, num_groups=10_000_000, seed=SEED, zipf=False):
    """key = mix64(u), u ~ U{0..num_groups-1}; v ~ U{0..10^6}; y = k/1024 with
    k ~ U{0..1023}: the sum of y^2 is exact in float8 too.  zipf: u is drawn
    with P(u = k) ~ 1/(k+1) instead (Zipf, s = 1, by inverting its continuous
    CDF ln(k)/ln(N)): a few keys take most of the rows."""
    raw = _raw(seed, 21, row0, n)
    if zipf:
        x = (raw >> np.uint64(11)).astype(np.float64) / float(1 << 53)     # [0, 1)
        u = np.minimum(np.floor(np.exp(x * np.log(float(num_groups)))).astype(np.uint64),
                       np.uint64(num_groups)) - np.uint64(1)
    else:
        u = raw % np.uint64(num_groups)
    with np.errstate(over="ignore"):
        key = _mix64(u).astype(np.int64)
    v = (_raw(seed, 22, row0, n) % np.uint64(1000001)).astype(np.int64)
    y = (_raw(seed, 23, row0, n) % np.uint64(1024)).astype(np.float64) / 1024.0
    return [(key, None), (v, None), (y, None)]


WORKLOADS = {
    "nogrp_agg": {"table": NOGRP_TABLE, "plan": nogrp_plan, "columns": nogrp_columns,
                  "row_bytes": 12},
    "where_agg": {"table": WHERE_TABLE, "plan": where_plan, "columns": where_columns,
                  "row_bytes": 20},
    "high_cardinality": {"table": HC_TABLE, "plan": hc_plan, "columns": hc_columns,
                         "row_bytes": 24},
}


def make_chunk(name, row0, n, **kw):
    from . import gpupreagg as gp
    w = WORKLOADS[name]
    cols = w["columns"](row0, n, **kw)
    coltypes = [t for _, t in w["table"].columns]
    return gp.DataStore(coltypes, cols, nrows=n), cols
