"""-m gpu: NaN, +-Infinity, -0 and huge float inputs through every scan kernel
(no GROUP BY: the register fast path must notice them and take the careful
path; GROUP BY: CTA-local table, with and without a WHERE clause).
PostgreSQL orders NaN above everything in min / max
(/root/reference/opencl_common.h:1553-1560), NaN poisons a sum, and a row
whose magnitude could overflow a sum in some summation order (|x| > 2^960, or
infinite) is re-checked on the host (kern_gpupreagg.cuh, PGS_PSUM_*_LIMIT).

The merge rules these kernels apply are checked on the CPU by
tests/test_codegen_hostsim.py; what only a GPU can show is that the kernels
route such rows to them."""
import math
import random
import struct

import numpy as np
import pytest

from oracle import bench_oracle, partial
from pg_strom_b200 import gpupreagg as gp
from pg_strom_b200 import pgplan as P

pytestmark = pytest.mark.gpu
GUCS = {"pg_strom.enabled": "on", "pg_strom.debug_force_gpupreagg": "on"}
TBL = P.Table("sf", [("k", "int4"), ("x", "float8"), ("y", "float4"), ("f", "int4")])
LIMIT8 = 2.0 ** 960
NAN, INF = float("nan"), float("inf")


def make_rows(n, seed):
    """Group k decides what kind of special value the group sees, so every
    expected result is exact: finite values are small dyadic numbers."""
    rng = random.Random(seed)
    special = {1: [NAN], 2: [INF], 3: [-INF], 4: [NAN, INF, -INF], 5: [1e300, -1e300],
               6: [-0.0, 0.0], 7: [INF, 1e300, NAN]}
    rows = []
    for i in range(n):
        k = rng.randrange(0, 9)
        x = rng.randrange(-4000, 4000) / 8.0
        y = rng.randrange(-400, 400) / 4.0
        if k in special and rng.random() < 0.2:
            x = rng.choice(special[k])
        if k in special and rng.random() < 0.2:
            v = rng.choice(special[k])
            if abs(v) != 1e300:         # (a large finite float4 would make the float
                y = v                   # sums depend on the summation order)
        y = struct.unpack("f", struct.pack("f", y))[0]         # what a real column holds
        rows.append((k, None if rng.random() < 0.05 else x,
                     None if rng.random() < 0.05 else y, rng.randrange(0, 100)))
    return rows


def canon(v):
    from oracle.partial import _canon_key
    return _canon_key(v)


def same(a, b):
    if a is None or b is None:
        return a is None and b is None
    if isinstance(a, float) or isinstance(b, float):
        a, b = float(a), float(b)
        if math.isnan(a) or math.isnan(b):
            return math.isnan(a) and math.isnan(b)
    return a == b


def run_and_check(tree, rows, chunk_rows):
    plan = gp.Plan(tree, gucs=GUCS)
    try:
        assert plan.num_gpupreagg == 1, plan.reject_reason
        desc = plan.describe()
        node = plan.tree()
        while node.get("custom_name") != "GpuPreAgg":
            node = node["lefttree"]
        coltypes = [t for _, t in TBL.columns]
        chunks = []
        for lo in range(0, len(rows), chunk_rows):
            part = rows[lo:lo + chunk_rows]
            cols = []
            for c, typ in enumerate(coltypes):
                raw = [r[c] for r in part]
                mask = np.array([v is None for v in raw], dtype=np.uint8)
                arr = np.array([0 if v is None else v for v in raw], dtype=gp.PGTYPES[typ][3])
                cols.append((arr, mask if mask.any() else None))
            chunks.append(gp.DataStore(coltypes, cols, nrows=len(part)))
        st = gp.GpuPreAggState(plan, chunks)
        try:
            device_rows = st.fetch_all()
            recheck = sorted(s * chunk_rows + r for s, r in st.recheck_rows())
        finally:
            st.end()
        for ds in chunks:
            ds.free()
    finally:
        plan.free()
    # rows the device must leave to the host: a value that feeds a float sum
    # and is infinite or beyond the limit (only among rows that pass the qual)
    # (avg(float4) sums (y)::float8 in a DOUBLE cell: the float8 limit applies,
    # so of the float4 values only the infinite ones are re-checked; 1e300 also
    # overflows x * x of the variance)
    sums8 = any(c["role"] == 2 and c["op"] == "PSUM" and c["cell_type"] == "DOUBLE"
                and "x" in c["text"] for c in desc["columns"])
    sums4 = any(c["role"] == 2 and c["op"] == "PSUM" and c["cell_type"] == "DOUBLE"
                and "y" in c["text"] for c in desc["columns"])
    quals = node.get("outer_quals") or []
    from oracle import pg_expr
    want = []
    for i, r in enumerate(rows):
        if not all(pg_expr.evaluate(q, r) is True for q in quals):
            continue
        if (sums8 and r[1] is not None and abs(r[1]) > LIMIT8) or \
                (sums4 and r[2] is not None and abs(r[2]) > LIMIT8):
            want.append(i)
    assert recheck == want, (len(recheck), len(want))
    skip = set(want)
    exp, _ = partial.partial_rows(node, [r for i, r in enumerate(rows) if i not in skip],
                                  len(TBL.columns))
    exp = {tuple(canon(k) for k in key): v for key, v in exp.items()}
    key_idx = [i for i, c in enumerate(desc["columns"]) if c["role"] == 1]
    drows = [tuple(canon(v) if i in key_idx else v for i, v in enumerate(r)) for r in device_rows]
    got = bench_oracle.combine_device_rows(desc, drows)
    assert set(got) == set(exp)
    for key, erow in exp.items():
        for i, c in enumerate(desc["columns"]):
            if c["role"] == 2:
                assert same(got[key][i], erow[i]), (key, c["text"], got[key][i], erow[i])
    return len(want)


def _targets(with_sums):
    t = TBL
    tg = [(P.Agg("count", star=True), "count"), (P.Agg("count", [t.col("x")]), "count"),
          (P.Agg("min", [t.col("x")]), "min"), (P.Agg("max", [t.col("x")]), "max"),
          (P.Agg("min", [t.col("y")]), "min"), (P.Agg("max", [t.col("y")]), "max")]
    if with_sums:
        tg += [(P.Agg("sum", [t.col("x")]), "sum"), (P.Agg("avg", [t.col("y")]), "avg"),
               (P.Agg("variance", [t.col("x")]), "variance")]
    return tg


@pytest.mark.parametrize("with_sums", [False, True])
@pytest.mark.parametrize("shape", ["nogroup", "nogroup_where", "group", "group_where"])
def test_special_floats(cuda, shape, with_sums):
    t = TBL
    rows = make_rows(60_000, seed=77)
    keyed = shape.startswith("group")
    where = [P.Op("<", t.col("f"), P.Const("int4", 40))] if shape.endswith("where") else []
    tree = P.make_agg_plan(t, ([(t.col("k"), "k")] if keyed else []) + _targets(with_sums),
                           group_by=["k"] if keyed else [], where=where, num_groups=16)
    nre = run_and_check(tree, rows, chunk_rows=25_000)
    assert (nre > 100) == with_sums


def test_special_floats_as_group_keys(cuda):
    """float8 keys: NaN = NaN and -0 = +0 form one group each (btree order)."""
    t = TBL
    rows = make_rows(20_000, seed=78)
    tree = P.make_agg_plan(t, [(t.col("x"), "x"), (P.Agg("count", star=True), "count"),
                               (P.Agg("max", [t.col("f")]), "max")],
                           group_by=["x"], num_groups=20000)
    run_and_check(tree, rows, chunk_rows=20_000)
