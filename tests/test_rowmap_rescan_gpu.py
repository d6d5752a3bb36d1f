"""-m gpu: the corners of the executor contract around the hot path -
kern_row_map input from a bulk-load child (/root/reference/opencl_common.h:
483-486, gpuscan.c:1425-1427: nvalids + rindex[], -1 = all rows), ReScan
(gpupreagg.c:2825-2857), EXPLAIN (gpupreagg.c:2859-2877), extern Params in
the kern_parambuf (datastore.c:41-148), empty inputs."""
import numpy as np
import pytest

from oracle import bench_oracle, partial
from pg_strom_b200 import gpupreagg as gp
from pg_strom_b200 import pgplan as P
from pg_strom_b200 import workloads as W

pytestmark = pytest.mark.gpu
GUCS = {"pg_strom.enabled": "on", "pg_strom.debug_force_gpupreagg": "on"}


def _subset(cols, idx):
    return [(np.asarray(v)[idx], None if m is None else np.asarray(m)[idx]) for v, m in cols]


@pytest.mark.parametrize("name,kw", [("nogrp_agg", {}), ("where_agg", {}),
                                     ("where_agg", {"with_nulls": True})])
def test_row_map_selects_rows(cuda, name, kw):
    """Only the rows the map lists take part; the map need not be sorted."""
    n = 300_000
    rng = np.random.default_rng(5)
    plan = gp.Plan(W.WORKLOADS[name]["plan"](), gucs=GUCS)
    sess = gp.Session(plan)
    try:
        ds, cols = W.make_chunk(name, 0, n, **kw)
        maps = [np.sort(rng.choice(n, size=n // 3, replace=False)),
                rng.permutation(n)[:12_345],
                np.arange(0, n, 7),
                np.array([0, n - 1, 5, 4]),
                np.array([], dtype=np.int64)]
        node = plan.tree()["lefttree"]
        for rm in maps:
            t = sess.submit(ds, rowmap=rm.astype(np.int32))
            assert sess.wait(t) == 0
            rows = sess.finish()
            if len(rm) == 0 and name != "nogrp_agg":
                assert rows == []
                continue
            if len(rm) == 0:
                funcs = [c["func"] for c in plan.describe()["columns"]]
                assert len(rows) == 1 and rows[0][funcs.index("nrows")] == 0
                continue
            bench_oracle.assert_partial_equal_node(plan.describe(), node, rows, _subset(cols, rm))
        ds.free()
    finally:
        sess.close()
        plan.free()


def test_row_map_all_rows_marker(cuda):
    """nvalids = -1 means "no map, every row" (gpuscan.c:1425-1427)."""
    n = 100_000
    plan = gp.Plan(W.WORKLOADS["where_agg"]["plan"](), gucs=GUCS)
    sess = gp.Session(plan)
    try:
        ds, cols = W.make_chunk("where_agg", 0, n)
        import ctypes as C
        marker = np.array([-1], dtype=np.int32)
        t = C.c_int64()
        gp.check(sess.lib.pgs_preagg_submit(sess.handle, ds.ptr, marker.ctypes.data, C.byref(t)))
        assert sess.wait(t.value) == 0
        rows = sess.finish()
        bench_oracle.assert_partial_equal_node(plan.describe(), plan.tree()["lefttree"], rows, cols)
        ds.free()
    finally:
        sess.close()
        plan.free()


def _chunks(name, sizes, **kw):
    out, allcols, r0 = [], [], 0
    for n in sizes:
        ds, cols = W.make_chunk(name, r0, n, **kw)
        out.append(ds)
        allcols.append(cols)
        r0 += (n + 3) // 4 * 4
    merged = []
    for c in range(len(allcols[0])):
        v = np.concatenate([a[c][0] for a in allcols])
        ms = [a[c][1] for a in allcols]
        m = None if all(x is None for x in ms) else np.concatenate(
            [np.zeros(len(a[c][0]), np.uint8) if a[c][1] is None else a[c][1] for a in allcols])
        merged.append((v, m))
    return out, merged


@pytest.mark.parametrize("name", ["nogrp_agg", "where_agg"])
def test_rescan_starts_over(cuda, name):
    """ReScan drops the state of the first scan: the second scan (other
    chunks) returns its own partial rows only, a third one over the first
    chunks returns the first result again."""
    plan = gp.Plan(W.WORKLOADS[name]["plan"](), gucs=GUCS)
    try:
        node = plan.tree()["lefttree"]
        desc = plan.describe()
        a, cols_a = _chunks(name, [40_000, 33_333])
        b, cols_b = _chunks(name, [25_001], seed=7)
        st = gp.GpuPreAggState(plan, a)
        try:
            rows_a = st.fetch_all()
            bench_oracle.assert_partial_equal_node(desc, node, rows_a, cols_a)
            assert "Bulkload:" in st.explain()
            st.rescan(b)
            rows_b = st.fetch_all()
            bench_oracle.assert_partial_equal_node(desc, node, rows_b, cols_b)
            st.rescan(a)
            rows_a2 = st.fetch_all()
            assert sorted(map(repr, rows_a2)) == sorted(map(repr, rows_a))
        finally:
            st.end()
        for ds in a + b:
            ds.free()
    finally:
        plan.free()


def test_no_input_at_all(cuda):
    """No chunk: no-group aggregation still returns its one row (count 0,
    everything else NULL), GROUP BY returns nothing (zero_agg.sql)."""
    for name in ("nogrp_agg", "where_agg"):
        plan = gp.Plan(W.WORKLOADS[name]["plan"](), gucs=GUCS)
        try:
            desc = plan.describe()
            st = gp.GpuPreAggState(plan, [])
            try:
                rows = st.fetch_all()
            finally:
                st.end()
            if name == "where_agg":
                assert rows == []
                continue
            assert len(rows) == 1
            for c, v in zip(desc["columns"], rows[0]):
                assert v == (0 if c["func"] == "nrows" else None), (c["text"], v)
        finally:
            plan.free()


def test_zero_row_chunk_between_chunks(cuda):
    plan = gp.Plan(W.WORKLOADS["where_agg"]["plan"](), gucs=GUCS)
    try:
        full, cols = _chunks("where_agg", [10_000])
        w = W.WORKLOADS["where_agg"]
        coltypes = [t for _, t in w["table"].columns]
        empty = gp.DataStore(coltypes, [(np.zeros(0, gp.PGTYPES[t][3]), None) for t in coltypes],
                             nrows=0)
        st = gp.GpuPreAggState(plan, [empty, full[0], empty])
        try:
            rows = st.fetch_all()
        finally:
            st.end()
        bench_oracle.assert_partial_equal_node(plan.describe(), plan.tree()["lefttree"], rows, cols)
        full[0].free()
    finally:
        plan.free()


def test_extern_param_in_qual(cuda):
    """A Param (PARAM_EXTERN) of the qual lives in the kern_parambuf next to
    the Consts (pgstrom_create_kern_parambuf, datastore.c:41-148); a NULL
    Param makes the qual NULL, i.e. no row passes."""
    t = W.WHERE_TABLE
    n = 50_000
    ds, cols = W.make_chunk("where_agg", 0, n)
    rows_py = list(zip(*[np.asarray(v).tolist() for v, _ in cols]))
    for value, isnull in ((25, False), (None, True)):
        param = {"node": "Param", "paramkind": "extern", "paramid": 1, "paramtype": "int4",
                 "value": None if isnull else str(value), "isnull": isnull}
        tree = P.make_agg_plan(
            t, [(t.col("key"), "key"), (P.Agg("count", star=True), "count"),
                (P.Agg("max", [t.col("w")]), "max")],
            group_by=["key"], num_groups=1000,
            where=[{"node": "OpExpr", "opname": "<", "opfuncname": "int4lt",
                    "opresulttype": "bool", "args": [t.col("f"), param]}])
        plan = gp.Plan(tree, gucs=GUCS)
        try:
            assert plan.num_gpupreagg == 1, plan.reject_reason
            desc = plan.describe()
            node = plan.tree()["lefttree"]
            assert node["custom_name"] == "GpuPreAgg"
            sess = gp.Session(plan)
            try:
                tk = sess.submit(ds)
                assert sess.wait(tk) == 0
                rows = sess.finish()
            finally:
                sess.close()
            exp, _ = partial.partial_rows(node, rows_py, len(t.columns))
            got = bench_oracle.combine_device_rows(desc, rows)
            if isnull:
                assert rows == [] and exp == {}
            else:
                assert len(exp) == 1000 and set(got) == set(exp)
                for k, e in exp.items():
                    assert list(got[k]) == list(e)
        finally:
            plan.free()
    ds.free()
