"""CPU: the planner half (pgstrom_grafter_json -> pgstrom_try_insert_gpupreagg
restated in csrc/gpupreagg_plan.cpp) against the reference's own EXPLAIN
goldens (expected/explain_agg.out -> tests/golden/explain_agg.json): with the
session's pg_strom.* settings of each statement the EXPLAIN (VERBOSE, COSTS
OFF) text - node names, the rewritten target lists with pgstrom.nrows / psum
/ pmin / pmax / psum_x2 / pcov_*, `Bulkload:` - must be identical line by
line.  That also pins WHICH statements are offloaded (aggfunc_catalog,
gpupreagg.c:134-333): a statement the reference leaves to PostgreSQL must be
left alone here too.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import harness  # noqa: E402
from pg_strom_b200 import gpupreagg as gp  # noqa: E402
from pg_strom_b200 import pgplan as P  # noqa: E402


NUMERIC_GAP_STATEMENTS = 0


def _statements():
    with open(os.path.join(HERE, "golden", "explain_agg.json")) as f:
        return json.load(f)


def test_explain_matches_reference_goldens(lib):
    bad = []
    n_off = 0
    stmts = _statements()
    for s in stmts:
        # (one line of explain_agg.sql carries a second statement behind the ';')
        sql = s["sql"].split(") ", 1)[1].split(";", 1)[0] + ";"
        q = P.parse_regression_sql(sql)
        table, _rows = harness.fixture_table(q["table"])
        plan = gp.Plan(P.plan_regression_sql(sql, table), gucs=s["gucs"])
        try:
            out = plan.explain()
            if any("GpuPreAgg" in ln for ln in out):
                n_off += 1
            if out != s["plan"]:
                bad.append((s["sql"], s["gucs"], out, s["plan"]))
        finally:
            plan.free()
    # (partial aggregates over a numeric column used to be a gap; they are on
    # the device now - kern_numeric.cuh - and all 612 plans must match)
    numeric_gap = [b for b in bad
                   if any("pgstrom." in ln and "nume_" in ln for ln in b[3])
                   and not any("GpuPreAgg" in ln for ln in b[2])]
    other = [b for b in bad if b not in numeric_gap]
    assert not other, "%d of %d plans differ, first: %r" % (len(other), len(stmts), other[0])
    assert len(numeric_gap) <= NUMERIC_GAP_STATEMENTS, len(numeric_gap)
    n_ref = sum(1 for s in stmts if any("GpuPreAgg" in ln for ln in s["plan"]))
    assert n_off == n_ref - len(numeric_gap)
    assert n_off >= 190


def _over_join(sql):
    """The statement's plan with a HashJoin (of the scan with itself, its
    output = the scan's) between the Agg and the scan."""
    q = P.parse_regression_sql(sql)
    table, rows = harness.fixture_table(q["table"])
    tree = P.plan_regression_sql(sql, table)
    agg = tree if tree["node"] == "Agg" else tree["lefttree"]
    scan = agg["lefttree"]
    agg["lefttree"] = {"node": "HashJoin", "targetlist": scan["targetlist"], "qual": [],
                       "hashclauses": ["(a.id = b.id)"], "lefttree": scan,
                       "righttree": {"node": "Hash", "targetlist": [], "lefttree": dict(scan)}}
    return q, table, rows, tree


def test_any_outer_plan_feeds_gpupreagg(lib):
    """gpupreagg.c:2031-2107: only a SeqScan below the Agg becomes a GpuScan
    whose quals move into the kernel; any other node (here a HashJoin) stays
    as it is, keeps its quals and feeds GpuPreAgg tuple by tuple
    (gpupreagg_load_next_outer, gpupreagg.c:2418-2505 -> "Bulkload: Off").
    The partial rows of the join's output, merged by the final aggregates,
    are the statement's golden result."""
    import json
    from oracle import partial
    with open(os.path.join(os.path.dirname(__file__), "golden", "where_agg.json")) as f:
        golden = {" ".join(s["sql"].split()): s["rows"] for s in json.load(f)}
    for sql in ("select avg(smlint_x) from gpupreagg_test where key=1 group by key order by key;",
                "select covar_pop(bigsrl_x,bigsrl_x) from gpupreagg_test where key=1 "
                "group by key order by key;"):
        q, table, rows, tree = _over_join(sql)
        plan = gp.Plan(tree, gucs=harness.GUCS)
        try:
            assert plan.num_gpupreagg == 1, plan.reject_reason
            out = plan.explain()
            node = harness.find_gpreagg_node(plan.tree())
            assert node["outer_bulkload"] is False and node["outer_quals"] == []
            child = node["lefttree"]
            assert child["node"] == "HashJoin"
            scan = child["lefttree"]
            assert scan["node"] == "SeqScan" and len(scan["qual"]) == 1    # untouched
            i = next(k for k, ln in enumerate(out) if "Custom (GpuPreAgg)" in ln)
            assert any("Bulkload: Off" in ln for ln in out[i:i + 4]), out
            assert any("HashJoin" in ln for ln in out[i:]), out
            # the kernel of this plan builds for sm_100a (its qual is empty)
            prog = plan.build_program()
            lib.pgs_program_release(prog)
            # the join's output = the scan's rows that pass the scan's filter
            tuples = [t for t, r in zip(harness.rows_as_tuples(table, rows), rows)
                      if r["key"] == q["where_key"]]
            desc = plan.describe()
            prs = []
            for row0 in range(0, len(tuples), 997):
                g, order = partial.partial_rows(node, tuples[row0:row0 + 997], len(table.columns))
                prs.extend(tuple(g[k]) for k in order)
            got, types, _ = harness.final_aggregate(desc, prs, q)
            exp = golden[" ".join(sql.split())]
            assert len(got) == len(exp)
            for gr, er in zip(got, exp):
                assert all(a == b or harness.cells_match(a, b, t)
                           for a, b, t in zip(gr, er, types)), (sql, gr, er)
        finally:
            plan.free()


def test_outer_plan_without_columns_is_left_alone(lib):
    """`select sum(1E+48)` (recheck_agg.sql): Agg over a Result without output
    columns - a chunk needs at least one column, the statement stays on the
    CPU (the numeric range it probes is pinned in test_numeric_device_code.py)."""
    agg = {"node": "Agg", "aggstrategy": "plain", "grpColIdx": [], "numGroups": 1.0,
           "targetlist": [{"node": "TargetEntry",
                           "expr": P.Agg("sum", [P.Const("numeric", "1E+48")]),
                           "resno": 1, "resname": "sum", "resjunk": False}],
           "qual": [], "lefttree": {"node": "Result", "targetlist": [], "qual": []}}
    plan = gp.Plan(agg, gucs=harness.GUCS)
    try:
        assert plan.num_gpupreagg == 0
        assert "no output column" in plan.reject_reason
        assert not any("GpuPreAgg" in ln for ln in plan.explain())
    finally:
        plan.free()


def test_cost_gpupreagg(lib):
    """cost_gpupreagg (gpupreagg.c:366-464): startup / total cost, rows and
    width of the GpuPreAgg node from the outer plan's estimates, restated here
    from the reference's formula; without estimates (the regression plans) the
    node carries none and the plan is rewritten like under
    pg_strom.debug_force_gpupreagg."""
    import math
    import pytest
    sql = "select key, avg(float_x) from gpupreagg_test group by key order by key;"
    q = P.parse_regression_sql(sql)
    table, _ = harness.fixture_table(q["table"])
    tree = P.plan_regression_sql(sql, table)
    plan = gp.Plan(tree, gucs=harness.GUCS)
    assert "total_cost" not in harness.find_gpreagg_node(plan.tree())
    plan.free()

    tree = P.plan_regression_sql(sql, table)
    agg = tree["lefttree"] if tree["node"] == "Sort" else tree
    scan = agg["lefttree"]["lefttree"] if agg["lefttree"]["node"] == "Sort" else agg["lefttree"]
    scan.update(startup_cost=0.0, total_cost=834.0, plan_rows=40000.0, plan_width=50)
    agg.update(plan_rows=31.0)
    for gucs in ({}, {"pg_strom.chunk_size": "4", "gpu_setup_cost": "100",
                      "gpu_operator_cost": "0.0001"}):
        plan = gp.Plan(tree, gucs=dict(harness.GUCS, **gucs))
        try:
            node = harness.find_gpreagg_node(plan.tree())
            chunk_mb = int(gucs.get("pg_strom.chunk_size", 15))
            setup = float(gucs.get("gpu_setup_cost", 500))
            gop = float(gucs.get("gpu_operator_cost", 0.000025))
            cop = 0.0025
            maxalign = lambda x: (x + 7) // 8 * 8
            rows_per_chunk = ((chunk_mb << 20) // 8192) * (8192 - maxalign(24)) / \
                (4 + maxalign(24 + 50))
            num_chunks = max(40000.0 / rows_per_chunk, 1.0)
            startup = 0.0 + setup + 2.0 * gop * math.log2(rows_per_chunk ** 2) * num_chunks
            run = 834.0 + gop * 40000.0
            # target list: 11 outer columns (key Var, NULL consts), nrows(float_x IS NOT NULL)
            # and psum(float_x): two function calls
            run += 2 * cop * gop / cop * math.log2(rows_per_chunk) * num_chunks
            width = 4 + 4 + 2 + 4 + 8 + 4 + 8 + 32 + 2 + 4 + 8 + 4 + 8
            assert node["plan_width"] == width
            assert node["plan_rows"] == pytest.approx(31.0 * num_chunks, rel=1e-12)
            assert node["startup_cost"] == pytest.approx(startup, rel=1e-12)
            assert node["total_cost"] == pytest.approx(startup + run, rel=1e-12)
        finally:
            plan.free()


def test_aggregates_below_a_join_are_rewritten_too(lib):
    """Agg -> HashJoin -> (Agg -> SeqScan): both aggregates get a GpuPreAgg;
    the inner one keeps its GpuScan with the quals pulled up."""
    sql = "select avg(smlint_x) from gpupreagg_test where key=1 group by key order by key;"
    q, table, rows, tree = _over_join(sql)
    agg = tree if tree["node"] == "Agg" else tree["lefttree"]
    join = agg["lefttree"]
    inner = P.plan_regression_sql("select key, max(integer_x) from gpupreagg_test where key=2 "
                                  "group by key order by key;", table)
    join["righttree"] = {"node": "Hash", "targetlist": [], "lefttree": inner}
    plan = gp.Plan(tree, gucs=harness.GUCS)
    try:
        assert plan.num_gpupreagg == 2, plan.reject_reason
        t = plan.tree()
        outer = harness.find_gpreagg_node(t)
        assert outer["lefttree"]["node"] == "HashJoin" and outer["outer_quals"] == []
        inner_g = harness.find_gpreagg_node(outer["lefttree"]["righttree"])
        assert inner_g is not None and len(inner_g["outer_quals"]) == 1
        assert inner_g["lefttree"]["custom_name"] == "GpuScan"
        for idx in range(2):
            assert "gpupreagg_projection" in plan.kernel_source(idx)
    finally:
        plan.free()
