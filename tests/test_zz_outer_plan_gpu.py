"""GPU: GpuPreAgg fed by a node that is not a scan (gpupreagg.c:2031-2107,
gpupreagg_load_next_outer :2418-2505): the plan keeps the HashJoin below the
Agg, its tuples arrive one by one in ROW_FLAT chunks, the kernel has no qual.
Results = the reference's where_agg goldens (the join delivers the rows the
scan's filter lets through)."""
import json
import os

import pytest

import harness
from test_planner_explain import _over_join

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
STATEMENTS = [
    "select avg(smlint_x) from gpupreagg_test where key=1 group by key order by key;",
    "select count(smlint_x) from gpupreagg_test where key=1 group by key order by key;",
    "select max(smlint_x) from gpupreagg_test where key=1 group by key order by key;",
    "select covar_pop(bigsrl_x,bigsrl_x) from gpupreagg_test where key=1 group by key order by key;",
]


@pytest.mark.parametrize("fmt", ["flat", "row", "column"])
def test_agg_over_join(cuda, fmt):
    with open(os.path.join(HERE, "golden", "where_agg.json")) as f:
        golden = {" ".join(s["sql"].split()): s["rows"] for s in json.load(f)}
    for sql in STATEMENTS:
        q, _table, _rows, tree = _over_join(sql)
        res = harness.run_statement_gpu(sql, chunk_rows=997, fmt=fmt, plan_tree=tree,
                                        outer_rows=lambda r: r["key"] == q["where_key"])
        assert res["offloaded"] and res["error"] is None, res
        exp = golden[" ".join(sql.split())]
        assert len(res["rows"]) == len(exp), (sql, res["rows"], exp)
        for gr, er, bnd in zip(res["rows"], exp, res["bounds"]):
            assert all(a == b or harness.cells_match(a, b, t, bd)
                       for a, b, t, bd in zip(gr, er, res["types"], bnd)), (sql, gr, er)
