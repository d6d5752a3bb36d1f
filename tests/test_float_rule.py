"""CPU: the float rule of the GPU parity suite (tests/harness.py:
check_partial_sums / final_interval / cells_match).

variance / stddev / corr / covar finals are differences of products of sums;
where the true result is (nearly) zero the sign of that difference depends on
the summation order, PostgreSQL clamps a negative one to 0 and any other order
may land on a tiny positive one (overflow_agg: every real_x of a key is the
constant 1e38 - the cell that failed on the heap-page kernel in round 1).
The rule: the partial SUMS the device returns must agree with PostgreSQL's
left-to-right sums to 1e-12; the final must lie in the interval that
tolerance allows.  Here a stand-in "device" (oracle/partial.py over shuffled,
chunked rows - another summation order, like another tile size) goes through
the same harness code as the GPU tests, against the reference's goldens."""
import json
import os
import random
import re

import pytest

import harness
from oracle import partial, pg_agg
from pg_strom_b200 import gpupreagg as gp
from pg_strom_b200 import pgplan as P

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _device_standin(sql, seed, perturb=None):
    q = P.parse_regression_sql(sql)
    table, rows = harness.fixture_table(q["table"])
    plan = gp.Plan(P.plan_regression_sql(sql, table), gucs=harness.GUCS)
    try:
        if plan.num_gpupreagg != 1:
            return None
        desc = plan.describe()
        node = harness.find_gpreagg_node(plan.tree())
        tuples = harness.table_tuples(table, rows)
        order = list(range(len(tuples)))
        random.Random(seed).shuffle(order)
        prs = []
        for lo in range(0, len(order), 2048):
            chunk = [tuples[i] for i in order[lo:lo + 2048]]
            g, keys = partial.partial_rows(node, chunk, len(table.columns))
            prs.extend(tuple(g[k]) for k in keys)
        if perturb:
            prs = [perturb(desc, pr) for pr in prs]
        oracle = harness.make_oracle_states(node, desc, tuples)
        return q, harness.final_aggregate(desc, prs, q, oracle=oracle)
    finally:
        plan.free()


def _statements(suite, pattern, limit):
    with open(os.path.join(GOLDEN, suite + ".json")) as f:
        stmts = [s for s in json.load(f) if re.search(pattern, s["sql"]) and not s["error"]]
    return stmts[:: max(1, len(stmts) // limit)]


@pytest.mark.parametrize("suite", ["overflow_agg", "group_agg", "nogrp_agg"])
def test_another_summation_order_passes(suite):
    checked = 0
    for seed, s in enumerate(_statements(suite, r"stddev|var|corr|covar", 5)):
        try:
            r = _device_standin(s["sql"], seed)
        except pg_agg.PgError:
            continue                        # raised on the host in either order
        if r is None:
            continue
        q, (rows, types, bounds) = r
        assert len(rows) == len(s["rows"]), s["sql"]
        for got, exp, bnd in zip(rows, s["rows"], bounds):
            for g, e, t, b in zip(got, exp, types, bnd):
                assert harness.cells_match(g, e, t, b), (s["sql"], got, exp, b)
                checked += 1
    assert checked >= 2


def test_the_cell_that_failed_in_round_1():
    sql = "select key, stddev_pop(real_x)::real from gpupreagg_overflow_test group by key order by key;"
    with open(os.path.join(GOLDEN, "overflow_agg.json")) as f:
        gold = [s for s in json.load(f) if " ".join(s["sql"].split()) == sql][0]
    saw_positive = False
    for seed in range(4):
        q, (rows, types, bounds) = _device_standin(sql, seed)
        for got, exp, bnd in zip(rows, gold["rows"], bounds):
            saw_positive |= (exp[1] == "0" and got[1] != "0")
            assert harness.cells_match(got[1], exp[1], types[1], bnd[1]), (seed, got, bnd)
            # ... while a bare relative tolerance on the final flips with the order
    assert saw_positive, "no order landed on the positive side: the test shows nothing"


def test_a_wrong_partial_sum_fails():
    """1e-9 off in one psum_x2: far inside what the final's interval would
    forgive near zero, caught by the check of the sums."""
    sql = "select key, stddev(float_x) from gpupreagg_test group by key order by key;"

    def perturb(desc, pr):
        i = [c["resno"] - 1 for c in desc["columns"] if c["func"] == "psum_x2"][0]
        pr = list(pr)
        if pr[i] is not None:
            pr[i] *= 1.0 + 1e-9
        return tuple(pr)
    with pytest.raises(AssertionError, match="partial sum"):
        _device_standin(sql, 1, perturb)


def test_final_outside_the_interval_fails():
    assert not harness.cells_match("0.5", "0.6", "float8", (0.49, 0.51, False))
    assert harness.cells_match("0.5", "0.500000000001", "float8", (0.5, 0.500000000001, False))
    assert not harness.cells_match("1", None, "float8", (0.9, 1.1, False))
    assert harness.cells_match("1", None, "float8", (-1.0, 1.0, True))
    assert harness.cells_match("3", "4", "int4", (3.4, 3.6, False))
    assert not harness.cells_match("3", "5", "int4", (3.4, 3.6, False))
