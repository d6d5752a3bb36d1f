"""CPU: the oracle (oracle/pg_agg.py, the restatement of PostgreSQL's CPU
aggregate executor, i.e. the reference's `pg_strom.enabled = off` side) is
pinned on the reference's own goldens: every statement of
input/sql/{nogrp,group,where,zero,overflow}_agg.sql must print exactly what
expected/*.out holds (tests/golden/*.json, extracted by make_golden.py) -
the same text for every type, floats included, because the oracle sums in
PostgreSQL's order.  The fixture tables come from oracle/pg_fixture.py, the
regeneration of agg_init.sql (glibc srandom(0) stream).
"""
import json
import os

import pytest

from oracle import pg_agg

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SUITES = ["nogrp_agg", "group_agg", "where_agg", "zero_agg", "overflow_agg"]


def _load(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


@pytest.mark.parametrize("suite", SUITES)
def test_oracle_reproduces_golden(suite):
    bad = []
    stmts = _load(suite)
    for s in stmts:
        rows, err = pg_agg.run_query_pg(s["sql"])
        if s["error"] or err:
            if (s["error"] or None) != (err or None):
                bad.append((s["sql"], "error", err, s["error"]))
            continue
        if rows != s["rows"]:
            bad.append((s["sql"], rows[:3], s["rows"][:3]))
    assert not bad, "%d of %d statements differ, first: %r" % (len(bad), len(stmts), bad[:3])


def test_fixture_known_answers():
    """Spot values SURVEY.md 8c lists for the regenerated gpupreagg_test."""
    from oracle import pg_fixture as fx
    rows = fx.table("gpupreagg_test")
    assert len(rows) == 40000
    ints = [r["integer_x"] for r in rows if r["integer_x"] is not None]
    assert (len(ints), sum(ints), max(ints), min(ints)) == (28511, 99027633, 2147112, -2147350)
    bigs = [r["bigint_x"] for r in rows if r["bigint_x"] is not None]
    assert (len(bigs), sum(bigs)) == (28502, -55757751021379520)
    assert sum(r["bigsrl_x"] for r in rows) == -44161238785585078
    k1 = [r["integer_x"] for r in rows if r["key"] == 1 and r["integer_x"] is not None]
    assert sum(k1) == 1042255413
