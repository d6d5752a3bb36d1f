"""-m gpu: the reference's pg_regress statements (input/sql/*_agg.sql) through
the product path, compared with the goldens PostgreSQL's CPU executor
produced (expected/*.out -> tests/golden/*.json).

Integer / count / numeric cells must be identical text; float8 within the
1e-12 relative tolerance of the north star (the suite prints 12 significant
digits); float4 to its 3 printed digits.  Statements the planner does not
offload (the reference does not either, see tests/test_planner_explain.py)
are skipped here: PostgreSQL runs them itself.
"""
import json
import os

import pytest

import harness

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


def _check(name, chunk_rows=None, every=1):
    stmts = _load(name)
    bad = []
    n_off = n_cpu = 0
    for i, s in enumerate(stmts):
        if i % every:
            continue
        r = harness.run_statement_gpu(s["sql"], chunk_rows=chunk_rows)
        if not r["offloaded"]:
            n_cpu += 1
            continue
        n_off += 1
        if s["error"] or r["error"]:
            if s["error"] != r["error"]:
                bad.append((s["sql"], "error", r["error"], s["error"]))
            continue
        if len(r["rows"]) != len(s["rows"]):
            bad.append((s["sql"], "row count", len(r["rows"]), len(s["rows"])))
            continue
        for got, exp in zip(r["rows"], s["rows"]):
            for g, e, t in zip(got, exp, r["types"]):
                if not harness.cells_match(g, e, t):
                    bad.append((s["sql"], t, got, exp))
                    break
            else:
                continue
            break
    assert not bad, "%d mismatches, first: %r" % (len(bad), bad[:3])
    return n_off, n_cpu


def test_nogrp_agg(cuda):
    n_off, n_cpu = _check("nogrp_agg")
    assert n_off >= 60


def test_group_agg(cuda):
    n_off, n_cpu = _check("group_agg")
    assert n_off >= 70


def test_where_agg(cuda):
    n_off, n_cpu = _check("where_agg")
    assert n_off >= 60


def test_zero_agg(cuda):
    n_off, n_cpu = _check("zero_agg")
    assert n_off >= 60


def test_overflow_agg(cuda):
    n_off, n_cpu = _check("overflow_agg")
    assert n_off >= 60


def test_group_agg_multi_chunk(cuda):
    """Same statements with the table cut into ragged chunks: the persistent
    device state must give the same answers."""
    _check("group_agg", chunk_rows=7001, every=5)
    _check("nogrp_agg", chunk_rows=12345, every=5)
