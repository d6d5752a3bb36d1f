"""-m gpu: the reference's pg_regress statements (input/sql/*_agg.sql) through
the product path, compared with the goldens PostgreSQL's CPU executor
produced (expected/*.out -> tests/golden/*.json).

Integer / count / numeric cells must be identical text; float8 sum / avg / min
/ max within the 1e-12 relative tolerance of the north star (the suite prints
12 significant digits); float4 to its 3 printed digits; variance / stddev /
corr / covar by the float rule of harness.cells_match: partial sums against
PostgreSQL's left-to-right sums at 1e-12, the final inside the interval that
tolerance allows.  Statements the planner does not
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


def _check(name, chunk_rows=None, every=1, fmt="column"):
    stmts = _load(name)
    bad = []
    n_off = n_cpu = 0
    for i, s in enumerate(stmts):
        if i % every:
            continue
        r = harness.run_statement_gpu(s["sql"], chunk_rows=chunk_rows, fmt=fmt)
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
        for got, exp, bnd in zip(r["rows"], s["rows"], r["bounds"]):
            for g, e, t, b in zip(got, exp, r["types"], bnd):
                if not harness.cells_match(g, e, t, b):
                    bad.append((s["sql"], t, got, exp, b))
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


def test_heap_page_chunks(cuda):
    """The reference's own input format: KDS_FORMAT_ROW chunks (heap pages +
    row items, datastore.c:556-710) de-formed on the device."""
    n_off, _ = _check("group_agg", every=4, fmt="row")
    assert n_off >= 15
    _check("nogrp_agg", every=4, fmt="row")
    _check("where_agg", every=6, fmt="row", chunk_rows=9000)
    _check("overflow_agg", every=6, fmt="row")


def test_flat_row_chunks(cuda):
    """KDS_FORMAT_ROW_FLAT (datastore.c:799-823)."""
    _check("group_agg", every=9, fmt="flat")
    _check("zero_agg", every=9, fmt="flat")


def test_numeric_on_device(cuda, monkeypatch):
    """NUMERIC partial aggregates computed on the device (kern_numeric.cuh;
    the reference: opencl_numeric.h).  The fixture's nume_x carries 20
    fractional digits, more than the 57-bit device mantissa holds, so the
    regression statements above re-check nearly every row on the host.  Here
    the same column holds money-like values (display scales 0..4, a few
    negative) and the device has to do the work: no row may be re-checked and
    every printed digit - display scale included - must equal what
    PostgreSQL's CPU aggregates (oracle/pg_agg.py) print."""
    from decimal import Decimal
    from oracle import pg_agg, pg_fixture

    orig = pg_fixture.table
    rows = []
    for i, r in enumerate(orig("gpupreagg_test")):
        r = dict(r)
        if r["nume_x"] is not None:
            scale = (i * 7) % 5
            cents = (i * 2654435761) % 2000000007 - 700000000
            r["nume_x"] = Decimal(cents).scaleb(-scale)
        rows.append(r)
    monkeypatch.setattr(pg_fixture, "table",
                        lambda name: rows if name == "gpupreagg_test" else orig(name))
    stmts = ["select sum(nume_x) from gpupreagg_test;",
             "select avg(nume_x) from gpupreagg_test;",
             "select min(nume_x) from gpupreagg_test;",
             "select max(nume_x) from gpupreagg_test;",
             "select count(nume_x) from gpupreagg_test;",
             "select corr(nume_x,nume_x) from gpupreagg_test;",
             "select key,sum(nume_x) from gpupreagg_test group by key order by key;",
             "select key,avg(nume_x) from gpupreagg_test group by key order by key;",
             "select key,min(nume_x) from gpupreagg_test group by key order by key;",
             "select key,max(nume_x) from gpupreagg_test group by key order by key;",
             "select key,covar_pop(nume_x,nume_x) from gpupreagg_test group by key order by key;",
             "select sum(nume_x) from gpupreagg_test where key=3;"]
    for fmt in ("column", "row"):
        for sql in (stmts if fmt == "column" else stmts[:3] + stmts[6:7]):
            exp, err = pg_agg.run_query_pg(sql)
            assert err is None
            r = harness.run_statement_gpu(sql, chunk_rows=15000, fmt=fmt)
            assert r["offloaded"], sql
            assert r["nrecheck"] == 0, (sql, r["nrecheck"])
            assert len(r["rows"]) == len(exp), sql
            for got, want, bnd in zip(r["rows"], exp, r["bounds"]):
                for g, e, t, b in zip(got, want, r["types"], bnd):
                    assert harness.cells_match(g, e, t, b), (sql, fmt, got, want, b)
