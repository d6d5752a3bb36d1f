"""GPU: SURVEY.md 8d C5 - an int8 column whose values sit near +-2^62, so that
every sum leaves int8 after a handful of rows: avg(int8) must equal
PostgreSQL's numeric results digit for digit (128-bit cells on the device,
sums beyond int8 emitted as several partial rows, merged by the numeric
accumulator), with and without GROUP BY, on column and heap chunks."""
import pytest

import harness

pytestmark = pytest.mark.gpu


def test_int8_sums_beyond_int8(cuda, monkeypatch):
    from oracle import pg_agg, pg_fixture

    orig = pg_fixture.table
    rows = []
    for i, r in enumerate(orig("gpupreagg_test")):
        r = dict(r)
        if r["bigint_x"] is not None:
            mag = (1 << 62) + (i * 2654435761) % (1 << 61)          # 2^62 .. 1.5 * 2^62
            r["bigint_x"] = mag if (i % 7) < 5 else -mag              # sums run away upwards
        rows.append(r)
    monkeypatch.setattr(pg_fixture, "table",
                        lambda name: rows if name == "gpupreagg_test" else orig(name))
    # (sum(int8) itself is not in the reference's aggfunc_catalog, gpupreagg.c:184-189:
    # the int8 partial sum travels under avg(int8) = pgstrom.avg_numeric(nrows, psum))
    stmts = ["select avg(bigint_x) from gpupreagg_test;",
             "select min(bigint_x) from gpupreagg_test;",
             "select max(bigint_x) from gpupreagg_test;",
             "select key,avg(bigint_x) from gpupreagg_test group by key order by key;",
             "select key,max(bigint_x) from gpupreagg_test group by key order by key;",
             "select avg(bigint_x) from gpupreagg_test where key=3;"]
    for fmt in ("column", "row"):
        for sql in stmts:
            exp, err = pg_agg.run_query_pg(sql)
            assert err is None
            r = harness.run_statement_gpu(sql, chunk_rows=15000, fmt=fmt)
            assert r["offloaded"] and r["error"] is None, (sql, r)
            assert r["nrecheck"] == 0, (sql, r["nrecheck"])
            assert len(r["rows"]) == len(exp), sql
            for got, want, bnd in zip(r["rows"], exp, r["bounds"]):
                for g, e, t, b in zip(got, want, r["types"], bnd):
                    assert harness.cells_match(g, e, t, b), (sql, fmt, got, want)
