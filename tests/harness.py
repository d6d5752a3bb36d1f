"""Test harness: runs a regression statement through the product path
(planner half -> chunk -> CUDA GpuPreAgg via the C ABI) and finishes it the
way PostgreSQL's final Agg node would, using the oracle's restatement of the
final aggregates.  The oracle is only the checker / the stand-in for
PostgreSQL; nothing here feeds oracle results into the device path.
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pg_agg, pg_expr, pg_fixture  # noqa: E402
from pg_strom_b200 import gpupreagg as gp       # noqa: E402
from pg_strom_b200 import pgplan as P           # noqa: E402

GUCS = {"pg_strom.enabled": "on", "pg_strom.debug_force_gpupreagg": "on"}


def fixture_table(name):
    rows = pg_fixture.table(name)
    if name == "gpupreagg_mix":
        cols = list(rows[0].keys()) if rows else \
            ["id", "key"] + [b + s for b in ("smlint", "integer", "bigint", "real", "float",
                                             "nume", "smlsrl", "serial", "bigsrl")
                             for s in ("_x", "_y", "_z")]
    else:
        cols = pg_fixture.COLUMNS
    coltypes = [pg_fixture.coltype(name, c) for c in cols]
    return P.Table(name, list(zip(cols, coltypes))), rows


def rows_as_tuples(table, rows):
    names = table.colnames()
    return [tuple(r[n] for n in names) for r in rows]


def make_datastore(table, rows, wanted_cols, nrows_slice=None, fmt="column"):
    """Chunk of the table.  fmt "column": KDS_FORMAT_COLUMN with only the
    referenced columns materialised; "row" / "flat": the reference's heap-page
    formats (KDS_FORMAT_ROW / ROW_FLAT) with every column in the tuples."""
    if fmt != "column":
        wanted_cols = set(range(len(table.columns)))
    names = table.colnames()
    coltypes = [t for _, t in table.columns]
    columns = []
    if nrows_slice is not None:
        rows = rows[nrows_slice[0]:nrows_slice[1]]
    for c, (name, typ) in enumerate(table.columns):
        if c not in wanted_cols:
            columns.append(None)
            continue
        raw = [r[name] for r in rows]
        mask = np.array([v is None for v in raw], dtype=np.uint8)
        attlen = gp.PGTYPES[typ][0]
        if attlen > 0:
            dt = gp.PGTYPES[typ][3]
            arr = np.array([0 if v is None else v for v in raw], dtype=dt)
            columns.append((arr, mask if mask.any() else None))
        else:
            vals = [None if v is None else gp.numeric_datum(format(v, "f")) for v in raw]
            columns.append((vals, None))
    if fmt != "column":
        return gp.HeapDataStore(coltypes, columns, nrows=len(rows), flat=(fmt == "flat"))
    return gp.DataStore(coltypes, columns, nrows=len(rows))


def final_aggregate(desc, partial_rows, q, extra_cast=None):
    """PostgreSQL's Agg node over the partial rows.  Returns text rows."""
    cols = desc["columns"]
    key_idx = [i for i, c in enumerate(cols) if c["role"] == 1]
    groups = {}
    order = []
    for pr in partial_rows:
        k = tuple(pr[i] for i in key_idx)
        if k not in groups:
            groups[k] = []
            order.append(k)
        groups[k].append(pr)
    if not key_idx and not groups:
        groups[()] = []
        order.append(())
    out = []
    types = []
    if key_idx:
        order = sorted([k for k in order if k[0] is not None]) + \
            [k for k in order if k[0] is None]
    for k in order:
        cells = []
        types = []
        for tle in desc["agg_targetlist"]:
            e = tle["expr"]
            if e["node"] == "Var":
                cells.append(None if k[0] is None else str(k[0]))
                types.append("int4")
                continue
            assert e["node"] == "Aggref", e
            fa = pg_agg.FinalAgg(e["orig_aggname"], e.get("orig_aggargtypes") or [])
            argcols = [a["varattno"] - 1 for a in e["args"]]
            for pr in groups[k]:
                fa.accum([pr[c] for c in argcols])
            check_extension_accum(e, argcols, groups[k], fa)
            v = fa.final()
            t = fa.rettype
            if q.get("cast"):
                dst = pg_agg.SQLTYPE[q["cast"].lower()]
                v = pg_agg.cast(v, t, dst)
                # a float8 aggregate printed through ::numeric (15 digits)
                # still carries the float8 tolerance of the north star
                t = "float8::numeric" if (t == "float8" and dst == "numeric") else dst
            cells.append(pg_agg.value_out(v, "numeric" if t == "float8::numeric" else t))
            types.append(t)
        out.append(cells)
    return out, types


def check_extension_accum(aggref, argcols, partial_rows, fa):
    """The extension's own transition functions (include/pgstrom_cuda.h
    section 7, pgstrom_*_accum of gpupreagg.c:4419-4773) over the same
    partial rows must hold the state PostgreSQL's final function is given."""
    from pg_strom_b200 import finalfn
    try:
        mine = finalfn.FinalAccum(aggref["orig_aggname"], aggref.get("orig_aggargtypes") or [])
    except KeyError:
        return              # count / min / max / sum(int8, float, numeric): PostgreSQL's own
    for pr in partial_rows:
        mine.accum([pr[c] for c in argcols])
    got = mine.state()
    if fa.agg == "sum":
        exp = [fa.nn, fa.v]
    elif fa.agg == "avg":
        exp = [fa.N, fa.S]
        got = got[:2]
    else:
        exp = list(fa.s)
    same = len(got) == len(exp) and all(
        a == b or (a != a and b != b) for a, b in zip(got, exp))     # NaN == NaN here
    assert same, (aggref["orig_aggname"], got, exp)


def recheck_partial_rows(gpreagg_node, table_rows, recheck):
    """What gpupreagg_next_tuple_fallback does for the flagged rows: evaluate
    the GpuPreAgg target list on the host, one partial row per input row."""
    out = []
    quals = gpreagg_node.get("outer_quals") or []
    for _seq, r in recheck:
        row = table_rows[r]
        if any(pg_expr.evaluate(qn, row) is not True for qn in quals):
            continue
        out.append(tuple(
            None if tle["expr"]["node"] == "Const" else pg_expr.evaluate(tle["expr"], row)
            for tle in gpreagg_node["targetlist"]))
    return out


def find_gpreagg_node(tree):
    n = tree
    while n is not None:
        if n.get("node") == "CustomPlan" and n.get("custom_name") == "GpuPreAgg":
            return n
        n = n.get("lefttree")
    return None


def run_statement_gpu(sql, chunk_rows=None, device=0, fmt="column", plan_tree=None,
                      outer_rows=None):
    """Returns dict(offloaded, rows, error, notices, nrecheck).  `plan_tree`
    replaces the statement's own plan (same table), `outer_rows(row) -> bool`
    says which table rows the node below GpuPreAgg delivers."""
    q = P.parse_regression_sql(sql)
    table, rows = fixture_table(q["table"])
    if outer_rows is not None:
        rows = [r for r in rows if outer_rows(r)]
    plan = gp.Plan(plan_tree or P.plan_regression_sql(sql, table), gucs=GUCS)
    try:
        if plan.num_gpupreagg == 0:
            return {"offloaded": False, "reason": plan.reject_reason}
        desc = plan.describe()
        node = find_gpreagg_node(plan.tree())
        wanted = set(desc["incol_index"])
        n = len(rows)
        if chunk_rows is None:
            chunk_rows = max(n, 1)
        chunks = []
        for lo in range(0, n, chunk_rows):
            chunks.append(make_datastore(table, rows, wanted, (lo, min(n, lo + chunk_rows)),
                                         fmt=fmt))
        st = gp.GpuPreAggState(plan, chunks, device=device)
        try:
            partial = st.fetch_all()
            recheck = st.recheck_rows()
        finally:
            notice = st.end()
        for ds in chunks:
            ds.free()
        tuples = rows_as_tuples(table, rows)
        recheck_abs = [(s, s * chunk_rows + r) for s, r in recheck]
        try:
            partial = list(partial) + recheck_partial_rows(node, tuples, recheck_abs)
            out, types = final_aggregate(desc, partial, q)
            err = None
        except pg_agg.PgError as e:
            out, types, err = None, None, str(e)
        return {"offloaded": True, "rows": out, "types": types, "error": err,
                "notices": [notice] if notice else [], "nrecheck": len(recheck),
                "npartial": len(partial)}
    finally:
        plan.free()


def cells_match(a, b, typ):
    """Compare a produced cell with the golden cell.  Integer, count and
    numeric results must be identical text.  float8 is printed with 12
    significant digits (extra_float_digits=-3): allow the <=1e-12 relative
    difference of the north star plus half a unit of the 12th digit; float4
    is printed with 3 digits: one unit of the last printed digit."""
    if a == b:
        return True
    if a is None or b is None or typ not in ("float4", "float8", "float8::numeric"):
        return False
    try:
        fa, fb = float(a), float(b)
    except ValueError:
        return False
    if math.isnan(fa) or math.isnan(fb) or math.isinf(fa) or math.isinf(fb):
        return False
    rel = 1.1e-2 if typ == "float4" else 6e-12
    return abs(fa - fb) <= rel * max(abs(fa), abs(fb))
