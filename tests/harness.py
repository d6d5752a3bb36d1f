"""Test harness: runs a regression statement through the product path
(planner half -> chunk -> CUDA GpuPreAgg via the C ABI) and finishes it the
way PostgreSQL's final Agg node would, using the oracle's restatement of the
final aggregates.  The oracle is only the checker / the stand-in for
PostgreSQL; nothing here feeds oracle results into the device path.
"""
import json
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


_TUPLE_CACHE = {}


def table_tuples(table, rows):
    """rows_as_tuples, once per row list (the oracle-side caches key on the
    identity of the result)."""
    hit = _TUPLE_CACHE.get(id(rows))
    if hit is None or hit[1] is not rows:
        hit = _TUPLE_CACHE[id(rows)] = (rows_as_tuples(table, rows), rows)
    return hit[0]


_COLUMN_CACHE = {}      # (table, column, slice) -> (numpy values | varlena images, mask)
_HEAP_CACHE = {}        # (table, fmt, slice) -> HeapDataStore, kept for the whole run


def _column_arrays(table, rows, c, nrows_slice, cacheable):
    """One table column as the (values, null mask) pair the chunk builders
    take.  The fixture tables never change, so the per-column conversion
    (a Python loop over 40 000 rows; numeric columns go through the
    library's numeric_in) is done once per column and slice."""
    name, typ = table.columns[c]
    key = (table.name, id(rows), c, nrows_slice)
    if cacheable and key in _COLUMN_CACHE:
        return _COLUMN_CACHE[key][0]
    part = rows if nrows_slice is None else rows[nrows_slice[0]:nrows_slice[1]]
    raw = [r[name] for r in part]
    mask = np.array([v is None for v in raw], dtype=np.uint8)
    attlen = gp.PGTYPES[typ][0]
    if attlen > 0:
        dt = gp.PGTYPES[typ][3]
        arr = np.array([0 if v is None else v for v in raw], dtype=dt)
        col = (arr, mask if mask.any() else None)
    else:
        vals = [None if v is None else gp.numeric_datum(format(v, "f")) for v in raw]
        col = (vals, None)
    if cacheable:
        _COLUMN_CACHE[key] = (col, rows)    # holding `rows` keeps its id() unique
    return col


def make_datastore(table, rows, wanted_cols, nrows_slice=None, fmt="column", cacheable=False):
    """Chunk of the table.  fmt "column": KDS_FORMAT_COLUMN with only the
    referenced columns materialised; "row" / "flat": the reference's heap-page
    formats (KDS_FORMAT_ROW / ROW_FLAT) with every column in the tuples.
    `cacheable`: `rows` is the unmodified fixture table (heap chunks are then
    built once per slice and shared by all statements; the caller must not
    free them)."""
    if fmt != "column":
        wanted_cols = set(range(len(table.columns)))
        hkey = (table.name, id(rows), fmt, nrows_slice)
        if cacheable and hkey in _HEAP_CACHE:
            return _HEAP_CACHE[hkey][0]
    coltypes = [t for _, t in table.columns]
    columns = []
    nrows = len(rows) if nrows_slice is None else \
        max(0, min(len(rows), nrows_slice[1]) - nrows_slice[0])
    for c in range(len(table.columns)):
        if c not in wanted_cols:
            columns.append(None)
            continue
        columns.append(_column_arrays(table, rows, c, nrows_slice, cacheable))
    if fmt != "column":
        ds = gp.HeapDataStore(coltypes, columns, nrows=nrows, flat=(fmt == "flat"))
        if cacheable:
            ds.shared = True
            _HEAP_CACHE[hkey] = (ds, rows)
        return ds
    return gp.DataStore(coltypes, columns, nrows=nrows)


# Aggregates whose final value is a difference of products of sums
# (N*SX2 - SX^2 and friends): the final is only as well conditioned as that
# difference, so the float rule below checks the SUMS and propagates their
# tolerance instead of comparing the final with a bare relative tolerance.
CANCELLING = ("stddev", "stddev_samp", "stddev_pop", "variance", "var_samp", "var_pop",
              "corr", "covar_pop", "covar_samp")
SUM_RTOL = 1e-12        # north star: float8 sum / avg / variance partials, relative
UNIT_ROUNDOFF = 2.0 ** -53


def final_aggregate(desc, partial_rows, q, extra_cast=None, oracle=None):
    """PostgreSQL's Agg node over the partial rows.  Returns (text rows,
    types, bounds): bounds[row][col] is None or the admissible interval
    (lo, hi, null_ok) of a cancelling float aggregate (see cells_match).
    `oracle(key, aggref) -> (sums, abs_sums)` gives PostgreSQL's own
    left-to-right transition state of a cancelling aggregate; the state built
    from the device's partial rows must agree with it (check_partial_sums)."""
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
    bounds = []
    if key_idx:
        order = sorted([k for k in order if k[0] is not None]) + \
            [k for k in order if k[0] is None]
    for k in order:
        cells = []
        types = []
        cbounds = []
        for tle in desc["agg_targetlist"]:
            e = tle["expr"]
            if e["node"] == "Var":
                cells.append(None if k[0] is None else str(k[0]))
                types.append("int4")
                cbounds.append(None)
                continue
            assert e["node"] == "Aggref", e
            fa = pg_agg.FinalAgg(e["orig_aggname"], e.get("orig_aggargtypes") or [])
            argcols = [a["varattno"] - 1 for a in e["args"]]
            for pr in groups[k]:
                fa.accum([pr[c] for c in argcols])
            check_extension_accum(e, argcols, groups[k], fa)
            bnd = None
            if fa.agg in CANCELLING:
                if oracle is not None:
                    check_partial_sums(fa, oracle(k, e), (k, e["orig_aggname"]))
                bnd = final_interval(fa)
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
            cbounds.append(bnd)
        out.append(cells)
        bounds.append(cbounds)
    return out, types, bounds


def check_partial_sums(fa, expected, what):
    """Float rule, part 1: the transition state the final function is given
    (N, SX, SX2[, SY, SY2, SXY]) - built from the device's partial rows - must
    equal PostgreSQL's own left-to-right state: N exactly, every sum within
    SUM_RTOL relative.  Where the sum itself is ill conditioned (sum(|x_i|) >>
    |sum(x_i)|: the fixture's blocks of positive and negative values cancel to
    a thousandth of their magnitude) no order of summation - PostgreSQL's
    included - is that close to the exact sum; two orders then differ by the
    forward-error bound of recursive summation, taken in its probabilistic
    form for both: 4 * sqrt(N) * 2^-53 * sum(|x_i|).  Summation order is the
    only freedom the device has, so nothing looser is accepted."""
    if expected is None:
        return
    exp, exp_abs = expected
    got = list(fa.s)
    assert got[0] == exp[0], ("row count", what, got[0], exp[0])
    for i in range(1, len(got)):
        g, x = got[i], exp[i]
        if g == x or (g != g and x != x):
            continue
        ok = (math.isfinite(g) and math.isfinite(x) and
              abs(g - x) <= max(SUM_RTOL * max(abs(g), abs(x)),
                                4.0 * math.sqrt(max(exp[0], 1.0)) * UNIT_ROUNDOFF * exp_abs[i]))
        assert ok, ("partial sum %d" % i, what, g, x)


def final_interval(fa):
    """Float rule, part 2: the values the final function can return when every
    sum of its state moves by SUM_RTOL relative.  The numerators are
    N*SX2 - SX*SX (and N*SY2 - SY*SY, N*SXY - SX*SY); |SX*SX| <= N*SX2, so a
    numerator moves by at most 4 * SUM_RTOL * N * SX2 (for the mixed one:
    sqrt of the product of both).  Negative numerators clamp to zero exactly
    as float8_var_* does; corr() returns NULL when a variance numerator is not
    positive, so NULL is admissible when its interval reaches zero.
    Returns (lo, hi, null_ok) or None when the state is not finite."""
    s = fa.s
    if not all(math.isfinite(x) for x in s) or s[0] == 0.0:
        return None
    N = s[0]
    a = fa.agg
    try:
        if len(s) == 3:
            num = N * s[2] - s[1] * s[1]
            d = 4.0 * SUM_RTOL * abs(N * s[2])
            sample = a in ("stddev", "stddev_samp", "variance", "var_samp")
            if sample and N <= 1.0:
                return None
            den = N * (N - 1.0) if sample else N * N
            lo, hi = max(num - d, 0.0) / den, max(num + d, 0.0) / den
            if a.startswith("stddev"):
                lo, hi = math.sqrt(lo), math.sqrt(hi)
            return (lo, hi, False)
        numx = N * s[2] - s[1] * s[1]
        numy = N * s[4] - s[3] * s[3]
        numxy = N * s[5] - s[1] * s[3]
        dx = 4.0 * SUM_RTOL * abs(N * s[2])
        dy = 4.0 * SUM_RTOL * abs(N * s[4])
        dxy = 4.0 * SUM_RTOL * N * math.sqrt(abs(s[2]) * abs(s[4]))
        if a == "corr":
            null_ok = (numx - dx <= 0.0) or (numy - dy <= 0.0)
            lox, loy = numx - dx, numy - dy
            if lox <= 0.0 or loy <= 0.0:
                return (-1.0, 1.0, null_ok)         # no information in the state
            cands = [(numxy + sxy * dxy) / math.sqrt((numx + sx * dx) * (numy + sy * dy))
                     for sxy in (-1, 1) for sx in (-1, 1) for sy in (-1, 1)]
            return (max(min(cands), -1.0), min(max(cands), 1.0), null_ok)
        if a == "covar_samp" and N <= 1.0:
            return None
        den = N * N if a == "covar_pop" else N * (N - 1.0)
        return ((numxy - dxy) / den, (numxy + dxy) / den, False)
    except (OverflowError, ValueError, ZeroDivisionError):
        return None


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


_EXPR_CACHE = {}
_EVAL_ERROR = object()


def _expr_column(cache_id, tuples, expr):
    """`expr` evaluated by the oracle (oracle/pg_expr.py) for every row of the
    table; a row whose evaluation raises is marked.  Cached per table: the
    regression statements share a handful of argument expressions."""
    key = (cache_id, json.dumps(expr, sort_keys=True))
    hit = _EXPR_CACHE.get(key)
    if hit is not None:
        return hit[0]
    vals = []
    for row in tuples:
        try:
            vals.append(pg_expr.evaluate(expr, row))
        except pg_agg.PgError:
            vals.append(_EVAL_ERROR)
    _EXPR_CACHE[key] = (vals, tuples)
    return vals


def make_oracle_states(node, desc, tuples):
    """PostgreSQL's own transition state of the cancelling aggregates, per
    group: the GpuPreAgg target list (nrows / psum / psum_x2 / pcov_*
    arguments, gpupreagg.c:1495-1748) evaluated row by row on the host and
    summed LEFT TO RIGHT in table order, as the CPU executor does
    (float8_accum / float8_regr_accum).  Returns f(key_tuple, aggref) ->
    (sums, sums of absolute values) or None when a row raises."""
    cache_id = id(tuples)
    n = len(tuples)
    passed = np.ones(n, dtype=bool)
    for qn in node.get("outer_quals") or []:
        col = _expr_column(cache_id, tuples, qn)
        passed &= np.array([v is True for v in col], dtype=bool)
    key_idx = [i for i, c in enumerate(desc["columns"]) if c["role"] == 1]
    keycols = [_expr_column(cache_id, tuples, node["targetlist"][i]["expr"]) for i in key_idx]
    per_agg = {}

    def state(k, aggref):
        ent = per_agg.get(id(aggref))
        if ent is None:
            argcols = [a["varattno"] - 1 for a in aggref["args"]]
            cols = [_expr_column(cache_id, tuples, node["targetlist"][c]["expr"])
                    for c in argcols]
            bad = any(v is _EVAL_ERROR for col in cols for v in col)
            isnull = np.zeros(n, dtype=bool)
            arrs = []
            for col in cols:
                isnull |= np.array([v is None or v is _EVAL_ERROR for v in col], dtype=bool)
                arrs.append(np.array([0.0 if (v is None or v is _EVAL_ERROR) else float(v)
                                      for v in col], dtype=np.float64))
            ent = per_agg[id(aggref)] = (bad, isnull, arrs, aggref)
        bad, isnull, arrs, _ = ent
        if bad:
            return None
        sel = passed & ~isnull
        for kc, kv in zip(keycols, k):
            sel &= np.array([v == kv if kv is not None else v is None for v in kc], dtype=bool)
        idx = np.nonzero(sel)[0]
        sums, abss = [], []
        for a in arrs:
            v = a[idx]
            # np.cumsum adds strictly left to right (np.sum is pairwise)
            with np.errstate(over="ignore", invalid="ignore"):
                sums.append(float(np.cumsum(v)[-1]) if len(v) else 0.0)
                abss.append(float(np.abs(v).sum()))
        return sums, abss
    return state


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
        cacheable = outer_rows is None
        for lo in range(0, n, chunk_rows):
            chunks.append(make_datastore(table, rows, wanted, (lo, min(n, lo + chunk_rows)),
                                         fmt=fmt, cacheable=cacheable))
        st = gp.GpuPreAggState(plan, chunks, device=device)
        try:
            partial = st.fetch_all()
            recheck = st.recheck_rows()
        finally:
            notice = st.end()
        for ds in chunks:
            if not getattr(ds, "shared", False):
                ds.free()
        tuples = table_tuples(table, rows)
        recheck_abs = [(s, s * chunk_rows + r) for s, r in recheck]
        cancelling = any(tle["expr"].get("orig_aggname") in CANCELLING
                         for tle in desc["agg_targetlist"])
        try:
            partial = list(partial) + recheck_partial_rows(node, tuples, recheck_abs)
            oracle = make_oracle_states(node, desc, tuples) if cancelling else None
            out, types, bounds = final_aggregate(desc, partial, q, oracle=oracle)
            err = None
        except pg_agg.PgError as e:
            out, types, bounds, err = None, None, None, str(e)
        return {"offloaded": True, "rows": out, "types": types, "bounds": bounds, "error": err,
                "notices": [notice] if notice else [], "nrecheck": len(recheck),
                "npartial": len(partial)}
    finally:
        plan.free()


def cells_match(a, b, typ, bound=None):
    """Compare a produced cell with the golden cell - the float rule of the
    suite:

      * integer, count and numeric results: identical text;
      * float8 sum / avg / min / max: printed with 12 significant digits
        (extra_float_digits=-3): the <=1e-12 relative difference of the north
        star plus half a unit of the 12th digit; float4: one unit of its 3rd
        printed digit;
      * variance / stddev / corr / covar (`bound` = final_interval() of the
        state the device's partial rows add up to): their partial sums were
        already held to 1e-12 against PostgreSQL's left-to-right sums
        (check_partial_sums); the final itself must lie in the interval those
        tolerances allow, widened by the print precision.  A cell is never
        accepted because the device's fold order happened to land on the
        clamped side of zero."""
    if a == b:
        return True
    if typ not in ("float4", "float8", "float8::numeric"):
        # a cancelling aggregate cast to an integer type: the interval decides
        if bound is None or a is None or b is None:
            return False
        try:
            fb = float(b)
        except ValueError:
            return False
        # the cast is rint(), monotonic: every integer between rint(lo) and
        # rint(hi) is the image of an admissible value
        lo, hi, _ = bound
        return round(lo) <= fb <= round(hi)
    rel = 1.1e-2 if typ == "float4" else 6e-12
    if bound is not None:
        lo, hi, null_ok = bound
        if b is None:
            return null_ok
        if a is None:
            # the device state clamped to NULL (corr): admissible when the
            # interval reaches a non-positive variance numerator
            return null_ok
        try:
            fb = float(b)
        except ValueError:
            return False
        pad_lo = rel * abs(lo) + (1e-300 if typ != "float4" else 1e-37)
        pad_hi = rel * abs(hi) + (1e-300 if typ != "float4" else 1e-37)
        return lo - pad_lo <= fb <= hi + pad_hi
    if a is None or b is None:
        return False
    try:
        fa, fb = float(a), float(b)
    except ValueError:
        return False
    if math.isnan(fa) or math.isnan(fb) or math.isinf(fa) or math.isinf(fb):
        return False
    return abs(fa - fb) <= rel * max(abs(fa), abs(fb))
