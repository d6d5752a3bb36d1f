"""numpy restatement of the GpuPreAgg partial aggregation for large inputs.

TEST INFRASTRUCTURE (oracle) - used by tests/, __graft_entry__.smoke() and
bench.py's result check only; never by the product path.

Same semantics as oracle/partial.py (which follows
/root/reference/gpupreagg.c:1495-1748 for the per-row initial values and
/root/reference/opencl_gpupreagg.h:862-987 for the merge rules), vectorised
so that 10^8-row inputs finish in seconds.  Integer work is exact (int64 /
python int); float sums are plain left-to-right float64 sums per group,
which is exact - hence order independent - on the dyadic grids the bench
tables use.
"""
import numpy as np

_NP = {"bool": np.bool_, "int2": np.int16, "int4": np.int32, "int8": np.int64,
       "float4": np.float32, "float8": np.float64}


def _etype(e):
    n = e["node"]
    if n == "Var":
        return e["vartype"]
    if n == "Const":
        return e["consttype"]
    if n == "FuncExpr":
        return e["funcresulttype"]
    if n == "OpExpr":
        return e.get("opresulttype", "bool")
    if n in ("NullTest", "BoolExpr"):
        return "bool"
    if n == "CaseExpr":
        return e["casetype"]
    raise KeyError(n)


def npeval(e, cols, n):
    """-> (values ndarray, isnull ndarray[bool]); cols[i] = (values, nullmask|None)."""
    node = e["node"]
    if node == "Var":
        v, m = cols[e["varattno"] - 1]
        return np.asarray(v), (np.zeros(n, bool) if m is None else np.asarray(m).astype(bool))
    if node == "Const":
        t = e["consttype"]
        if e.get("constisnull"):
            return np.zeros(n, _NP.get(t, np.int64)), np.ones(n, bool)
        raw = e["constvalue"]
        val = (raw in ("t", "true")) if t == "bool" else \
            (float(raw) if t in ("float4", "float8") else int(raw))
        return np.full(n, val, _NP[t]), np.zeros(n, bool)
    if node == "NullTest":
        _, m = npeval(e["arg"], cols, n)
        r = m if e["nulltesttype"] == "IS_NULL" else ~m
        return r, np.zeros(n, bool)
    if node == "BoolExpr":
        parts = [npeval(a, cols, n) for a in e["args"]]
        if e["boolop"] == "NOT":
            return ~parts[0][0].astype(bool), parts[0][1]
        vs = [p[0].astype(bool) for p in parts]
        ms = [p[1] for p in parts]
        if e["boolop"] == "AND":
            anyfalse = np.zeros(n, bool)
            for v, m in zip(vs, ms):
                anyfalse |= (~v & ~m)
            anynull = np.logical_or.reduce(ms)
            return ~anyfalse & ~anynull, ~anyfalse & anynull
        anytrue = np.zeros(n, bool)
        for v, m in zip(vs, ms):
            anytrue |= (v & ~m)
        anynull = np.logical_or.reduce(ms)
        return anytrue, ~anytrue & anynull
    if node in ("FuncExpr", "OpExpr"):
        name = e["funcname"] if node == "FuncExpr" else e["opfuncname"]
        args = [npeval(a, cols, n) for a in e.get("args", [])]
        rtype = _etype(e)
        if node == "FuncExpr" and e.get("funcschema") == "pgstrom":
            raise ValueError("placeholder functions are evaluated by expected_partial")
        m = np.logical_or.reduce([a[1] for a in args]) if args else np.zeros(n, bool)
        if name in _NP and len(args) == 1:       # cast
            return args[0][0].astype(_NP[name]), m
        a = args[0][0]
        b = args[1][0] if len(args) > 1 else None
        for sfx, fn in (("eq", np.equal), ("ne", np.not_equal), ("le", np.less_equal),
                        ("ge", np.greater_equal), ("lt", np.less), ("gt", np.greater)):
            if name.endswith(sfx):
                return fn(a, b), m
        for sfx, fn in (("pl", np.add), ("mi", np.subtract), ("mul", np.multiply)):
            if name.endswith(sfx):
                return fn(a.astype(_NP[rtype]), b.astype(_NP[rtype])), m
        raise NotImplementedError(name)
    raise NotImplementedError(node)


def expected_partial_node(gpreagg_node, cols):
    """-> (keys: list of tuples, columns: list over target list of python lists)"""
    n = len(np.asarray(cols[[i for i, c in enumerate(cols) if c is not None][0]][0]))
    tlist = gpreagg_node["targetlist"]
    keep = np.ones(n, bool)
    for q in gpreagg_node.get("outer_quals") or []:
        v, m = npeval(q, cols, n)
        keep &= (v.astype(bool) & ~m)
    key_cols = [i for i, t in enumerate(tlist) if t["expr"]["node"] == "Var"]
    idx = np.nonzero(keep)[0]
    if key_cols:
        kv = []
        for i in key_cols:
            v, m = npeval(tlist[i]["expr"], cols, n)
            assert not m[idx].any(), "NULL keys: use oracle/partial.py"
            kv.append(v[idx])
        if len(kv) == 1:
            uniq, inv = np.unique(kv[0], return_inverse=True)
            keys = [(int(k),) for k in uniq]
        else:
            stacked = np.stack(kv, axis=1)
            uniq, inv = np.unique(stacked, axis=0, return_inverse=True)
            keys = [tuple(int(x) for x in k) for k in uniq]
        ngroups = len(keys)
    else:
        inv = np.zeros(len(idx), np.int64)
        keys = [()]
        ngroups = 1
    order = np.argsort(inv, kind="stable")
    sinv = inv[order]
    starts = np.searchsorted(sinv, np.arange(ngroups))
    out = []
    for i, tle in enumerate(tlist):
        e = tle["expr"]
        if e["node"] == "Const":
            out.append([None] * ngroups)
            continue
        if e["node"] == "Var":
            out.append([k[key_cols.index(i)] for k in keys])
            continue
        f = e["funcname"]
        args = e.get("args", [])
        if f == "nrows":
            ok = np.ones(n, bool)
            for a in args:
                v, m = npeval(a, cols, n)
                ok &= (v.astype(bool) & ~m)
            cnt = np.bincount(inv, weights=None, minlength=ngroups) if not args else \
                np.bincount(inv[ok[idx]], minlength=ngroups)
            out.append([int(c) for c in cnt])
            continue
        if f in ("psum", "pmin", "pmax", "psum_x2"):
            v, m = npeval(args[0], cols, n)
            v, m = v[idx], m[idx]
            if f == "psum_x2":
                v = v.astype(np.float64) * v.astype(np.float64)
        elif f.startswith("pcov_"):
            fv, fm = npeval(args[0], cols, n)
            x, xm = npeval(args[1], cols, n)
            y, ym = npeval(args[2], cols, n)
            m = (xm | ym | fm | ~fv.astype(bool))[idx]
            x, y = x[idx].astype(np.float64), y[idx].astype(np.float64)
            v = {"pcov_x": x, "pcov_y": y, "pcov_x2": x * x, "pcov_y2": y * y,
                 "pcov_xy": x * y}[f]
        else:
            raise NotImplementedError(f)
        res = [None] * ngroups
        if len(idx) and ngroups:
            vs, ms = v[order], m[order]
            nn = np.add.reduceat((~ms).astype(np.int64), starts) if len(vs) else np.zeros(ngroups, np.int64)
            # groups with no member at all cannot exist (they come from rows)
            if f in ("pmin", "pmax"):
                isf = np.issubdtype(vs.dtype, np.floating)
                fill = (np.inf if f == "pmin" else -np.inf) if isf else \
                    (np.iinfo(vs.dtype).max if f == "pmin" else np.iinfo(vs.dtype).min)
                vv = np.where(ms, fill, vs)
                red = (np.minimum if f == "pmin" else np.maximum).reduceat(vv, starts)
            else:
                if np.issubdtype(vs.dtype, np.floating):
                    vv = np.where(ms, 0.0, vs.astype(np.float64))
                else:
                    vv = np.where(ms, 0, vs.astype(np.int64))
                red = np.add.reduceat(vv, starts)
            for g in range(ngroups):
                if nn[g] > 0:
                    res[g] = red[g].item()
        out.append(res)
    if not key_cols and len(idx) == 0:
        out = [[0] if t["expr"]["node"] == "FuncExpr" and t["expr"]["funcname"] == "nrows"
               else [None] for t in tlist]
    return keys, out


def combine_device_rows(desc, rows):
    """Merge the partial rows the device emitted for one group (a group is
    split when a count exceeds int4 or a 128-bit sum exceeds int8)."""
    cols = desc["columns"]
    key_idx = [i for i, c in enumerate(cols) if c["role"] == 1]
    groups = {}
    for r in rows:
        k = tuple(r[i] for i in key_idx)
        acc = groups.get(k)
        if acc is None:
            groups[k] = list(r)
            continue
        for i, c in enumerate(cols):
            if c["role"] != 2 or r[i] is None:
                continue
            if acc[i] is None:
                acc[i] = r[i]
            elif c["op"] == "PSUM":
                acc[i] = acc[i] + r[i]
            elif c["op"] == "PMIN":
                acc[i] = min(acc[i], r[i])
            else:
                acc[i] = max(acc[i], r[i])
    return groups


def merge_expected(desc, parts):
    """parts: [(keys, columns)] as expected_partial_node returns them, one per
    rank / shard -> one (keys, columns) for the union (PSUM adds, PMIN / PMAX
    take the extreme, NULL = no non-NULL input in that shard)."""
    dcols = desc["columns"]
    merged = {}
    for keys, out in parts:
        for g, k in enumerate(keys):
            vals = [out[i][g] for i in range(len(dcols))]
            acc = merged.get(k)
            if acc is None:
                merged[k] = vals
                continue
            for i, c in enumerate(dcols):
                if c["role"] != 2 or vals[i] is None:
                    continue
                if acc[i] is None:
                    acc[i] = vals[i]
                elif c["op"] == "PSUM":
                    acc[i] = acc[i] + vals[i]
                elif c["op"] == "PMIN":
                    acc[i] = min(acc[i], vals[i])
                else:
                    acc[i] = max(acc[i], vals[i])
    keys = sorted(merged)
    return keys, [[merged[k][i] for k in keys] for i in range(len(dcols))]


def assert_rows_equal_expected(desc, device_rows, keys, exp, rel_tol=0.0):
    got = combine_device_rows(desc, device_rows)
    assert len(got) == len(keys), "device produced %d groups, oracle %d" % (len(got), len(keys))
    dcols = desc["columns"]
    for g, k in enumerate(keys):
        assert k in got, "group %r missing on the device" % (k,)
        row = got[k]
        for i, c in enumerate(dcols):
            e = exp[i][g]
            d = row[i]
            if c["role"] == 0:
                assert d is None
                continue
            if e is None or d is None:
                assert e is None and d is None, (k, c["text"], d, e)
            elif isinstance(e, float) and rel_tol > 0:
                assert abs(d - e) <= rel_tol * max(abs(d), abs(e)), (k, c["text"], d, e)
            else:
                assert d == e, "group %r column %s: device %r != oracle %r" % (k, c["text"], d, e)
    return len(keys)


def assert_partial_equal_node(desc, gpreagg_node, device_rows, cols, rel_tol=0.0):
    keys, exp = expected_partial_node(gpreagg_node, cols)
    return assert_rows_equal_expected(desc, device_rows, keys, exp, rel_tol)
