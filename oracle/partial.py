"""CPU restatement of the GpuPreAgg partial aggregation (the hot path).

TEST INFRASTRUCTURE (oracle) - never imported by the product path.

Follows the reference's device algorithm, not its launch structure:
  - per input row: gpupreagg_qual_eval, then gpupreagg_projection producing
    the initial partial values (/root/reference/gpupreagg.c:1495-1748:
    nrows = AND of EVAL(arg); psum/pmin/pmax = argument as-is; psum_x2 =
    float8mul(x,x); pcov_* NULL unless filter true and both args non-NULL)
  - rows are grouped on the key columns, NULL keys forming one group
    (gpupreagg.c:1234-1243)
  - values of one group are merged with the GPUPREAGG_AGGCALC_* rules
    (/root/reference/opencl_gpupreagg.h:862-987): PMIN/PMAX ignore NULL and
    adopt the first value, PSUM starts NULL and adds non-NULL inputs.
What the reference leaves to StromError_CpuReCheck (integer overflow of a
partial sum) is computed exactly here with python integers; the test
harness compares *final* results, so it does not matter into how many
partial rows the device splits a group.

partial_rows(desc, tree_node, rows) -> {group key tuple: [col values...]}
"""
import math

from . import pg_expr
from .pg_agg import float8_cmp, f4

ROLE_NULL, ROLE_KEY, ROLE_AGG = 0, 1, 2


def _merge(func, typ, acc, val):
    if val is None:
        return acc
    if func in ("pmin", "pmax"):
        if acc is None:
            return val
        if typ in ("float4", "float8"):
            c = float8_cmp(acc, val)
        else:
            c = (acc > val) - (acc < val)
        if func == "pmax":
            return acc if c > 0 else val
        return acc if c < 0 else val
    # nrows / psum / psum_x2 / pcov_*
    if acc is None:
        return val
    if typ == "float4":
        # the device accumulates float4 sums in double and rounds once
        return acc + val
    return acc + val


def partial_rows(gpreagg_node, outer_rows, ncols_outer):
    """gpreagg_node: the CustomPlan(GpuPreAgg) JSON node of the rewritten
    plan; outer_rows: list of row tuples of the outer relation."""
    tlist = gpreagg_node["targetlist"]
    quals = gpreagg_node.get("outer_quals") or []
    roles = []
    for tle in tlist:
        e = tle["expr"]
        if e["node"] == "Var":
            roles.append((ROLE_KEY, None))
        elif e["node"] == "Const":
            roles.append((ROLE_NULL, None))
        else:
            roles.append((ROLE_AGG, e["funcname"]))
    groups = {}
    order = []
    for row in outer_rows:
        ok = True
        for q in quals:
            if pg_expr.evaluate(q, row) is not True:
                ok = False
                break
        if not ok:
            continue
        vals = [pg_expr.evaluate(tle["expr"], row) if roles[i][0] != ROLE_NULL else None
                for i, tle in enumerate(tlist)]
        key = tuple(_canon_key(vals[i]) for i, r in enumerate(roles) if r[0] == ROLE_KEY)
        acc = groups.get(key)
        if acc is None:
            acc = groups[key] = [None] * len(tlist)
            order.append(key)
            for i, r in enumerate(roles):
                if r[0] == ROLE_KEY:
                    acc[i] = vals[i]
        for i, r in enumerate(roles):
            if r[0] == ROLE_AGG:
                acc[i] = _merge(r[1], pg_expr.etype(tlist[i]["expr"]), acc[i], vals[i])
    if not any(r[0] == ROLE_KEY for r in roles) and not groups:
        # no GROUP BY and no surviving row: one all-initial row (nrows = 0)
        groups[()] = [0 if r[1] == "nrows" else None for r in roles]
        order.append(())
    return groups, order


def _canon_key(v):
    if isinstance(v, float):
        if math.isnan(v):
            return "NaN"
        if v == 0.0:
            return 0.0
    return v
