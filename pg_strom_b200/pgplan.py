"""Plan-tree construction helpers: what PostgreSQL's parser + planner hand to
the extension, written down as the JSON node format the C ABI takes
(include/pgstrom_cuda.h section 2).

This is harness code standing in for PostgreSQL itself (there is none in the
build image): it resolves operators / implicit casts the way the PostgreSQL
parser does for the handful of types the device supports, and builds the
plain plans (`[Sort ->] Agg -> SeqScan`) that the planner half
(pgstrom_grafter_json) rewrites.  No device logic lives here.
"""
import json
import re

SQLNAME = {"bool": "boolean", "int2": "smallint", "int4": "integer", "int8": "bigint",
           "float4": "real", "float8": "double precision", "numeric": "numeric",
           "date": "date", "text": "text", "time": "time without time zone",
           "timestamp": "timestamp without time zone", "bpchar": "character"}
_INTS = ("int2", "int4", "int8")
_FLOATS = ("float4", "float8")
_OPNAME = {"=": "eq", "<>": "ne", "<": "lt", "<=": "le", ">": "gt", ">=": "ge",
           "+": "pl", "-": "mi", "*": "mul", "/": "div", "%": "mod"}


_CMP = ("eq", "ne", "lt", "le", "gt", "ge")
# pg_operator rows of the date / time arithmetic the device runs
_DATE_OPS = {("date", "int4", "pl"): ("date_pli", "date"),
             ("date", "int4", "mi"): ("date_mii", "date"),
             ("date", "date", "mi"): ("date_mi", "int4"),
             ("int4", "date", "pl"): ("integer_pl_date", "date"),
             ("date", "time", "pl"): ("datetime_pl", "timestamp"),
             ("time", "date", "pl"): ("timedate_pl", "timestamp")}


def Var(attno, typ, typmod=None):
    v = {"node": "Var", "varattno": attno, "vartype": typ}
    if typmod is not None:
        v["vartypmod"] = typmod         # atttypmod, e.g. VARHDRSZ + n for character(n)
    return v


def Const(typ, value, isnull=False):
    if isnull or value is None:
        return {"node": "Const", "consttype": typ, "constisnull": True}
    if isinstance(value, bool):
        value = "t" if value else "f"
    return {"node": "Const", "consttype": typ, "constisnull": False,
            "constvalue": str(value)}


def Cast(expr, typ, fmt="cast"):
    """Function cast (pg_cast COERCION_METHOD_FUNCTION), e.g. int4 -> int8."""
    if etype(expr) == typ:
        return expr
    return {"node": "FuncExpr", "funcname": typ, "funcresulttype": typ,
            "funcformat": fmt, "args": [expr]}


def etype(e):
    n = e["node"]
    if n == "Var":
        return e["vartype"]
    if n == "Const":
        return e["consttype"]
    if n == "Param":
        return e["paramtype"]
    if n == "FuncExpr":
        return e["funcresulttype"]
    if n in ("OpExpr", "DistinctExpr"):
        return e.get("opresulttype", "bool")
    if n in ("NullTest", "BooleanTest", "BoolExpr"):
        return "bool"
    if n == "RelabelType":
        return e["resulttype"]
    if n == "CaseExpr":
        return e["casetype"]
    if n == "Aggref":
        return e["aggtype"]
    raise KeyError(n)


def _opfunc(op, lt, rt):
    sfx = _OPNAME[op]
    if lt in _INTS and rt in _INTS:
        name = "int" + lt[3:] + ("" if lt == rt else rt[3:]) + sfx
        arith = {"int2": 2, "int4": 4, "int8": 8}
        res = "int%d" % max(arith[lt], arith[rt])
    elif lt in _FLOATS and rt in _FLOATS:
        name = "float" + lt[5:] + ("" if lt == rt else rt[5:]) + sfx
        res = "float8" if "float8" in (lt, rt) else "float4"
    elif lt == rt == "bool":
        name, res = "bool" + sfx, "bool"
    elif lt == rt and lt in ("date", "time", "timestamp") and sfx in _CMP:
        name, res = lt + "_" + sfx, "bool"
    elif (lt, rt) in (("date", "timestamp"), ("timestamp", "date")) and sfx in _CMP:
        name, res = "%s_%s_%s" % (lt, sfx, rt), "bool"
    elif (lt, rt, sfx) in _DATE_OPS:
        name, res = _DATE_OPS[(lt, rt, sfx)]
    elif lt == rt == "text" and sfx in _CMP:
        name, res = ("text" + sfx if sfx in ("eq", "ne") else "text_" + sfx), "bool"
    elif lt == rt == "bpchar" and sfx in _CMP:
        name, res = "bpchar" + sfx, "bool"
    elif lt == rt == "numeric":
        name = {"eq": "numeric_eq", "ne": "numeric_ne", "lt": "numeric_lt",
                "le": "numeric_le", "gt": "numeric_gt", "ge": "numeric_ge",
                "pl": "numeric_add", "mi": "numeric_sub", "mul": "numeric_mul"}[sfx]
        res = "numeric"
    else:
        return None, None
    if sfx in ("eq", "ne", "lt", "le", "gt", "ge"):
        res = "bool"
    return name, res


def Op(op, left, right, collation=None):
    """Binary operator with PostgreSQL-like resolution: exact match among the
    cross-type integer / float operators, else promote int -> float8 /
    numeric.  `collation`: name of the input collation of a text operator
    (what the PostgreSQL glue resolves inputcollid to)."""
    lt, rt = etype(left), etype(right)
    name, res = _opfunc(op, lt, rt)
    if name is None:
        # implicit casts: integer with float -> float8; anything with numeric
        if lt in _INTS and rt in _FLOATS:
            left = Cast(left, "float8", "implicit")
        elif lt in _FLOATS and rt in _INTS:
            right = Cast(right, "float8", "implicit")
        elif lt == "numeric" and rt in _INTS:
            right = Cast(right, "numeric", "implicit")
        elif rt == "numeric" and lt in _INTS:
            left = Cast(left, "numeric", "implicit")
        lt, rt = etype(left), etype(right)
        name, res = _opfunc(op, lt, rt)
        if name is None:
            raise TypeError("operator does not exist: %s %s %s" % (lt, op, rt))
    node = {"node": "OpExpr", "opname": op, "opfuncname": name,
            "opresulttype": res, "args": [left, right]}
    if collation is not None:
        node["inputcollid"] = collation
    return node


def Distinct(left, right, collation=None):
    """left IS DISTINCT FROM right: a DistinctExpr carries the `=` operator of
    its argument types (parse_oper.c make_distinct_op)."""
    node = Op("=", left, right, collation)
    node["node"] = "DistinctExpr"
    return node


def And(*args):
    return {"node": "BoolExpr", "boolop": "AND", "args": list(args)}


def Or(*args):
    return {"node": "BoolExpr", "boolop": "OR", "args": list(args)}


def Not(arg):
    return {"node": "BoolExpr", "boolop": "NOT", "args": [arg]}


def IsNull(arg, notnull=False):
    return {"node": "NullTest", "arg": arg,
            "nulltesttype": "IS_NOT_NULL" if notnull else "IS_NULL", "argisrow": False}


def Case(whens, default, typ):
    return {"node": "CaseExpr", "casetype": typ, "arg": None,
            "args": [{"node": "CaseWhen", "expr": w, "result": r} for w, r in whens],
            "defresult": default}


# result type of PostgreSQL's own aggregates, by argument type
def _aggtype(name, argtypes):
    t = argtypes[0] if argtypes else None
    if name == "count":
        return "int8"
    if name in ("min", "max"):
        return t
    if name == "sum":
        return {"int2": "int8", "int4": "int8", "int8": "numeric", "float4": "float4",
                "float8": "float8", "numeric": "numeric"}[t]
    if name == "avg":
        return "float8" if t in _FLOATS else "numeric"
    if name in ("stddev", "stddev_samp", "stddev_pop", "variance", "var_samp", "var_pop"):
        return "float8" if t in _FLOATS else "numeric"
    if name in ("corr", "covar_pop", "covar_samp"):
        return "float8"
    raise KeyError(name)


def Agg(name, args=(), filter=None, star=False):
    """Aggref as the PostgreSQL parser produces it: corr/covar arguments are
    implicitly cast to float8 (there is only a float8 signature)."""
    args = list(args)
    if name in ("corr", "covar_pop", "covar_samp"):
        args = [Cast(a, "float8", "implicit") for a in args]
    argtypes = [etype(a) for a in args]
    return {"node": "Aggref", "aggname": name, "aggargtypes": argtypes,
            "aggtype": _aggtype(name, argtypes),
            "args": [{"node": "TargetEntry", "expr": a, "resno": i + 1}
                     for i, a in enumerate(args)],
            "aggfilter": filter, "aggstar": bool(star)}


class Table:
    """Column catalogue of a relation: [(name, type), ...]."""

    def __init__(self, name, columns, schema="public", typmods=None):
        self.name = name
        self.schema = schema
        self.columns = list(columns)
        self.typmods = dict(typmods or {})      # column name -> atttypmod

    def col(self, name):
        for i, (n, t) in enumerate(self.columns):
            if n == name:
                return Var(i + 1, t, self.typmods.get(n))
        raise KeyError(name)

    def colnames(self):
        return [n for n, _ in self.columns]

    def scan_tlist(self):
        return [{"node": "TargetEntry", "expr": Var(i + 1, t, self.typmods.get(n)),
                 "resno": i + 1,
                 "resname": n, "resjunk": False}
                for i, (n, t) in enumerate(self.columns)]


def make_agg_plan(table, targets, group_by=(), where=(), order_by_keys=False,
                  num_groups=None, strategy=None):
    """[Sort ->] Agg -> SeqScan, the plan the standard planner produces.

    targets : list of (expr, resname); exprs are Var (group keys) or Aggref
    group_by: list of column names
    where   : list of qual expressions (implicitly AND-ed)
    """
    grp_idx = [table.colnames().index(g) + 1 for g in group_by]
    scan = {"node": "SeqScan", "relname": table.name, "schema": table.schema,
            "alias": table.name, "targetlist": table.scan_tlist(), "qual": list(where)}
    if strategy is None:
        strategy = "hashed" if grp_idx else "plain"
    agg = {"node": "Agg", "aggstrategy": strategy, "grpColIdx": grp_idx,
           "numGroups": float(num_groups if num_groups is not None else (200 if grp_idx else 1)),
           "groupkeys": ["%s.%s" % (table.name, g) for g in group_by],
           "targetlist": [{"node": "TargetEntry", "expr": e, "resno": i + 1,
                           "resname": n, "resjunk": False}
                          for i, (e, n) in enumerate(targets)],
           "qual": [], "lefttree": scan}
    if not order_by_keys:
        return agg
    sort = {"node": "Sort",
            "sortkeys": ["%s.%s" % (table.name, g) for g in group_by],
            # after set_plan_references() an upper node refers to its child's
            # output columns by OUTER_VAR Vars
            "targetlist": [{"node": "TargetEntry",
                            "expr": {"node": "Var", "varno": "OUTER",
                                     "varattno": t["resno"], "vartype": etype(t["expr"])},
                            "resno": t["resno"],
                            "resname": t["resname"], "resjunk": False}
                           for t in agg["targetlist"]],
            "lefttree": agg}
    return sort


# ---- the regression suite's SQL subset -------------------------------------
_Q = re.compile(
    r"^(?:explain\s*\([^)]*\)\s*)?select\s+(?P<key>key\s*,)?\s*(?P<agg>\w+)\((?P<args>[^)]*)\)"
    r"(?:::(?P<cast>\w+))?\s+from\s+(?P<table>\w+)\s*"
    r"(?P<where>where\s+key\s*=\s*(?P<wkey>\d+))?\s*"
    r"(?P<group>group\s+by\s+key\s+order\s+by\s+key)?\s*;?$", re.I)


def parse_regression_sql(sql):
    m = _Q.match(" ".join(sql.split()))
    if not m:
        raise ValueError("unsupported regression query: %r" % sql)
    return {"agg": m.group("agg").lower(),
            "args": [a.strip() for a in m.group("args").split(",") if a.strip()],
            "cast": m.group("cast"),
            "table": m.group("table"),
            "where_key": int(m.group("wkey")) if m.group("where") else None,
            "group": bool(m.group("group")),
            "show_key": bool(m.group("key"))}


def plan_regression_sql(sql, table):
    """Plan tree for one statement of input/sql/*_agg.sql."""
    q = parse_regression_sql(sql)
    star = (q["args"] == ["*"])
    args = [] if star else [table.col(a) for a in q["args"]]
    aggref = Agg(q["agg"], args, star=star)
    targets = []
    if q["show_key"]:
        targets.append((table.col("key"), "key"))
    targets.append((aggref, q["agg"]))
    where = []
    if q["where_key"] is not None:
        where.append(Op("=", table.col("key"), Const("int4", q["where_key"])))
    return make_agg_plan(table, targets,
                         group_by=["key"] if q["group"] else [],
                         where=where, order_by_keys=q["group"],
                         num_groups=31 if q["group"] else 1)


def dumps(plan):
    return json.dumps(plan, separators=(",", ":"))
