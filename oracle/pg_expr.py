"""CPU evaluation of expression trees with PostgreSQL semantics.

TEST INFRASTRUCTURE (oracle) - never imported by the product path.

Plays the role of PostgreSQL's ExecEvalExpr / ExecProject for the node kinds
the device code generator accepts (/root/reference/codegen.c:1065-1392) and
of the partial-aggregate placeholder functions the reference installs in
schema pgstrom (/root/reference/gpupreagg.c:4251-4412,
pg_strom--1.0.sql:99-226).  The host uses exactly this evaluation for rows
the device flags StromError_CpuReCheck (gpupreagg.c:2507-2607).
"""
import math
import struct
from decimal import Decimal

from . import pg_typelib
from .pg_agg import PgError, cast as pg_cast, float8pl, float8mul, check_float8, f4
from .pg_agg import numeric_add as pg_numeric_add, numeric_mul as pg_numeric_mul

_INT_RANGE = {"int2": (-(1 << 15), (1 << 15) - 1, "smallint out of range"),
              "int4": (-(1 << 31), (1 << 31) - 1, "integer out of range"),
              "int8": (-(1 << 63), (1 << 63) - 1, "bigint out of range")}


def etype(e):
    n = e["node"]
    return {"Var": lambda: e["vartype"], "Const": lambda: e["consttype"],
            "Param": lambda: e["paramtype"], "FuncExpr": lambda: e["funcresulttype"],
            "OpExpr": lambda: e.get("opresulttype", "bool"),
            "DistinctExpr": lambda: "bool",
            "NullTest": lambda: "bool", "BooleanTest": lambda: "bool",
            "BoolExpr": lambda: "bool", "RelabelType": lambda: e["resulttype"],
            "CaseExpr": lambda: e["casetype"], "Aggref": lambda: e["aggtype"]}[n]()


def _const_value(e):
    if e.get("constisnull"):
        return None
    t = e["consttype"]
    v = e.get("constvalue")
    if t == "bool":
        return v in ("t", "true", True)
    if t in _INT_RANGE or t in ("date", "time", "timestamp"):
        return int(v)
    if t in ("text", "bpchar"):
        return v.encode() if isinstance(v, str) else bytes(v)
    if t in ("float4", "float8"):
        x = float({"NaN": "nan", "Infinity": "inf", "-Infinity": "-inf"}.get(v, v))
        return f4(x) if t == "float4" else x
    if t == "numeric":
        return Decimal(v)
    return v


def _int_arith(op, a, b, rtype):
    lo, hi, msg = _INT_RANGE[rtype]
    if op == "pl":
        r = a + b
    elif op == "mi":
        r = a - b
    elif op == "mul":
        r = a * b
    elif op == "div":
        if b == 0:
            raise PgError("division by zero")
        r = abs(a) // abs(b) * (1 if (a >= 0) == (b >= 0) else -1)   # trunc toward 0
    elif op == "mod":
        if b == 0:
            raise PgError("division by zero")
        r = 0 if b == -1 else int(math.fmod(a, b)) if abs(a) < 2**53 else a - b * int(Decimal(a) / Decimal(b))
    else:
        raise KeyError(op)
    if not (lo <= r <= hi):
        raise PgError(msg)
    return r


def _float_arith(op, a, b, rtype):
    if rtype == "float4":
        a, b = f4(a), f4(b)
    if op == "pl":
        r = a + b
        ok_inf, ok_zero = (math.isinf(a) or math.isinf(b)), True
    elif op == "mi":
        r = a - b
        ok_inf, ok_zero = (math.isinf(a) or math.isinf(b)), True
    elif op == "mul":
        r = a * b
        ok_inf, ok_zero = (math.isinf(a) or math.isinf(b)), (a == 0 or b == 0)
    elif op == "div":
        if b == 0.0:
            raise PgError("division by zero")
        r = a / b
        ok_inf, ok_zero = (math.isinf(a) or math.isinf(b)), (a == 0)
    else:
        raise KeyError(op)
    if rtype == "float4":
        r = f4(r)
    check_float8(r, ok_inf, ok_zero)
    return r


def _fcmp(a, b):
    if isinstance(a, float) or isinstance(b, float):
        a, b = float(a), float(b)
        if math.isnan(a):
            return 0 if math.isnan(b) else 1
        if math.isnan(b):
            return -1
    return (a > b) - (a < b)


_CMP = {"eq": lambda c: c == 0, "ne": lambda c: c != 0, "lt": lambda c: c < 0,
        "le": lambda c: c <= 0, "gt": lambda c: c > 0, "ge": lambda c: c >= 0}


# float.c: dsqrt, dexp, dlog1, dlog10, dcbrt, dceil ..., dacos ... (9.4 era)
def _domain(check, msg, fn, inf_ok=None, zero_ok=None):
    def f(x):
        if not math.isnan(x) and check(x):
            raise PgError(msg)
        try:
            r = fn(x)
        except OverflowError:
            r = math.inf
        except ValueError:
            r = math.nan
        if inf_ok is not None:
            check_float8(r, inf_ok(x), zero_ok(x))
        return r
    return f


def _dpow(a, b):
    fl = b if (math.isinf(b) or math.isnan(b)) else float(math.floor(b))    # C floor()
    if (a == 0.0 and b < 0.0) or (a < 0.0 and (fl != b)):
        raise PgError("invalid argument for power function")
    try:
        r = math.pow(a, b)
    except OverflowError:
        # |result| too large; odd negative powers of a negative base are negative
        neg = a < 0.0 and not math.isinf(b) and int(b) % 2 == 1
        r = -math.inf if neg else math.inf
    check_float8(r, math.isinf(a) or math.isinf(b), a == 0.0)
    return r


def _scale(factor):
    def f(x):
        r = x * factor
        check_float8(r, math.isinf(x), x == 0.0)
        return r
    return f


def _roundf(fn):
    return lambda x: x if (math.isnan(x) or math.isinf(x)) else float(fn(x))


_never = lambda x: False
_MATH = {
    "sqrt": _domain(lambda x: x < 0, "cannot take square root of a negative number", math.sqrt),
    "exp": _domain(_never, "", math.exp, math.isinf, _never),
    "ln": _domain(lambda x: x <= 0, "cannot take logarithm of zero or a negative number", math.log),
    "log": _domain(lambda x: x <= 0, "cannot take logarithm of zero or a negative number", math.log10),
    "cbrt": _domain(_never, "", lambda x: math.copysign(abs(x) ** (1.0 / 3.0), x)),
    "acos": _domain(lambda x: abs(x) > 1, "input is out of range", math.acos),
    "asin": _domain(lambda x: abs(x) > 1, "input is out of range", math.asin),
    "atan": math.atan, "atan2": math.atan2,
    "cos": _domain(math.isinf, "input is out of range", math.cos),
    "sin": _domain(math.isinf, "input is out of range", math.sin),
    "tan": _domain(math.isinf, "input is out of range", math.tan),
    "ceil": _roundf(math.ceil), "floor": _roundf(math.floor), "trunc": _roundf(math.trunc),
    "round": _roundf(lambda x: float.__round__(x)),      # rint(): half to even
    "sign": lambda x: (x > 0) - (x < 0) + 0.0,
    "degrees": _scale(180.0 / math.pi), "radians": _scale(math.pi / 180.0),
    "power": _dpow, "pow": _dpow, "dpow": _dpow, "pi": lambda: math.pi,
}
for _a, _b in (("dsqrt", "sqrt"), ("dexp", "exp"), ("dlog1", "ln"), ("dlog10", "log"),
               ("dcbrt", "cbrt"), ("ceiling", "ceil"), ("dround", "round"), ("dtrunc", "trunc")):
    _MATH[_a] = _MATH[_b]


def _call(name, argtypes, rettype, args):
    """strict functions: NULL in -> NULL out (handled by the caller)."""
    fn = pg_typelib.lookup(name, argtypes)
    if fn is not None:
        return fn(*args)
    # casts: function named after the target type
    if name in ("int2", "int4", "int8", "float4", "float8", "numeric") and len(args) == 1:
        return pg_cast(args[0], argtypes[0], name)
    for sfx in ("pl", "mi", "mul", "div", "mod"):
        if name.endswith(sfx) and name[:-len(sfx)].rstrip("0123456789") in ("int", "float"):
            if name.startswith("int"):
                return _int_arith(sfx, args[0], args[1], rettype)
            return _float_arith(sfx, args[0], args[1], rettype)
    for sfx, fn in _CMP.items():
        if name.endswith(sfx) and (name.startswith("int") or name.startswith("float")
                                   or name.startswith("bool") or name.startswith("date_")
                                   or name.startswith("numeric_")):
            return fn(_fcmp(args[0], args[1]))
    if name.endswith("um"):
        r = -args[0]
        if rettype in _INT_RANGE and not (_INT_RANGE[rettype][0] <= r <= _INT_RANGE[rettype][1]):
            raise PgError(_INT_RANGE[rettype][2])
        return r
    if name.endswith("abs") or name == "abs":
        r = abs(args[0])
        if rettype in _INT_RANGE and r > _INT_RANGE[rettype][1]:
            raise PgError(_INT_RANGE[rettype][2])       # int2abs / int4abs / int8abs
        return r
    if name in _MATH:
        return _MATH[name](*args)
    # numeric.c: add / sub keep the larger display scale, mul adds them up
    if name == "numeric_add":
        return pg_numeric_add(args[0], args[1])
    if name == "numeric_sub":
        return pg_numeric_add(args[0], -args[1])
    if name == "numeric_mul":
        return pg_numeric_mul(args[0], args[1])
    if name == "numeric_uminus":
        return -args[0]
    if name == "numeric_uplus":
        return args[0]
    if name == "numeric_abs":
        return abs(args[0])
    raise NotImplementedError("oracle: function %s(%s)" % (name, ",".join(argtypes)))


def evaluate(e, row):
    """row: list/tuple of column values (None = NULL), Var.varattno is 1-based."""
    if e is None:
        return None
    n = e["node"]
    if n == "Var":
        return row[e["varattno"] - 1]
    if n == "Const":
        return _const_value(e)
    if n == "Param":
        return None if e.get("isnull") else _const_value(
            {"consttype": e["paramtype"], "constvalue": e.get("value")})
    if n == "RelabelType":
        return evaluate(e["arg"], row)
    if n == "NullTest":
        v = evaluate(e["arg"], row)
        return (v is None) if e["nulltesttype"] == "IS_NULL" else (v is not None)
    if n == "BooleanTest":
        v = evaluate(e["arg"], row)
        t = e["booltesttype"]
        return {"IS_TRUE": v is True, "IS_NOT_TRUE": v is not True,
                "IS_FALSE": v is False, "IS_NOT_FALSE": v is not False,
                "IS_UNKNOWN": v is None, "IS_NOT_UNKNOWN": v is not None}[t]
    if n == "BoolExpr":
        op = e["boolop"]
        vals = [evaluate(a, row) for a in e["args"]]
        if op == "NOT":
            return None if vals[0] is None else (not vals[0])
        if op == "AND":
            if any(v is False for v in vals):
                return False
            return None if any(v is None for v in vals) else True
        if any(v is True for v in vals):
            return True
        return None if any(v is None for v in vals) else False
    if n == "CaseExpr":
        base = evaluate(e["arg"], row) if e.get("arg") else None
        for w in e["args"]:
            if e.get("arg"):
                c = evaluate(w["expr"], row)
                hit = (base is not None and c is not None and _fcmp(base, c) == 0)
            else:
                hit = evaluate(w["expr"], row) is True
            if hit:
                return evaluate(w["result"], row)
        return evaluate(e.get("defresult"), row)
    if n == "DistinctExpr":
        # execQual.c ExecEvalDistinct: never NULL; NULL is not distinct from NULL
        a, b = [evaluate(x, row) for x in e["args"]]
        if a is None or b is None:
            return (a is None) != (b is None)
        return not _call(e["opfuncname"], [etype(x) for x in e["args"]], "bool", [a, b])
    if n in ("FuncExpr", "OpExpr"):
        name = e["funcname"] if n == "FuncExpr" else e["opfuncname"]
        args = [evaluate(a, row) for a in e.get("args", [])]
        argtypes = [etype(a) for a in e.get("args", [])]
        if n == "FuncExpr" and e.get("funcschema") == "pgstrom":
            return _partial_placeholder(name, args)
        if any(a is None for a in args):
            return None
        return _call(name, argtypes, etype(e), args)
    raise NotImplementedError(n)


def _partial_placeholder(name, args):
    """gpupreagg.c:4251-4412 (with the documented fix: pcov_* read the value
    arguments 1/2, not the filter)."""
    if name == "nrows":
        return 1 if all(a is True for a in args) else 0
    if name in ("pmin", "pmax", "psum"):
        return args[0]
    if name == "psum_x2":
        if args[0] is None:
            return None
        if isinstance(args[0], Decimal):
            return args[0] * args[0]
        return float8mul(args[0], args[0])
    if name.startswith("pcov_"):
        flt, x, y = args
        if flt is not True or x is None or y is None:
            return None
        return {"pcov_x": lambda: x, "pcov_y": lambda: y,
                "pcov_x2": lambda: float8mul(x, x), "pcov_y2": lambda: float8mul(y, y),
                "pcov_xy": lambda: float8mul(x, y)}[name]()
    raise NotImplementedError(name)
