"""CPU restatement of PostgreSQL's aggregate executor for the regression suite.

TEST INFRASTRUCTURE (oracle) - never imported by the product path.

This is the "pg_strom.enabled = off" side of the reference's differential
test (/root/reference/input/make_expected.sh:22-28, input/disable.conf:9):
HashAggregate/Agg over SeqScan with PostgreSQL 9.4-era transition and final
functions.  PostgreSQL core is not vendored in the reference tree; the
formulas below restate utils/adt/{numeric,float,int8}.c of that era and are
pinned by the goldens in tests/golden/*.json (see tests/test_oracle_golden.py).

It also restates the reference's *final* aggregates over GpuPreAgg partial
rows (pg_strom--1.0.sql:247-401, gpupreagg.c:4430-4773), which is what the
PostgreSQL Agg node on top of GpuPreAgg computes.
"""
import math
import re
import struct
from decimal import Decimal, ROUND_HALF_UP, getcontext, localcontext

from . import pg_fixture as fx

NBASE = 10000
DEC_DIGITS = 4
NUMERIC_MIN_SIG_DIGITS = 16
NUMERIC_MAX_DISPLAY_SCALE = 1000


class PgError(Exception):
    pass


# ---------------------------------------------------------------- numeric
def dscale(d):
    e = d.as_tuple().exponent
    return -e if e < 0 else 0


# all exact operations (add/mul/quantize) run in one wide default context;
# only division and sqrt narrow it locally.
getcontext().prec = 2500


def _nbase_head(d):
    """(weight, first base-10000 digit) of |d|; (0, 0) for zero.
    numeric.c select_div_scale() looks at the first nonzero NBASE digit."""
    d = abs(d)
    if d == 0:
        return 0, 0
    # weight = floor(log_10000(d))
    adj = d.adjusted()                    # floor(log10(d))
    w = adj // DEC_DIGITS                 # floor division handles negatives
    first = int(d.scaleb(-w * DEC_DIGITS))
    assert 1 <= first < NBASE, (d, w, first)
    return w, first


def select_div_scale(a, b):
    w1, f1 = _nbase_head(a)
    w2, f2 = _nbase_head(b)
    qweight = w1 - w2
    if f1 <= f2:
        qweight -= 1
    rscale = NUMERIC_MIN_SIG_DIGITS - qweight * DEC_DIGITS
    rscale = max(rscale, dscale(a), dscale(b), 0)
    return min(rscale, NUMERIC_MAX_DISPLAY_SCALE)


def nquant(d, scale):
    r = d.quantize(Decimal(1).scaleb(-scale), rounding=ROUND_HALF_UP)
    if r.is_zero() and r.is_signed():
        r = -r                      # numeric has no negative zero
    return r


def _div_round(a, b, rscale):
    """a / b rounded half away from zero at `rscale` decimals (div_var with
    round=true computes the exact quotient digits and then rounds)."""
    if a == 0:
        return nquant(Decimal(0), rscale)
    with localcontext() as ctx:
        # enough digits that the guard digits cannot hide a tie: the quotient
        # of two finite decimals is either exact within this many digits or
        # never lands exactly on a half-unit boundary.
        ctx.prec = max(a.adjusted() - b.adjusted(), 0) + rscale + \
            len(a.as_tuple().digits) + len(b.as_tuple().digits) + 30
        return nquant(a / b, rscale)


def numeric_div(a, b):
    if b == 0:
        raise PgError("division by zero")
    return _div_round(a, b, select_div_scale(a, b))


def numeric_mul(a, b, rscale=None):
    r = a * b
    if rscale is None:
        rscale = dscale(a) + dscale(b)
    return nquant(r, rscale)


def numeric_add(a, b):
    return nquant(a + b, max(dscale(a), dscale(b)))


def numeric_sqrt(a, rscale):
    with localcontext() as ctx:
        ctx.prec = max(a.adjusted(), 0) // 2 + rscale + 40
        return nquant(a.sqrt(), rscale)


def numeric_out(d):
    if d is None:
        return None
    s = format(d, "f")
    return s


def int_numeric(i):
    return Decimal(i)


# ------------------------------------------------------------ float helpers
def f4(x):
    try:
        return struct.unpack("f", struct.pack("f", x))[0]
    except OverflowError:
        return math.copysign(math.inf, x)


def float8_out(v, extra_float_digits=-3):
    if v is None:
        return None
    if math.isnan(v):
        return "NaN"
    if math.isinf(v):
        return "Infinity" if v > 0 else "-Infinity"
    nd = max(1, 15 + extra_float_digits)
    return _fmt_g(v, nd)


def float4_out(v, extra_float_digits=-3):
    if v is None:
        return None
    if math.isnan(v):
        return "NaN"
    if math.isinf(v):
        return "Infinity" if v > 0 else "-Infinity"
    nd = max(1, 6 + extra_float_digits)
    return _fmt_g(v, nd)


def _fmt_g(v, nd):
    s = "%.*g" % (nd, v)
    return s


def check_float8(val, inf_is_valid, zero_is_valid=True):
    """CHECKFLOATVAL (float.c)"""
    if math.isinf(val) and not inf_is_valid:
        raise PgError("value out of range: overflow")
    if val == 0.0 and not zero_is_valid:
        raise PgError("value out of range: underflow")


def float8pl(a, b):
    r = a + b
    check_float8(r, math.isinf(a) or math.isinf(b))
    return r


def float8mul(a, b):
    r = a * b
    check_float8(r, math.isinf(a) or math.isinf(b), a == 0 or b == 0)
    return r


def float4pl(a, b):
    r = f4(a + b)
    if math.isinf(r) and not (math.isinf(a) or math.isinf(b)):
        raise PgError("value out of range: overflow")
    return r


def float8_cmp(a, b):
    """float8_cmp_internal: NaN sorts above everything, NaN == NaN."""
    if math.isnan(a):
        return 0 if math.isnan(b) else 1
    if math.isnan(b):
        return -1
    return (a > b) - (a < b)


# -------------------------------------------------------------------- casts
INT_RANGE = {"int2": (-(1 << 15), (1 << 15) - 1, "smallint out of range"),
             "int4": (-(1 << 31), (1 << 31) - 1, "integer out of range"),
             "int8": (-(1 << 63), (1 << 63) - 1, "bigint out of range")}
SQLTYPE = {"smallint": "int2", "integer": "int4", "bigint": "int8",
           "real": "float4", "float": "float8", "numeric": "numeric",
           "double precision": "float8", "int": "int4"}


def cast(value, src, dst):
    """SQL cast of an aggregate result (overflow_agg.sql `agg(x)::type`)."""
    if value is None or src == dst:
        return value
    if dst in INT_RANGE:
        lo, hi, msg = INT_RANGE[dst]
        if src in INT_RANGE:
            r = value
        elif src in ("float4", "float8"):
            if math.isnan(value) or math.isinf(value):
                raise PgError(msg)
            r = int(float.__round__(value))       # rint()
            # dtoi*: range check happens on the double
            if not (lo <= r <= hi):
                raise PgError(msg)
        elif src == "numeric":
            r = int(nquant(value, 0))             # half away from zero
        else:
            raise NotImplementedError((src, dst))
        if not (lo <= r <= hi):
            raise PgError(msg)
        return r
    if dst == "float8":
        if src in INT_RANGE:
            return float(value)
        if src == "float4":
            return float(value)
        if src == "numeric":
            return float(str(value))
    if dst == "float4":
        if src in INT_RANGE:
            return f4(float(value))
        if src == "float8":
            r = f4(value)
            if math.isinf(r) and not math.isinf(value):
                raise PgError("value out of range: overflow")
            if r == 0.0 and value != 0.0:
                raise PgError("value out of range: underflow")
            return r
        if src == "numeric":
            return f4(float(str(value)))
    if dst == "numeric":
        if src in INT_RANGE:
            return Decimal(value)
        if src == "float8":
            if math.isnan(value):
                return Decimal("NaN")
            if math.isinf(value):
                raise PgError("cannot convert infinity to numeric")
            return Decimal("%.15g" % value)
        if src == "float4":
            if math.isinf(value):
                raise PgError("cannot convert infinity to numeric")
            return Decimal("%.6g" % value)
    raise NotImplementedError((src, dst))


def value_out(v, typ):
    if v is None:
        return None
    if typ in INT_RANGE:
        return str(v)
    if typ == "float8":
        return float8_out(v)
    if typ == "float4":
        return float4_out(v)
    if typ == "numeric":
        return numeric_out(v)
    raise NotImplementedError(typ)


# ------------------------------------------------- PostgreSQL's own aggregates
# Each aggregate: result type by argument type, and a class with
# accum(value...) / final().  NULL inputs are skipped by the executor for
# strict transition functions (all of these are strict).

class _Count:
    rettype = "int8"

    def __init__(self, argtypes):
        self.n = 0

    def accum(self, *a):
        self.n += 1

    def final(self):
        return self.n


class _MinMax:
    def __init__(self, argtypes, is_max):
        self.t = argtypes[0]
        self.rettype = self.t
        self.v = None
        self.is_max = is_max

    def accum(self, x):
        if self.v is None:
            self.v = x
            return
        if self.t in ("float4", "float8"):
            c = float8_cmp(self.v, x)
        else:
            c = (self.v > x) - (self.v < x)
        # float8larger: (cmp(a,b) > 0) ? a : b ; float8smaller: (cmp<0) ? a : b
        if self.is_max:
            self.v = self.v if c > 0 else x
        else:
            self.v = self.v if c < 0 else x

    def final(self):
        return self.v


class _Sum:
    def __init__(self, argtypes):
        self.t = argtypes[0]
        self.rettype = {"int2": "int8", "int4": "int8", "int8": "numeric",
                        "float4": "float4", "float8": "float8",
                        "numeric": "numeric"}[self.t]
        self.v = None

    def accum(self, x):
        if self.v is None:
            if self.t == "int8":
                self.v = Decimal(x)
            else:
                self.v = x
            return
        if self.t in ("int2", "int4"):
            self.v += x
            if not (-(1 << 63) <= self.v < (1 << 63)):
                raise PgError("bigint out of range")
        elif self.t == "int8":
            self.v = self.v + Decimal(x)
        elif self.t == "float4":
            self.v = float4pl(self.v, x)
        elif self.t == "float8":
            self.v = float8pl(self.v, x)
        else:
            self.v = numeric_add(self.v, x)

    def final(self):
        return self.v


class _NumericAggState:
    """numeric.c NumericAggState / do_numeric_accum (9.4)."""

    def __init__(self, calc_x2):
        self.N = 0
        self.sumX = Decimal(0)
        self.sumX2 = Decimal(0)
        self.calc_x2 = calc_x2

    def accum(self, x):
        self.N += 1
        self.sumX = numeric_add(self.sumX, x)
        if self.calc_x2:
            self.sumX2 = numeric_add(self.sumX2, numeric_mul(x, x))


def numeric_avg_final(N, sumX):
    if N == 0:
        return None
    return numeric_div(sumX, Decimal(N))


def numeric_stddev_internal(N, sumX, sumX2, variance, sample):
    """numeric.c numeric_stddev_internal"""
    if N == 0 or (sample and N <= 1):
        return None
    vN = Decimal(N)
    sx = numeric_mul(sumX, sumX, dscale(sumX) * 2)
    sx2 = numeric_mul(vN, sumX2, dscale(sumX2))
    num = numeric_add(sx2, -sx)
    if num <= 0:
        return Decimal(0)
    den = vN * (vN - 1) if sample else vN * vN
    rscale = select_div_scale(num, den)
    v = _div_round(num, den, rscale)
    if not variance:
        v = numeric_sqrt(v, rscale)
    return v


class _Avg:
    def __init__(self, argtypes):
        self.t = argtypes[0]
        if self.t in ("float4", "float8"):
            self.rettype = "float8"
            self.N = 0.0
            self.sumX = 0.0
        else:
            self.rettype = "numeric"
            self.N = 0
            self.sumX = 0 if self.t in ("int2", "int4") else Decimal(0)

    def accum(self, x):
        if self.t in ("float4", "float8"):
            self.N += 1.0
            self.sumX = float8pl(self.sumX, float(x))
        elif self.t in ("int2", "int4"):
            self.N += 1
            self.sumX += x                 # int8 state (int4_avg_accum)
        elif self.t == "int8":
            self.N += 1
            self.sumX = numeric_add(self.sumX, Decimal(x))
        else:
            self.N += 1
            self.sumX = numeric_add(self.sumX, x)

    def final(self):
        if self.t in ("float4", "float8"):
            if self.N == 0.0:
                return None
            return self.sumX / self.N
        if self.N == 0:
            return None
        return numeric_div(Decimal(self.sumX), Decimal(self.N))


def float8_var_final(N, sumX, sumX2, variance, sample):
    """float.c float8_var_samp/pop, float8_stddev_samp/pop (9.4)."""
    if N == 0.0 or (sample and N <= 1.0):
        return None
    numerator = N * sumX2 - sumX * sumX
    check_float8(numerator, math.isinf(sumX2) or math.isinf(sumX))
    if numerator <= 0.0:
        return 0.0
    v = numerator / (N * (N - 1.0)) if sample else numerator / (N * N)
    return v if variance else math.sqrt(v)


class _Var:
    def __init__(self, argtypes, variance, sample):
        self.t = argtypes[0]
        self.variance = variance
        self.sample = sample
        if self.t in ("float4", "float8"):
            self.rettype = "float8"
            self.N = 0.0
            self.sumX = 0.0
            self.sumX2 = 0.0
        else:
            self.rettype = "numeric"
            self.st = _NumericAggState(True)

    def accum(self, x):
        if self.t in ("float4", "float8"):
            x = float(x)
            self.N += 1.0
            self.sumX = float8pl(self.sumX, x)
            self.sumX2 = float8pl(self.sumX2, float8mul(x, x))
        else:
            self.st.accum(Decimal(x) if not isinstance(x, Decimal) else x)

    def final(self):
        if self.t in ("float4", "float8"):
            return float8_var_final(self.N, self.sumX, self.sumX2,
                                    self.variance, self.sample)
        return numeric_stddev_internal(self.st.N, self.st.sumX, self.st.sumX2,
                                       self.variance, self.sample)


def float8_regr_final(kind, N, sumX, sumX2, sumY, sumY2, sumXY):
    """float.c float8_corr / float8_covar_pop / float8_covar_samp (9.4)."""
    if kind == "corr":
        if N < 1.0:
            return None
        numX = N * sumX2 - sumX * sumX
        check_float8(numX, math.isinf(sumX2) or math.isinf(sumX))
        numY = N * sumY2 - sumY * sumY
        check_float8(numY, math.isinf(sumY2) or math.isinf(sumY))
        numXY = N * sumXY - sumX * sumY
        check_float8(numXY, math.isinf(sumXY) or math.isinf(sumX) or math.isinf(sumY))
        if numX <= 0 or numY <= 0:
            return None
        return numXY / math.sqrt(numX * numY)
    if kind == "covar_pop":
        if N < 1.0:
            return None
        num = N * sumXY - sumX * sumY
        check_float8(num, math.isinf(sumXY) or math.isinf(sumX) or math.isinf(sumY))
        return num / (N * N)
    if kind == "covar_samp":
        if N < 2.0:
            return None
        num = N * sumXY - sumX * sumY
        check_float8(num, math.isinf(sumXY) or math.isinf(sumX) or math.isinf(sumY))
        return num / (N * (N - 1.0))
    raise KeyError(kind)


class _Regr:
    rettype = "float8"

    def __init__(self, argtypes, kind):
        self.kind = kind
        self.argtypes = argtypes
        self.s = [0.0] * 6            # N, sumX, sumX2, sumY, sumY2, sumXY

    def accum(self, y, x):
        # SQL: corr(Y, X); float8_regr_accum(transvalues, newvalY, newvalX)
        y = cast(y, self.argtypes[0], "float8")
        x = cast(x, self.argtypes[1], "float8")
        s = self.s
        s[0] += 1.0
        s[1] = float8pl(s[1], x)
        s[2] = float8pl(s[2], float8mul(x, x))
        s[3] = float8pl(s[3], y)
        s[4] = float8pl(s[4], float8mul(y, y))
        s[5] = float8pl(s[5], float8mul(x, y))

    def final(self):
        return float8_regr_final(self.kind, *self.s)


def make_pg_agg(name, argtypes):
    if name == "count":
        return _Count(argtypes)
    if name == "max":
        return _MinMax(argtypes, True)
    if name == "min":
        return _MinMax(argtypes, False)
    if name == "sum":
        return _Sum(argtypes)
    if name == "avg":
        return _Avg(argtypes)
    if name in ("stddev", "stddev_samp"):
        return _Var(argtypes, False, True)
    if name == "stddev_pop":
        return _Var(argtypes, False, False)
    if name in ("variance", "var_samp"):
        return _Var(argtypes, True, True)
    if name == "var_pop":
        return _Var(argtypes, True, False)
    if name in ("corr", "covar_pop", "covar_samp"):
        return _Regr(argtypes, name)
    raise KeyError(name)


# --------------------------------------------------------------- SQL subset
_Q = re.compile(
    r"^select\s+(?P<key>key\s*,)?\s*(?P<agg>\w+)\((?P<args>[^)]*)\)"
    r"(?:::(?P<cast>\w+))?\s+from\s+(?P<table>\w+)\s*"
    r"(?P<where>where\s+key\s*=\s*(?P<wkey>\d+))?\s*"
    r"(?P<group>group\s+by\s+key\s+order\s+by\s+key)?\s*;$", re.I)


def parse_query(sql):
    m = _Q.match(" ".join(sql.split()))
    if not m:
        raise ValueError("unsupported regression query: %r" % sql)
    args = [a.strip() for a in m.group("args").split(",") if a.strip()]
    return {"agg": m.group("agg").lower(), "args": args,
            "cast": SQLTYPE.get(m.group("cast").lower()) if m.group("cast") else None,
            "table": m.group("table"),
            "where_key": int(m.group("wkey")) if m.group("where") else None,
            "group": bool(m.group("group")),
            "show_key": bool(m.group("key"))}


def run_query_pg(sql):
    """Executes one regression statement with PostgreSQL CPU semantics.
    Returns (rows_as_text, error_message)."""
    q = parse_query(sql)
    rows = fx.table(q["table"])
    argtypes = [fx.coltype(q["table"], a) for a in q["args"]]
    groups = {}
    order = []
    try:
        for r in rows:
            if q["where_key"] is not None and r["key"] != q["where_key"]:
                continue
            k = r["key"] if q["group"] else "__all__"
            st = groups.get(k)
            if st is None:
                st = groups[k] = make_pg_agg(q["agg"], argtypes)
                order.append(k)
            vals = [r[a] for a in q["args"]]
            if q["args"] and any(v is None for v in vals):
                continue
            st.accum(*vals)
        if not q["group"] and not groups:
            groups["__all__"] = make_pg_agg(q["agg"], argtypes)
            order.append("__all__")
        out = []
        keys = order
        if q["group"]:
            keys = sorted([k for k in order if k is not None]) + \
                ([None] if None in groups else [])
        for k in keys:
            st = groups[k]
            v = st.final()
            t = st.rettype
            if q["cast"]:
                v = cast(v, t, q["cast"])
                t = q["cast"]
            cell = value_out(v, t)
            if q["show_key"]:
                out.append([None if k is None else str(k), cell])
            else:
                out.append([cell])
        return out, None
    except PgError as e:
        return None, str(e)


# ------------------------------------------ final aggregates over partial rows
# What PostgreSQL's Agg node computes on top of GpuPreAgg
# (gpupreagg.c:134-333 catalog; pg_strom--1.0.sql:247-401).

class FinalAgg:
    """Merges partial rows (dict of partial-column -> value|None) of one
    group.  `spec` = {"agg": name, "argtypes": [...]} and the partial columns
    are delivered positionally as produced by the planner half:
      count        : [nrows]
      min/max      : [pmin|pmax]
      sum          : [psum]
      avg          : [nrows, psum]
      stddev/var.. : [nrows, psum, psum_x2]
      corr/covar.. : [nrows, pcov_x, pcov_x2, pcov_y, pcov_y2, pcov_xy]
    """

    def __init__(self, agg, argtypes):
        self.agg = agg
        self.argtypes = argtypes
        t = argtypes[0] if argtypes else None
        self.t = t
        a = agg
        if a == "count":
            self.rettype = "int8"            # c:sum(int4) -> int8
            self.v = None
        elif a in ("min", "max"):
            self.rettype = t
            self.mm = _MinMax([t], a == "max")
        elif a == "sum":
            if t in ("int2", "int4"):
                self.rettype = "int8"        # pgstrom.sum(int8): state {0,0}
                self.v = 0
                self.nn = 0
            else:
                self.rettype = t
                self.sm = _Sum([t])
        elif a == "avg":
            if t in ("int2", "int4"):
                self.rettype = "numeric"     # pgstrom.avg(int4,int8) -> int8_avg
                self.N = 0
                self.S = 0
            elif t == "int8" or t == "numeric":
                self.rettype = "numeric"     # numeric state, N += nrows
                self.N = 0
                self.S = Decimal(0)
            else:
                self.rettype = "float8"      # pgstrom.avg(int4,float8)
                self.N = 0.0
                self.S = 0.0
        elif a in ("stddev", "stddev_samp", "stddev_pop", "variance",
                   "var_samp", "var_pop"):
            self.rettype = "float8"
            self.s = [0.0, 0.0, 0.0]
        elif a in ("corr", "covar_pop", "covar_samp"):
            self.rettype = "float8"
            self.s = [0.0] * 6
        else:
            raise KeyError(a)

    def accum(self, p):
        a = self.agg
        if a == "count":
            # sum(int4) -> int8 ; strict: NULL partials skipped
            if p[0] is not None:
                self.v = (self.v or 0) + p[0]
        elif a in ("min", "max"):
            if p[0] is not None:
                self.mm.accum(p[0])
        elif a == "sum":
            if self.t in ("int2", "int4"):
                if p[0] is not None:           # sfunc is STRICT
                    self.v += p[0]
                    self.nn += 1
            elif p[0] is not None:
                self.sm.accum(p[0])
        elif a == "avg":
            nrows, psum = p
            if self.t in ("int2", "int4"):
                if nrows is None or psum is None:
                    return                     # STRICT sfunc
                self.N += nrows
                self.S += psum
            elif self.t in ("int8", "numeric"):
                # pgstrom_int8_avg_accum / pgstrom_numeric_avg_accum
                # (gpupreagg.c:4540-4588): int8_avg_accum(psum), which counts
                # one row, then "if (nrows > 0) N += nrows - 1".  For a
                # partial row with nrows = 0 and a non-NULL psum (FILTERed
                # psum defaults to 0; a group split over several partial rows)
                # that leaves N one too high.  The glue's accum function adds
                # exactly nrows instead - same catalog signature.
                if psum is None:
                    return
                self.S = numeric_add(self.S, Decimal(psum))
                self.N += nrows
            else:
                if nrows is None or psum is None:
                    return
                self.N += float(nrows)
                self.S = float8pl(self.S, float(psum))
        elif len(self.s) == 3:
            if any(x is None for x in p):
                return
            self.s[0] += float(p[0])
            self.s[1] = float8pl(self.s[1], p[1])
            self.s[2] = float8pl(self.s[2], p[2])
        else:
            if any(x is None for x in p):
                return
            self.s[0] += float(p[0])
            for i in range(1, 6):
                self.s[i] = float8pl(self.s[i], p[i])

    def final(self):
        a = self.agg
        if a == "count":
            return self.v
        if a in ("min", "max"):
            return self.mm.final()
        if a == "sum":
            if self.t in ("int2", "int4"):
                # the reference returns 0 here for an all-NULL group
                # (SURVEY.md section 2 defects); the fixed final returns NULL.
                return self.v if self.nn > 0 else None
            return self.sm.final()
        if a == "avg":
            if self.t in ("int2", "int4"):
                if self.N == 0:
                    return None
                return numeric_div(Decimal(self.S), Decimal(self.N))
            if self.t in ("int8", "numeric"):
                if self.N == 0:
                    return None
                return numeric_div(self.S, Decimal(self.N))
            if self.N == 0.0:
                return None
            return self.S / self.N
        if len(self.s) == 3:
            variance = a in ("variance", "var_samp", "var_pop")
            sample = a in ("stddev", "stddev_samp", "variance", "var_samp")
            return float8_var_final(self.s[0], self.s[1], self.s[2],
                                    variance, sample)
        # partial columns arrive as (nrows, pcov_x, pcov_x2, pcov_y, pcov_y2,
        # pcov_xy) = float8_regr state order (N, sumX, sumX2, sumY, sumY2, sumXY)
        return float8_regr_final(a, *self.s)
