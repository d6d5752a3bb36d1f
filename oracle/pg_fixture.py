"""Regenerates the reference's regression tables without PostgreSQL.

TEST INFRASTRUCTURE (oracle) - never imported by the product path.

Follows /root/reference/input/sql/agg_init.sql:
  gpupreagg_test           :18-108   (setseed(0) at :43)
  gpupreagg_zero_test      :113-114  (empty)
  gpupreagg_overflow_test  :117-205  (setseed(0) at :141)
  gpupreagg_mix            :218-264  (self-join of the three 10000-row blocks)

PostgreSQL core behaviour restated here (PG <= 11, not vendored in the
reference tree):
  setseed(s)  = srandom((unsigned int)(s * MAX_RANDOM_VALUE))
  random()    = (double) random() / ((double) MAX_RANDOM_VALUE + 1)
  float8 -> int2/int4/int8 cast = rint() (ties to even) + range check
  float8 -> numeric cast        = sprintf("%.15g")
  round(numeric, n)             = half away from zero
  numeric -> float4/float8      = strtod of the decimal text
  ExecTargetList with a set-returning function evaluates every non-SRF
  target expression one extra time after the last row (the SRF reports
  "done" only then), so each INSERT ... SELECT generate_series(..) consumes
  one extra row worth of random() draws.

The columns are (agg_init.sql:18-30):
  id, key, smlint_x, integer_x, bigint_x, real_x, float_x, nume_x,
  smlsrl_x, serial_x, bigsrl_x
Values: python int for integer columns, float for real_x (already rounded to
float4) and float_x, decimal.Decimal for nume_x, None for NULL.
"""
import ctypes
import struct
from decimal import Decimal, ROUND_HALF_UP

COLUMNS = ["id", "key", "smlint_x", "integer_x", "bigint_x", "real_x",
           "float_x", "nume_x", "smlsrl_x", "serial_x", "bigsrl_x"]
# PostgreSQL type of each column (names as in pg_type.typname)
COLTYPES = {"id": "int4", "key": "int4", "smlint_x": "int2",
            "integer_x": "int4", "bigint_x": "int8", "real_x": "float4",
            "float_x": "float8", "nume_x": "numeric", "smlsrl_x": "int2",
            "serial_x": "int4", "bigsrl_x": "int8"}

_libc = ctypes.CDLL("libc.so.6")
_libc.random.restype = ctypes.c_long
_libc.srandom.argtypes = [ctypes.c_uint]


def setseed(s):
    _libc.srandom(ctypes.c_uint(int(s * 0x7FFFFFFF)))


def rnd():
    return _libc.random() / (0x7FFFFFFF + 1.0)


def rint(x):
    """dtoi2/dtoi4/dtoi8: rint() = round half to even."""
    return int(float.__round__(x))


def _nozero_sign(d):
    return -d if (d.is_zero() and d.is_signed()) else d


def f8_numeric(v):
    """float8::numeric goes through "%.15g" (numeric has no negative zero)."""
    return _nozero_sign(Decimal("%.15g" % v))


def nround(d, s):
    return _nozero_sign(d.quantize(Decimal(1).scaleb(-s), rounding=ROUND_HALF_UP))


def f4(x):
    return struct.unpack("f", struct.pack("f", x))[0]


def numeric_f4(d):
    return f4(float(str(d)))


def numeric_f8(d):
    return float(str(d))


def _null_row(i, nkeys=None):
    return {"id": i, "key": None, "smlint_x": None, "integer_x": None,
            "bigint_x": None, "real_x": None, "float_x": None, "nume_x": None,
            "smlsrl_x": 0, "serial_x": 0, "bigsrl_x": 0}


def build_gpupreagg_test():
    rows = []

    def block(id0, n, key0, g):
        for i in range(n + 1):          # +1: the extra evaluation pass
            r = {"id": id0 + i, "key": key0 + (i % 10)}
            r["smlint_x"] = None if rnd() > 0.95 else rint(g(rnd()) * 32767 / 1000)
            r["integer_x"] = None if rnd() > 0.95 else rint(g(rnd()) * 2147483647 / 1000)
            r["bigint_x"] = None if rnd() > 0.95 else rint(g(rnd()) * 9223372036854775807 / 1000)
            r["real_x"] = None if rnd() > 0.95 else numeric_f4(nround(f8_numeric(g(rnd())), 4))
            r["float_x"] = None if rnd() > 0.95 else numeric_f8(nround(f8_numeric(g(rnd())), 13))
            r["nume_x"] = None if rnd() > 0.95 else f8_numeric(g(rnd()))
            r["smlsrl_x"] = rint(g(rnd()) * 32767 / 1000)
            r["serial_x"] = rint(g(rnd()) * 2147483647 / 1000)
            r["bigsrl_x"] = rint(g(rnd()) * 9223372036854775807 / 1000)
            if i < n:
                rows.append(r)

    setseed(0)
    block(1, 10000, 1, lambda u: u)
    # `random()*-32767/1000`: the parser folds -32767 into a constant, so the
    # expression is (u * -32767) / 1000 -- same double as (u*-1)*32767/1000
    # only by sign symmetry of IEEE multiplication, which holds exactly.
    block(10001, 10000, 11, lambda u: u * -1)
    block(20001, 10000, 21, lambda u: u * 2 - 1)
    for i in range(30001, 40001):
        rows.append(_null_row(i))
    return rows


def build_gpupreagg_overflow_test():
    rows = []
    E21 = 1000000000000000000000.0

    def nfloor(v):
        # floor(float8) is float8 floor(); the column is numeric, so the
        # assignment cast is float8 -> numeric ("%.15g")
        import math
        return f8_numeric(math.floor(v))

    def block(id0, n, key0, kind):
        for i in range(n + 1):
            r = {"id": id0 + i, "key": key0 + (i % 10)}
            if kind == "+":
                r["smlint_x"] = None if rnd() > 0.95 else 32767
                r["integer_x"] = None if rnd() > 0.95 else 2147483647
                r["bigint_x"] = None if rnd() > 0.95 else 9223372036854775807
                r["real_x"] = None if rnd() > 0.95 else f4(1.0e38)
                r["float_x"] = None if rnd() > 0.95 else 1.0e308
                r["nume_x"] = None if rnd() > 0.95 else nfloor(rnd() * E21)
                r["smlsrl_x"] = rint(rnd() * 32767)
                r["serial_x"] = rint(rnd() * 2147483647)
                r["bigsrl_x"] = _clamp_i8(rnd() * 9223372036854775807)
            elif kind == "-":
                r["smlint_x"] = None if rnd() > 0.95 else -32768
                r["integer_x"] = None if rnd() > 0.95 else -2147483648
                r["bigint_x"] = None if rnd() > 0.95 else -9223372036854775808
                r["real_x"] = None if rnd() > 0.95 else f4(-1.0e38)
                r["float_x"] = None if rnd() > 0.95 else -1.0e308
                r["nume_x"] = None if rnd() > 0.95 else nfloor(rnd() * E21) * -1
                r["smlsrl_x"] = rint(rnd() * -32767)
                r["serial_x"] = rint(rnd() * -2147483647)
                r["bigsrl_x"] = _clamp_i8(rnd() * -9223372036854775807)
            else:
                r["smlint_x"] = None if rnd() > 0.95 else rint((rnd() * 2 - 1) * 32767)
                r["integer_x"] = None if rnd() > 0.95 else rint((rnd() * 2 - 1) * 2147483647)
                r["bigint_x"] = None if rnd() > 0.95 else _clamp_i8((rnd() * 2 - 1) * 9223372036854775807)
                r["real_x"] = None if rnd() > 0.95 else f4((rnd() * 2 - 1) * 1.0e38)
                r["float_x"] = None if rnd() > 0.95 else (rnd() * 2 - 1) * 1.0e308
                r["nume_x"] = None if rnd() > 0.95 else nfloor((rnd() * 2 - 1) * E21)
                r["smlsrl_x"] = rint((rnd() * 2 - 1) * 32767)
                r["serial_x"] = rint((rnd() * 2 - 1) * 2147483647)
                r["bigsrl_x"] = _clamp_i8((rnd() * 2 - 1) * 9223372036854775807)
            if i < n:
                rows.append(r)

    setseed(0)
    block(1, 10000, 1, "+")
    block(10001, 10000, 11, "-")
    block(20001, 10000, 21, "+-")
    for i in range(30001, 40001):
        rows.append(_null_row(i))
    return rows


def _clamp_i8(v):
    r = rint(v)
    assert -(1 << 63) <= r < (1 << 63), "bigint out of range in fixture"
    return r


def build_gpupreagg_mix(test_rows=None):
    """agg_init.sql:218-264: x = rows with id<=10000, y = keys 11..20 shifted,
    z = keys 21..30 shifted, joined on id."""
    t = test_rows if test_rows is not None else build_gpupreagg_test()
    xs = {r["id"]: r for r in t if r["id"] <= 10000}
    ys = {r["id"] - 10000: r for r in t if r["key"] is not None and 11 <= r["key"] <= 20}
    zs = {r["id"] - 20000: r for r in t if r["key"] is not None and 21 <= r["key"] <= 30}
    out = []
    for i in sorted(xs):
        if i in ys and i in zs:
            x, y, z = xs[i], ys[i], zs[i]
            r = {"id": x["id"], "key": x["key"]}
            for c in COLUMNS[2:]:
                base = c[:-2]
                r[base + "_x"] = x[c]
                r[base + "_y"] = y[c]
                r[base + "_z"] = z[c]
            out.append(r)
    return out


def mix_coltype(col):
    return COLTYPES[col[:-2] + "_x"] if col not in ("id", "key") else "int4"


_cache = {}


def table(name):
    """Returns (rows, coltype-lookup) for a regression table name."""
    if name not in _cache:
        if name == "gpupreagg_test":
            _cache[name] = build_gpupreagg_test()
        elif name == "gpupreagg_zero_test":
            _cache[name] = []
        elif name == "gpupreagg_overflow_test":
            _cache[name] = build_gpupreagg_overflow_test()
        elif name == "gpupreagg_mix":
            _cache[name] = build_gpupreagg_mix(table("gpupreagg_test"))
        else:
            raise KeyError(name)
    return _cache[name]


def coltype(tablename, col):
    if tablename == "gpupreagg_mix":
        return mix_coltype(col)
    return COLTYPES[col]
