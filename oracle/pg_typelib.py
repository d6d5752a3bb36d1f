"""PostgreSQL's date / time / timestamp operators, text comparison and
float -> numeric casts, restated in python.

TEST INFRASTRUCTURE (oracle) - the checker of the device runtime's
kern_timelib.cuh / kern_textlib.cuh / pgs_float_to_numeric; never on the
product path.

These functions are PostgreSQL core (9.4 era, not vendored in the reference):
utils/adt/date.c (date_pli, date_mii, date_mi, datetime_timestamp,
date2timestamp, timestamp_date, timestamp_time, date_cmp_timestamp ...),
utils/adt/timestamp.c (timestamp2tm, j2date, date2j), utils/adt/varlena.c
(varstr_cmp under the C collation, bpchar: bcTruelen) and utils/adt/numeric.c
(float8_numeric / float4_numeric: "%.15g" / "%.6g" + numeric_in).  The
reference's device versions are /root/reference/opencl_timelib.h:123-664 and
opencl_textlib.h:148-410; where they differ from PostgreSQL (signed byte
compare, isnull left unset in pgfn_timestamp_date) PostgreSQL is followed,
because the oracle is pg_strom.enabled=off.

Parity: no golden vectors exist for these functions in the reference's test
suite (its regression SQL only uses integer / float / numeric columns), so
this part of the oracle is "parity unpinned" against the reference; it is
anchored on the known answers PostgreSQL's documentation publishes
(tests/test_typelib_device_code.py::test_documented_examples: epochs, Julian
day numbers, the date / time operator examples) and on the identity
date2j(j2date(d)) = d.
"""
from decimal import Decimal

from .pg_agg import PgError

DATE_NOBEGIN, DATE_NOEND = -2 ** 31, 2 ** 31 - 1
DT_NOBEGIN, DT_NOEND = -2 ** 63, 2 ** 63 - 1
USECS_PER_DAY = 86400000000
POSTGRES_EPOCH_JDATE = 2451545
INT_MAX = 2 ** 31 - 1


class PgRangeError(PgError):
    """PostgreSQL would ereport(ERROR) (or, for 9.4's unchecked int4
    arithmetic, wrap): the device leaves such rows to the host."""


def date2j(y, m, d):
    if m > 2:
        m += 1
        y += 4800
    else:
        m += 13
        y += 4799
    century = y // 100
    julian = y * 365 - 32167
    julian += y // 4 - century + century // 4
    julian += 7834 * m // 256 + d
    return julian


def j2date(jd):
    julian = jd + 32044
    quad = julian // 146097
    extra = (julian - quad * 146097) * 4 + 3
    julian += 60 + quad * 3 + extra // 146097
    quad = julian // 1461
    julian -= quad * 1461
    y = julian * 4 // 1461
    julian = ((julian + 305) % 365 if y != 0 else (julian + 306) % 366) + 123
    y += quad * 4
    year = y - 4800
    quad = julian * 2141 // 65536
    day = julian - 7834 * quad // 256
    month = (quad + 10) % 12 + 1
    return year, month, day


def _timestamp2tm(ts):
    """-> (year, month, day, usecs of the day) or PgRangeError."""
    date, time = divmod(ts, USECS_PER_DAY)      # floor: time in [0, day)
    date += POSTGRES_EPOCH_JDATE
    if date < 0 or date > INT_MAX:
        raise PgRangeError("timestamp out of range")
    y, m, d = j2date(date)
    return y, m, d, time


def date_timestamp(d):
    if d == DATE_NOBEGIN:
        return DT_NOBEGIN
    if d == DATE_NOEND:
        return DT_NOEND
    r = d * USECS_PER_DAY
    if not (DT_NOBEGIN <= r <= DT_NOEND):
        raise PgRangeError("date out of range for timestamp")
    return r


def timestamp_date(ts):
    if ts == DT_NOBEGIN:
        return DATE_NOBEGIN
    if ts == DT_NOEND:
        return DATE_NOEND
    y, m, d, _ = _timestamp2tm(ts)
    return date2j(y, m, d) - POSTGRES_EPOCH_JDATE


def timestamp_time(ts):
    if ts in (DT_NOBEGIN, DT_NOEND):
        return None
    return _timestamp2tm(ts)[3]


def _int4(v, what):
    if not (-2 ** 31 < v < 2 ** 31 - 1):
        # beyond int4, or exactly on an infinity mark: 9.4 wraps / later
        # versions raise "date out of range"
        raise PgRangeError(what)
    return v


def date_pli(d, n):
    if d in (DATE_NOBEGIN, DATE_NOEND):
        return d
    return _int4(d + n, "date out of range")


def date_mii(d, n):
    if d in (DATE_NOBEGIN, DATE_NOEND):
        return d
    return _int4(d - n, "date out of range")


def date_mi(a, b):
    if a in (DATE_NOBEGIN, DATE_NOEND) or b in (DATE_NOBEGIN, DATE_NOEND):
        raise PgRangeError("cannot subtract infinite dates")
    v = a - b
    if not (-2 ** 31 <= v <= 2 ** 31 - 1):
        raise PgRangeError("integer out of range")
    return v


def datetime_pl(d, t):
    r = date_timestamp(d)
    if r in (DT_NOBEGIN, DT_NOEND):
        return r
    r += t
    if r > DT_NOEND:
        raise PgRangeError("timestamp out of range")
    return r


def date_cmp_timestamp(d, ts):
    a = date_timestamp(d)
    return (a > ts) - (a < ts)


def timestamp_cmp_date(ts, d):
    return -date_cmp_timestamp(d, ts)


# ---- text ----
def text_cmp(a, b):
    """varstr_cmp under the C collation: memcmp, then the shorter first."""
    return (a > b) - (a < b)            # python compares bytes unsigned


def bpchar_cmp(a, b):
    return text_cmp(a.rstrip(b" "), b.rstrip(b" "))


def varlena(payload, short=None):
    """varlena image of a text payload: 1-byte header when it fits (as a heap
    tuple stores it) unless short=False."""
    if short is None:
        short = len(payload) + 1 <= 127
    if short:
        assert len(payload) + 1 <= 127
        return bytes([((len(payload) + 1) << 1) | 1]) + payload
    return ((len(payload) + 4) << 2).to_bytes(4, "little") + payload


# ---- float -> numeric ----
def float_numeric(v, ndigits=15):
    """float8_numeric (ndigits = DBL_DIG) / float4_numeric (FLT_DIG): the
    Decimal and the display scale numeric_in gives the "%.*g" text.  NaN / Inf
    -> None (numeric NaN / error: host only)."""
    if v != v or v in (float("inf"), float("-inf")):
        return None
    text = "%.*g" % (ndigits, v)
    mant, _, exp = text.partition("e")
    exp = int(exp) if exp else 0
    ndec = len(mant.partition(".")[2])
    dscale = max(0, ndec - exp)
    return Decimal(text), dscale


# ---- dispatch by pg_proc name, for oracle/pg_expr.py ----
_CMP = {"eq": lambda c: c == 0, "ne": lambda c: c != 0, "lt": lambda c: c < 0,
        "le": lambda c: c <= 0, "gt": lambda c: c > 0, "ge": lambda c: c >= 0}


def _icmp(a, b):
    return (a > b) - (a < b)


def lookup(name, argtypes):
    """-> python callable for the strict function `name(argtypes)` or None."""
    at = tuple(argtypes)
    if len(at) == 1:
        if name in ("date", "time", "timestamp") and at[0] == name:
            return lambda a: a
        return {("date", "timestamp"): timestamp_date, ("time", "timestamp"): timestamp_time,
                ("timestamp", "date"): date_timestamp}.get((name, at[0]))
    table = {"date_pli": date_pli, "date_mii": date_mii, "date_mi": date_mi,
             "datetime_pl": datetime_pl, "integer_pl_date": lambda n, d: date_pli(d, n),
             "timedate_pl": lambda t, d: datetime_pl(d, t),
             "date_cmp_timestamp": date_cmp_timestamp,
             "timestamp_cmp_date": timestamp_cmp_date,
             "date_cmp": _icmp, "time_cmp": _icmp, "timestamp_cmp": _icmp,
             "bttextcmp": text_cmp, "bpcharcmp": bpchar_cmp}
    if name in table:
        return table[name]
    for sfx, test in _CMP.items():
        for pfx, cmpfn in (("date_%s_timestamp", date_cmp_timestamp),
                           ("timestamp_%s_date", timestamp_cmp_date),
                           ("date_%s", _icmp), ("time_%s", _icmp), ("timestamp_%s", _icmp),
                           ("text_%s", text_cmp), ("text%s", text_cmp),
                           ("bpchar%s", bpchar_cmp)):
            if name == pfx % sfx:
                return lambda a, b, cmpfn=cmpfn, test=test: test(cmpfn(a, b))
    return None
