"""CPU: the date/time, text and float->numeric device runtime
(pg_strom_b200/csrc/kern_timelib.cuh, kern_textlib.cuh, pgs_float_to_numeric
in kern_numeric.cuh - the counterparts of the reference's opencl_timelib.h,
opencl_textlib.h and the float casts of its numeric catalogue,
codegen.c:519-629) compiled with g++ through a shim and checked against the
oracle's restatement of PostgreSQL's functions (oracle/pg_typelib.py); plus:
queries that use them plan as GpuPreAgg and their generated device programs
compile for sm_100a (NVRTC, no GPU needed)."""
import ctypes as C
import os
import random
import struct
import subprocess
from decimal import Decimal

import pytest

from oracle import pg_typelib as T
from pg_strom_b200 import gpupreagg as gp
from pg_strom_b200 import pgplan as P

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CPU_RECHECK = 2
GUCS = {"pg_strom.enabled": "on", "pg_strom.debug_force_gpupreagg": "on"}


@pytest.fixture(scope="module")
def shim(lib):
    out = os.path.join(HERE, "native", "_typelib_shim.so")
    src = os.path.join(HERE, "native", "typelib_host_shim.cpp")
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared",
                    "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(ROOT, "pg_strom_b200", "csrc"),
                    "-o", out, src], check=True)
    so = C.CDLL(out)
    so.shim_time_cast.argtypes = [C.c_int, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int)]
    so.shim_time_binop.argtypes = [C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_int64),
                                   C.POINTER(C.c_int)]
    so.shim_text_op.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_char_p,
                                C.POINTER(C.c_int), C.POINTER(C.c_int)]
    so.shim_float_numeric.argtypes = [C.c_double, C.c_int, C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_int)]
    so.shim_text_keybits.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_uint64),
                                     C.POINTER(C.c_int)]
    so.shim_keyheap_set.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_uint64,
                                    C.c_void_p, C.c_uint]
    so.shim_keyheap_set.restype = None
    return so


def _expect(fn, *args):
    """-> ("ok", value|None) or ("recheck", None)."""
    try:
        return "ok", fn(*args)
    except T.PgRangeError:
        return "recheck", None


def _check(err, isnull, out, exp, what):
    kind, val = exp
    if kind == "recheck":
        assert err == CPU_RECHECK and isnull, what
    else:
        assert err == 0, what
        if val is None:
            assert isnull, what
        else:
            assert not isnull and out == val, (what, out, val)


DATES = [T.DATE_NOBEGIN, T.DATE_NOEND, T.DATE_NOBEGIN + 1, T.DATE_NOEND - 1, 0, 1, -1,
         7305, -730119, 106751991, 106751992, -106751991, -106751992, -2451545, -2451546,
         2 ** 31 - 2451545 - 1]
STAMPS = [T.DT_NOBEGIN, T.DT_NOEND, T.DT_NOBEGIN + 1, T.DT_NOEND - 1, 0, 1, -1,
          T.USECS_PER_DAY, T.USECS_PER_DAY - 1, -T.USECS_PER_DAY, -T.USECS_PER_DAY - 1,
          631152000000000, -211813488000000000, -211813488000000001,
          -2451545 * T.USECS_PER_DAY, -2451545 * T.USECS_PER_DAY - 1]


def test_julian_identity():
    """The device turns a timestamp into a date with one floor division; the
    oracle goes through j2date / date2j like PostgreSQL: same thing."""
    rng = random.Random(1)
    for jd in [0, 1, 59, 60, 61, 365, 366, 1721060, 2451545, T.INT_MAX - 1, T.INT_MAX] + \
            [rng.randrange(0, T.INT_MAX) for _ in range(20000)]:
        assert T.date2j(*T.j2date(jd)) == jd


def test_dates_against_python_datetime(shim):
    """oracle/pg_typelib.py has no golden vectors in the reference (its
    regression SQL has no date column); beyond the documented examples below it
    is pinned on an independent implementation of the proleptic Gregorian
    calendar, Python's datetime: EVERY day from 0001-01-01 to 9999-12-31
    (3.65 M days) through j2date / date2j, and the device's casts (timestamp ->
    date / time, date -> timestamp; CPU build of kern_timelib.cuh) on every
    37th day at a random time of day."""
    import datetime
    first = datetime.date(1, 1, 1)
    jd0 = T.date2j(1, 1, 1)
    assert jd0 == first.toordinal() + 1721425       # JDN of 0001-01-01 is 1721426
    n = datetime.date(9999, 12, 31).toordinal() - first.toordinal() + 1
    y, m, d = 1, 1, 1
    mdays = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
    for i in range(n):
        # (walking the calendar by hand is 10x faster than date.fromordinal and
        # just as independent; every year boundary is checked against datetime)
        assert T.j2date(jd0 + i) == (y, m, d), (i, y, m, d)
        if m == 1 and d == 1:
            assert datetime.date(y, 1, 1).toordinal() - first.toordinal() == i
            assert T.date2j(y, 1, 1) == jd0 + i
        leap = (y % 4 == 0 and (y % 100 != 0 or y % 400 == 0))
        d += 1
        if d > mdays[m - 1] + (1 if (m == 2 and leap) else 0):
            d = 1
            m += 1
            if m > 12:
                m = 1
                y += 1
    assert (y, m, d) == (10000, 1, 1)
    rng = random.Random(7)
    out, isnull = C.c_int64(), C.c_int()
    epoch = datetime.datetime(2000, 1, 1)
    for i in range(0, n, 37):
        day = first + datetime.timedelta(days=i)
        usec = rng.randrange(0, T.USECS_PER_DAY)
        pgdate = day.toordinal() - datetime.date(2000, 1, 1).toordinal()
        ts = pgdate * T.USECS_PER_DAY + usec
        if 1 < day.year < 9999:
            dt = epoch + datetime.timedelta(microseconds=ts)
            assert (dt.year, dt.month, dt.day) == (day.year, day.month, day.day)
        assert T.timestamp_date(ts) == pgdate and T.timestamp_time(ts) == usec
        assert shim.shim_time_cast(0, ts, C.byref(out), C.byref(isnull)) == 0
        assert not isnull.value and out.value == pgdate
        assert shim.shim_time_cast(1, ts, C.byref(out), C.byref(isnull)) == 0 and out.value == usec


def test_documented_examples(shim):
    """Known answers published in PostgreSQL's documentation (Date/Time
    Operators table and the datatype-datetime chapter): the epoch of date and
    timestamp is 2000-01-01, date '2001-09-28' + 7 = date '2001-10-05',
    date '2001-10-01' - date '2001-09-28' = 3, date '2001-09-28' + time
    '03:00' = timestamp '2001-09-28 03:00:00', date '2001-09-28' - 7 =
    date '2001-09-21'; Julian day 0 is 4714-11-24 BC (year -4713)."""
    def d(y, m, dd):
        return T.date2j(y, m, dd) - T.POSTGRES_EPOCH_JDATE
    assert d(2000, 1, 1) == 0 and T.date2j(2000, 1, 1) == 2451545
    assert T.j2date(0) == (-4713, 11, 24)
    assert T.date2j(1970, 1, 1) == 2440588          # UNIX_EPOCH_JDATE
    out, isnull = C.c_int64(), C.c_int()
    hour = 3600 * 1000000
    cases = [(0, d(2001, 9, 28), 7, d(2001, 10, 5)),                # date + integer
             (1, d(2001, 9, 28), 7, d(2001, 9, 21)),                # date - integer
             (2, d(2001, 10, 1), d(2001, 9, 28), 3),                # date - date
             (3, d(2001, 9, 28), 3 * hour, d(2001, 9, 28) * T.USECS_PER_DAY + 3 * hour),
             (4, 7, d(2001, 9, 28), d(2001, 10, 5))]                # integer + date
    for fn, a, b, want in cases:
        assert shim.shim_time_binop(fn, a, b, C.byref(out), C.byref(isnull)) == 0
        assert not isnull.value and out.value == want, (fn, a, b, out.value, want)
    # timestamp '2001-09-28 03:00:00' :: date / :: time
    ts = d(2001, 9, 28) * T.USECS_PER_DAY + 3 * hour
    assert shim.shim_time_cast(0, ts, C.byref(out), C.byref(isnull)) == 0
    assert out.value == d(2001, 9, 28) and T.j2date(out.value + T.POSTGRES_EPOCH_JDATE) == (2001, 9, 28)
    assert shim.shim_time_cast(1, ts, C.byref(out), C.byref(isnull)) == 0 and out.value == 3 * hour
    # a timestamp before the epoch: 1999-12-31 23:00 is day -1, 23:00
    assert shim.shim_time_cast(0, -hour, C.byref(out), C.byref(isnull)) == 0 and out.value == -1
    assert shim.shim_time_cast(1, -hour, C.byref(out), C.byref(isnull)) == 0 and out.value == 23 * hour


def test_time_casts(shim):
    rng = random.Random(2)
    out, isnull = C.c_int64(), C.c_int()
    stamps = STAMPS + [rng.randrange(-2 ** 63, 2 ** 63) for _ in range(3000)] + \
        [rng.randrange(-10 ** 17, 10 ** 17) for _ in range(3000)]
    for ts in stamps:
        err = shim.shim_time_cast(0, ts, C.byref(out), C.byref(isnull))
        _check(err, isnull.value, out.value, _expect(T.timestamp_date, ts), ("timestamp_date", ts))
        err = shim.shim_time_cast(1, ts, C.byref(out), C.byref(isnull))
        _check(err, isnull.value, out.value, _expect(T.timestamp_time, ts), ("timestamp_time", ts))
    dates = DATES + [rng.randrange(-2 ** 31, 2 ** 31) for _ in range(3000)] + \
        [rng.randrange(-10 ** 6, 10 ** 6) for _ in range(1000)]
    for d in dates:
        err = shim.shim_time_cast(2, d, C.byref(out), C.byref(isnull))
        _check(err, isnull.value, out.value, _expect(T.date_timestamp, d), ("date_timestamp", d))


def test_time_operators(shim):
    rng = random.Random(3)
    out, isnull = C.c_int64(), C.c_int()
    ints = [0, 1, -1, 2 ** 31 - 1, -2 ** 31, 365, -365]
    times = [0, 1, T.USECS_PER_DAY - 1, 43200000000]

    def pick(pool, lo, hi):
        return rng.choice(pool) if rng.random() < 0.4 else rng.randrange(lo, hi)

    cmp_ops = {"eq": lambda c: c == 0, "ne": lambda c: c != 0, "lt": lambda c: c < 0,
               "le": lambda c: c <= 0, "gt": lambda c: c > 0, "ge": lambda c: c >= 0}
    for _ in range(6000):
        d = pick(DATES, -2 ** 31, 2 ** 31)
        d2 = pick(DATES, -2 ** 31, 2 ** 31)
        n = pick(ints, -2 ** 31, 2 ** 31)
        t = pick(times, 0, T.USECS_PER_DAY)
        ts = pick(STAMPS, -2 ** 63, 2 ** 63)
        if rng.random() < 0.3:      # a timestamp right at a date: equality cases
            try:
                ts = T.date_timestamp(d)
            except T.PgRangeError:
                pass
        cases = [(0, d, n, T.date_pli, (d, n)), (1, d, n, T.date_mii, (d, n)),
                 (2, d, d2, T.date_mi, (d, d2)), (3, d, t, T.datetime_pl, (d, t)),
                 (4, n, d, T.date_pli, (d, n)), (5, t, d, T.datetime_pl, (d, t)),
                 (6, d, ts, T.date_cmp_timestamp, (d, ts)),
                 (7, ts, d, T.timestamp_cmp_date, (ts, d))]
        for i, (name, test) in enumerate(cmp_ops.items()):
            cases.append((10 + i, d, ts,
                          (lambda a, b, test=test: int(test(T.date_cmp_timestamp(a, b)))), (d, ts)))
            cases.append((20 + i, ts, d,
                          (lambda a, b, test=test: int(test(T.timestamp_cmp_date(a, b)))), (ts, d)))
        for fn, a, b, ofn, oargs in cases:
            err = shim.shim_time_binop(fn, a, b, C.byref(out), C.byref(isnull))
            _check(err, isnull.value, out.value, _expect(ofn, *oargs), (fn, a, b))


def test_text_compare(shim):
    rng = random.Random(4)
    out, isnull = C.c_int(), C.c_int()
    alphabet = [b"a", b"b", b"A", b" ", b"z", b"\x7f", b"\x80", b"\xc3\xa9", b"\xff", b"0"]

    def rand_text():
        s = b"".join(rng.choice(alphabet) for _ in range(rng.choice([0, 1, 2, 3, 5, 8, 130])))
        if rng.random() < 0.4:
            s += b" " * rng.randrange(0, 4)
        return s

    ops = [lambda c: c == 0, lambda c: c != 0, lambda c: c < 0, lambda c: c <= 0,
           lambda c: c > 0, lambda c: c >= 0]
    for _ in range(4000):
        a = rand_text()
        b = rand_text() if rng.random() < 0.6 else a + rng.choice([b"", b" ", b"  ", b"a"])
        ia = T.varlena(a, short=None if rng.random() < 0.5 else False)
        ib = T.varlena(b, short=None if rng.random() < 0.5 else False)
        for bp, cmpfn in ((0, T.text_cmp), (1, T.bpchar_cmp)):
            c = cmpfn(a, b)
            for fn in range(6):
                err = shim.shim_text_op(fn, bp, ia, ib, C.byref(out), C.byref(isnull))
                assert err == 0 and not isnull.value
                assert bool(out.value) == ops[fn](c), (a, b, bp, fn)
            err = shim.shim_text_op(6, bp, ia, ib, C.byref(out), C.byref(isnull))
            assert err == 0 and not isnull.value and out.value == c, (a, b, bp)
    # compressed (4-byte header, bit 1) and out-of-line (0x01) datums: host only
    ok = T.varlena(b"abc")
    for bad in (bytes([0x02 | (20 << 2), 0, 0, 0]) + b"x" * 16, bytes([0x01, 18]) + b"p" * 16):
        err = shim.shim_text_op(0, 0, bad, ok, C.byref(out), C.byref(isnull))
        assert err == CPU_RECHECK and isnull.value


def _unpack_numeric(v):
    exp = v >> 58
    if exp >= 32:
        exp -= 64
    mant = v & ((1 << 57) - 1)
    d = Decimal(mant).scaleb(exp)
    return (-d if (v >> 57) & 1 else d), -exp


def test_float_to_numeric(shim):
    rng = random.Random(5)
    out, isnull = C.c_uint64(), C.c_int()
    f4 = lambda x: struct.unpack("f", struct.pack("f", x))[0]
    vals = [0.0, -0.0, 1.0, -1.0, 0.1, 0.5, 1.5, 2.5, 1e15, 1e16, 123456789012345.0,
            1234567890123456.0, 9.999999999999995e-5, 0.30000000000000004, 1e-5, 1e-17,
            1e-18, 1e-19, 1e-30, 5e-324, 1e22, 1e23, 1.5e17, 1e38, 1.7e38, 1e39, 1e300,
            99999999999999.95, 999999999999999.5, 0.1 + 0.2, 2.0 ** 53, 2.0 ** 57, 2.0 ** 60,
            float("nan"), float("inf"), float("-inf"), 8.5, 0.000123456789012345678]
    for _ in range(20000):
        kind = rng.random()
        if kind < 0.3:
            vals.append(rng.uniform(-1000, 1000))
        elif kind < 0.5:
            vals.append(rng.randrange(-10 ** 9, 10 ** 9) / 10.0 ** rng.randrange(0, 12))
        elif kind < 0.7:
            vals.append(struct.unpack("d", struct.pack("Q", rng.randrange(0, 2 ** 64)))[0])
        else:
            vals.append(rng.uniform(-1, 1) * 10.0 ** rng.randrange(-25, 45))
    handled = 0
    for v in vals:
        for is_f4, nd in ((0, 15), (1, 6)):
            x = f4(v) if is_f4 else v
            if is_f4 and (x != x or abs(x) == float("inf")) and v == v and abs(v) != float("inf"):
                continue                    # does not fit a float4: not this function's input
            err = shim.shim_float_numeric(x, is_f4, C.byref(out), C.byref(isnull))
            exp = T.float_numeric(x, nd)
            if err == CPU_RECHECK:
                assert isnull.value
                # the device may only decline what is outside its ranges
                if exp is not None:
                    d, dscale = exp
                    mant = abs(int(d.scaleb(dscale)))
                    small = x != 0 and abs(x) < (1e-17 if not is_f4 else 1e-26)
                    assert mant > (1 << 57) - 2 or dscale > 32 or small or abs(x) >= 2.0 ** 127, x
                continue
            assert err == 0 and not isnull.value, x
            assert exp is not None, x
            got, gscale = _unpack_numeric(out.value)
            assert got == exp[0] and gscale == exp[1], (x, is_f4, got, gscale, exp)
            handled += 1
    assert handled > 20000


# ---- planner + NVRTC ------------------------------------------------------
EVENTS = P.Table("events", [("d", "date"), ("ts", "timestamp"), ("tm", "time"),
                            ("s", "text"), ("c", "bpchar"), ("k", "int4"), ("v", "int8"),
                            ("f", "float8")])


def typelib_queries():
    """(name, plan) of queries exercising the date/time, text and
    float->numeric catalogue entries in WHERE, FILTER-less aggregate arguments
    and CASE."""
    t = EVENTS
    d, ts, tm, s, c, k, v, f = (t.col(n) for n in ("d", "ts", "tm", "s", "c", "k", "v", "f"))
    cnt = (P.Agg("count", star=True), "count")
    sumv = (P.Agg("sum", [P.Cast(v, "numeric")]), "sum")
    q = []
    q.append(("date_arith", P.make_agg_plan(
        t, [(k, "k"), cnt, (P.Agg("max", [P.Op("-", d, P.Const("date", 7000))]), "max")],
        group_by=["k"], num_groups=16,
        where=[P.Op(">", P.Op("+", d, P.Const("int4", 30)), P.Const("date", 7400)),
               P.Op("<=", P.Op("-", d, P.Const("int4", 5)), P.Const("date", 7700))])))
    q.append(("date_vs_timestamp", P.make_agg_plan(
        t, [cnt, (P.Agg("min", [v]), "min")],
        where=[P.Op("<", d, ts), P.Op(">=", ts, P.Const("date", 7300)),
               P.Op("<>", P.Cast(ts, "date"), d)])))
    q.append(("timestamp_parts", P.make_agg_plan(
        t, [(k, "k"), cnt], group_by=["k"], num_groups=16,
        where=[P.Op("<", P.Cast(ts, "time"), tm),
               P.Op(">", P.Op("+", d, tm), ts),
               P.Op("=", P.Cast(d, "timestamp"), P.Op("+", P.Cast(ts, "date"),
                                                      P.Const("time", 0)))])))
    q.append(("text_eq", P.make_agg_plan(
        t, [(k, "k"), cnt, (P.Agg("sum", [f]), "sum")], group_by=["k"], num_groups=16,
        where=[P.Op("=", s, P.Const("text", "bbb"))])))
    q.append(("text_order", P.make_agg_plan(
        t, [cnt, (P.Agg("max", [v]), "max")],
        where=[P.Op(">=", s, P.Const("text", "ccc"), collation="C"),
               P.Op("<", s, P.Const("text", "xx"), collation="C"),
               P.Op("<>", c, P.Const("bpchar", "ab   "))])))
    q.append(("bpchar_eq", P.make_agg_plan(
        t, [(k, "k"), cnt], group_by=["k"], num_groups=16,
        where=[P.Op("=", c, P.Const("bpchar", "ab"))])))
    q.append(("text_isnull_case", P.make_agg_plan(
        t, [cnt, (P.Agg("avg", [P.Case([(P.Op("=", s, P.Const("text", "aaa")), v)],
                                       P.Const("int8", 0), "int8")]), "avg")],
        where=[P.IsNull(s, notnull=True)])))
    q.append(("float_numeric", P.make_agg_plan(
        t, [(k, "k"), (P.Agg("sum", [P.Cast(f, "numeric")]), "sum"),
            (P.Agg("max", [P.Cast(f, "numeric")]), "max"), sumv],
        group_by=["k"], num_groups=16)))
    return q


@pytest.mark.parametrize("name", [n for n, _ in typelib_queries()])
def test_queries_plan_and_compile(lib, name):
    plan_tree = dict(typelib_queries())[name]
    plan = gp.Plan(plan_tree, gucs=GUCS)
    try:
        assert plan.num_gpupreagg == 1, plan.reject_reason
        src = plan.kernel_source()
        if name.startswith(("date", "timestamp")):
            assert '#include "kern_timelib.cuh"' in src
        if name.startswith(("text", "bpchar")):
            assert '#include "kern_textlib.cuh"' in src
        prog = plan.build_program()         # NVRTC for sm_100a
        plan.lib.pgs_program_release(prog)
    finally:
        plan.free()


def test_text_ordering_needs_c_collation(lib):
    t = EVENTS
    tree = P.make_agg_plan(t, [(P.Agg("count", star=True), "count")],
                           where=[P.Op("<", t.col("s"), P.Const("text", "m"),
                                       collation="en_US.utf8")])
    plan = gp.Plan(tree, gucs=GUCS)
    try:
        # the qual stays on the scan node (PostgreSQL's executor filters the
        # rows); GpuPreAgg aggregates what the scan returns
        assert plan.num_gpupreagg == 1, plan.reject_reason
        assert "#define GPUPREAGG_HAS_QUAL 0" in plan.kernel_source()
        assert "kern_textlib" not in plan.kernel_source()
    finally:
        plan.free()
    # equality does not depend on the collation
    tree = P.make_agg_plan(t, [(P.Agg("count", star=True), "count")],
                           where=[P.Op("=", t.col("s"), P.Const("text", "m"),
                                       collation="en_US.utf8")])
    plan = gp.Plan(tree, gucs=GUCS)
    try:
        assert plan.num_gpupreagg == 1, plan.reject_reason
        assert "#define GPUPREAGG_HAS_QUAL 1" in plan.kernel_source()
    finally:
        plan.free()


def test_text_group_key_is_kernel_text(lib):
    """text / bpchar grouping keys are offloaded as "kernel text" (<= 7 bytes
    by value); the generated program compiles; numeric keys stay on the host."""
    t = P.Table("cats", [("cat", "text"), ("code", "bpchar"), ("n", "numeric"), ("v", "int4")],
                typmods={"code": 4 + 5})
    tree = P.make_agg_plan(t, [(t.col("cat"), "cat"), (t.col("code"), "code"),
                               (P.Agg("count", star=True), "count"),
                               (P.Agg("sum", [t.col("v")]), "sum")],
                           group_by=["cat", "code"], num_groups=26)
    plan = gp.Plan(tree, gucs=GUCS)
    try:
        assert plan.num_gpupreagg == 1, plan.reject_reason
        src = plan.kernel_source()
        assert "pg_text_vstore(kds_src,kds_in,errcode,0,rowidx_out,KVAR_1)" in src
        assert "pg_bpchar_vstore(kds_src,kds_in,errcode,1,rowidx_out,KVAR_2)" in src
        cols = plan.describe()["columns"]
        assert cols[0]["type"] == "text" and cols[1]["type"] == "bpchar"
        assert cols[1]["typmod"] == 9 and "typmod" not in cols[0]
        prog = plan.build_program()
        plan.lib.pgs_program_release(prog)
    finally:
        plan.free()
    tree = P.make_agg_plan(t, [(t.col("n"), "n"), (P.Agg("count", star=True), "count")],
                           group_by=["n"], num_groups=26)
    plan = gp.Plan(tree, gucs=GUCS)
    try:
        assert plan.num_gpupreagg == 0
        assert "not supported" in plan.reject_reason
    finally:
        plan.free()


def test_kernel_text_fixup(lib, shim):
    """device packing (pgs_text_keybits) and host fix-up
    (pgstrom_fixup_kernel_text) are inverse; bpchar loses its trailing blanks
    on the device and gets them back from the typmod."""
    out, isnull = C.c_uint64(), C.c_int()
    buf = C.create_string_buffer(64)
    cases = [b"", b"a", b"abc", b"abcdefg", b"ab  ", b"   ", b"caf\xc3\xa9", b"\xe3\x81\x82",
             b"a b", b"\xff\x00x"[:1] + b"x"]
    for s in cases:
        for short in (None, False):
            img = T.varlena(s, short=short)
            err = shim.shim_text_keybits(img, 0, C.byref(out), C.byref(isnull))
            assert err == 0 and not isnull.value
            n = lib.pgstrom_fixup_kernel_text(out.value, -1, buf, len(buf))
            assert buf.raw[:n] == T.varlena(s, short=False), s
            # bpchar: blanks are not part of the key ...
            err = shim.shim_text_keybits(img, 1, C.byref(out), C.byref(isnull))
            assert err == 0 and not isnull.value
            word2 = C.c_uint64()
            shim.shim_text_keybits(T.varlena(s.rstrip(b" ")), 1, C.byref(word2), C.byref(isnull))
            assert word2.value == out.value
            n = lib.pgstrom_fixup_kernel_text(out.value, -1, buf, len(buf))
            assert buf.raw[4:n] == s.rstrip(b" ")
            # ... and come back from character(8)
            n = lib.pgstrom_fixup_kernel_text(out.value, 4 + 8, buf, len(buf))
            nchars = len(s.rstrip(b" ").decode("utf-8", "replace"))
            assert buf.raw[4:n] == s.rstrip(b" ") + b" " * max(0, 8 - nchars), s
    # two strings, one key <=> equal
    words = {}
    for s in cases + [b"abcdefg", b"abcdef", b"b", b"ab"]:
        shim.shim_text_keybits(T.varlena(s), 0, C.byref(out), C.byref(isnull))
        assert words.setdefault(out.value, s) == s
    # longer than 7 bytes and no key heap: a row for the host
    shim.shim_keyheap_set(None, 0, None, 0, None, 0)
    err = shim.shim_text_keybits(T.varlena(b"abcdefgh"), 0, C.byref(out), C.byref(isnull))
    assert err == CPU_RECHECK and isnull.value
    err = shim.shim_text_keybits(T.varlena(b"abcdefg   "), 1, C.byref(out), C.byref(isnull))
    assert err == 0 and not isnull.value
    assert lib.pgstrom_fixup_kernel_text(out.value, -1, buf, 8) == 0     # buffer too small


class _KeyHeap:
    """Host memory laid out like the key heap of a session (cuda_layer.cpp:
    session_alloc_keyheap) for the CPU build of pgs_keyheap_intern."""

    def __init__(self, shim, nslots, heap_bytes, max_probe=512):
        self.slots = (C.c_uint64 * (2 * nslots))()
        self.heap = C.create_string_buffer(max(heap_bytes, 8))
        self.used = C.c_uint64(0)
        self.heap_bytes = heap_bytes
        shim.shim_keyheap_set(self.slots, nslots, self.heap, heap_bytes, C.byref(self.used),
                              max_probe)


def test_key_heap_interns_long_keys(lib, shim):
    """text / bpchar keys of more than 7 bytes: the device stores the string
    once in the session's key heap and groups by the 8-byte word (top byte
    0x80 + heap offset); equal strings <=> equal words, the host gets the
    varlena back from the heap copy (pgstrom_fixup_kernel_text_heap) - the
    counterpart of the reference's varlena key move + pointer fix-up
    (opencl_gpupreagg.h:326-366)."""
    rng = random.Random(5)
    kh = _KeyHeap(shim, 1 << 12, 1 << 20)
    out, isnull = C.c_uint64(), C.c_int()
    pool = [b"abcdefgh", b"abcdefgi", b"category-with-a-long-name", b"x" * 300, b"y" * 3000,
            b"caf\xc3\xa9 au lait", b"12345678", b"\x00" * 9, b"trailing blank  "]
    pool += [bytes(rng.randrange(256) for _ in range(rng.randrange(8, 60))) for _ in range(1500)]
    pool = list(dict.fromkeys(pool))
    words = {}
    for rnd in range(3):
        order = pool[:]
        rng.shuffle(order)
        for sval in order:
            img = T.varlena(sval, short=(None if rnd else False))
            err = shim.shim_text_keybits(img, 0, C.byref(out), C.byref(isnull))
            assert err == 0 and not isnull.value, sval
            assert out.value >> 56 == 0x80
            assert words.setdefault(sval, out.value) == out.value       # same string, same word
    assert len(set(words.values())) == len(pool)                        # other string, other word
    # every string was stored exactly once
    assert kh.used.value == sum(8 + (len(s) + 7) // 8 * 8 for s in pool)
    # host side: the word + the heap copy -> the varlena PostgreSQL expects
    heap = C.string_at(kh.heap, kh.used.value)
    buf = C.create_string_buffer(4096)
    for sval, w in words.items():
        n = lib.pgstrom_fixup_kernel_text_heap(w, -1, heap, len(heap), buf, len(buf))
        assert buf.raw[:n] == T.varlena(sval, short=False), sval
    # without the heap, with a truncated heap, with a word that points nowhere: refused
    w = words[b"y" * 3000]
    assert lib.pgstrom_fixup_kernel_text_heap(w, -1, None, 0, buf, len(buf)) == 0
    assert lib.pgstrom_fixup_kernel_text(w, -1, buf, len(buf)) == 0
    off = w & ((1 << 56) - 1)
    assert lib.pgstrom_fixup_kernel_text_heap(w, -1, heap, off + 100, buf, len(buf)) == 0
    assert lib.pgstrom_fixup_kernel_text_heap(w | 4, -1, heap, len(heap), buf, len(buf)) == 0
    assert lib.pgstrom_fixup_kernel_text_heap(w, -1, heap, len(heap), buf, 3000) == 0    # buf too small
    # bpchar: trailing blanks are not part of the key, the typmod pads them back
    err = shim.shim_text_keybits(T.varlena(b"abcdefgh   "), 1, C.byref(out), C.byref(isnull))
    assert err == 0 and out.value == words[b"abcdefgh"]
    n = lib.pgstrom_fixup_kernel_text_heap(out.value, 4 + 12, heap, len(heap), buf, len(buf))
    assert buf.raw[4:n] == b"abcdefgh    "
    # a key of more than 64 KB is not worth a thread's time: a row for the host
    err = shim.shim_text_keybits(T.varlena(b"z" * 65535, short=False), 0, C.byref(out),
                                 C.byref(isnull))
    assert err == 0 and out.value >> 56 == 0x80
    err = shim.shim_text_keybits(T.varlena(b"z" * 65536, short=False), 0, C.byref(out),
                                 C.byref(isnull))
    assert err == CPU_RECHECK and isnull.value
    # short keys never touch the heap
    before = kh.used.value
    err = shim.shim_text_keybits(T.varlena(b"abcdefg"), 0, C.byref(out), C.byref(isnull))
    assert err == 0 and out.value >> 56 == 7 and kh.used.value == before
    shim.shim_keyheap_set(None, 0, None, 0, None, 0)


def test_key_heap_full_means_recheck(lib, shim):
    """A heap without room (or a lookup table whose probe limit is hit) turns
    rows with an unseen long key into rows for the host; keys stored before
    keep working, and so does a later, shorter string only if it fits."""
    out, isnull = C.c_uint64(), C.c_int()
    kh = _KeyHeap(shim, 64, 80)                     # room for 24 + 24 + 24 bytes
    a, b, c, d = b"first-key", b"second-long-key", b"third-long-keyyy", b"fourth-key"
    for sval in (a, b, c):
        assert shim.shim_text_keybits(T.varlena(sval), 0, C.byref(out), C.byref(isnull)) == 0
    wa = C.c_uint64()
    assert shim.shim_text_keybits(T.varlena(a), 0, C.byref(wa), C.byref(isnull)) == 0
    assert kh.used.value == 72
    err = shim.shim_text_keybits(T.varlena(d), 0, C.byref(out), C.byref(isnull))
    assert err == CPU_RECHECK and isnull.value
    # again: the slot says "no room", still a re-check; the others are still found
    err = shim.shim_text_keybits(T.varlena(d), 0, C.byref(out), C.byref(isnull))
    assert err == CPU_RECHECK and isnull.value
    w2 = C.c_uint64()
    assert shim.shim_text_keybits(T.varlena(a), 0, C.byref(w2), C.byref(isnull)) == 0
    assert w2.value == wa.value
    # probe limit: 8 slots, all taken by other strings
    kh = _KeyHeap(shim, 8, 4096, max_probe=8)
    got = 0
    for i in range(40):
        err = shim.shim_text_keybits(T.varlena(b"key-number-%04d" % i), 0, C.byref(out),
                                     C.byref(isnull))
        assert err in (0, CPU_RECHECK)
        got += (err == 0)
    assert got == 8
    shim.shim_keyheap_set(None, 0, None, 0, None, 0)


def test_key_heap_under_concurrency():
    """The claim / publish / wait protocol of pgs_keyheap_intern with real
    threads (tests/native/keyheap_stress.cpp: the device atomics mapped to
    __atomic builtins): every thread gets the same word for the same string,
    different strings get different words, every string is stored once; with
    a heap that is too small the strings that found no room are refused for
    every thread alike and nothing else changes."""
    import json
    exe = os.path.join(HERE, "native", "_keyheap_stress")
    subprocess.run(["g++", "-std=c++17", "-O2", "-pthread", "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(ROOT, "pg_strom_b200", "csrc"), "-o", exe,
                    os.path.join(HERE, "native", "keyheap_stress.cpp")], check=True)
    for args in (["8", "4000", "4", "16384", str(8 << 20)], ["32", "1500", "12", "4096", str(1 << 20)]):
        r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=120)
        d = json.loads(r.stdout)
        assert r.returncode == 0, d
        assert d["failed_lookups"] == 0 and d["mismatch"] == 0 and d["duplicate_words"] == 0
        assert d["heap_used"] == d["expect_used"]
    r = subprocess.run([exe, "8", "4000", "3", "16384", "40000"], capture_output=True, text=True,
                       timeout=120)
    d = json.loads(r.stdout)
    assert d["failed_lookups"] > 0 and d["missing"] > 0
    assert d["mismatch"] == 0 and d["duplicate_words"] == 0
    # the same under ThreadSanitizer (acquire / release as mapped above): no data race
    tsan = os.path.join(HERE, "native", "_keyheap_stress_tsan")
    c = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-pthread", "-w",
                        "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(ROOT, "pg_strom_b200", "csrc"), "-o", tsan,
                        os.path.join(HERE, "native", "keyheap_stress.cpp")], capture_output=True)
    if c.returncode == 0:
        r = subprocess.run([tsan, "8", "2000", "3", "8192", str(4 << 20)], capture_output=True,
                           text=True, timeout=300)
        assert "ThreadSanitizer" not in r.stderr, r.stderr[:2000]
        assert r.returncode == 0, r.stdout
