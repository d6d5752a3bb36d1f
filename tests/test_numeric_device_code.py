"""CPU: the device NUMERIC code (pg_strom_b200/csrc/kern_numeric.cuh, the
counterpart of the reference's opencl_numeric.h) compiled with g++ through a
small shim and checked against python's Decimal: varlena image -> 64-bit
device format (exponent = -display scale, 57-bit mantissa, CpuReCheck beyond
that), comparison, add / sub / mul with PostgreSQL's display-scale rules,
casts."""
import ctypes as C
import os
import random
import subprocess
from decimal import Decimal, getcontext

import pytest

from pg_strom_b200 import gpupreagg as gp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
MANT_LIMIT = (1 << 57) - 2


@pytest.fixture(scope="module")
def shim(lib):
    out = os.path.join(HERE, "native", "_numeric_shim.so")
    src = os.path.join(HERE, "native", "numeric_host_shim.cpp")
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared",
                    "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(ROOT, "pg_strom_b200", "csrc"),
                    "-o", out, src], check=True)
    so = C.CDLL(out)
    so.shim_from_varlena.argtypes = [C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
    so.shim_cmp.argtypes = [C.c_uint64, C.c_uint64]
    so.shim_binop.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64),
                              C.POINTER(C.c_int)]
    so.shim_to_float8.argtypes = [C.c_uint64]
    so.shim_to_float8.restype = C.c_double
    so.shim_to_int8.argtypes = [C.c_uint64, C.POINTER(C.c_int64)]
    return so


def unpack(v):
    """64-bit device numeric -> Decimal (what pgstrom_fixup_kernel_numeric +
    numeric_in do, datastore.c:150-167)."""
    exp = v >> 58
    if exp >= 32:
        exp -= 64
    mant = v & ((1 << 57) - 1)
    d = Decimal(mant).scaleb(exp)
    return -d if (v >> 57) & 1 else d


def rand_decimal(rng):
    ndig = rng.choice([1, 2, 5, 9, 12, 15, 17, 18, 20, 25])
    scale = rng.choice([0, 0, 1, 2, 2, 4, 6, 10, 16, 20])
    mant = rng.randrange(10 ** (ndig - 1) if ndig > 1 else 0, 10 ** ndig)
    if rng.random() < 0.15:
        mant = mant // 10 ** rng.randrange(1, 6) * 10 ** rng.randrange(1, 6)   # trailing zeros
    d = Decimal(mant).scaleb(-scale)
    return -d if rng.random() < 0.4 else d


def to_device(shim, d, short=False):
    img = gp.numeric_datum(format(d, "f"))
    if short:
        assert len(img) - 3 <= 127
        img = bytes([((len(img) - 3) << 1) | 1]) + img[4:]
    out, isnull = C.c_uint64(), C.c_int()
    err = shim.shim_from_varlena(img, C.byref(out), C.byref(isnull))
    return err, out.value, bool(isnull.value)


def fits(d):
    sign, digits, exp = d.as_tuple()
    scale = max(0, -exp)
    mant = int(abs(d).scaleb(scale))
    return mant <= MANT_LIMIT and scale <= 32, mant, scale


def test_from_varlena(shim):
    rng = random.Random(7)
    nfit = nre = 0
    for i in range(4000):
        d = rand_decimal(rng)
        ok, mant, scale = fits(d)
        err, v, isnull = to_device(shim, d, short=(i % 2 == 1))
        if ok:
            assert err == 0 and not isnull, d
            got = unpack(v)
            assert got == d and got.as_tuple().exponent == -scale, (d, got)
            nfit += 1
        else:
            assert err == 2 and isnull, d      # StromError_CpuReCheck
            nre += 1
    assert nfit > 1500 and nre > 300
    for text in ("0", "0.00", "-0.0", "1", "10000", "100000000", "0.0001", "0.00010",
                 "123456789012345678", "0.5", "99999999.9999"):
        d = Decimal(text)
        err, v, isnull = to_device(shim, d)
        assert err == 0
        assert unpack(v) == d and -unpack(v).as_tuple().exponent == max(0, -d.as_tuple().exponent)


def test_compare_and_arithmetic(shim):
    getcontext().prec = 120
    rng = random.Random(11)
    vals = []
    while len(vals) < 300:
        d = rand_decimal(rng)
        if fits(d)[0]:
            vals.append((d, to_device(shim, d)[1]))
    for i in range(3000):
        (a, va), (b, vb) = rng.choice(vals), rng.choice(vals)
        assert shim.shim_cmp(va, vb) == (a > b) - (a < b), (a, b)
        for op, fn in ((0, lambda x, y: x + y), (1, lambda x, y: x - y), (2, lambda x, y: x * y)):
            out, isnull = C.c_uint64(), C.c_int()
            err = shim.shim_binop(op, va, vb, C.byref(out), C.byref(isnull))
            want = fn(a, b)
            sa, sb = -a.as_tuple().exponent, -b.as_tuple().exponent
            wscale = sa + sb if op == 2 else max(sa, sb)
            want = want.quantize(Decimal(1).scaleb(-wscale)) if want == want else want
            if err == 0:
                got = unpack(out.value)
                assert got == want and -got.as_tuple().exponent == wscale, (op, a, b, got, want)
            else:
                assert err == 2
                mant = int(abs(want).scaleb(wscale))
                assert mant > MANT_LIMIT or wscale > 32 or abs(sa - sb) > 20


def test_casts(shim):
    rng = random.Random(5)
    for i in range(2000):
        d = rand_decimal(rng)
        ok, mant, scale = fits(d)
        if not ok:
            continue
        v = to_device(shim, d)[1]
        f = shim.shim_to_float8(v)
        if mant < (1 << 53) and scale <= 22:
            assert f == float(d), d            # correctly rounded, like strtod
        else:
            assert abs(f - float(d)) <= 4e-16 * abs(float(d))
        out = C.c_int64()
        rc = shim.shim_to_int8(v, C.byref(out))
        want = int(d.quantize(Decimal(1), rounding="ROUND_HALF_UP"))
        if -2 ** 63 <= want < 2 ** 63:
            assert rc == 0 and out.value == want, d
        else:
            assert rc == -1


def test_recheck_agg_known_answers(shim, lib):
    """The reference's only direct known-answer test of the device numeric
    range: input/sql/recheck_agg.sql / expected/recheck_agg.out
    (tests/golden/recheck_agg.json).  `select sum(<literal>)`: the literal
    reaches the device as a numeric varlena (kern_parambuf); a value the 64-bit
    device numeric holds comes back through pgstrom_fixup_kernel_numeric() as
    the golden's text, one it does not hold raises CpuReCheck and PostgreSQL
    computes the golden's text itself (the NOTICE of the golden).

    One deliberate difference: here the exponent of a parsed
    value is minus its display scale, so that sums / min / max print like
    PostgreSQL's; 1E+48 (49 digits at scale 0) is therefore re-checked, where
    the reference stores it as 10^17 x 10^31 and loses the display scale.  The
    result text is the same either way."""
    import json
    with open(os.path.join(HERE, "golden", "recheck_agg.json")) as f:
        stmts = json.load(f)
    assert len(stmts) == 7
    rechecked_here = []
    for s in stmts:
        lit = s["sql"].split("sum(")[1].split(")")[0]
        d = Decimal(lit)
        err, v, isnull = to_device(shim, d)
        ref_rechecks = any("re-checked by CPU" in n for n in s["notices"])
        if err == 0:
            assert not isnull and not ref_rechecks, lit
            buf = C.create_string_buffer(128)
            assert lib.pgstrom_fixup_kernel_numeric(v, buf, len(buf)) == 0
            # numeric_in() of the fixed-up text, numeric_out() of that
            assert format(Decimal(buf.value.decode()), "f") == s["rows"][0][0], lit
        else:
            assert err == 2 and isnull, lit
            rechecked_here.append(lit)
            assert format(d, "f") == s["rows"][0][0], lit      # what the host computes
            assert ref_rechecks or lit == "1E+48", lit
    assert rechecked_here == ["1E+48", "1E-33", "1E+49", "1E+1000", "1E-1000"]
