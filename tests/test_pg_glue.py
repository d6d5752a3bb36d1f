"""CPU: the fmgr V1 functions the SQL catalog names (pg_strom--1.0.sql:99-401;
reference bodies gpupreagg.c:4251-4773), i.e. pg_glue/gpupreagg_fmgr.c,
compiled against a stand-in for the PostgreSQL headers
(tests/native/pg_stub/: PG_FUNCTION_ARGS, the ARR_* macros, ereport) and
called the way the executor calls them: as plain functions (the transition
array is copied) and as an aggregate's sfunc (updated in place).  The two
numeric-state wrappers need PostgreSQL's int8_avg_accum / numeric_avg_accum
and are compiled out here; their N rule is tested on pgs_numeric_avg_accum in
test_finalfn.py."""
import ctypes as C
import math
import os
import struct
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
INT8OID, FLOAT8OID = 20, 701
NARGS = 100


class FCInfo(C.Structure):
    _fields_ = [("nargs", C.c_short), ("isnull", C.c_bool), ("context", C.c_void_p),
                ("arg", C.c_uint64 * NARGS), ("argnull", C.c_bool * NARGS)]


def make_array(elemtype, values):
    fmt = "q" if elemtype == INT8OID else "d"
    raw = struct.pack("<ii?xxxI", 1, len(values), False, elemtype) + \
        struct.pack("<%d%s" % (len(values), fmt), *values)
    assert len(raw) == 16 + 8 * len(values)
    return C.create_string_buffer(raw, len(raw))


def read_array(ptr, elemtype, n):
    raw = C.string_at(ptr, 16 + 8 * n)
    return list(struct.unpack("<%d%s" % (n, "q" if elemtype == INT8OID else "d"), raw[16:]))


def f8(v):
    return struct.unpack("<Q", struct.pack("<d", v))[0]


def datum_f8(d):
    return struct.unpack("<d", struct.pack("<Q", d))[0]


@pytest.fixture(scope="module")
def glue(lib):
    out = os.path.join(HERE, "native", "_pg_glue.so")
    subprocess.run(["gcc", "-std=gnu11", "-Wall", "-Werror", "-O1", "-fPIC", "-shared",
                    "-DPGSTROM_GLUE_NO_NUMERIC",
                    "-I", os.path.join(HERE, "native", "pg_stub"),
                    "-I", os.path.join(ROOT, "include"), "-o", out,
                    os.path.join(ROOT, "pg_glue", "gpupreagg_fmgr.c"),
                    os.path.join(HERE, "native", "pg_stub", "pg_stub.c"),
                    "-L", os.path.join(ROOT, "pg_strom_b200"), "-lpgstrom_cuda",
                    "-Wl,-rpath," + os.path.join(ROOT, "pg_strom_b200")], check=True)
    return C.CDLL(out)


def call(glue, name, args, aggregate=False):
    """args: list of Datum (int) or None for SQL NULL.  Returns (datum, isnull, error)."""
    fc = FCInfo()
    fc.nargs = len(args)
    fc.context = 1 if aggregate else None
    for i, a in enumerate(args):
        fc.argnull[i] = a is None
        fc.arg[i] = 0 if a is None else (a & 0xFFFFFFFFFFFFFFFF)
    fn = getattr(glue, name)
    fn.restype = C.c_uint64
    fn.argtypes = [C.POINTER(FCInfo)]
    C.c_int.in_dll(glue, "pg_stub_error_code").value = 0
    d = fn(C.byref(fc))
    code = C.c_int.in_dll(glue, "pg_stub_error_code").value
    msg = C.string_at(C.addressof((C.c_char * 256).in_dll(glue, "pg_stub_error_message"))).decode()
    return d, bool(fc.isnull), (code, msg) if code else None


def test_catalog_symbols_exist(glue):
    """every C symbol of the catalog's GpuPreAgg section that does not need
    PostgreSQL's numeric internals"""
    for name in ("gpupreagg_partial_nrows", "gpupreagg_pseudo_expr", "gpupreagg_psum_int",
                 "gpupreagg_psum_float4", "gpupreagg_psum_float8", "gpupreagg_psum_numeric",
                 "gpupreagg_psum_x2_float", "gpupreagg_corr_psum_x", "gpupreagg_corr_psum_y",
                 "gpupreagg_corr_psum_x2", "gpupreagg_corr_psum_y2", "gpupreagg_corr_psum_xy",
                 "pgstrom_avg_int8_accum", "pgstrom_sum_int8_accum", "pgstrom_sum_int8_final",
                 "pgstrom_sum_float8_accum", "pgstrom_variance_float8_accum",
                 "pgstrom_covariance_float8_accum"):
        assert getattr(glue, name)
    src = open(os.path.join(ROOT, "pg_glue", "gpupreagg_fmgr.c")).read()
    for name in ("gpupreagg_psum_x2_numeric", "pgstrom_int8_avg_accum",
                 "pgstrom_numeric_avg_accum"):
        assert "PG_FUNCTION_INFO_V1(%s)" % name in src


def test_placeholders(glue):
    assert call(glue, "gpupreagg_partial_nrows", [])[0] == 1
    assert call(glue, "gpupreagg_partial_nrows", [1, 1])[0] == 1
    assert call(glue, "gpupreagg_partial_nrows", [1, 0])[0] == 0
    assert call(glue, "gpupreagg_partial_nrows", [1, None])[0] == 0
    assert call(glue, "gpupreagg_pseudo_expr", [12345])[:2] == (12345, False)
    for fn in ("gpupreagg_psum_int", "gpupreagg_psum_float4", "gpupreagg_psum_float8",
               "gpupreagg_psum_numeric"):
        assert call(glue, fn, [None])[1] is True
        assert call(glue, fn, [777])[:2] == (777, False)
    d, isnull, _ = call(glue, "gpupreagg_psum_x2_float", [f8(-1.5)])
    assert not isnull and datum_f8(d) == 2.25
    assert call(glue, "gpupreagg_psum_x2_float", [None])[1] is True
    want = {"x": 3.0, "y": -0.5, "x2": 9.0, "y2": 0.25, "xy": -1.5}
    for k, v in want.items():
        d, isnull, _ = call(glue, "gpupreagg_corr_psum_" + k, [1, f8(3.0), f8(-0.5)])
        assert not isnull and datum_f8(d) == v, k
        for args in ([0, f8(3.0), f8(-0.5)], [None, f8(3.0), f8(-0.5)],
                     [1, None, f8(-0.5)], [1, f8(3.0), None]):
            assert call(glue, "gpupreagg_corr_psum_" + k, args)[1] is True


@pytest.mark.parametrize("aggregate", [False, True])
def test_accumulators(glue, aggregate):
    # int8[2]
    arr = make_array(INT8OID, [0, 0])
    p = C.addressof(arr)
    for nrows, psum in ((3, 10), (2 ** 31 - 1, -2 ** 40), (0, 5)):
        d, isnull, err = call(glue, "pgstrom_avg_int8_accum", [p, nrows, psum], aggregate)
        assert err is None and not isnull
        assert (d == p) == aggregate          # in place only as an aggregate's sfunc
        p = d
    assert read_array(p, INT8OID, 2) == [3 + 2 ** 31 - 1, 15 - 2 ** 40]
    if not aggregate:
        assert read_array(C.addressof(arr), INT8OID, 2) == [0, 0]
    arr = make_array(INT8OID, [0, 0])
    p = C.addressof(arr)
    assert call(glue, "pgstrom_sum_int8_final", [p])[1] is True       # no row: NULL
    for psum in (7, -9, 2 ** 50):
        p = call(glue, "pgstrom_sum_int8_accum", [p, psum], aggregate)[0]
    d, isnull, _ = call(glue, "pgstrom_sum_int8_final", [p])
    assert not isnull and d == 2 ** 50 - 2
    # float8[3]
    p = C.addressof(make_array(FLOAT8OID, [0.0, 0.0, 0.0]))
    keep = []
    for nrows, ps, ps2 in ((2, 1.5, 2.25), (5, -0.25, 8.0)):
        a = make_array(FLOAT8OID, read_array(p, FLOAT8OID, 3))
        keep.append(a)
        p = call(glue, "pgstrom_variance_float8_accum", [C.addressof(a), nrows, f8(ps), f8(ps2)],
                 aggregate)[0]
    assert read_array(p, FLOAT8OID, 3) == [7.0, 1.25, 10.25]
    a = make_array(FLOAT8OID, [1.0, 1.7e308, 0.0])
    d, isnull, err = call(glue, "pgstrom_sum_float8_accum", [C.addressof(a), 1, f8(1.7e308)],
                          aggregate)
    assert err == (0x2203, "value out of range: overflow")
    assert read_array(C.addressof(a), FLOAT8OID, 3) == [1.0, 1.7e308, 0.0]
    d, isnull, err = call(glue, "pgstrom_sum_float8_accum",
                          [C.addressof(a), 1, f8(float("inf"))], aggregate)
    assert err is None and math.isinf(read_array(d, FLOAT8OID, 3)[1])
    # float8[6]
    a = make_array(FLOAT8OID, [0.0] * 6)
    d, isnull, err = call(glue, "pgstrom_covariance_float8_accum",
                          [C.addressof(a), 4] + [f8(v) for v in (1.0, 2.0, 3.0, 4.0, 5.0)],
                          aggregate)
    assert err is None and read_array(d, FLOAT8OID, 6) == [4.0, 1.0, 2.0, 3.0, 4.0, 5.0]
    # a transition array of the wrong shape is refused like in the reference
    bad = make_array(FLOAT8OID, [0.0, 0.0])
    d, isnull, err = call(glue, "pgstrom_variance_float8_accum",
                          [C.addressof(bad), 1, f8(1.0), f8(1.0)], aggregate)
    assert err is not None and "3-elements array is expected" in err[1]
