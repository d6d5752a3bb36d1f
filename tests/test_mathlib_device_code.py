"""CPU: the arithmetic / cast / math runtime of the device code
(pg_strom_b200/csrc/kern_mathlib.cuh, the counterpart of the reference's
opencl_mathlib.h:34-818 - PostgreSQL-compatible overflow, division-by-zero and
range detection, the result being NULL + StromError_CpuReCheck) compiled with
g++ through a generated shim and checked against the oracle's restatement of
PostgreSQL's operators (oracle/pg_expr.py).  Also: the domain checks codegen
puts around CUDA's math built-ins (sqrt, ln, exp ...)."""
import ctypes as C
import math
import os
import random
import struct
import subprocess

import pytest

from oracle import pg_expr
from oracle.pg_agg import PgError
from pg_strom_b200 import gpupreagg as gp
from pg_strom_b200 import pgplan as P

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CPU_RECHECK = 2
BASE = {"bool": "cl_bool", "int2": "cl_short", "int4": "cl_int", "int8": "cl_long",
        "float4": "cl_float", "float8": "cl_double"}
INTS = ("int2", "int4", "int8")
FLOATS = ("float4", "float8")
GUCS = {"pg_strom.enabled": "on", "pg_strom.debug_force_gpupreagg": "on"}

PREAMBLE = r'''
#include <cstdint>
#include <cmath>
#include <climits>
#include "pgstrom_kds.h"
using std::isinf; using std::isnan;
#define DEVFN static inline
static inline void STROM_SET_ERROR(cl_int *p_error, cl_int errcode)
{
    cl_int oldcode = *p_error;
    if (StromErrorIsSignificant(errcode))
    {
        if (!StromErrorIsSignificant(oldcode))
            *p_error = errcode;
    }
    else if (errcode > oldcode)
        *p_error = errcode;
}
static inline long long __mul64hi(long long a, long long b)
{ return (long long)(((__int128)a * (__int128)b) >> 64); }
typedef struct { cl_bool value; bool isnull; } pg_bool_t;
typedef struct { cl_short value; bool isnull; } pg_int2_t;
typedef struct { cl_int value; bool isnull; } pg_int4_t;
typedef struct { cl_long value; bool isnull; } pg_int8_t;
typedef struct { cl_float value; bool isnull; } pg_float4_t;
typedef struct { cl_double value; bool isnull; } pg_float8_t;
#undef LONG_MAX
#undef LONG_MIN
#define LONG_MAX    9223372036854775807LL
#define LONG_MIN    (-LONG_MAX-1LL)
#include "kern_mathlib.cuh"
'''


def catalogue():
    """(pgfn name, argtypes, rettype, pg_proc name) of kern_mathlib.cuh."""
    out = []
    width = {"int2": 2, "int4": 4, "int8": 8}
    for sfx in ("pl", "mi", "mul", "div"):
        for a in INTS:
            for b in INTS:
                n = "int" + a[3:] + ("" if a == b else b[3:]) + sfx
                out.append((n, [a, b], "int%d" % max(width[a], width[b]), n))
        for n, a, b in (("float4", "float4", "float4"), ("float48", "float4", "float8"),
                        ("float84", "float8", "float4"), ("float8", "float8", "float8")):
            out.append((n + sfx, [a, b], "float8" if "float8" in (a, b) else "float4", n + sfx))
    for a in INTS:
        out.append((a + "mod", [a, a], a, a + "mod"))
        out.append((a + "um", [a], a, a + "um"))
        out.append((a + "abs", [a], a, a + "abs"))
    nums = INTS + FLOATS
    for r in nums:
        for a in nums:
            if r != a:
                out.append(("%s_%s" % (a, r), [a], r, r))      # cast: function named after target
    out.append(("dpow", ["float8", "float8"], "float8", "dpow"))
    out.append(("dsign", ["float8"], "float8", "sign"))
    out.append(("degrees", ["float8"], "float8", "degrees"))
    out.append(("radians", ["float8"], "float8", "radians"))
    return out


def shim_source(cat):
    s = [PREAMBLE, 'extern "C" int shim_call(int fn, const double *fa, const long long *ia,',
         '                          double *fout, long long *iout, int *isnull)',
         '{', '    cl_int e = 0;', '    switch (fn)', '    {']
    for k, (name, args, ret, _) in enumerate(cat):
        s.append("        case %d: {" % k)
        for i, t in enumerate(args):
            src = ("fa[%d]" if t in FLOATS else "ia[%d]") % i
            s.append("            pg_%s_t a%d = { (%s)%s, false };" % (t, i, BASE[t], src))
        s.append("            pg_%s_t r = pgfn_%s(&e%s);" %
                 (ret, name, "".join(", a%d" % i for i in range(len(args)))))
        s.append("            *isnull = r.isnull; *fout = (double)r.value; *iout = (long long)r.value;")
        s.append("            return e; }")
    s += ['    }', '    return -1;', '}']
    return "\n".join(s) + "\n"


@pytest.fixture(scope="module")
def shim(lib):
    cat = catalogue()
    src = os.path.join(HERE, "native", "_mathlib_shim_generated.cpp")
    out = os.path.join(HERE, "native", "_mathlib_shim.so")
    with open(src, "w") as f:
        f.write(shim_source(cat))
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-ffp-contract=off",
                    "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(ROOT, "pg_strom_b200", "csrc"),
                    "-o", out, src], check=True)
    so = C.CDLL(out)
    so.shim_call.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong),
                             C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
    return so, cat


def f4(x):
    return struct.unpack("f", struct.pack("f", x))[0]


def samples(t, rng):
    lim = {"int2": 15, "int4": 31, "int8": 63}
    if t in lim:
        b = lim[t]
        edge = [0, 1, -1, 2, -2, 2 ** b - 1, -2 ** b, 2 ** b - 2, -2 ** b + 1, 2 ** (b // 2),
                -2 ** (b // 2), 3, 7, -7, 10]
        return rng.choice(edge) if rng.random() < 0.5 else \
            rng.randrange(-2 ** b, 2 ** b) >> rng.choice([0, 0, b // 2, b - 4])
    edge = [0.0, -0.0, 1.0, -1.0, 0.5, 1.5, 2.5, -2.5, 1e-300, 1e300, -1e300, 1e-45, 3.4e38,
            -3.4e38, 1e38, 1e-38, 32767.5, 32767.4, -32768.5, 2147483647.5, 2147483647.4,
            -2147483648.5, 9.3e18, -9.3e18, 9223372036854775807.0, float("inf"), float("-inf"),
            float("nan"), 65504.0, 1e10, 123456789.125]
    v = rng.choice(edge) if rng.random() < 0.5 else \
        rng.uniform(-1, 1) * 10.0 ** rng.randrange(-40, 40)
    if t == "float4":
        try:
            v = f4(v)
        except OverflowError:
            v = math.copysign(math.inf, v)
    return v


def same(a, b, t):
    if t in FLOATS:
        if math.isnan(a) or math.isnan(b):
            return math.isnan(a) and math.isnan(b)
        return a == b and math.copysign(1, a) == math.copysign(1, b) or (a == b == 0)
    return a == b


def test_mathlib_against_oracle(shim):
    so, cat = shim
    rng = random.Random(9)
    fa = (C.c_double * 2)()
    ia = (C.c_longlong * 2)()
    fout, iout, isnull = C.c_double(), C.c_longlong(), C.c_int()
    nfail = nok = 0
    for k, (name, args, ret, pgname) in enumerate(cat):
        for _ in range(1500):
            vals = [samples(t, rng) for t in args]
            for i, (t, v) in enumerate(zip(args, vals)):
                if t in FLOATS:
                    fa[i] = v
                else:
                    ia[i] = v
            err = so.shim_call(k, fa, ia, C.byref(fout), C.byref(iout), C.byref(isnull))
            try:
                exp = pg_expr._call(pgname, args, ret, vals)
                failed = False
            except PgError:
                failed = True
            if failed:
                assert err == CPU_RECHECK and isnull.value, (name, vals)
                nfail += 1
            else:
                assert err == 0 and not isnull.value, (name, vals, exp)
                got = fout.value if ret in FLOATS else iout.value
                assert same(got, exp, ret), (name, vals, got, exp)
                nok += 1
    assert nfail > 5000 and nok > 50000


MATHT = P.Table("m", [("x", "float8"), ("k", "int4")])


def _fn(name, *args):
    return {"node": "FuncExpr", "funcname": name, "funcresulttype": "float8",
            "funcformat": "call", "args": list(args)}


def test_builtin_domain_checks_are_generated(lib):
    """sqrt / ln / log / exp / acos / asin / cos / sin / tan / cbrt go through
    CUDA's built-ins; PostgreSQL raises where those return NaN / Inf / 0, so
    the generated wrappers must re-check exactly there."""
    x = MATHT.col("x")
    one = P.Const("float8", "1")
    quals = [P.Op(">", _fn(f, x), one) for f in
             ("sqrt", "ln", "log", "exp", "acos", "asin", "cos", "sin", "tan", "cbrt",
              "atan", "floor", "ceil", "round", "trunc", "sign", "degrees", "radians")]
    quals.append(P.Op(">", _fn("power", x, one), one))
    quals.append(P.Op(">", _fn("atan2", x, one), one))
    tree = P.make_agg_plan(MATHT, [(P.Agg("count", star=True), "count")], where=quals)
    plan = gp.Plan(tree, gucs=GUCS)
    try:
        assert plan.num_gpupreagg == 1, plan.reject_reason
        src = plan.kernel_source()
        assert "#define GPUPREAGG_HAS_QUAL 1" in src

        def body(fname):
            i = src.index("pgfn_%s(cl_int *errcode" % fname)
            return src[i:src.index("\n}\n", i)]
        assert "arg1.value < 0.0" in body("sqrt") and "PGS_MATH_FAIL" in body("sqrt")
        for f in ("ln", "log"):
            assert "arg1.value <= 0.0" in body(f)
        assert "CHECKFLOATVAL(result.value, isinf(arg1.value), false)" in body("exp")
        for f in ("acos", "asin"):
            assert "fabs(arg1.value) > 1.0" in body(f)
        for f in ("cos", "sin", "tan"):
            assert "isinf(arg1.value)" in body(f)
        for f in ("atan", "floor", "ceil"):
            assert "PGS_MATH_FAIL" not in body(f)
        prog = plan.build_program()         # NVRTC, sm_100a
        plan.lib.pgs_program_release(prog)
    finally:
        plan.free()
