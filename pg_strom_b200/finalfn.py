"""Host-side mirror of the SQL-side half of GpuPreAgg (include/pgstrom_cuda.h
section 7, csrc/gpupreagg_finalfn.cpp): the transition functions of the
`pgstrom.*` final aggregates (pg_strom--1.0.sql:247-401, gpupreagg.c:4419-4773)
driven through the C ABI, one object per (group, aggregate) like PostgreSQL's
per-group transition value.  The final functions proper (`int8_avg`,
`float8_var_samp`, ... - PostgreSQL built-ins) are not part of the extension
and not part of this module; `state()` returns what they read.
"""
import ctypes as C
from decimal import Decimal

from . import _capi

VARIANCE_AGGS = ("stddev", "stddev_samp", "stddev_pop", "variance", "var_samp", "var_pop")
COVARIANCE_AGGS = ("corr", "covar_pop", "covar_samp")
ERRORS = {1: "value out of range: overflow",
          2: "Bug? NULL or negative nrows was given",
          3: "invalid input syntax for type numeric"}


class FinalFnError(RuntimeError):
    def __init__(self, code):
        self.code = code
        super().__init__(ERRORS.get(code, "error %d" % code))


def _check(rc):
    if rc != 0:
        raise FinalFnError(rc)


def partial_nrows(*args):
    """pgstrom.nrows(bool, ...): args are True / False / None."""
    lib = _capi.load()
    vals = bytes(1 if a else 0 for a in args)
    nulls = bytes(1 if a is None else 0 for a in args)
    return lib.pgs_partial_nrows(len(args), vals, nulls)


def psum_x2(x):
    lib = _capi.load()
    out = C.c_double()
    isnull = lib.pgs_psum_x2_float8(0.0 if x is None else x, x is None, C.byref(out))
    return None if isnull else out.value


def pcov(kind, filt, x, y):
    """pgstrom.pcov_{x,y,x2,y2,xy}(bool, float8, float8)."""
    lib = _capi.load()
    out = C.c_double()
    isnull = lib.pgs_pcov_float8(
        ("x", "y", "x2", "y2", "xy").index(kind), bool(filt), filt is None,
        0.0 if x is None else x, x is None, 0.0 if y is None else y, y is None, C.byref(out))
    return None if isnull else out.value


class FinalAccum:
    """Transition state of one `pgstrom.*` final aggregate.  Partial columns
    arrive positionally as the planner half lays them out:
      sum(int2/int4)      [psum]                    pgstrom.sum(int8)       int8[2]
      avg(int2/int4)      [nrows, psum]             pgstrom.avg(int4,int8)  int8[2]
      avg(int8/numeric)   [nrows, psum]             pgstrom.avg(int4,numeric) internal
      sum/avg(float)      [nrows, psum] / [psum]    float8[3]
      stddev/variance     [nrows, psum, psum_x2]    float8[3]
      corr/covar_*        [nrows, pcov_x, pcov_x2, pcov_y, pcov_y2, pcov_xy]  float8[6]
    count / min / max / sum(int8, float, numeric) are merged by PostgreSQL's
    own sum / min / max (gpupreagg.c:134-333) and have no pgstrom.* function.
    """

    def __init__(self, agg, argtypes):
        self.lib = _capi.load()
        self.agg = agg
        t = argtypes[0] if argtypes else None
        self.t = t
        self.num = None
        if agg == "sum" and t in ("int2", "int4"):
            self.kind, self.trans = "sum_int8", (C.c_int64 * 2)(0, 0)
        elif agg == "avg" and t in ("int2", "int4"):
            self.kind, self.trans = "avg_int8", (C.c_int64 * 2)(0, 0)
        elif agg == "avg" and t in ("int8", "numeric"):
            self.kind, self.trans = "avg_numeric", None
            self.num = C.c_void_p(self.lib.pgs_numeric_avg_init())
        elif agg == "avg" and t in ("float4", "float8"):
            self.kind, self.trans = "sum_float8", (C.c_double * 3)(0.0, 0.0, 0.0)
        elif agg in VARIANCE_AGGS and t in ("float4", "float8"):
            self.kind, self.trans = "variance_float8", (C.c_double * 3)(0.0, 0.0, 0.0)
        elif agg in COVARIANCE_AGGS:
            self.kind, self.trans = "covariance_float8", (C.c_double * 6)(*([0.0] * 6))
        else:
            raise KeyError("%s(%s) has no pgstrom.* final aggregate" % (agg, t))

    def __del__(self):
        if getattr(self, "num", None):
            self.lib.pgs_numeric_avg_free(self.num)
            self.num = None

    def accum(self, p):
        k = self.kind
        if k == "avg_numeric":
            # not STRICT (pg_strom--1.0.sql:271-285): NULL psum is skipped inside
            nrows, psum = p
            text = None if psum is None else format(Decimal(psum), "f").encode()
            _check(self.lib.pgs_numeric_avg_accum(self.num, 0 if nrows is None else int(nrows),
                                                  nrows is None, text))
            return
        if any(x is None for x in p):
            return                      # STRICT transition functions
        if k == "sum_int8":
            _check(self.lib.pgs_sum_int8_accum(self.trans, int(p[0])))
        elif k == "avg_int8":
            _check(self.lib.pgs_avg_int8_accum(self.trans, int(p[0]), int(p[1])))
        elif k == "sum_float8":
            _check(self.lib.pgs_sum_float8_accum(self.trans, int(p[0]), float(p[1])))
        elif k == "variance_float8":
            _check(self.lib.pgs_variance_float8_accum(self.trans, int(p[0]), float(p[1]),
                                                      float(p[2])))
        else:
            ps = (C.c_double * 5)(*[float(x) for x in p[1:6]])
            _check(self.lib.pgs_covariance_float8_accum(self.trans, int(p[0]), ps))

    def state(self):
        if self.kind == "avg_numeric":
            buf = C.create_string_buffer(1 << 16)
            n = self.lib.pgs_numeric_avg_sum_text(self.num, buf, len(buf))
            assert n > 0
            return [self.lib.pgs_numeric_avg_count(self.num), Decimal(buf.value.decode())]
        return list(self.trans)

    def sum_int8_final(self):
        """pgstrom_sum_int8_final (gpupreagg.c:4508): None = NULL."""
        assert self.kind == "sum_int8"
        out = C.c_int64()
        return None if self.lib.pgs_sum_int8_final(self.trans, C.byref(out)) else out.value
