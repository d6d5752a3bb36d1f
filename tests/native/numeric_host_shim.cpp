/*
 * Test shim: compiles pg_strom_b200/csrc/kern_numeric.cuh (device code, plain
 * C++ with __int128) with g++ so that the numeric conversion and arithmetic
 * the kernels run can be checked on the CPU against python's Decimal
 * (tests/test_numeric_device_code.py).  Test infrastructure only.
 */
#include <cstdint>
#include <cstring>
#include "pgstrom_kds.h"
#define DEVFN static inline
#ifndef LONG_MAX
#define LONG_MAX    9223372036854775807LL
#define LONG_MIN    (-LONG_MAX-1LL)
#endif
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
typedef struct { cl_bool value; bool isnull; } pg_bool_t;
typedef struct { cl_short value; bool isnull; } pg_int2_t;
typedef struct { cl_int value; bool isnull; } pg_int4_t;
typedef struct { cl_long value; bool isnull; } pg_int8_t;
typedef struct { cl_float value; bool isnull; } pg_float4_t;
typedef struct { cl_double value; bool isnull; } pg_float8_t;
#define STROMCL_SIMPLE_NULLTEST_TEMPLATE(NAME)
#define GPUPREAGG_INCOL_SLOT(colidx) 0
#include "kern_numeric.cuh"

extern "C" {
/* returns the errcode; *out = packed device numeric */
int shim_from_varlena(const unsigned char *p, uint64_t *out, int *isnull)
{
    cl_int e = 0;
    pg_numeric_t r = pg_numeric_from_varlena(&e, p);
    *out = r.value; *isnull = r.isnull;
    return e;
}
int shim_cmp(uint64_t a, uint64_t b) { return pgs_numeric_cmp(a, b); }
int shim_binop(int op, uint64_t a, uint64_t b, uint64_t *out, int *isnull)
{
    cl_int e = 0;
    pg_numeric_t x, y, r;
    x.value = a; x.isnull = false; y.value = b; y.isnull = false;
    r = (op == 0 ? pgfn_numeric_add(&e, x, y) : op == 1 ? pgfn_numeric_sub(&e, x, y)
                                                        : pgfn_numeric_mul(&e, x, y));
    *out = r.value; *isnull = r.isnull;
    return e;
}
double shim_to_float8(uint64_t a)
{
    cl_int e = 0; pg_numeric_t x; x.value = a; x.isnull = false;
    return pgfn_numeric_float8(&e, x).value;
}
int shim_to_int8(uint64_t a, int64_t *out)
{
    cl_int e = 0; pg_numeric_t x; x.value = a; x.isnull = false;
    pg_int8_t r = pgfn_numeric_int8(&e, x);
    *out = r.value;
    return r.isnull ? -1 : e;
}
}
