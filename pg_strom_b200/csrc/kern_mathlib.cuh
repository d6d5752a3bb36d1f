/*
 * kern_mathlib.cuh - arithmetic operators, casts and math functions of the
 * device runtime with PostgreSQL's overflow / division-by-zero / domain
 * checks (the counterpart of the reference's opencl_mathlib.h:34-818 and the
 * cast functions of opencl_common.h).  Included by kern_common.cuh; plain C++
 * over the pg_<type>_t structs, so tests/native/mathlib_host_shim.cpp can
 * compile it with g++ and run it against the oracle on the CPU.
 */
#ifndef KERN_MATHLIB_CUH
#define KERN_MATHLIB_CUH

/* ------------------------------------------------------------------
 * mathlib: PostgreSQL-compatible overflow / division-by-zero detection.
 * The host raises the error after re-checking the row, so the device
 * result is NULL + CpuReCheck (opencl_mathlib.h).
 * ------------------------------------------------------------------ */
#define CHECKFLOATVAL(val, inf_is_valid, zero_is_valid)         \
    ((isinf(val) && !(inf_is_valid)) ||                         \
     ((val) == 0.0 && !(zero_is_valid)))
#define SAMESIGN(a,b)   (((a) < 0) == ((b) < 0))

#define PGS_MATH_FAIL(result)                                   \
    do {                                                        \
        (result).isnull = true;                                 \
        (result).value = 0;                                     \
        STROM_SET_ERROR(errcode, StromError_CpuReCheck);        \
    } while (0)

/* integer +,-,* : compute in the next wider type and range-check (int2/int4),
 * sign rules for int8 */
#define BASIC_INT_ARITH_NARROW(name,op,r_type,R_BASE,R_MIN,R_MAX,x_type,y_type) \
    DEVFN pg_##r_type##_t                                               \
    pgfn_##name(cl_int *errcode, pg_##x_type##_t arg1, pg_##y_type##_t arg2) \
    {                                                                   \
        pg_##r_type##_t result;                                         \
        result.value = 0;                                               \
        result.isnull = arg1.isnull | arg2.isnull;                      \
        if (!result.isnull)                                             \
        {                                                               \
            cl_long v = (cl_long)arg1.value op (cl_long)arg2.value;     \
            if (v < (cl_long)(R_MIN) || v > (cl_long)(R_MAX))           \
                PGS_MATH_FAIL(result);                                  \
            else                                                        \
                result.value = (R_BASE)v;                               \
        }                                                               \
        return result;                                                  \
    }

#define BASIC_INT8_ADDSUB(name,is_sub,x_type,y_type)                    \
    DEVFN pg_int8_t                                                     \
    pgfn_##name(cl_int *errcode, pg_##x_type##_t arg1, pg_##y_type##_t arg2) \
    {                                                                   \
        pg_int8_t result;                                               \
        result.value = 0;                                               \
        result.isnull = arg1.isnull | arg2.isnull;                      \
        if (!result.isnull)                                             \
        {                                                               \
            cl_long a = (cl_long)arg1.value;                            \
            cl_long b = (cl_long)arg2.value;                            \
            cl_long r;                                                  \
            bool ovf;                                                   \
            if (is_sub)                                                 \
            {                                                           \
                r = (cl_long)((cl_ulong)a - (cl_ulong)b);               \
                ovf = (!SAMESIGN(a, b) && !SAMESIGN(r, a));             \
            }                                                           \
            else                                                        \
            {                                                           \
                r = (cl_long)((cl_ulong)a + (cl_ulong)b);               \
                ovf = (SAMESIGN(a, b) && !SAMESIGN(r, a));              \
            }                                                           \
            if (ovf)                                                    \
                PGS_MATH_FAIL(result);                                  \
            else                                                        \
                result.value = r;                                       \
        }                                                               \
        return result;                                                  \
    }

#define BASIC_INT8_MUL(name,x_type,y_type)                              \
    DEVFN pg_int8_t                                                     \
    pgfn_##name(cl_int *errcode, pg_##x_type##_t arg1, pg_##y_type##_t arg2) \
    {                                                                   \
        pg_int8_t result;                                               \
        result.value = 0;                                               \
        result.isnull = arg1.isnull | arg2.isnull;                      \
        if (!result.isnull)                                             \
        {                                                               \
            cl_long a = (cl_long)arg1.value;                            \
            cl_long b = (cl_long)arg2.value;                            \
            cl_long hi = __mul64hi(a, b);                               \
            cl_long lo = (cl_long)((cl_ulong)a * (cl_ulong)b);          \
            if (hi != (lo >> 63))                                       \
                PGS_MATH_FAIL(result);                                  \
            else                                                        \
                result.value = lo;                                      \
        }                                                               \
        return result;                                                  \
    }

#define BASIC_FLOAT_ARITH(name,op,r_type,R_BASE,x_type,y_type,zero_ok)  \
    DEVFN pg_##r_type##_t                                               \
    pgfn_##name(cl_int *errcode, pg_##x_type##_t arg1, pg_##y_type##_t arg2) \
    {                                                                   \
        pg_##r_type##_t result;                                         \
        result.value = 0;                                               \
        result.isnull = arg1.isnull | arg2.isnull;                      \
        if (!result.isnull)                                             \
        {                                                               \
            R_BASE a = (R_BASE)arg1.value;                              \
            R_BASE b = (R_BASE)arg2.value;                              \
            R_BASE r = a op b;                                          \
            if (CHECKFLOATVAL(r, isinf(a) || isinf(b), zero_ok))        \
                PGS_MATH_FAIL(result);                                  \
            else                                                        \
                result.value = r;                                       \
        }                                                               \
        return result;                                                  \
    }

/* '+' */
BASIC_INT_ARITH_NARROW(int2pl,  +, int2, cl_short, SHRT_MIN, SHRT_MAX, int2, int2)
BASIC_INT_ARITH_NARROW(int24pl, +, int4, cl_int,   INT_MIN,  INT_MAX,  int2, int4)
BASIC_INT_ARITH_NARROW(int42pl, +, int4, cl_int,   INT_MIN,  INT_MAX,  int4, int2)
BASIC_INT_ARITH_NARROW(int4pl,  +, int4, cl_int,   INT_MIN,  INT_MAX,  int4, int4)
BASIC_INT8_ADDSUB(int28pl, false, int2, int8)
BASIC_INT8_ADDSUB(int48pl, false, int4, int8)
BASIC_INT8_ADDSUB(int82pl, false, int8, int2)
BASIC_INT8_ADDSUB(int84pl, false, int8, int4)
BASIC_INT8_ADDSUB(int8pl,  false, int8, int8)
BASIC_FLOAT_ARITH(float4pl,  +, float4, cl_float,  float4, float4, true)
BASIC_FLOAT_ARITH(float48pl, +, float8, cl_double, float4, float8, true)
BASIC_FLOAT_ARITH(float84pl, +, float8, cl_double, float8, float4, true)
BASIC_FLOAT_ARITH(float8pl,  +, float8, cl_double, float8, float8, true)
/* '-' */
BASIC_INT_ARITH_NARROW(int2mi,  -, int2, cl_short, SHRT_MIN, SHRT_MAX, int2, int2)
BASIC_INT_ARITH_NARROW(int24mi, -, int4, cl_int,   INT_MIN,  INT_MAX,  int2, int4)
BASIC_INT_ARITH_NARROW(int42mi, -, int4, cl_int,   INT_MIN,  INT_MAX,  int4, int2)
BASIC_INT_ARITH_NARROW(int4mi,  -, int4, cl_int,   INT_MIN,  INT_MAX,  int4, int4)
BASIC_INT8_ADDSUB(int28mi, true, int2, int8)
BASIC_INT8_ADDSUB(int48mi, true, int4, int8)
BASIC_INT8_ADDSUB(int82mi, true, int8, int2)
BASIC_INT8_ADDSUB(int84mi, true, int8, int4)
BASIC_INT8_ADDSUB(int8mi,  true, int8, int8)
BASIC_FLOAT_ARITH(float4mi,  -, float4, cl_float,  float4, float4, true)
BASIC_FLOAT_ARITH(float48mi, -, float8, cl_double, float4, float8, true)
BASIC_FLOAT_ARITH(float84mi, -, float8, cl_double, float8, float4, true)
BASIC_FLOAT_ARITH(float8mi,  -, float8, cl_double, float8, float8, true)
/* '*' */
BASIC_INT_ARITH_NARROW(int2mul,  *, int2, cl_short, SHRT_MIN, SHRT_MAX, int2, int2)
BASIC_INT_ARITH_NARROW(int24mul, *, int4, cl_int,   INT_MIN,  INT_MAX,  int2, int4)
BASIC_INT_ARITH_NARROW(int42mul, *, int4, cl_int,   INT_MIN,  INT_MAX,  int4, int2)
BASIC_INT_ARITH_NARROW(int4mul,  *, int4, cl_int,   INT_MIN,  INT_MAX,  int4, int4)
BASIC_INT8_MUL(int28mul, int2, int8)
BASIC_INT8_MUL(int48mul, int4, int8)
BASIC_INT8_MUL(int82mul, int8, int2)
BASIC_INT8_MUL(int84mul, int8, int4)
BASIC_INT8_MUL(int8mul,  int8, int8)
BASIC_FLOAT_ARITH(float4mul,  *, float4, cl_float,  float4, float4, (a == 0 || b == 0))
BASIC_FLOAT_ARITH(float48mul, *, float8, cl_double, float4, float8, (a == 0 || b == 0))
BASIC_FLOAT_ARITH(float84mul, *, float8, cl_double, float8, float4, (a == 0 || b == 0))
BASIC_FLOAT_ARITH(float8mul,  *, float8, cl_double, float8, float8, (a == 0 || b == 0))

/* '/' : division by zero => re-check (the host raises "division by zero");
 * INT_MIN / -1 overflows */
#define BASIC_INT_DIV(name,r_type,R_BASE,R_MIN,x_type,y_type)           \
    DEVFN pg_##r_type##_t                                               \
    pgfn_##name(cl_int *errcode, pg_##x_type##_t arg1, pg_##y_type##_t arg2) \
    {                                                                   \
        pg_##r_type##_t result;                                         \
        result.value = 0;                                               \
        result.isnull = arg1.isnull | arg2.isnull;                      \
        if (!result.isnull)                                             \
        {                                                               \
            cl_long a = (cl_long)arg1.value;                            \
            cl_long b = (cl_long)arg2.value;                            \
            if (b == 0 || (b == -1 && a == (cl_long)(R_MIN)))           \
                PGS_MATH_FAIL(result);                                  \
            else                                                        \
                result.value = (R_BASE)(a / b);                         \
        }                                                               \
        return result;                                                  \
    }
BASIC_INT_DIV(int2div,  int2, cl_short, SHRT_MIN, int2, int2)
BASIC_INT_DIV(int24div, int4, cl_int,   INT_MIN,  int2, int4)
BASIC_INT_DIV(int28div, int8, cl_long,  LONG_MIN, int2, int8)
BASIC_INT_DIV(int42div, int4, cl_int,   INT_MIN,  int4, int2)
BASIC_INT_DIV(int4div,  int4, cl_int,   INT_MIN,  int4, int4)
BASIC_INT_DIV(int48div, int8, cl_long,  LONG_MIN, int4, int8)
BASIC_INT_DIV(int82div, int8, cl_long,  LONG_MIN, int8, int2)
BASIC_INT_DIV(int84div, int8, cl_long,  LONG_MIN, int8, int4)
BASIC_INT_DIV(int8div,  int8, cl_long,  LONG_MIN, int8, int8)

#define BASIC_FLOAT_DIV(name,r_type,R_BASE,x_type,y_type)               \
    DEVFN pg_##r_type##_t                                               \
    pgfn_##name(cl_int *errcode, pg_##x_type##_t arg1, pg_##y_type##_t arg2) \
    {                                                                   \
        pg_##r_type##_t result;                                         \
        result.value = 0;                                               \
        result.isnull = arg1.isnull | arg2.isnull;                      \
        if (!result.isnull)                                             \
        {                                                               \
            R_BASE a = (R_BASE)arg1.value;                              \
            R_BASE b = (R_BASE)arg2.value;                              \
            R_BASE r;                                                   \
            if (b == 0.0)                                               \
                PGS_MATH_FAIL(result);                                  \
            else                                                        \
            {                                                           \
                r = a / b;                                              \
                if (CHECKFLOATVAL(r, isinf(a) || isinf(b), a == 0.0))   \
                    PGS_MATH_FAIL(result);                              \
                else                                                    \
                    result.value = r;                                   \
            }                                                           \
        }                                                               \
        return result;                                                  \
    }
BASIC_FLOAT_DIV(float4div,  float4, cl_float,  float4, float4)
BASIC_FLOAT_DIV(float48div, float8, cl_double, float4, float8)
BASIC_FLOAT_DIV(float84div, float8, cl_double, float8, float4)
BASIC_FLOAT_DIV(float8div,  float8, cl_double, float8, float8)

/* '%' : x % -1 is defined as 0 (avoids INT_MIN % -1 trap semantics) */
#define BASIC_INT_MODFUNC_TEMPLATE(name,d_type,D_BASE)                  \
    DEVFN pg_##d_type##_t                                               \
    pgfn_##name(cl_int *errcode, pg_##d_type##_t arg1, pg_##d_type##_t arg2) \
    {                                                                   \
        pg_##d_type##_t result;                                         \
        result.value = 0;                                               \
        result.isnull = arg1.isnull | arg2.isnull;                      \
        if (!result.isnull)                                             \
        {                                                               \
            if (arg2.value == 0)                                        \
                PGS_MATH_FAIL(result);                                  \
            else if (arg2.value == -1)                                  \
                result.value = 0;                                       \
            else                                                        \
                result.value = (D_BASE)(arg1.value % arg2.value);       \
        }                                                               \
        return result;                                                  \
    }
BASIC_INT_MODFUNC_TEMPLATE(int2mod, int2, cl_short)
BASIC_INT_MODFUNC_TEMPLATE(int4mod, int4, cl_int)
BASIC_INT_MODFUNC_TEMPLATE(int8mod, int8, cl_long)

/* unary minus / abs on integers overflow at the minimum value */
#define BASIC_INT_UNARY(name,d_type,D_BASE,D_MIN,expr)                  \
    DEVFN pg_##d_type##_t                                               \
    pgfn_##name(cl_int *errcode, pg_##d_type##_t arg)                   \
    {                                                                   \
        pg_##d_type##_t result;                                         \
        result.value = 0;                                               \
        result.isnull = arg.isnull;                                     \
        if (!result.isnull)                                             \
        {                                                               \
            if (arg.value == (D_BASE)(D_MIN))                           \
                PGS_MATH_FAIL(result);                                  \
            else                                                        \
                result.value = (D_BASE)(expr);                          \
        }                                                               \
        return result;                                                  \
    }
BASIC_INT_UNARY(int2um, int2, cl_short, SHRT_MIN, -arg.value)
BASIC_INT_UNARY(int4um, int4, cl_int,   INT_MIN,  -arg.value)
BASIC_INT_UNARY(int8um, int8, cl_long,  LONG_MIN, -arg.value)
BASIC_INT_UNARY(int2abs, int2, cl_short, SHRT_MIN, (arg.value < 0 ? -arg.value : arg.value))
BASIC_INT_UNARY(int4abs, int4, cl_int,   INT_MIN,  (arg.value < 0 ? -arg.value : arg.value))
BASIC_INT_UNARY(int8abs, int8, cl_long,  LONG_MIN, (arg.value < 0 ? -arg.value : arg.value))

DEVFN pg_float8_t
pgfn_dpi(cl_int *errcode)
{
    pg_float8_t result;
    result.isnull = false;
    result.value = 3.14159265358979323846;
    return result;
}

/* float.c dsign(), degrees(), radians() */
DEVFN pg_float8_t
pgfn_dsign(cl_int *errcode, pg_float8_t arg1)
{
    pg_float8_t result;

    result.isnull = arg1.isnull;
    result.value = (arg1.isnull ? 0.0 :
                    arg1.value > 0.0 ? 1.0 : (arg1.value < 0.0 ? -1.0 : 0.0));
    return result;
}

#define PGS_FLOAT8_SCALE_TEMPLATE(name,FACTOR)                          \
    DEVFN pg_float8_t                                                   \
    pgfn_##name(cl_int *errcode, pg_float8_t arg1)                      \
    {                                                                   \
        pg_float8_t result;                                             \
                                                                        \
        result.value = 0.0;                                             \
        result.isnull = arg1.isnull;                                    \
        if (!arg1.isnull)                                               \
        {                                                               \
            double r = arg1.value * (FACTOR);                           \
            if (CHECKFLOATVAL(r, isinf(arg1.value), arg1.value == 0.0)) \
                PGS_MATH_FAIL(result);                                  \
            else                                                        \
                result.value = r;                                       \
        }                                                               \
        return result;                                                  \
    }
PGS_FLOAT8_SCALE_TEMPLATE(degrees, 180.0 / 3.14159265358979323846)
PGS_FLOAT8_SCALE_TEMPLATE(radians, 3.14159265358979323846 / 180.0)

DEVFN pg_float8_t
pgfn_dpow(cl_int *errcode, pg_float8_t arg1, pg_float8_t arg2)
{
    pg_float8_t result;

    result.value = 0;
    result.isnull = arg1.isnull | arg2.isnull;
    if (!result.isnull)
    {
        /* float.c dpow(): 0 ^ negative and negative ^ non-integer are errors */
        if ((arg1.value == 0.0 && arg2.value < 0.0) ||
            (arg1.value < 0.0 && floor(arg2.value) != arg2.value))
            PGS_MATH_FAIL(result);
        else
        {
            double r = pow(arg1.value, arg2.value);
            if (CHECKFLOATVAL(r, isinf(arg1.value) || isinf(arg2.value),
                              arg1.value == 0.0))
                PGS_MATH_FAIL(result);
            else
                result.value = r;
        }
    }
    return result;
}

/* ------------------------------------------------------------------
 * type casts between the basic numeric types with PostgreSQL's range
 * checks (int84(), dtoi4(), ftoi2(), dtof() ...): float -> int rounds half
 * to even (rint), out of range => re-check.
 * ------------------------------------------------------------------ */
#define CAST_INT_NARROW(name,r_type,R_BASE,R_MIN,R_MAX,x_type)          \
    DEVFN pg_##r_type##_t                                               \
    pgfn_##name(cl_int *errcode, pg_##x_type##_t arg)                   \
    {                                                                   \
        pg_##r_type##_t result;                                         \
        result.value = 0;                                               \
        result.isnull = arg.isnull;                                     \
        if (!result.isnull)                                             \
        {                                                               \
            if ((cl_long)arg.value < (cl_long)(R_MIN) ||                \
                (cl_long)arg.value > (cl_long)(R_MAX))                  \
                PGS_MATH_FAIL(result);                                  \
            else                                                        \
                result.value = (R_BASE)arg.value;                       \
        }                                                               \
        return result;                                                  \
    }
#define CAST_SIMPLE(name,r_type,R_BASE,x_type)                          \
    DEVFN pg_##r_type##_t                                               \
    pgfn_##name(cl_int *errcode, pg_##x_type##_t arg)                   \
    {                                                                   \
        pg_##r_type##_t result;                                         \
        result.value = (R_BASE)arg.value;                               \
        result.isnull = arg.isnull;                                     \
        return result;                                                  \
    }
/* float -> int: rint() then range check on the floating value */
#define CAST_FLOAT_INT(name,r_type,R_BASE,LOWER,UPPER,x_type)           \
    DEVFN pg_##r_type##_t                                               \
    pgfn_##name(cl_int *errcode, pg_##x_type##_t arg)                   \
    {                                                                   \
        pg_##r_type##_t result;                                         \
        result.value = 0;                                               \
        result.isnull = arg.isnull;                                     \
        if (!result.isnull)                                             \
        {                                                               \
            double r = rint((double)arg.value);                         \
            if (isnan(r) || r < (LOWER) || r >= (UPPER))                \
                PGS_MATH_FAIL(result);                                  \
            else                                                        \
                result.value = (R_BASE)r;                               \
        }                                                               \
        return result;                                                  \
    }
CAST_INT_NARROW(int4_int2, int2, cl_short, SHRT_MIN, SHRT_MAX, int4)
CAST_INT_NARROW(int8_int2, int2, cl_short, SHRT_MIN, SHRT_MAX, int8)
CAST_INT_NARROW(int8_int4, int4, cl_int,   INT_MIN,  INT_MAX,  int8)
CAST_FLOAT_INT(float4_int2, int2, cl_short, -32768.0, 32768.0, float4)
CAST_FLOAT_INT(float8_int2, int2, cl_short, -32768.0, 32768.0, float8)
CAST_FLOAT_INT(float4_int4, int4, cl_int, -2147483648.0, 2147483648.0, float4)
CAST_FLOAT_INT(float8_int4, int4, cl_int, -2147483648.0, 2147483648.0, float8)
CAST_FLOAT_INT(float4_int8, int8, cl_long, -9223372036854775808.0, 9223372036854775808.0, float4)
CAST_FLOAT_INT(float8_int8, int8, cl_long, -9223372036854775808.0, 9223372036854775808.0, float8)
CAST_SIMPLE(bool_int4, int4, cl_int,   bool)
CAST_SIMPLE(int2_int4, int4, cl_int,   int2)
CAST_SIMPLE(int2_int8, int8, cl_long,  int2)
CAST_SIMPLE(int4_int8, int8, cl_long,  int4)
CAST_SIMPLE(int2_float4, float4, cl_float, int2)
CAST_SIMPLE(int4_float4, float4, cl_float, int4)
CAST_SIMPLE(int8_float4, float4, cl_float, int8)
CAST_SIMPLE(int2_float8, float8, cl_double, int2)
CAST_SIMPLE(int4_float8, float8, cl_double, int4)
CAST_SIMPLE(int8_float8, float8, cl_double, int8)
CAST_SIMPLE(float4_float8, float8, cl_double, float4)
/* dtof(): overflow / underflow are errors */
DEVFN pg_float4_t
pgfn_float8_float4(cl_int *errcode, pg_float8_t arg)
{
    pg_float4_t result;

    result.value = 0;
    result.isnull = arg.isnull;
    if (!result.isnull)
    {
        float r = (float)arg.value;
        if (CHECKFLOATVAL(r, isinf(arg.value), arg.value == 0.0))
            PGS_MATH_FAIL(result);
        else
            result.value = r;
    }
    return result;
}

#endif  /* KERN_MATHLIB_CUH */
