/*
 * kern_common.cuh - device-side runtime shared by every generated GpuPreAgg
 * program.  Compiled by NVRTC for sm_100a together with pgstrom_kds.h, the
 * codegen output and kern_gpupreagg.cuh (no libc / libcu++ headers here).
 *
 * What it restates from the reference (behaviour, not text):
 *   STROM_SET_ERROR priority rule           opencl_common.h:132-144
 *   pg_<type>_t / _vref / _param / nulltest opencl_common.h:530-670
 *   EVAL + BooleanTest + NOT                opencl_common.h:1539-1622
 *   devfunc_int_comp / devfunc_float_comp   opencl_common.h:1550-1560
 *   overflow-checked arithmetic             opencl_mathlib.h:34-853
 *     (every overflow / division by zero => NULL + StromError_CpuReCheck)
 * The column accessors read a *tile view* (shared-memory staging filled by
 * cp.async.bulk, or global memory for the row-map gather path) instead of
 * walking heap tuples (opencl_common.h:817-981).
 */
#ifndef KERN_COMMON_CUH
#define KERN_COMMON_CUH

#define DEVFN   static __device__ __forceinline__

#ifndef NULL
#define NULL    0
#endif
#ifndef INT_MAX
#define SHRT_MAX    32767
#define SHRT_MIN    (-32767-1)
#define INT_MAX     2147483647
#define INT_MIN     (-INT_MAX-1)
#define LONG_MAX    9223372036854775807LL
#define LONG_MIN    (-LONG_MAX-1LL)
#endif
#ifndef DBL_MAX
#define DBL_MAX     1.7976931348623157e+308
#define FLT_MAX     3.402823466e+38F
#endif

/* dynamic shared memory of the CTA; every staged access goes through this
 * symbol so that the compiler emits LDS/ATOMS, not generic accesses */
extern __shared__ __align__(1024) unsigned char __pgs_smem[];

/*
 * It sets an error code unless a significant error code is already set.
 * CpuReCheck outranks RowFiltered.
 */
DEVFN void
STROM_SET_ERROR(cl_int *p_error, cl_int errcode)
{
    cl_int  oldcode = *p_error;

    if (StromErrorIsSignificant(errcode))
    {
        if (!StromErrorIsSignificant(oldcode))
            *p_error = errcode;
    }
    else if (errcode > oldcode)
        *p_error = errcode;
}

/* ------------------------------------------------------------------
 * tile views
 *
 * A generated function receives `kds` as one of these and never looks at
 * the chunk header: column `colidx` of the outer relation is staged in slot
 * GPUPREAGG_INCOL_SLOT(colidx) (a constexpr switch emitted by codegen).
 * ------------------------------------------------------------------ */
#define KERN_TILE_NO_NULLMAP    0xffffffffU

struct kern_tile_smem
{
    cl_uint     row0;                           /* first row of the tile */
    cl_uint     val_off[GPUPREAGG_NUM_INCOLS];  /* smem offset of values */
    cl_uint     nul_off[GPUPREAGG_NUM_INCOLS];  /* smem offset of bitmap */

    template <typename T>
    __device__ __forceinline__ bool
    fetch(int slot, cl_uint rowidx, T &out) const
    {
        cl_uint i = rowidx - row0;

        /* every staged array starts on a 128-byte boundary: lets the
         * compiler merge the loads of adjacent rows into LDS.64 / LDS.128 */
        __builtin_assume((val_off[slot] & 127U) == 0);
        /* the value slot of a NULL row holds padding that is safe to read:
         * load first (unconditional loads of adjacent rows vectorise), then
         * look at the validity bit */
        out = *((const T *)(__pgs_smem + val_off[slot]) + i);
        if (nul_off[slot] != KERN_TILE_NO_NULLMAP)
        {
            cl_uint w = *((const cl_uint *)(__pgs_smem + nul_off[slot]) + (i >> 5));
            if (((w >> (i & 31)) & 1U) == 0)
                return false;
        }
        return true;
    }
};

struct kern_tile_gmem
{
    const char     *val_ptr[GPUPREAGG_NUM_INCOLS];  /* column arrays in HBM */
    const cl_uint  *nul_ptr[GPUPREAGG_NUM_INCOLS];  /* NULL = no bitmap */

    template <typename T>
    __device__ __forceinline__ bool
    fetch(int slot, cl_uint rowidx, T &out) const
    {
        if (nul_ptr[slot])
        {
            cl_uint w = __ldg(nul_ptr[slot] + (rowidx >> 5));
            if (((w >> (rowidx & 31)) & 1U) == 0)
                return false;
        }
        out = __ldg((const T *)val_ptr[slot] + rowidx);
        return true;
    }
};

/*
 * kern_row_regs - one row whose staged columns were already pulled into
 * registers (the staged kernel loads PGS_ROWS_PER_THREAD adjacent rows with
 * 128-bit shared memory loads first, then evaluates them one by one, so the
 * generated code never does address arithmetic).
 */
struct kern_row_regs
{
    cl_ulong    v[GPUPREAGG_NUM_INCOLS > 0 ? GPUPREAGG_NUM_INCOLS : 1];
    /* bit `shift` = validity (NOT NULL) of this row, per staged column */
    cl_uint     vbits[GPUPREAGG_NUM_INCOLS > 0 ? GPUPREAGG_NUM_INCOLS : 1];
    int         shift;

    template <typename T>
    __device__ __forceinline__ bool
    fetch(int slot, cl_uint rowidx, T &out) const
    {
        union { cl_ulong u; T t; } cv;
        cv.u = v[slot];
        out = cv.t;
        return (vbits[slot] & (1U << shift)) != 0;
    }
};

/* loads of PGS_ROWS_PER_THREAD (= 4) adjacent values of one staged column;
 * `p` is 16-byte aligned for every attlen because the row index is a
 * multiple of 4 ... for attlen >= 4; narrower columns use narrower loads */
template <int ATTLEN>
struct pgs_rowload;
template <>
struct pgs_rowload<8>
{
    static __device__ __forceinline__ void
    load4(const unsigned char *p, cl_ulong &a, cl_ulong &b, cl_ulong &c, cl_ulong &d)
    {
        ulonglong2 lo = *((const ulonglong2 *)p);
        ulonglong2 hi = *((const ulonglong2 *)p + 1);
        a = lo.x; b = lo.y; c = hi.x; d = hi.y;
    }
};
template <>
struct pgs_rowload<4>
{
    static __device__ __forceinline__ void
    load4(const unsigned char *p, cl_ulong &a, cl_ulong &b, cl_ulong &c, cl_ulong &d)
    {
        uint4 q = *((const uint4 *)p);
        a = q.x; b = q.y; c = q.z; d = q.w;
    }
};
template <>
struct pgs_rowload<2>
{
    static __device__ __forceinline__ void
    load4(const unsigned char *p, cl_ulong &a, cl_ulong &b, cl_ulong &c, cl_ulong &d)
    {
        uint2 q = *((const uint2 *)p);
        a = q.x & 0xffffU; b = q.x >> 16; c = q.y & 0xffffU; d = q.y >> 16;
    }
};
template <>
struct pgs_rowload<1>
{
    static __device__ __forceinline__ void
    load4(const unsigned char *p, cl_ulong &a, cl_ulong &b, cl_ulong &c, cl_ulong &d)
    {
        cl_uint q = *((const cl_uint *)p);
        a = q & 0xffU; b = (q >> 8) & 0xffU; c = (q >> 16) & 0xffU; d = q >> 24;
    }
};

/* ------------------------------------------------------------------
 * PostgreSQL data types on the device: { BASE value; bool isnull; }
 * ------------------------------------------------------------------ */
#define STROMCL_SIMPLE_DATATYPE_TEMPLATE(NAME,BASE)             \
    typedef struct {                                            \
        BASE    value;                                          \
        bool    isnull;                                         \
    } pg_##NAME##_t;

#define STROMCL_SIMPLE_VARREF_TEMPLATE(NAME,BASE)               \
    template <typename KDS>                                     \
    DEVFN pg_##NAME##_t                                         \
    pg_##NAME##_vref(const KDS &kds, const void *ktoast,        \
                     cl_int *errcode,                           \
                     cl_uint colidx, cl_uint rowidx)            \
    {                                                           \
        pg_##NAME##_t result;                                   \
        result.value = 0;                                       \
        result.isnull = !kds.template fetch<BASE>(              \
            GPUPREAGG_INCOL_SLOT(colidx), rowidx, result.value);\
        return result;                                          \
    }

#define STROMCL_SIMPLE_PARAMREF_TEMPLATE(NAME,BASE)             \
    DEVFN pg_##NAME##_t                                         \
    pg_##NAME##_param(const kern_parambuf *kparams,             \
                      cl_int *errcode, cl_uint param_id)        \
    {                                                           \
        pg_##NAME##_t result;                                   \
        if (param_id < kparams->nparams &&                      \
            kparams->poffset[param_id] > 0)                     \
        {                                                       \
            result.value = *((const BASE *)                     \
                             ((const char *)kparams +           \
                              kparams->poffset[param_id]));     \
            result.isnull = false;                              \
        }                                                       \
        else                                                    \
        {                                                       \
            result.value = 0;                                   \
            result.isnull = true;                               \
        }                                                       \
        return result;                                          \
    }

#define STROMCL_SIMPLE_NULLTEST_TEMPLATE(NAME)                  \
    DEVFN pg_bool_t                                             \
    pgfn_##NAME##_isnull(cl_int *errcode, pg_##NAME##_t arg)    \
    {                                                           \
        pg_bool_t result;                                       \
        result.isnull = false;                                  \
        result.value = arg.isnull;                              \
        return result;                                          \
    }                                                           \
    DEVFN pg_bool_t                                             \
    pgfn_##NAME##_isnotnull(cl_int *errcode, pg_##NAME##_t arg) \
    {                                                           \
        pg_bool_t result;                                       \
        result.isnull = false;                                  \
        result.value = !arg.isnull;                             \
        return result;                                          \
    }

#define STROMCL_SIMPLE_TYPE_TEMPLATE(NAME,BASE)     \
    STROMCL_SIMPLE_DATATYPE_TEMPLATE(NAME,BASE)     \
    STROMCL_SIMPLE_VARREF_TEMPLATE(NAME,BASE)       \
    STROMCL_SIMPLE_PARAMREF_TEMPLATE(NAME,BASE)

STROMCL_SIMPLE_DATATYPE_TEMPLATE(bool, cl_bool)
STROMCL_SIMPLE_VARREF_TEMPLATE(bool, cl_bool)
STROMCL_SIMPLE_PARAMREF_TEMPLATE(bool, cl_bool)
STROMCL_SIMPLE_NULLTEST_TEMPLATE(bool)
STROMCL_SIMPLE_TYPE_TEMPLATE(int2, cl_short)
STROMCL_SIMPLE_NULLTEST_TEMPLATE(int2)
STROMCL_SIMPLE_TYPE_TEMPLATE(int4, cl_int)
STROMCL_SIMPLE_NULLTEST_TEMPLATE(int4)
STROMCL_SIMPLE_TYPE_TEMPLATE(int8, cl_long)
STROMCL_SIMPLE_NULLTEST_TEMPLATE(int8)
STROMCL_SIMPLE_TYPE_TEMPLATE(float4, cl_float)
STROMCL_SIMPLE_NULLTEST_TEMPLATE(float4)
STROMCL_SIMPLE_TYPE_TEMPLATE(float8, cl_double)
STROMCL_SIMPLE_NULLTEST_TEMPLATE(float8)
/* date = int4 days, time/timestamp = int8 microseconds (HAVE_INT64_TIMESTAMP) */
STROMCL_SIMPLE_TYPE_TEMPLATE(date, cl_int)
STROMCL_SIMPLE_NULLTEST_TEMPLATE(date)
STROMCL_SIMPLE_TYPE_TEMPLATE(time, cl_long)
STROMCL_SIMPLE_NULLTEST_TEMPLATE(time)
STROMCL_SIMPLE_TYPE_TEMPLATE(timestamp, cl_long)
STROMCL_SIMPLE_NULLTEST_TEMPLATE(timestamp)

/* bytea parameter (KPARAM_0 of GpuPreAgg is one): value = offset in kparams */
typedef struct {
    cl_uint value;
    bool    isnull;
} pg_bytea_t;

/*
 * A utility function to evaluate pg_bool_t value as if built-in bool.
 */
DEVFN bool
EVAL(pg_bool_t arg)
{
    return (!arg.isnull && arg.value != 0);
}

/* make a pg_<type>_t from a C value; used by generated NULL / zero consts */
#define PG_MAKE_NULL(NAME)      pg_##NAME##_null()
#define STROMCL_NULLCONST_TEMPLATE(NAME)                \
    DEVFN pg_##NAME##_t pg_##NAME##_null(void)          \
    {                                                   \
        pg_##NAME##_t r;                                \
        r.value = 0;                                    \
        r.isnull = true;                                \
        return r;                                       \
    }
STROMCL_NULLCONST_TEMPLATE(bool)
STROMCL_NULLCONST_TEMPLATE(int2)
STROMCL_NULLCONST_TEMPLATE(int4)
STROMCL_NULLCONST_TEMPLATE(int8)
STROMCL_NULLCONST_TEMPLATE(float4)
STROMCL_NULLCONST_TEMPLATE(float8)
STROMCL_NULLCONST_TEMPLATE(date)
STROMCL_NULLCONST_TEMPLATE(time)
STROMCL_NULLCONST_TEMPLATE(timestamp)

/*
 * macros for general binary compare functions
 */
#define devfunc_int_comp(x,y)                   \
    ((x) < (y) ? -1 : ((x) > (y) ? 1 : 0))

#define devfunc_float_comp(x,y)                 \
    (isnan(x)                                   \
     ? (isnan(y)                                \
        ? 0     /* NAN = NAN */                 \
        : 1)    /* NAN > non-NAN */             \
     : (isnan(y)                                \
        ? -1    /* non-NAN < NAN */             \
        : devfunc_int_comp((x),(y))))

/*
 * Functions for BooleanTest
 */
DEVFN pg_bool_t
pgfn_bool_is_true(cl_int *errcode, pg_bool_t result)
{
    result.value = (!result.isnull && result.value);
    result.isnull = false;
    return result;
}
DEVFN pg_bool_t
pgfn_bool_is_not_true(cl_int *errcode, pg_bool_t result)
{
    result.value = (result.isnull || !result.value);
    result.isnull = false;
    return result;
}
DEVFN pg_bool_t
pgfn_bool_is_false(cl_int *errcode, pg_bool_t result)
{
    result.value = (!result.isnull && !result.value);
    result.isnull = false;
    return result;
}
DEVFN pg_bool_t
pgfn_bool_is_not_false(cl_int *errcode, pg_bool_t result)
{
    result.value = (result.isnull || result.value);
    result.isnull = false;
    return result;
}
DEVFN pg_bool_t
pgfn_bool_is_unknown(cl_int *errcode, pg_bool_t result)
{
    result.value = result.isnull;
    result.isnull = false;
    return result;
}
DEVFN pg_bool_t
pgfn_bool_is_not_unknown(cl_int *errcode, pg_bool_t result)
{
    result.value = !result.isnull;
    result.isnull = false;
    return result;
}
/* NOT: NULL stays NULL */
DEVFN pg_bool_t
pgfn_boolop_not(cl_int *errcode, pg_bool_t result)
{
    result.value = !result.value;
    return result;
}

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

#endif  /* KERN_COMMON_CUH */
