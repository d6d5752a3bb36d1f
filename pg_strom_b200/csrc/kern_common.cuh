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

#include "kern_mathlib.cuh"

#endif  /* KERN_COMMON_CUH */
