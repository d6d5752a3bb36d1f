/*
 * gpupreagg_fmgr.c - the fmgr V1 functions pg_strom--1.0.sql:99-401 names
 * ('MODULE_PATHNAME', 'gpupreagg_partial_nrows' ... 'pgstrom_covariance_float8_accum'),
 * as thin wrappers over libpgstrom_cuda.so (include/pgstrom_cuda.h section 7).
 * Replaces the bodies at gpupreagg.c:4251-4773 of the reference; symbol
 * names, argument lists and strictness are the catalog's, unchanged.
 *
 * Part of the PostgreSQL-side glue: it needs postgres.h and
 * is built with the extension, not with the library.  No PostgreSQL tree
 * exists in the build image of this repository; tests/test_pg_glue.py
 * compiles this file against tests/native/pg_stub/ (a stand-in for the few
 * fmgr / array macros used here) and calls every function through it.
 */
#include "postgres.h"
#include "fmgr.h"
#include "catalog/pg_type.h"
#include "utils/array.h"
#include "utils/builtins.h"

#include "pgstrom_cuda.h"

#define FINALFN_CHECK(rc)                                               \
    do {                                                                \
        int __rc = (rc);                                                \
        if (__rc == PGS_FINALFN_OVERFLOW)                               \
            ereport(ERROR,                                              \
                    (errcode(ERRCODE_NUMERIC_VALUE_OUT_OF_RANGE),       \
                     errmsg("value out of range: overflow")));          \
        else if (__rc != 0)                                             \
            elog(ERROR, "Bug? NULL or negative nrows was given");       \
    } while (0)

/* ---- partial placeholders ------------------------------------------------ */

PG_FUNCTION_INFO_V1(gpupreagg_partial_nrows);
Datum
gpupreagg_partial_nrows(PG_FUNCTION_ARGS)
{
    char    values[FUNC_MAX_ARGS];
    char    isnull[FUNC_MAX_ARGS];
    int     i;

    for (i = 0; i < PG_NARGS(); i++)
    {
        isnull[i] = PG_ARGISNULL(i);
        values[i] = !isnull[i] && PG_GETARG_BOOL(i);
    }
    PG_RETURN_INT32(pgs_partial_nrows(PG_NARGS(), values, isnull));
}

/* pgstrom.pmin / pgstrom.pmax: the argument as it is */
PG_FUNCTION_INFO_V1(gpupreagg_pseudo_expr);
Datum
gpupreagg_pseudo_expr(PG_FUNCTION_ARGS)
{
    PG_RETURN_DATUM(PG_GETARG_DATUM(0));
}

/* pgstrom.psum(int8 | float4 | float8 | numeric): NULL stays NULL */
#define PSUM_PASS_THROUGH(fname)                    \
    PG_FUNCTION_INFO_V1(fname);                     \
    Datum fname(PG_FUNCTION_ARGS)                   \
    {                                               \
        if (PG_ARGISNULL(0))                        \
            PG_RETURN_NULL();                       \
        PG_RETURN_DATUM(PG_GETARG_DATUM(0));        \
    }
PSUM_PASS_THROUGH(gpupreagg_psum_int)
PSUM_PASS_THROUGH(gpupreagg_psum_float4)
PSUM_PASS_THROUGH(gpupreagg_psum_float8)
PSUM_PASS_THROUGH(gpupreagg_psum_numeric)

PG_FUNCTION_INFO_V1(gpupreagg_psum_x2_float);
Datum
gpupreagg_psum_x2_float(PG_FUNCTION_ARGS)
{
    double  result;

    if (pgs_psum_x2_float8(PG_ARGISNULL(0) ? 0.0 : PG_GETARG_FLOAT8(0),
                           PG_ARGISNULL(0), &result))
        PG_RETURN_NULL();
    PG_RETURN_FLOAT8(result);
}

#ifndef PGSTROM_GLUE_NO_NUMERIC
PG_FUNCTION_INFO_V1(gpupreagg_psum_x2_numeric);
Datum
gpupreagg_psum_x2_numeric(PG_FUNCTION_ARGS)
{
    if (PG_ARGISNULL(0))
        PG_RETURN_NULL();
    PG_RETURN_DATUM(DirectFunctionCall2(numeric_mul,
                                        PG_GETARG_DATUM(0),
                                        PG_GETARG_DATUM(0)));
}
#endif

/* pgstrom.pcov_*(bool filter, float8 X, float8 Y) */
#define PCOV_FUNCTION(fname, kind)                                          \
    PG_FUNCTION_INFO_V1(fname);                                             \
    Datum fname(PG_FUNCTION_ARGS)                                           \
    {                                                                       \
        double  result;                                                     \
        if (pgs_pcov_float8(kind,                                           \
                            !PG_ARGISNULL(0) && PG_GETARG_BOOL(0),          \
                            PG_ARGISNULL(0),                                \
                            PG_ARGISNULL(1) ? 0.0 : PG_GETARG_FLOAT8(1),    \
                            PG_ARGISNULL(1),                                \
                            PG_ARGISNULL(2) ? 0.0 : PG_GETARG_FLOAT8(2),    \
                            PG_ARGISNULL(2), &result))                      \
            PG_RETURN_NULL();                                               \
        PG_RETURN_FLOAT8(result);                                           \
    }
PCOV_FUNCTION(gpupreagg_corr_psum_x, PGS_PCOV_X)
PCOV_FUNCTION(gpupreagg_corr_psum_y, PGS_PCOV_Y)
PCOV_FUNCTION(gpupreagg_corr_psum_x2, PGS_PCOV_X2)
PCOV_FUNCTION(gpupreagg_corr_psum_y2, PGS_PCOV_Y2)
PCOV_FUNCTION(gpupreagg_corr_psum_xy, PGS_PCOV_XY)

/* ---- final accumulators (all STRICT) -------------------------------------- */

/* the transition array, modifiable: in place inside an aggregate, a copy
 * when called as a plain function */
static ArrayType *
transarray(FunctionCallInfo fcinfo, int nitems, Oid elemtype)
{
    ArrayType  *arr = (AggCheckCallContext(fcinfo, NULL)
                       ? PG_GETARG_ARRAYTYPE_P(0)
                       : PG_GETARG_ARRAYTYPE_P_COPY(0));

    if (ARR_NDIM(arr) != 1 ||
        ARR_DIMS(arr)[0] != nitems ||
        ARR_HASNULL(arr) ||
        ARR_ELEMTYPE(arr) != elemtype)
        return NULL;
    return arr;
}
#define TRANSARRAY(var, nitems, elemtype)                               \
    do {                                                                \
        (var) = transarray(fcinfo, nitems, elemtype);                   \
        if (!(var))                                                     \
            elog(ERROR, "%d-elements array is expected", nitems);       \
    } while (0)

PG_FUNCTION_INFO_V1(pgstrom_avg_int8_accum);
Datum
pgstrom_avg_int8_accum(PG_FUNCTION_ARGS)
{
    ArrayType  *arr;

    TRANSARRAY(arr, 2, INT8OID);
    FINALFN_CHECK(pgs_avg_int8_accum((int64_t *) ARR_DATA_PTR(arr),
                                     PG_GETARG_INT32(1), PG_GETARG_INT64(2)));
    PG_RETURN_ARRAYTYPE_P(arr);
}

PG_FUNCTION_INFO_V1(pgstrom_sum_int8_accum);
Datum
pgstrom_sum_int8_accum(PG_FUNCTION_ARGS)
{
    ArrayType  *arr;

    TRANSARRAY(arr, 2, INT8OID);
    FINALFN_CHECK(pgs_sum_int8_accum((int64_t *) ARR_DATA_PTR(arr), PG_GETARG_INT64(1)));
    PG_RETURN_ARRAYTYPE_P(arr);
}

PG_FUNCTION_INFO_V1(pgstrom_sum_int8_final);
Datum
pgstrom_sum_int8_final(PG_FUNCTION_ARGS)
{
    ArrayType  *arr = PG_GETARG_ARRAYTYPE_P(0);
    int64_t     result;

    if (ARR_NDIM(arr) != 1 || ARR_DIMS(arr)[0] != 2 ||
        ARR_HASNULL(arr) || ARR_ELEMTYPE(arr) != INT8OID)
        elog(ERROR, "Two elements int8 array is expected");
    if (pgs_sum_int8_final((const int64_t *) ARR_DATA_PTR(arr), &result))
        PG_RETURN_NULL();
    PG_RETURN_INT64(result);
}

PG_FUNCTION_INFO_V1(pgstrom_sum_float8_accum);
Datum
pgstrom_sum_float8_accum(PG_FUNCTION_ARGS)
{
    ArrayType  *arr;

    TRANSARRAY(arr, 3, FLOAT8OID);
    FINALFN_CHECK(pgs_sum_float8_accum((double *) ARR_DATA_PTR(arr),
                                       PG_GETARG_INT32(1), PG_GETARG_FLOAT8(2)));
    PG_RETURN_ARRAYTYPE_P(arr);
}

PG_FUNCTION_INFO_V1(pgstrom_variance_float8_accum);
Datum
pgstrom_variance_float8_accum(PG_FUNCTION_ARGS)
{
    ArrayType  *arr;

    TRANSARRAY(arr, 3, FLOAT8OID);
    FINALFN_CHECK(pgs_variance_float8_accum((double *) ARR_DATA_PTR(arr),
                                            PG_GETARG_INT32(1),
                                            PG_GETARG_FLOAT8(2), PG_GETARG_FLOAT8(3)));
    PG_RETURN_ARRAYTYPE_P(arr);
}

PG_FUNCTION_INFO_V1(pgstrom_covariance_float8_accum);
Datum
pgstrom_covariance_float8_accum(PG_FUNCTION_ARGS)
{
    ArrayType  *arr;
    double      psum[5];
    int         i;

    TRANSARRAY(arr, 6, FLOAT8OID);
    for (i = 0; i < 5; i++)
        psum[i] = PG_GETARG_FLOAT8(2 + i);     /* pcov_x, pcov_x2, pcov_y, pcov_y2, pcov_xy */
    FINALFN_CHECK(pgs_covariance_float8_accum((double *) ARR_DATA_PTR(arr),
                                              PG_GETARG_INT32(1), psum));
    PG_RETURN_ARRAYTYPE_P(arr);
}

#ifndef PGSTROM_GLUE_NO_NUMERIC
/*
 * pgstrom.int8_avg_accum(internal, int4 nrows, int8 psum) and
 * pgstrom.numeric_avg_accum(internal, int4 nrows, numeric psum): PostgreSQL's
 * own NumericAggState through int8_avg_accum / numeric_avg_accum, like the
 * reference (gpupreagg.c:4540-4588) - but N grows by exactly nrows (the rule
 * pgs_numeric_avg_accum states): the built-in counts one row per call, so
 * nrows - 1 is added whenever a sum was added, also for nrows = 0.
 */
typedef struct
{
    bool            calcSumX2;
    MemoryContext   agg_context;
    int64           N;
} NumericAggStateHead;      /* leading members of numeric.c's NumericAggState */

static Datum
numeric_avg_accum_common(FunctionCallInfo fcinfo, PGFunction accum)
{
    int32       nrows = PG_GETARG_INT32(1);
    bool        psum_isnull = PG_ARGISNULL(2);
    NumericAggStateHead *state;

    if (PG_ARGISNULL(1) || nrows < 0)
        elog(ERROR, "Bug? NULL or negative nrows was given");
    fcinfo->nargs = 2;
    fcinfo->arg[1] = fcinfo->arg[2];
    fcinfo->argnull[1] = fcinfo->argnull[2];
    state = (NumericAggStateHead *) DatumGetPointer(accum(fcinfo));
    if (state && !psum_isnull)
        state->N += (int64) nrows - 1;
    PG_RETURN_POINTER(state);
}

PG_FUNCTION_INFO_V1(pgstrom_int8_avg_accum);
Datum
pgstrom_int8_avg_accum(PG_FUNCTION_ARGS)
{
    return numeric_avg_accum_common(fcinfo, int8_avg_accum);
}

PG_FUNCTION_INFO_V1(pgstrom_numeric_avg_accum);
Datum
pgstrom_numeric_avg_accum(PG_FUNCTION_ARGS)
{
    return numeric_avg_accum_common(fcinfo, numeric_avg_accum);
}
#endif  /* PGSTROM_GLUE_NO_NUMERIC */
