/* TEST STAND-IN, see postgres.h in this directory */
#ifndef PG_STUB_FMGR_H
#define PG_STUB_FMGR_H
#define PG_FUNCTION_ARGS        FunctionCallInfo fcinfo
#define PG_FUNCTION_INFO_V1(f)  extern Datum f(PG_FUNCTION_ARGS)
#define PG_NARGS()              (fcinfo->nargs)
#define PG_ARGISNULL(n)         (fcinfo->argnull[n])
#define PG_GETARG_DATUM(n)      (fcinfo->arg[n])
#define PG_GETARG_BOOL(n)       ((bool) (fcinfo->arg[n] != 0))
#define PG_GETARG_INT32(n)      ((int32) fcinfo->arg[n])
#define PG_GETARG_INT64(n)      ((int64) fcinfo->arg[n])
#define PG_GETARG_FLOAT8(n)     DatumGetFloat8(fcinfo->arg[n])
#define PG_RETURN_NULL()        do { fcinfo->isnull = true; return (Datum) 0; } while (0)
#define PG_RETURN_DATUM(x)      return (x)
#define PG_RETURN_INT32(x)      return (Datum) (uint32_t) (x)
#define PG_RETURN_INT64(x)      return (Datum) (x)
#define PG_RETURN_FLOAT8(x)     return Float8GetDatum(x)
#define PG_RETURN_POINTER(x)    return PointerGetDatum(x)
#define AggCheckCallContext(fcinfo, p)  ((fcinfo)->context != NULL)
#endif
