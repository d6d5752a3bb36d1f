/*
 * TEST STAND-IN for PostgreSQL's postgres.h / fmgr.h / utils/array.h - only
 * what pg_glue/gpupreagg_fmgr.c uses, so that the glue can be compiled and
 * called by tests/test_pg_glue.py in an image without a PostgreSQL tree.
 * Not PostgreSQL code and not shipped: the real extension is built against
 * the real headers.  ereport(ERROR) / elog(ERROR) record the message and
 * return from the calling function (the real ones do not return).
 */
#ifndef PG_STUB_POSTGRES_H
#define PG_STUB_POSTGRES_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>

typedef uintptr_t   Datum;
typedef int32_t     int32;
typedef int64_t     int64;
typedef double      float8;
typedef unsigned int Oid;
typedef void       *MemoryContext;
#define FUNC_MAX_ARGS   100
#define ERROR           20
#define ERRCODE_NUMERIC_VALUE_OUT_OF_RANGE  0x2203

typedef struct FunctionCallInfoData
{
    short   nargs;
    bool    isnull;
    void   *context;            /* non-NULL = called as an aggregate's sfunc */
    Datum   arg[FUNC_MAX_ARGS];
    bool    argnull[FUNC_MAX_ARGS];
} FunctionCallInfoData, *FunctionCallInfo;
typedef Datum (*PGFunction)(FunctionCallInfo fcinfo);

extern char pg_stub_error_message[256];
extern int  pg_stub_error_code;

#define errcode(c)      (pg_stub_error_code = (c))
#define errmsg(...)     snprintf(pg_stub_error_message, sizeof(pg_stub_error_message), __VA_ARGS__)
#define ereport(lev, rest)  do { rest; fcinfo->isnull = true; return (Datum) 0; } while (0)
#define elog(lev, ...)  do { pg_stub_error_code = -1; errmsg(__VA_ARGS__); \
                             fcinfo->isnull = true; return (Datum) 0; } while (0)

static inline Datum Float8GetDatum(double v) { Datum d; memcpy(&d, &v, 8); return d; }
static inline double DatumGetFloat8(Datum d) { double v; memcpy(&v, &d, 8); return v; }
#define DatumGetPointer(d)      ((void *) (d))
#define PointerGetDatum(p)      ((Datum) (p))
#endif
