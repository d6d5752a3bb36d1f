/*
 * Test shim: compiles the date/time and text device runtime
 * (pg_strom_b200/csrc/kern_timelib.cuh, kern_textlib.cuh) and the
 * float -> numeric cast of kern_numeric.cuh with g++, so that what the
 * kernels evaluate per row can be checked on the CPU against python
 * restatements of PostgreSQL's functions (tests/test_typelib_device_code.py).
 * Test infrastructure only.
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
typedef struct { cl_int value; bool isnull; } pg_date_t;
typedef struct { cl_long value; bool isnull; } pg_time_t;
typedef struct { cl_long value; bool isnull; } pg_timestamp_t;
#define STROMCL_SIMPLE_NULLTEST_TEMPLATE(NAME)
#define GPUPREAGG_INCOL_SLOT(colidx) 0
#define devfunc_int_comp(x,y)   ((x) < (y) ? -1 : ((x) > (y) ? 1 : 0))
#define CAST_SIMPLE(name,r_type,R_BASE,x_type)                          \
    DEVFN pg_##r_type##_t                                               \
    pgfn_##name(cl_int *errcode, pg_##x_type##_t arg)                   \
    {                                                                   \
        pg_##r_type##_t result;                                         \
        result.value = (R_BASE)arg.value;                               \
        result.isnull = arg.isnull;                                     \
        return result;                                                  \
    }
/* single-threaded stand-ins of the atomics of the key heap */
template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicCAS(T *p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
template <typename T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
#include "kern_numeric.cuh"
#include "kern_timelib.cuh"
#include "kern_textlib.cuh"

extern "C" {

/* fn: 0 timestamp_date, 1 timestamp_time, 2 date_timestamp.
 * returns errcode; *isnull / *out = result */
int shim_time_cast(int fn, int64_t a, int64_t *out, int *isnull)
{
    cl_int e = 0;
    if (fn == 0)
    {
        pg_timestamp_t x = { (cl_long)a, false };
        pg_date_t r = pgfn_timestamp_date(&e, x);
        *out = r.value; *isnull = r.isnull;
    }
    else if (fn == 1)
    {
        pg_timestamp_t x = { (cl_long)a, false };
        pg_time_t r = pgfn_timestamp_time(&e, x);
        *out = r.value; *isnull = r.isnull;
    }
    else
    {
        pg_date_t x = { (cl_int)a, false };
        pg_timestamp_t r = pgfn_date_timestamp(&e, x);
        *out = r.value; *isnull = r.isnull;
    }
    return e;
}

/* fn: 0 date_pli, 1 date_mii, 2 date_mi, 3 datetime_pl, 4 integer_pl_date,
 *     5 timedata_pl, 6 date_cmp_timestamp, 7 timestamp_cmp_date,
 *     10..15 date_{eq,ne,lt,le,gt,ge}_timestamp, 20..25 timestamp_.._date */
int shim_time_binop(int fn, int64_t a, int64_t b, int64_t *out, int *isnull)
{
    cl_int e = 0;
    pg_date_t da = { (cl_int)a, false }, db = { (cl_int)b, false };
    pg_int4_t ia = { (cl_int)a, false }, ib = { (cl_int)b, false };
    pg_time_t ta = { (cl_long)a, false }, tb = { (cl_long)b, false };
    pg_timestamp_t sa = { (cl_long)a, false }, sb = { (cl_long)b, false };
#define RET(r) do { *out = (int64_t)(r).value; *isnull = (r).isnull; return e; } while (0)
    switch (fn)
    {
        case 0: { pg_date_t r = pgfn_date_pli(&e, da, ib); RET(r); }
        case 1: { pg_date_t r = pgfn_date_mii(&e, da, ib); RET(r); }
        case 2: { pg_int4_t r = pgfn_date_mi(&e, da, db); RET(r); }
        case 3: { pg_timestamp_t r = pgfn_datetime_pl(&e, da, tb); RET(r); }
        case 4: { pg_date_t r = pgfn_integer_pl_date(&e, ia, db); RET(r); }
        case 5: { pg_timestamp_t r = pgfn_timedata_pl(&e, ta, db); RET(r); }
        case 6: { pg_int4_t r = pgfn_date_cmp_timestamp(&e, da, sb); RET(r); }
        case 7: { pg_int4_t r = pgfn_timestamp_cmp_date(&e, sa, db); RET(r); }
        case 10: { pg_bool_t r = pgfn_date_eq_timestamp(&e, da, sb); RET(r); }
        case 11: { pg_bool_t r = pgfn_date_ne_timestamp(&e, da, sb); RET(r); }
        case 12: { pg_bool_t r = pgfn_date_lt_timestamp(&e, da, sb); RET(r); }
        case 13: { pg_bool_t r = pgfn_date_le_timestamp(&e, da, sb); RET(r); }
        case 14: { pg_bool_t r = pgfn_date_gt_timestamp(&e, da, sb); RET(r); }
        case 15: { pg_bool_t r = pgfn_date_ge_timestamp(&e, da, sb); RET(r); }
        case 20: { pg_bool_t r = pgfn_timestamp_eq_date(&e, sa, db); RET(r); }
        case 21: { pg_bool_t r = pgfn_timestamp_ne_date(&e, sa, db); RET(r); }
        case 22: { pg_bool_t r = pgfn_timestamp_lt_date(&e, sa, db); RET(r); }
        case 23: { pg_bool_t r = pgfn_timestamp_le_date(&e, sa, db); RET(r); }
        case 24: { pg_bool_t r = pgfn_timestamp_gt_date(&e, sa, db); RET(r); }
        case 25: { pg_bool_t r = pgfn_timestamp_ge_date(&e, sa, db); RET(r); }
    }
#undef RET
    return -1;
}

/* a, b: varlena images.  fn 0..5 = eq ne lt le gt ge, 6 = cmp; bpchar != 0
 * selects the blank-padded semantics.  returns errcode */
int shim_text_op(int fn, int bpchar, const unsigned char *a, const unsigned char *b,
                 int *out, int *isnull)
{
    cl_int e = 0;
    pg_varlena_t x = pgs_varlena_make(&e, a);
    pg_varlena_t y = pgs_varlena_make(&e, b);
    pg_bool_t r = { 0, true };
    pg_int4_t c = { 0, true };
    if (bpchar)
    {
        switch (fn)
        {
            case 0: r = pgfn_bpchareq(&e, x, y); break;
            case 1: r = pgfn_bpcharne(&e, x, y); break;
            case 2: r = pgfn_bpcharlt(&e, x, y); break;
            case 3: r = pgfn_bpcharle(&e, x, y); break;
            case 4: r = pgfn_bpchargt(&e, x, y); break;
            case 5: r = pgfn_bpcharge(&e, x, y); break;
            default: c = pgfn_bpcharcmp(&e, x, y); break;
        }
    }
    else
    {
        switch (fn)
        {
            case 0: r = pgfn_texteq(&e, x, y); break;
            case 1: r = pgfn_textne(&e, x, y); break;
            case 2: r = pgfn_text_lt(&e, x, y); break;
            case 3: r = pgfn_text_le(&e, x, y); break;
            case 4: r = pgfn_text_gt(&e, x, y); break;
            case 5: r = pgfn_text_ge(&e, x, y); break;
            default: c = pgfn_text_cmp(&e, x, y); break;
        }
    }
    if (fn <= 5) { *out = r.value; *isnull = r.isnull; }
    else { *out = c.value; *isnull = c.isnull; }
    return e;
}

/* varlena image -> "kernel text" key word (bpchar != 0: without trailing blanks) */
int shim_text_keybits(const unsigned char *a, int bpchar, uint64_t *out, int *isnull)
{
    cl_int e = 0;
    bool n = false;
    pg_varlena_t x = pgs_varlena_make(&e, a);
    *out = pgs_text_keybits(&e, x, bpchar != 0, &n);
    *isnull = n;
    return e;
}

/* the key heap of the "session": nslots (power of 2, 0 = none) lookup slots
 * and heap_bytes of strings in caller memory (slots: 16 bytes each, zeroed
 * by the caller; used: the allocation cursor) */
void shim_keyheap_set(void *slots, unsigned int nslots, void *heap, uint64_t heap_bytes,
                      uint64_t *used, unsigned int max_probe)
{
    pgs_keyheap.slots = (cl_ulong *)slots;
    pgs_keyheap.nslots = nslots;
    pgs_keyheap.heap = (unsigned char *)heap;
    pgs_keyheap.heap_bytes = heap_bytes;
    pgs_keyheap.heap_used = (cl_ulong *)used;
    pgs_keyheap.max_probe = max_probe;
}

/* float8 (is_f4 = 0) or float4 value -> packed device numeric */
int shim_float_numeric(double v, int is_f4, uint64_t *out, int *isnull)
{
    cl_int e = 0;
    pg_numeric_t r;
    if (is_f4)
    {
        pg_float4_t a = { (float)v, false };
        r = pgfn_float4_numeric(&e, a);
    }
    else
    {
        pg_float8_t a = { v, false };
        r = pgfn_float8_numeric(&e, a);
    }
    *out = r.value; *isnull = r.isnull;
    return e;
}

}
