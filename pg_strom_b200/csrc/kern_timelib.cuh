/*
 * kern_timelib.cuh - date / time / timestamp functions of the device runtime
 * (the counterpart of the reference's opencl_timelib.h; catalogue entries
 * codegen.c:540-606).
 *
 * Representation (HAVE_INT64_TIMESTAMP): date = int4 days since 2000-01-01,
 * time = int8 microseconds since midnight, timestamp = int8 microseconds
 * since 2000-01-01 00:00; INT_MIN / INT_MAX and LONG_MIN / LONG_MAX are
 * -infinity / +infinity.
 *
 * The results are PostgreSQL's (utils/adt/date.c, timestamp.c).  Where
 * PostgreSQL raises an error ("date out of range for timestamp", "cannot
 * subtract infinite dates", "timestamp out of range") or where 9.4 silently
 * wraps an int4 (date + integer), the device returns NULL and flags the row
 * StromError_CpuReCheck: the host evaluates that row and raises / wraps as it
 * always did.  The reference converts a timestamp to a date through a
 * broken-down struct pg_tm (timestamp2tm + date2j, opencl_timelib.h:189-258);
 * date2j(j2date(d)) = d, so a floor division by the microseconds of a day is
 * the same function and that is what runs here.
 */
#ifndef KERN_TIMELIB_CUH
#define KERN_TIMELIB_CUH

#define PGS_DATEVAL_NOBEGIN     ((cl_int)(-0x7fffffff - 1))
#define PGS_DATEVAL_NOEND       ((cl_int)0x7fffffff)
#define PGS_DATE_NOT_FINITE(d)  ((d) == PGS_DATEVAL_NOBEGIN || (d) == PGS_DATEVAL_NOEND)
#define PGS_DT_NOBEGIN          ((cl_long)(-0x7fffffffffffffffLL - 1))
#define PGS_DT_NOEND            ((cl_long)0x7fffffffffffffffLL)
#define PGS_TS_NOT_FINITE(t)    ((t) == PGS_DT_NOBEGIN || (t) == PGS_DT_NOEND)
#define PGS_USECS_PER_DAY       86400000000LL
#define PGS_POSTGRES_EPOCH_JDATE 2451545    /* date2j(2000, 1, 1) */
/* |days| beyond this do not fit a timestamp */
#define PGS_DATE_MAX_FOR_TS     (0x7fffffffffffffffLL / PGS_USECS_PER_DAY)

/* floor(t / day), t mod day in [0, day) */
DEVFN void
pgs_ts_split(cl_long ts, cl_long *days, cl_long *usecs)
{
    cl_long     d = ts / PGS_USECS_PER_DAY;
    cl_long     u = ts - d * PGS_USECS_PER_DAY;

    if (u < 0)
    {
        u += PGS_USECS_PER_DAY;
        d -= 1;
    }
    *days = d;
    *usecs = u;
}

/* timestamp2tm() refuses Julian days outside [0, INT_MAX] */
DEVFN bool
pgs_ts_days_in_range(cl_long days)
{
    cl_long     jd = days + PGS_POSTGRES_EPOCH_JDATE;

    return (jd >= 0 && jd <= 0x7fffffffLL);
}

/* ---- type casts ---- */
CAST_SIMPLE(date_date, date, cl_int, date)
CAST_SIMPLE(time_time, time, cl_long, time)
CAST_SIMPLE(timestamp_timestamp, timestamp, cl_long, timestamp)

DEVFN pg_date_t
pgfn_timestamp_date(cl_int *errcode, pg_timestamp_t arg1)
{
    pg_date_t   result;
    cl_long     days, usecs;

    result.value = 0;
    result.isnull = arg1.isnull;
    if (arg1.isnull)
        return result;
    if (arg1.value == PGS_DT_NOBEGIN)
        result.value = PGS_DATEVAL_NOBEGIN;
    else if (arg1.value == PGS_DT_NOEND)
        result.value = PGS_DATEVAL_NOEND;
    else
    {
        pgs_ts_split(arg1.value, &days, &usecs);
        if (!pgs_ts_days_in_range(days))
        {
            result.isnull = true;
            STROM_SET_ERROR(errcode, StromError_CpuReCheck);
        }
        else
            result.value = (cl_int)days;
    }
    return result;
}

DEVFN pg_time_t
pgfn_timestamp_time(cl_int *errcode, pg_timestamp_t arg1)
{
    pg_time_t   result;
    cl_long     days, usecs;

    result.value = 0;
    result.isnull = true;
    if (arg1.isnull || PGS_TS_NOT_FINITE(arg1.value))
        return result;          /* time of an infinite timestamp is NULL */
    pgs_ts_split(arg1.value, &days, &usecs);
    if (!pgs_ts_days_in_range(days))
        STROM_SET_ERROR(errcode, StromError_CpuReCheck);
    else
    {
        result.isnull = false;
        result.value = usecs;
    }
    return result;
}

DEVFN pg_timestamp_t
pgfn_date_timestamp(cl_int *errcode, pg_date_t arg1)
{
    pg_timestamp_t result;

    result.value = 0;
    result.isnull = arg1.isnull;
    if (arg1.isnull)
        return result;
    if (arg1.value == PGS_DATEVAL_NOBEGIN)
        result.value = PGS_DT_NOBEGIN;
    else if (arg1.value == PGS_DATEVAL_NOEND)
        result.value = PGS_DT_NOEND;
    else if ((cl_long)arg1.value > PGS_DATE_MAX_FOR_TS ||
             (cl_long)arg1.value < -PGS_DATE_MAX_FOR_TS)
    {
        /* date's range is wider than timestamp's */
        result.isnull = true;
        STROM_SET_ERROR(errcode, StromError_CpuReCheck);
    }
    else
        result.value = (cl_long)arg1.value * PGS_USECS_PER_DAY;
    return result;
}

/* ---- date / time operators ---- */
DEVFN pg_date_t
pgs_date_add_days(cl_int *errcode, pg_date_t arg1, pg_int4_t arg2, bool subtract)
{
    pg_date_t   result;

    result.value = 0;
    result.isnull = (arg1.isnull | arg2.isnull);
    if (result.isnull)
        return result;
    if (PGS_DATE_NOT_FINITE(arg1.value))
        result.value = arg1.value;      /* can't change infinity */
    else
    {
        cl_long     v = (subtract ? (cl_long)arg1.value - (cl_long)arg2.value
                                  : (cl_long)arg1.value + (cl_long)arg2.value);
        /* leaving int4, or landing on an infinity mark: host decides */
        if (v <= (cl_long)PGS_DATEVAL_NOBEGIN || v >= (cl_long)PGS_DATEVAL_NOEND)
        {
            result.isnull = true;
            STROM_SET_ERROR(errcode, StromError_CpuReCheck);
        }
        else
            result.value = (cl_int)v;
    }
    return result;
}

DEVFN pg_date_t
pgfn_date_pli(cl_int *errcode, pg_date_t arg1, pg_int4_t arg2)
{
    return pgs_date_add_days(errcode, arg1, arg2, false);
}

DEVFN pg_date_t
pgfn_date_mii(cl_int *errcode, pg_date_t arg1, pg_int4_t arg2)
{
    return pgs_date_add_days(errcode, arg1, arg2, true);
}

DEVFN pg_date_t
pgfn_integer_pl_date(cl_int *errcode, pg_int4_t arg1, pg_date_t arg2)
{
    return pgs_date_add_days(errcode, arg2, arg1, false);
}

DEVFN pg_int4_t
pgfn_date_mi(cl_int *errcode, pg_date_t arg1, pg_date_t arg2)
{
    pg_int4_t   result;

    result.value = 0;
    result.isnull = (arg1.isnull | arg2.isnull);
    if (result.isnull)
        return result;
    cl_long     v = (cl_long)arg1.value - (cl_long)arg2.value;
    if (PGS_DATE_NOT_FINITE(arg1.value) || PGS_DATE_NOT_FINITE(arg2.value) ||
        v < -0x80000000LL || v > 0x7fffffffLL)
    {
        result.isnull = true;
        STROM_SET_ERROR(errcode, StromError_CpuReCheck);
    }
    else
        result.value = (cl_int)v;
    return result;
}

DEVFN pg_timestamp_t
pgfn_datetime_pl(cl_int *errcode, pg_date_t arg1, pg_time_t arg2)
{
    pg_timestamp_t result;

    result.value = 0;
    result.isnull = (arg1.isnull | arg2.isnull);
    if (result.isnull)
        return result;
    result = pgfn_date_timestamp(errcode, arg1);
    if (!result.isnull && !PGS_TS_NOT_FINITE(result.value))
    {
        if (arg2.value > 0 && result.value > PGS_DT_NOEND - arg2.value)
        {
            result.value = 0;
            result.isnull = true;
            STROM_SET_ERROR(errcode, StromError_CpuReCheck);
        }
        else
            result.value += arg2.value;
    }
    return result;
}

DEVFN pg_timestamp_t
pgfn_timedata_pl(cl_int *errcode, pg_time_t arg1, pg_date_t arg2)
{
    return pgfn_datetime_pl(errcode, arg2, arg1);
}

/* ---- date <-> timestamp comparison (date.c: date_cmp_timestamp ...) ---- */
DEVFN pg_int4_t
pgfn_date_cmp_timestamp(cl_int *errcode, pg_date_t arg1, pg_timestamp_t arg2)
{
    pg_int4_t       result;
    pg_timestamp_t  dt1;

    /* strict: with a NULL argument PostgreSQL does not call the function,
     * so an out-of-range date must not raise anything either */
    result.value = 0;
    result.isnull = true;
    if (arg1.isnull | arg2.isnull)
        return result;
    dt1 = pgfn_date_timestamp(errcode, arg1);
    result.isnull = dt1.isnull;
    result.value = (result.isnull ? 0 : devfunc_int_comp(dt1.value, arg2.value));
    return result;
}

DEVFN pg_int4_t
pgfn_timestamp_cmp_date(cl_int *errcode, pg_timestamp_t arg1, pg_date_t arg2)
{
    pg_int4_t       result;
    pg_timestamp_t  dt2;

    result.value = 0;
    result.isnull = true;
    if (arg1.isnull | arg2.isnull)
        return result;
    dt2 = pgfn_date_timestamp(errcode, arg2);
    result.isnull = dt2.isnull;
    result.value = (result.isnull ? 0 : devfunc_int_comp(arg1.value, dt2.value));
    return result;
}

#define PGS_CROSS_COMPARE_TEMPLATE(LNAME,RNAME,SFX,OPER)                    \
    DEVFN pg_bool_t                                                         \
    pgfn_##LNAME##_##SFX##_##RNAME(cl_int *errcode,                         \
                                   pg_##LNAME##_t arg1, pg_##RNAME##_t arg2)\
    {                                                                       \
        pg_int4_t   c = pgfn_##LNAME##_cmp_##RNAME(errcode, arg1, arg2);    \
        pg_bool_t   result;                                                 \
                                                                            \
        result.isnull = c.isnull;                                           \
        result.value = (cl_bool)(!c.isnull && (c.value OPER 0));            \
        return result;                                                      \
    }
PGS_CROSS_COMPARE_TEMPLATE(date, timestamp, eq, ==)
PGS_CROSS_COMPARE_TEMPLATE(date, timestamp, ne, !=)
PGS_CROSS_COMPARE_TEMPLATE(date, timestamp, lt, <)
PGS_CROSS_COMPARE_TEMPLATE(date, timestamp, le, <=)
PGS_CROSS_COMPARE_TEMPLATE(date, timestamp, gt, >)
PGS_CROSS_COMPARE_TEMPLATE(date, timestamp, ge, >=)
PGS_CROSS_COMPARE_TEMPLATE(timestamp, date, eq, ==)
PGS_CROSS_COMPARE_TEMPLATE(timestamp, date, ne, !=)
PGS_CROSS_COMPARE_TEMPLATE(timestamp, date, lt, <)
PGS_CROSS_COMPARE_TEMPLATE(timestamp, date, le, <=)
PGS_CROSS_COMPARE_TEMPLATE(timestamp, date, gt, >)
PGS_CROSS_COMPARE_TEMPLATE(timestamp, date, ge, >=)

#endif  /* KERN_TIMELIB_CUH */
