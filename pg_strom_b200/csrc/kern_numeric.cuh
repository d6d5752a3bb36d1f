/*
 * kern_numeric.cuh - NUMERIC on the device.
 *
 * The reference keeps a numeric in 64 bits on the device: 6-bit exponent of
 * ten, sign, 57-bit mantissa (opencl_numeric.h:141-162), turns PostgreSQL's
 * varlena image into it (pg_numeric_from_varlena, :166-307), and whatever
 * does not fit is left to the CPU (StromError_CpuReCheck).  The partial
 * results travel in that format and the host turns them back into a varlena
 * with numeric_in() (pgstrom_fixup_kernel_numeric, datastore.c:150-167).
 *
 * Same format and same contract here, with one rule added so that the
 * results print exactly like PostgreSQL's own: the exponent of a value read
 * from a tuple is always minus its display scale (1.50 is 150 * 10^-2, never
 * 15 * 10^-1).  numeric_in() gives "mantissa e exponent" the display scale
 * -exponent, sums take the largest display scale of their inputs and min /
 * max return one of their inputs unchanged - PostgreSQL's rules - so the
 * final aggregates see numerics with the display scales PostgreSQL's own
 * partial sums would have.  Values whose digits or scale do not fit
 * (mantissa >= 2^57 - 1, display scale > 32, NaN, toasted) are re-checked
 * per row.
 *
 * Partial sums (PSUM NUMERIC) do not live in this format: three cells hold a
 * 128-bit integer at the fixed scale PGS_NUMERIC_SUM_SCALE and the largest
 * display scale seen; a sum therefore never overflows on the device, and the
 * flush writes it - divided back to its display scale - as one or several
 * partial rows of 57-bit mantissas (PostgreSQL's final sum adds them up).
 */
#ifndef KERN_NUMERIC_CUH
#define KERN_NUMERIC_CUH

typedef struct {
    cl_ulong    value;
    bool        isnull;
} pg_numeric_t;

#define PGS_NUMERIC_MANT_LIMIT      (PG_NUMERIC_MANTISSA_MAX - 1)   /* all ones: "no value" */
#define PGS_NUMERIC_EMPTY           0xFFFFFFFFFFFFFFFFULL
#define PGS_NUMERIC_MAX_DSCALE      32
#define PGS_NUMERIC_SUM_SCALE       16

typedef unsigned __int128   pgs_u128;
typedef __int128            pgs_s128;

DEVFN cl_ulong
pgs_pow10_u64(int n)        /* 0 <= n <= 19 */
{
    cl_ulong v = 1;
    for (int i = 0; i < n; i++)
        v *= 10;
    return v;
}

DEVFN pgs_u128
pgs_pow10_u128(int n)       /* 0 <= n <= 38 */
{
    pgs_u128 v = 1;
    for (int i = 0; i < n; i++)
        v *= 10;
    return v;
}

/* value = (neg ? -1 : 1) * mant * 10^-dscale */
DEVFN int
pgs_numeric_dscale(cl_ulong v)
{
    return -(int)PG_NUMERIC_EXPONENT(v);
}

DEVFN pg_numeric_t
pgs_numeric_make(cl_int *errcode, bool neg, pgs_u128 mant, int dscale)
{
    pg_numeric_t    r;

    if (mant > (pgs_u128)PGS_NUMERIC_MANT_LIMIT ||
        dscale > PGS_NUMERIC_MAX_DSCALE || dscale < -PG_NUMERIC_EXPONENT_MAX)
    {
        STROM_SET_ERROR(errcode, StromError_CpuReCheck);
        r.isnull = true;
        r.value = 0;
        return r;
    }
    r.isnull = false;
    r.value = PG_NUMERIC_SET(-dscale, neg && mant != 0, (cl_ulong)mant);
    return r;
}

/* signed mantissa of `v` rescaled to display scale `dscale` (>= its own) */
DEVFN pgs_s128
pgs_numeric_scaled(cl_ulong v, int dscale)
{
    pgs_s128    m = (pgs_s128)PG_NUMERIC_MANTISSA(v);
    int         up = dscale - pgs_numeric_dscale(v);

    m *= (pgs_s128)pgs_pow10_u128(up);      /* < 2^57 * 10^63: callers keep up <= 38 */
    return PG_NUMERIC_SIGN(v) ? -m : m;
}

/* -1 / 0 / +1; display scales do not take part (1.50 = 1.5) */
DEVFN int
pgs_numeric_cmp(cl_ulong a, cl_ulong b)
{
    int         sa = pgs_numeric_dscale(a), sb = pgs_numeric_dscale(b);
    int         s = (sa > sb ? sa : sb);
    pgs_s128    x, y;

    /* |sa - sb| <= 63; 2^57 * 10^63 does not fit 127 bits: compare in two
     * steps when the scales are far apart */
    if (s - sa > 20 || s - sb > 20)
    {
        /* bring the one with the larger scale down instead (truncating),
         * ties broken by the remainder */
        int     lo = (sa < sb ? sa : sb);
        pgs_u128 d = pgs_pow10_u128((s - lo) > 38 ? 38 : (s - lo));
        pgs_s128 big = (pgs_s128)PG_NUMERIC_MANTISSA(sa > sb ? a : b);
        pgs_s128 q = big / (pgs_s128)d, rem = big % (pgs_s128)d;
        if ((s - lo) > 38)
        {
            rem = (big != 0);       /* below one unit of the other's scale */
            q = 0;
        }
        if (PG_NUMERIC_SIGN(sa > sb ? a : b)) { q = -q; rem = -rem; }
        pgs_s128 other = (pgs_s128)PG_NUMERIC_MANTISSA(sa > sb ? b : a);
        if (PG_NUMERIC_SIGN(sa > sb ? b : a)) other = -other;
        int c = (q < other ? -1 : (q > other ? 1 : (rem < 0 ? -1 : (rem > 0 ? 1 : 0))));
        return (sa > sb ? c : -c);
    }
    x = pgs_numeric_scaled(a, s);
    y = pgs_numeric_scaled(b, s);
    return (x < y ? -1 : (x > y ? 1 : 0));
}

/*
 * PostgreSQL's varlena image (utils/adt/numeric.c: NumericShort / NumericLong,
 * base-10000 digits, weight = exponent of the first digit, display scale) ->
 * device format.  `p` points at the varlena header (1-byte or 4-byte form).
 */
DEVFN pg_numeric_t
pg_numeric_from_varlena(cl_int *errcode, const unsigned char *p)
{
    pg_numeric_t    r;
    const unsigned char *data;
    cl_uint     len, n_header, ndigits;
    int         dscale, weight, shift;
    bool        neg;
    pgs_u128    mant = 0;

    r.isnull = true;
    r.value = 0;
    if (!p)
        return r;
    if (p[0] == 0x01)
        goto recheck;                   /* external TOAST pointer */
    if (p[0] & 0x01)
    {
        len = (p[0] >> 1) & 0x7fU;
        if (len < 1 + 2)
            goto recheck;
        data = p + 1;
        len -= 1;
    }
    else
    {
        cl_uint hdr = (cl_uint)p[0] | ((cl_uint)p[1] << 8) | ((cl_uint)p[2] << 16) | ((cl_uint)p[3] << 24);
        if ((hdr & 0x03) != 0)
            goto recheck;               /* compressed in line */
        len = (hdr >> 2) & 0x3fffffffU;
        if (len < 4 + 2)
            goto recheck;
        data = p + 4;
        len -= 4;
    }
    n_header = (cl_uint)data[0] | ((cl_uint)data[1] << 8);
    if ((n_header & 0xC000) == 0xC000)
        goto recheck;                   /* NaN */
    if ((n_header & 0xC000) == 0x8000)
    {
        neg = (n_header & 0x2000) != 0;
        dscale = (int)((n_header & 0x1F80) >> 7);
        weight = (int)(n_header & 0x003F);
        if (n_header & 0x0040)
            weight |= ~0x3F;
        data += 2;
        len -= 2;
    }
    else
    {
        if (len < 4)
            goto recheck;
        neg = (n_header & 0xC000) == 0x4000;
        dscale = (int)(n_header & 0x3FFF);
        weight = (int)(short)((cl_uint)data[2] | ((cl_uint)data[3] << 8));
        data += 4;
        len -= 4;
    }
    ndigits = len / 2;
    if (dscale > PGS_NUMERIC_MAX_DSCALE || ndigits > 8)
        goto recheck;                   /* 8 digits of 10000 = 32 decimal digits > 2^57 */
    for (cl_uint i = 0; i < ndigits; i++)
        mant = mant * 10000 + ((cl_uint)data[2 * i] | ((cl_uint)data[2 * i + 1] << 8));
    /* value = mant * 10^(4 * (weight - ndigits + 1)); wanted: mant' * 10^-dscale */
    shift = 4 * (weight - (int)ndigits + 1) + dscale;
    if (mant == 0)
        shift = 0;
    if (shift > 0)
    {
        if (shift > 18 || mant > (pgs_u128)PGS_NUMERIC_MANT_LIMIT)
            goto recheck;
        mant *= pgs_pow10_u128(shift);
    }
    else if (shift < 0)
    {
        pgs_u128 d;
        if (shift < -38)
            goto recheck;
        d = pgs_pow10_u128(-shift);
        if (mant % d != 0)
            goto recheck;               /* digits below the display scale */
        mant /= d;
    }
    return pgs_numeric_make(errcode, neg, mant, dscale);
recheck:
    STROM_SET_ERROR(errcode, StromError_CpuReCheck);
    return r;
}

/* pg_numeric_vref: the staged / de-formed value of a varlena column is the
 * offset of the datum from the head of the chunk (0 = NULL) */
template <typename KDS>
DEVFN pg_numeric_t
pg_numeric_vref(const KDS &kds, const void *ktoast, cl_int *errcode,
                cl_uint colidx, cl_uint rowidx)
{
    cl_uint     offset = 0;
    bool        ok = kds.template fetch<cl_uint>(GPUPREAGG_INCOL_SLOT(colidx), rowidx, offset);

    if (!ok || offset == 0)
    {
        pg_numeric_t r;
        r.isnull = true;
        r.value = 0;
        return r;
    }
    return pg_numeric_from_varlena(errcode, (const unsigned char *)ktoast + offset);
}

DEVFN pg_numeric_t
pg_numeric_param(const kern_parambuf *kparams, cl_int *errcode, cl_uint param_id)
{
    if (param_id < kparams->nparams && kparams->poffset[param_id] > 0)
        return pg_numeric_from_varlena(errcode, (const unsigned char *)kparams +
                                       kparams->poffset[param_id]);
    pg_numeric_t r;
    r.isnull = true;
    r.value = 0;
    return r;
}

STROMCL_SIMPLE_NULLTEST_TEMPLATE(numeric)

/* NULL of a CASE without ELSE (also what an aggregate FILTER becomes) */
DEVFN pg_numeric_t
pg_numeric_null(void)
{
    pg_numeric_t r;

    r.isnull = true;
    r.value = 0;
    return r;
}

/* ---- arithmetic (opencl_numeric.h:816-1094): exact, or CpuReCheck ---- */
DEVFN pg_numeric_t
pgs_numeric_addsub(cl_int *errcode, pg_numeric_t a, pg_numeric_t b, bool sub)
{
    pg_numeric_t r;
    int     sa, sb, s;
    pgs_s128 x, y;

    r.isnull = a.isnull | b.isnull;
    r.value = 0;
    if (r.isnull)
        return r;
    sa = pgs_numeric_dscale(a.value);
    sb = pgs_numeric_dscale(b.value);
    s = (sa > sb ? sa : sb);
    if (s - sa > 20 || s - sb > 20)
    {
        STROM_SET_ERROR(errcode, StromError_CpuReCheck);
        r.isnull = true;
        return r;
    }
    x = pgs_numeric_scaled(a.value, s);
    y = pgs_numeric_scaled(b.value, s);
    x = (sub ? x - y : x + y);
    return pgs_numeric_make(errcode, x < 0, (pgs_u128)(x < 0 ? -x : x), s);
}
DEVFN pg_numeric_t
pgfn_numeric_add(cl_int *errcode, pg_numeric_t a, pg_numeric_t b)
{ return pgs_numeric_addsub(errcode, a, b, false); }
DEVFN pg_numeric_t
pgfn_numeric_sub(cl_int *errcode, pg_numeric_t a, pg_numeric_t b)
{ return pgs_numeric_addsub(errcode, a, b, true); }
DEVFN pg_numeric_t
pgfn_numeric_mul(cl_int *errcode, pg_numeric_t a, pg_numeric_t b)
{
    pg_numeric_t r;

    r.isnull = a.isnull | b.isnull;
    r.value = 0;
    if (r.isnull)
        return r;
    /* numeric_mul: display scale = sum of the scales */
    return pgs_numeric_make(errcode,
                            PG_NUMERIC_SIGN(a.value) != PG_NUMERIC_SIGN(b.value),
                            (pgs_u128)PG_NUMERIC_MANTISSA(a.value) *
                            (pgs_u128)PG_NUMERIC_MANTISSA(b.value),
                            pgs_numeric_dscale(a.value) + pgs_numeric_dscale(b.value));
}
DEVFN pg_numeric_t
pgfn_numeric_uplus(cl_int *errcode, pg_numeric_t a)
{ return a; }
DEVFN pg_numeric_t
pgfn_numeric_uminus(cl_int *errcode, pg_numeric_t a)
{
    if (!a.isnull && PG_NUMERIC_MANTISSA(a.value) != 0)
        a.value ^= PG_NUMERIC_SIGN_MASK;
    return a;
}
DEVFN pg_numeric_t
pgfn_numeric_abs(cl_int *errcode, pg_numeric_t a)
{
    if (!a.isnull)
        a.value &= ~PG_NUMERIC_SIGN_MASK;
    return a;
}

#define PGS_NUMERIC_COMPARE_TEMPLATE(NAME,OPER)                         \
    DEVFN pg_bool_t                                                     \
    pgfn_numeric_##NAME(cl_int *errcode, pg_numeric_t a, pg_numeric_t b) \
    {                                                                   \
        pg_bool_t r;                                                    \
        r.isnull = a.isnull | b.isnull;                                 \
        r.value = (!r.isnull && pgs_numeric_cmp(a.value, b.value) OPER 0); \
        return r;                                                       \
    }
PGS_NUMERIC_COMPARE_TEMPLATE(eq, ==)
PGS_NUMERIC_COMPARE_TEMPLATE(ne, !=)
PGS_NUMERIC_COMPARE_TEMPLATE(lt, <)
PGS_NUMERIC_COMPARE_TEMPLATE(le, <=)
PGS_NUMERIC_COMPARE_TEMPLATE(gt, >)
PGS_NUMERIC_COMPARE_TEMPLATE(ge, >=)
DEVFN pg_int4_t
pgfn_numeric_cmp(cl_int *errcode, pg_numeric_t a, pg_numeric_t b)
{
    pg_int4_t r;
    r.isnull = a.isnull | b.isnull;
    r.value = (r.isnull ? 0 : pgs_numeric_cmp(a.value, b.value));
    return r;
}

/* ---- casts (opencl_numeric.h:399-779) ---- */
DEVFN pg_float8_t
pgfn_numeric_float8(cl_int *errcode, pg_numeric_t a)
{
    pg_float8_t r;
    int     ds;

    r.isnull = a.isnull;
    r.value = 0.0;
    if (a.isnull)
        return r;
    /* PostgreSQL goes through the decimal text and strtod (correctly
     * rounded).  mantissa / 10^scale is the same whenever both are exact in
     * binary64 (one rounding); beyond that it may differ in the last bit */
    ds = pgs_numeric_dscale(a.value);
    r.value = (double)PG_NUMERIC_MANTISSA(a.value);
    if (ds > 0)
    {
        if (ds <= 22)
            r.value /= (double)pgs_pow10_u128(ds);
        else
            r.value = r.value / 1e22 / (double)pgs_pow10_u128(ds - 22);
    }
    else if (ds < 0)
        r.value *= (double)pgs_pow10_u128(-ds);
    if (PG_NUMERIC_SIGN(a.value))
        r.value = -r.value;
    return r;
}
DEVFN pg_float4_t
pgfn_numeric_float4(cl_int *errcode, pg_numeric_t a)
{
    pg_float8_t d = pgfn_numeric_float8(errcode, a);
    pg_float4_t r;
    r.isnull = d.isnull;
    r.value = (float)d.value;
    return r;
}
#define PGS_NUMERIC_FROM_INT_TEMPLATE(NAME)                             \
    DEVFN pg_numeric_t                                                  \
    pgfn_##NAME##_numeric(cl_int *errcode, pg_##NAME##_t a)             \
    {                                                                   \
        pg_numeric_t r;                                                 \
        cl_long v = (cl_long)a.value;                                   \
        r.isnull = a.isnull;                                            \
        r.value = 0;                                                    \
        if (a.isnull)                                                   \
            return r;                                                   \
        return pgs_numeric_make(errcode, v < 0,                         \
                                (pgs_u128)(v < 0 ? -(pgs_s128)v : (pgs_s128)v), 0); \
    }
PGS_NUMERIC_FROM_INT_TEMPLATE(int2)
PGS_NUMERIC_FROM_INT_TEMPLATE(int4)
PGS_NUMERIC_FROM_INT_TEMPLATE(int8)
/*
 * float8 / float4 -> numeric (numeric.c float8_numeric / float4_numeric):
 * PostgreSQL prints the value with "%.15g" (float4: "%.6g") and reads the
 * text back with numeric_in().  printf rounds the EXACT binary value to P
 * significant digits (ties to even), %g drops trailing zeros, and numeric_in
 * gives the display scale of what is left.  Here the same decimal is
 * computed with 128-bit integers:  |v| = m * 2^e,  q = floor(|v| * 10^k)
 * with k chosen so that 10^(P-1) <= q < 10^P, remainder kept as a fraction
 * num/den for the tie test.  Ranges that would need more than 128 bits
 * (|v| < ~1e-18 for float8, |v| >= 2^127) are re-checked on the host, like
 * everything that does not fit the 57-bit device format, NaN and +-Inf.
 */
DEVFN pg_numeric_t
pgs_float_to_numeric(cl_int *errcode, double v, int ndigits)
{
    pg_numeric_t r;
    union { double d; cl_ulong u; } cv;
    cv.d = v;
    cl_ulong    bits = cv.u;
    bool        neg = (bits >> 63) != 0;
    int         bexp = (int)((bits >> 52) & 0x7ff);
    cl_ulong    frac = bits & 0x000fffffffffffffULL;
    cl_ulong    m;
    int         e, x;
    pgs_u128    lo = pgs_pow10_u128(ndigits - 1), hi = lo * 10;
    pgs_u128    q = 0, num = 0, den = 1;
    int         k = 0;
    bool        found = false;

    r.isnull = true;
    r.value = 0;
    if (bexp == 0x7ff)
        goto recheck;               /* NaN (numeric NaN) / Inf (PostgreSQL error) */
    if (bexp == 0 && frac == 0)
        return pgs_numeric_make(errcode, false, 0, 0);
    if (bexp == 0)
        goto recheck;               /* subnormal: far below 10^-32 */
    m = frac | (1ULL << 52);
    e = bexp - 1075;
    /* floor(log10 |v|) estimated from the binary exponent (30103/100000 ~
     * log10 2), then corrected by the exact test below */
    x = (int)(((cl_long)(e + 52) * 30103LL) / 100000LL);
    if (e + 52 < 0)
        x -= 1;
    for (int attempt = 0; attempt < 4 && !found; attempt++)
    {
        k = ndigits - 1 - x;
        if (k >= 0)
        {
            int     s = e + k;
            if (k > 32)
                goto recheck;
            num = (pgs_u128)m * (pgs_pow10_u128(k) >> k);   /* m * 5^k < 2^128 */
            if (s >= 0)
            {
                if (s > 20)
                    goto recheck;   /* cannot happen for a correct x */
                q = num << s;
                num = 0;
                den = 1;
            }
            else
            {
                if (-s >= 127)
                    goto recheck;
                den = (pgs_u128)1 << (-s);
                q = num >> (-s);
                num = num & (den - 1);
            }
        }
        else
        {
            if (-k > 36)
                goto recheck;
            den = pgs_pow10_u128(-k);
            if (e >= 0)
            {
                if (e > 74)
                    goto recheck;   /* |v| >= 2^127 */
                num = (pgs_u128)m << e;
            }
            else
            {
                /* 10^-k <= |v| < 2^(53+e): den stays below 2^58 */
                if (-e > 52)
                    goto recheck;   /* cannot happen: |v| >= 10^P */
                num = (pgs_u128)m;
                den <<= (-e);
            }
            q = num / den;
            num = num % den;
        }
        if (q >= hi)
            x += 1;
        else if (q < lo)
            x -= 1;
        else
            found = true;
    }
    if (!found)
        goto recheck;
    /* round to nearest, ties to even: 2*num <=> den without overflow */
    if (num > den - num || (num == den - num && (q & 1)))
        q += 1;
    while (k > 0 && (q % 10) == 0)
    {
        q /= 10;
        k--;
    }
    if (k < 0)
    {
        if (-k > 17)
            goto recheck;           /* >= 10^(P-1+18): beyond 57 bits */
        q *= pgs_pow10_u128(-k);
        k = 0;
    }
    return pgs_numeric_make(errcode, neg, q, k);
recheck:
    STROM_SET_ERROR(errcode, StromError_CpuReCheck);
    return r;
}

DEVFN pg_numeric_t
pgfn_float8_numeric(cl_int *errcode, pg_float8_t a)
{
    pg_numeric_t r;

    r.isnull = true;
    r.value = 0;
    if (a.isnull)
        return r;
    return pgs_float_to_numeric(errcode, a.value, 15);      /* DBL_DIG */
}

DEVFN pg_numeric_t
pgfn_float4_numeric(cl_int *errcode, pg_float4_t a)
{
    pg_numeric_t r;

    r.isnull = true;
    r.value = 0;
    if (a.isnull)
        return r;
    return pgs_float_to_numeric(errcode, (double)a.value, 6);   /* FLT_DIG */
}

/* numeric -> integer: round half away from zero (numeric.c numericvar_to_int8) */
DEVFN cl_long
pgs_numeric_to_long(cl_int *errcode, cl_ulong v, cl_long lo, cl_long hi, bool *isnull)
{
    int         ds = pgs_numeric_dscale(v);
    pgs_s128    m = (pgs_s128)PG_NUMERIC_MANTISSA(v);

    if (ds > 0)
    {
        pgs_s128 d = (pgs_s128)pgs_pow10_u128(ds);
        pgs_s128 q = m / d, rem = m % d;
        if (rem * 2 >= d)
            q += 1;
        m = q;
    }
    else if (ds < 0)
        m *= (pgs_s128)pgs_pow10_u128(-ds);
    if (PG_NUMERIC_SIGN(v))
        m = -m;
    if (m < (pgs_s128)lo || m > (pgs_s128)hi)
    {
        /* PostgreSQL raises "out of range": let it */
        STROM_SET_ERROR(errcode, StromError_CpuReCheck);
        *isnull = true;
        return 0;
    }
    return (cl_long)m;
}
#define PGS_NUMERIC_TO_INT_TEMPLATE(NAME,BASE,LO,HI)                    \
    DEVFN pg_##NAME##_t                                                 \
    pgfn_numeric_##NAME(cl_int *errcode, pg_numeric_t a)                \
    {                                                                   \
        pg_##NAME##_t r;                                                \
        r.isnull = a.isnull;                                            \
        r.value = 0;                                                    \
        if (!a.isnull)                                                  \
            r.value = (BASE)pgs_numeric_to_long(errcode, a.value, LO, HI, &r.isnull); \
        return r;                                                       \
    }
PGS_NUMERIC_TO_INT_TEMPLATE(int2, cl_short, -32768LL, 32767LL)
PGS_NUMERIC_TO_INT_TEMPLATE(int4, cl_int, -2147483648LL, 2147483647LL)
PGS_NUMERIC_TO_INT_TEMPLATE(int8, cl_long, LONG_MIN, LONG_MAX)

#endif  /* KERN_NUMERIC_CUH */
