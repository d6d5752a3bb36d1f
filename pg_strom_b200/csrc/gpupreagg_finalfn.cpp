/*
 * gpupreagg_finalfn.cpp - the SQL-side half of GpuPreAgg: the arithmetic of
 * the partial placeholder functions and of the final accumulators
 * (gpupreagg.c:4251-4773, catalog pg_strom--1.0.sql:99-401), as plain C.
 *
 * In the extension every function here sits under a one-line fmgr V1 wrapper
 *  which unpacks PG_FUNCTION_ARGS / the transition
 * array and turns a non-zero return code into ereport(ERROR).  The
 * placeholders run on the host only for rows the device hands back
 * (gpupreagg_recheck_rows(), the reference's gpupreagg_next_tuple_fallback,
 * gpupreagg.c:2507-2607); the accumulators run for every partial row in
 * PostgreSQL's final Agg node.  Nothing here touches the device.
 *
 * Deliberate differences from the reference, each towards PostgreSQL's own
 * result:
 *   - pcov_*: the reference reads argument 0 (the FILTER flag) and 1 where
 *     it means X = argument 1 and Y = argument 2 (gpupreagg.c:4344-4417); the
 *     device code it generates (gpupreagg.c:1413-1427) uses X and Y, and so
 *     does this file.
 *   - covariance accum: the range check after "newSumY = ..." tests newSumX
 *     in the reference (gpupreagg.c:4729); newSumY is checked here.
 *   - int8 / numeric avg: N grows by exactly nrows; the reference lets
 *     int8_avg_accum count one row and then adds nrows - 1 only if nrows > 0
 *     (gpupreagg.c:4556-4561,4581-4586), one too many for a partial row with
 *     nrows = 0 and a non-NULL sum.
 */
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <algorithm>

#include "pgstrom_cuda.h"

/* ------------------------------------------------------------------------
 * partial placeholders
 * ------------------------------------------------------------------------ */

/* gpupreagg.c:4251-4262 gpupreagg_partial_nrows(): 1 if every argument is a
 * non-NULL true, else 0; no argument = count(*) */
extern "C" int32_t
pgs_partial_nrows(int nargs, const char *values, const char *isnull)
{
    for (int i = 0; i < nargs; i++)
        if ((isnull && isnull[i]) || !values[i])
            return 0;
    return 1;
}

/* gpupreagg.c:4308-4316 gpupreagg_psum_x2_float(); returns the isnull flag */
extern "C" int
pgs_psum_x2_float8(double x, int x_isnull, double *result)
{
    if (x_isnull)
        return 1;
    *result = x * x;
    return 0;
}

/* gpupreagg.c:4344-4417 gpupreagg_corr_psum_{x,y,x2,y2,xy}(filter, X, Y):
 * NULL unless the FILTER flag is a non-NULL true and both X and Y are
 * non-NULL.  Returns the isnull flag. */
extern "C" int
pgs_pcov_float8(int kind, int filter, int filter_isnull,
                double x, int x_isnull, double y, int y_isnull, double *result)
{
    if (filter_isnull || !filter || x_isnull || y_isnull)
        return 1;
    switch (kind)
    {
        case PGS_PCOV_X:  *result = x;     break;
        case PGS_PCOV_Y:  *result = y;     break;
        case PGS_PCOV_X2: *result = x * x; break;
        case PGS_PCOV_Y2: *result = y * y; break;
        case PGS_PCOV_XY: *result = x * y; break;
        default: return 1;
    }
    return 0;
}

/* ------------------------------------------------------------------------
 * final accumulators over int8[2] / float8[3] / float8[6] transition arrays
 * (all STRICT in the catalog: PostgreSQL skips a partial row with a NULL in
 * it before the call)
 * ------------------------------------------------------------------------ */

/* gpupreagg.c:4434-4468 pgstrom_avg_int8_accum(): {N += nrows, S += psum} */
extern "C" int
pgs_avg_int8_accum(int64_t *trans, int32_t nrows, int64_t psum)
{
    trans[0] += nrows;
    trans[1] = (int64_t) ((uint64_t) trans[1] + (uint64_t) psum);
    return 0;
}

/* gpupreagg.c:4470-4500 pgstrom_sum_int8_accum(): {dummy, S += psum}.
 * trans[0] counts the partial rows instead of staying 0: pgs_sum_int8_final
 * needs it to return NULL, not 0, for a group without a non-NULL input
 * (what PostgreSQL's own sum(int4) returns; the reference returns 0) */
extern "C" int
pgs_sum_int8_accum(int64_t *trans, int64_t psum)
{
    trans[0] += 1;
    trans[1] = (int64_t) ((uint64_t) trans[1] + (uint64_t) psum);
    return 0;
}

/* gpupreagg.c:4508-4517 pgstrom_sum_int8_final(); returns the isnull flag */
extern "C" int
pgs_sum_int8_final(const int64_t *trans, int64_t *result)
{
    if (trans[0] == 0)
        return 1;
    *result = trans[1];
    return 0;
}

/* float.c check_float8_valid() as copied at gpupreagg.c:4608-4620 with
 * zero_is_valid = true */
static inline int
float8_sum_check(double newval, double oldval, double addend)
{
    if (std::isinf(newval) && !(std::isinf(oldval) || std::isinf(addend)))
        return PGS_FINALFN_OVERFLOW;
    return 0;
}

/* gpupreagg.c:4622-4661 pgstrom_sum_float8_accum(): {N, SX, 0} */
extern "C" int
pgs_sum_float8_accum(double *trans, int32_t nrows, double psum)
{
    double  newN = trans[0] + (double) nrows;
    double  newSumX = trans[1] + psum;
    int     rc = float8_sum_check(newSumX, trans[1], psum);

    if (rc)
        return rc;
    trans[0] = newN;
    trans[1] = newSumX;
    trans[2] = 0.0;
    return 0;
}

/* gpupreagg.c:4668-4713 pgstrom_variance_float8_accum(): {N, SX, SX2} */
extern "C" int
pgs_variance_float8_accum(double *trans, int32_t nrows, double psum, double psum_x2)
{
    double  newN = trans[0] + (double) nrows;
    double  newSumX = trans[1] + psum;
    double  newSumX2 = trans[2] + psum_x2;
    int     rc;

    if ((rc = float8_sum_check(newSumX, trans[1], psum)) != 0 ||
        (rc = float8_sum_check(newSumX2, trans[2], psum_x2)) != 0)
        return rc;
    trans[0] = newN;
    trans[1] = newSumX;
    trans[2] = newSumX2;
    return 0;
}

/* gpupreagg.c:4719-4787 pgstrom_covariance_float8_accum():
 * {N, SX, SX2, SY, SY2, SXY}; psum[] = {pcov_x, pcov_x2, pcov_y, pcov_y2, pcov_xy} */
extern "C" int
pgs_covariance_float8_accum(double *trans, int32_t nrows, const double *psum)
{
    double  next[6];

    next[0] = trans[0] + (double) nrows;
    for (int i = 1; i < 6; i++)
    {
        int rc;

        next[i] = trans[i] + psum[i - 1];
        if ((rc = float8_sum_check(next[i], trans[i], psum[i - 1])) != 0)
            return rc;
    }
    memcpy(trans, next, sizeof(next));
    return 0;
}

/* ------------------------------------------------------------------------
 * numeric transition state of avg(int8) / avg(numeric)
 * (gpupreagg.c:4519-4588; in the extension this is PostgreSQL's own
 * NumericAggState driven through int8_avg_accum / numeric_avg_accum - the
 * exact decimal sum below is what a harness without PostgreSQL uses, and the
 * executable statement of the N rule)
 * ------------------------------------------------------------------------ */
struct pgs_numeric_avg_state
{
    int64_t     N = 0;
    bool        neg = false;
    bool        nan = false;
    std::string digits = "0";   /* |sum| * 10^scale, no leading zeros */
    int         scale = 0;      /* largest display scale seen */
};

static int
mag_cmp(const std::string &a, const std::string &b)
{
    if (a.size() != b.size())
        return a.size() < b.size() ? -1 : 1;
    return a.compare(b) < 0 ? -1 : (a == b ? 0 : 1);
}

static std::string
mag_strip(std::string s)
{
    size_t p = s.find_first_not_of('0');
    return p == std::string::npos ? std::string("0") : s.substr(p);
}

static std::string
mag_add(const std::string &a, const std::string &b)
{
    std::string r;
    int carry = 0;
    for (size_t i = 0; i < std::max(a.size(), b.size()) || carry; i++)
    {
        int d = carry;
        if (i < a.size()) d += a[a.size() - 1 - i] - '0';
        if (i < b.size()) d += b[b.size() - 1 - i] - '0';
        r.push_back((char) ('0' + d % 10));
        carry = d / 10;
    }
    std::reverse(r.begin(), r.end());
    return mag_strip(r);
}

static std::string
mag_sub(const std::string &a, const std::string &b)    /* a >= b */
{
    std::string r;
    int borrow = 0;
    for (size_t i = 0; i < a.size(); i++)
    {
        int d = (a[a.size() - 1 - i] - '0') - borrow;
        if (i < b.size()) d -= b[b.size() - 1 - i] - '0';
        borrow = d < 0;
        r.push_back((char) ('0' + (d + 10) % 10));
    }
    std::reverse(r.begin(), r.end());
    return mag_strip(r);
}

/* "[-]ddd[.ddd][e[+-]n]" (what numeric_out and pgstrom_fixup_kernel_numeric
 * print) -> sign, digit string, scale */
static bool
parse_decimal(const char *text, bool &neg, std::string &digits, int &scale, bool &nan)
{
    const char *p = text;
    std::string ip, fp;
    long e = 0;

    neg = nan = false;
    while (*p == ' ') p++;
    if (!strncmp(p, "NaN", 3) || !strncmp(p, "nan", 3))
    {
        nan = true;
        return true;
    }
    if (*p == '+' || *p == '-')
        neg = (*p++ == '-');
    while (*p >= '0' && *p <= '9') ip.push_back(*p++);
    if (*p == '.')
        for (p++; *p >= '0' && *p <= '9'; p++) fp.push_back(*p);
    if (ip.empty() && fp.empty())
        return false;
    if (*p == 'e' || *p == 'E')
    {
        char *end;
        e = strtol(p + 1, &end, 10);
        if (end == p + 1 || e > 100000 || e < -100000)
            return false;
        p = end;
    }
    while (*p == ' ') p++;
    if (*p)
        return false;
    digits = ip + fp;
    long sc = (long) fp.size() - e;
    if (sc < 0)
    {
        digits.append((size_t) -sc, '0');
        sc = 0;
    }
    scale = (int) sc;
    digits = mag_strip(digits);
    if (digits == "0")
        neg = false;
    return true;
}

extern "C" pgs_numeric_avg_state *
pgs_numeric_avg_init(void)
{
    return new pgs_numeric_avg_state();
}

extern "C" void
pgs_numeric_avg_free(pgs_numeric_avg_state *state)
{
    delete state;
}

/* pgstrom_int8_avg_accum / pgstrom_numeric_avg_accum (nrows int4, psum
 * numeric; not STRICT: a NULL psum leaves the state alone like
 * int8_avg_accum does, a NULL or negative nrows is the reference's
 * "Bug? NULL or negative nrows was given") */
extern "C" int
pgs_numeric_avg_accum(pgs_numeric_avg_state *state, int32_t nrows, int nrows_isnull,
                      const char *psum_text)
{
    bool neg, nan;
    std::string digits;
    int scale;

    if (nrows_isnull || nrows < 0)
        return PGS_FINALFN_BAD_NROWS;
    if (!psum_text)
        return 0;
    if (!parse_decimal(psum_text, neg, digits, scale, nan))
        return PGS_FINALFN_BAD_NUMERIC;
    state->N += nrows;
    if (nan)
    {
        state->nan = true;
        return 0;
    }
    if (scale > state->scale)
    {
        if (state->digits != "0")
            state->digits.append((size_t) (scale - state->scale), '0');
        state->scale = scale;
    }
    else if (scale < state->scale && digits != "0")
        digits.append((size_t) (state->scale - scale), '0');
    if (neg == state->neg)
        state->digits = mag_add(state->digits, digits);
    else if (mag_cmp(state->digits, digits) >= 0)
        state->digits = mag_sub(state->digits, digits);
    else
    {
        state->digits = mag_sub(digits, state->digits);
        state->neg = neg;
    }
    if (state->digits == "0")
        state->neg = false;
    return 0;
}

extern "C" int64_t
pgs_numeric_avg_count(const pgs_numeric_avg_state *state)
{
    return state->N;
}

/* the sum as numeric_out prints it (display scale = the largest seen);
 * returns the length, 0 if buf is too small */
extern "C" size_t
pgs_numeric_avg_sum_text(const pgs_numeric_avg_state *state, char *buf, size_t buflen)
{
    std::string out;

    if (state->nan)
        out = "NaN";
    else
    {
        std::string d = state->digits;
        if ((int) d.size() <= state->scale)
            d.insert(0, (size_t) (state->scale - (int) d.size() + 1), '0');
        if (state->neg)
            out = "-";
        out += d.substr(0, d.size() - (size_t) state->scale);
        if (state->scale > 0)
            out += "." + d.substr(d.size() - (size_t) state->scale);
    }
    if (out.size() + 1 > buflen)
        return 0;
    memcpy(buf, out.c_str(), out.size() + 1);
    return out.size();
}
