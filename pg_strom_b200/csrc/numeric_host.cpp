/*
 * numeric_host.cpp - PostgreSQL NUMERIC varlena <-> decimal text on the host.
 *
 * The chunk loader copies numeric datums verbatim from heap tuples; this
 * file exists for the places where the library itself has to make one
 * (numeric Const in kern_parambuf, datastore.c:74-81) and for test
 * harnesses that have no PostgreSQL to do it.  Format: utils/adt/numeric.c
 * (NumericShort / NumericLong, base-10000 digits), the same the device
 * parser pg_numeric_from_varlena reads (opencl_numeric.h:166-307).
 */
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#define NBASE           10000
#define DEC_DIGITS      4
#define NUMERIC_POS     0x0000
#define NUMERIC_NEG     0x4000
#define NUMERIC_SHORT   0x8000
#define NUMERIC_NAN     0xC000
#define NUMERIC_SHORT_SIGN_MASK         0x2000
#define NUMERIC_SHORT_DSCALE_SHIFT      7
#define NUMERIC_SHORT_DSCALE_MAX        63
#define NUMERIC_SHORT_WEIGHT_SIGN_MASK  0x0040
#define NUMERIC_SHORT_WEIGHT_MASK       0x003F
#define NUMERIC_SHORT_WEIGHT_MAX        63
#define NUMERIC_SHORT_WEIGHT_MIN        (-64)
#define NUMERIC_DSCALE_MASK             0x3FFF

namespace pgs {

/* decimal text -> varlena image with a 4-byte header */
bool
pgs_numeric_from_text(const std::string &text, std::vector<unsigned char> &out)
{
    size_t i = 0, n = text.size();
    bool neg = false;
    std::string intpart, fracpart;
    long exp10 = 0;

    while (i < n && text[i] == ' ') i++;
    if (i < n && (text[i] == '+' || text[i] == '-'))
        neg = (text[i++] == '-');
    if (n - i >= 3 && (text.compare(i, 3, "NaN") == 0 || text.compare(i, 3, "nan") == 0))
    {
        uint32_t hdr = (uint32_t)((4 + 2) << 2);
        uint16_t h = NUMERIC_NAN;
        out.resize(6);
        memcpy(&out[0], &hdr, 4);
        memcpy(&out[4], &h, 2);
        return true;
    }
    while (i < n && text[i] >= '0' && text[i] <= '9')
        intpart += text[i++];
    if (i < n && text[i] == '.')
    {
        i++;
        while (i < n && text[i] >= '0' && text[i] <= '9')
            fracpart += text[i++];
    }
    if (intpart.empty() && fracpart.empty())
        return false;
    if (i < n && (text[i] == 'e' || text[i] == 'E'))
    {
        char *end;
        exp10 = strtol(text.c_str() + i + 1, &end, 10);
        i = end - text.c_str();
    }
    while (i < n && text[i] == ' ') i++;
    if (i != n)
        return false;
    /* apply the exponent by moving the decimal point */
    std::string digits = intpart + fracpart;
    long point = (long)intpart.size() + exp10;      /* digits before the point */
    long dscale = (long)digits.size() - point;
    if (dscale < 0)
    {
        digits.append((size_t)(-dscale), '0');
        dscale = 0;
    }
    if (point < 0)
    {
        digits.insert(0, (size_t)(-point), '0');
        point = 0;
    }
    if (dscale > NUMERIC_DSCALE_MASK)
        return false;
    /* align to base-10000 digit boundaries around the decimal point */
    long lead = (DEC_DIGITS - (point % DEC_DIGITS)) % DEC_DIGITS;
    std::string padded(lead, '0');
    padded += digits;
    long ipart_len = point + lead;
    while ((padded.size() - ipart_len) % DEC_DIGITS != 0)
        padded += '0';
    std::vector<int16_t> nd;
    for (size_t k = 0; k < padded.size(); k += DEC_DIGITS)
        nd.push_back((int16_t)atoi(padded.substr(k, DEC_DIGITS).c_str()));
    long weight = ipart_len / DEC_DIGITS - 1;
    /* strip leading / trailing zero digits */
    size_t first = 0;
    while (first < nd.size() && nd[first] == 0)
    {
        first++;
        weight--;
    }
    size_t last = nd.size();
    while (last > first && nd[last - 1] == 0)
        last--;
    size_t ndigits = last - first;
    if (ndigits == 0)
    {
        weight = 0;
        neg = false;
    }
    bool can_short = (dscale <= NUMERIC_SHORT_DSCALE_MAX &&
                      weight <= NUMERIC_SHORT_WEIGHT_MAX &&
                      weight >= NUMERIC_SHORT_WEIGHT_MIN);
    size_t hdrsz = can_short ? 2 : 4;
    size_t len = 4 + hdrsz + ndigits * 2;
    uint32_t vl = (uint32_t)(len << 2);
    out.assign(len, 0);
    memcpy(&out[0], &vl, 4);
    if (can_short)
    {
        uint16_t h = (uint16_t)(NUMERIC_SHORT |
                                (neg ? NUMERIC_SHORT_SIGN_MASK : 0) |
                                (dscale << NUMERIC_SHORT_DSCALE_SHIFT) |
                                (weight < 0 ? NUMERIC_SHORT_WEIGHT_SIGN_MASK : 0) |
                                (weight & NUMERIC_SHORT_WEIGHT_MASK));
        memcpy(&out[4], &h, 2);
    }
    else
    {
        uint16_t sd = (uint16_t)((neg ? NUMERIC_NEG : NUMERIC_POS) | (dscale & NUMERIC_DSCALE_MASK));
        int16_t w = (int16_t)weight;
        memcpy(&out[4], &sd, 2);
        memcpy(&out[6], &w, 2);
    }
    for (size_t k = 0; k < ndigits; k++)
        memcpy(&out[4 + hdrsz + 2 * k], &nd[first + k], 2);
    return true;
}

/* varlena image (4-byte or 1-byte header) -> decimal text */
bool
pgs_numeric_to_text(const unsigned char *vl, std::string &out)
{
    const unsigned char *data;
    size_t len;

    if (vl[0] & 0x01)
    {
        len = ((vl[0] >> 1) & 0x7F) - 1;
        data = vl + 1;
    }
    else
    {
        uint32_t hdr;
        memcpy(&hdr, vl, 4);
        len = ((hdr >> 2) & 0x3FFFFFFF) - 4;
        data = vl + 4;
    }
    if (len < 2)
        return false;
    uint16_t h;
    memcpy(&h, data, 2);
    int sign, dscale, weight;
    size_t off;
    if ((h & 0xC000) == NUMERIC_NAN)
    {
        out = "NaN";
        return true;
    }
    if ((h & 0xC000) == NUMERIC_SHORT)
    {
        sign = (h & NUMERIC_SHORT_SIGN_MASK) ? NUMERIC_NEG : NUMERIC_POS;
        dscale = (h >> NUMERIC_SHORT_DSCALE_SHIFT) & 0x3F;
        weight = (h & NUMERIC_SHORT_WEIGHT_SIGN_MASK ? ~NUMERIC_SHORT_WEIGHT_MASK : 0) |
            (h & NUMERIC_SHORT_WEIGHT_MASK);
        off = 2;
    }
    else
    {
        int16_t w;
        sign = h & 0xC000;
        dscale = h & NUMERIC_DSCALE_MASK;
        memcpy(&w, data + 2, 2);
        weight = w;
        off = 4;
    }
    size_t ndigits = (len - off) / 2;
    std::string s;
    if (sign == NUMERIC_NEG)
        s += '-';
    /* integer part */
    if (weight < 0)
        s += '0';
    else
        for (int d = 0; d <= weight; d++)
        {
            int16_t dig = 0;
            char buf[8];
            if ((size_t)d < ndigits)
                memcpy(&dig, data + off + 2 * d, 2);
            if (d == 0)
                snprintf(buf, sizeof(buf), "%d", dig);
            else
                snprintf(buf, sizeof(buf), "%04d", dig);
            s += buf;
        }
    if (dscale > 0)
    {
        std::string frac;
        for (int d = weight + 1; (int)frac.size() < dscale; d++)
        {
            int16_t dig = 0;
            char buf[8];
            if (d >= 0 && (size_t)d < ndigits)
                memcpy(&dig, data + off + 2 * d, 2);
            snprintf(buf, sizeof(buf), "%04d", dig);
            frac += buf;
        }
        s += '.';
        s += frac.substr(0, dscale);
    }
    out = s;
    return true;
}

}   /* namespace pgs */

extern "C" {

/* returns the datum length, 0 on a syntax error or if buf is too small */
size_t
pgstrom_numeric_from_text(const char *text, void *buf, size_t buflen)
{
    std::vector<unsigned char> v;
    if (!text || !pgs::pgs_numeric_from_text(text, v) || v.size() > buflen)
        return 0;
    memcpy(buf, v.data(), v.size());
    return v.size();
}

size_t
pgstrom_numeric_to_text(const void *varlena, char *buf, size_t buflen)
{
    std::string s;
    if (!varlena || !pgs::pgs_numeric_to_text((const unsigned char *)varlena, s) ||
        s.size() + 1 > buflen)
        return 0;
    memcpy(buf, s.c_str(), s.size() + 1);
    return s.size();
}

}
