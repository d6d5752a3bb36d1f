/*
 * datastore.cpp - kern_data_store builders and readers (host side).
 *
 * Mirrors datastore.c of the reference: init_kern_data_store (:312-380),
 * the TUPSLOT creator (:501-529), pgstrom_fetch_data_store (:169-242) and
 * pgstrom_fixup_kernel_numeric (:150-167).  The column-format builder is new
 * (see pgstrom_kds.h); it is what a chunk loader calls instead of
 * pgstrom_data_store_insert_block (:556-710) when it de-forms heap tuples on
 * the host.
 */
#include <cstdio>
#include <cstring>
#include <string>
#include "../../include/pgstrom_cuda.h"

namespace pgs { extern thread_local std::string last_error; }

static inline size_t
align_up(size_t v, size_t a)
{
    return (v + a - 1) & ~(a - 1);
}

/* size of a varlena datum as it sits in a tuple (VARSIZE_ANY) */
static inline size_t
varsize_any(const unsigned char *p)
{
    if (p[0] == 0x01)                   /* external (1B_E): tag decides */
        return 2 + (p[1] == 1 ? 8 : 16);
    if (p[0] & 0x01)                    /* short 1-byte header */
        return (p[0] >> 1) & 0x7F;
    uint32_t hdr;
    memcpy(&hdr, p, 4);
    return (hdr >> 2) & 0x3FFFFFFF;
}

extern "C" {

size_t
pgstrom_kds_head_length(int ncols)
{
    return KERN_DATA_STORE_HEAD_LENGTH(ncols);
}

static void
init_kern_data_store(kern_data_store *kds, int ncols, const kern_colmeta *colmeta,
                     size_t length, uint32_t nrooms, int format)
{
    memset(kds, 0, offsetof(kern_data_store, colmeta));
    kds->hostptr = (hostptr_t)(uintptr_t)kds;
    kds->length = (cl_uint)length;
    kds->usage = 0;
    kds->ncols = (cl_uint)ncols;
    kds->nitems = 0;
    kds->nrooms = nrooms;
    kds->nblocks = 0;
    kds->maxblocks = 0;
    kds->format = (cl_char)format;
    kds->tdhasoid = 0;
    kds->tdtypeid = 2249;       /* RECORDOID */
    kds->tdtypmod = -1;
    memcpy(kds->colmeta, colmeta, sizeof(kern_colmeta) * ncols);
}

size_t
pgstrom_kds_column_length(int ncols, const kern_colmeta *colmeta, uint32_t nrows,
                          const void *const *values, const uint8_t *const *isnull)
{
    size_t len = KERN_DATA_STORE_COLUMN_HEAD_LENGTH(ncols);
    /* arrays are padded to whole tiles of KDS_COLUMN_ROW_QUANTUM rows so a
     * 16-byte granular bulk copy of the last tile never leaves the chunk */
    size_t prows = align_up(nrows, KDS_COLUMN_ROW_QUANTUM);

    for (int c = 0; c < ncols; c++)
    {
        int attlen = colmeta[c].attlen;
        bool has_null = false;

        /* a column the query does not reference is simply not loaded:
         * values[c] == NULL => values_offset = 0, no bytes cross the bus */
        if (!values || !values[c])
            continue;
        if (attlen > 0)
            len += align_up(prows * (size_t)attlen, KDS_COLUMN_ALIGN);
        else
            len += align_up(prows * sizeof(cl_uint), KDS_COLUMN_ALIGN);
        if (isnull && isnull[c])
            for (uint32_t r = 0; r < nrows && !has_null; r++)
                has_null = (isnull[c][r] != 0);
        if (attlen < 0 && values && values[c])
        {
            const void *const *ptrs = (const void *const *)values[c];
            for (uint32_t r = 0; r < nrows; r++)
            {
                if (!ptrs[r])
                    has_null = true;
                else if (!(isnull && isnull[c] && isnull[c][r]))
                    len += align_up(varsize_any((const unsigned char *)ptrs[r]), 4);
            }
        }
        if (has_null)
            len += align_up(prows / 8, KDS_COLUMN_ALIGN);
    }
    return align_up(len, KDS_COLUMN_ALIGN);
}

int
pgstrom_kds_column_build(void *buffer, size_t buflen, int ncols,
                         const kern_colmeta *colmeta, uint32_t nrows,
                         const void *const *values, const uint8_t *const *isnull)
{
    size_t need = pgstrom_kds_column_length(ncols, colmeta, nrows, values, isnull);
    kern_data_store *kds = (kern_data_store *)buffer;
    size_t prows = align_up(nrows, KDS_COLUMN_ROW_QUANTUM);
    size_t pos;

    if (need > buflen || need > 0xffffffffULL)
    {
        pgs::last_error = "column store does not fit the buffer";
        return StromError_DataStoreNoSpace;
    }
    if (((uintptr_t)buffer & (KDS_COLUMN_ALIGN - 1)) != 0)
    {
        pgs::last_error = "chunk buffer must be 128-byte aligned";
        return StromError_BadRequestMessage;
    }
    memset(buffer, 0, need);
    init_kern_data_store(kds, ncols, colmeta, need, nrows, KDS_FORMAT_COLUMN);
    kds->nitems = nrows;
    pos = KERN_DATA_STORE_COLUMN_HEAD_LENGTH(ncols);
    /* pass 1: fixed arrays and bitmaps; pass 2: varlena pool */
    for (int c = 0; c < ncols; c++)
    {
        kern_colpos *cpos = KERN_DATA_STORE_COLPOS(kds, c);
        int     attlen = colmeta[c].attlen;
        bool    has_null = false;
        const uint8_t *nulls = (isnull ? isnull[c] : NULL);

        if (!values || !values[c])
        {
            cpos->values_offset = 0;
            cpos->nullmap_offset = 0;
            continue;
        }
        cpos->values_offset = (cl_uint)pos;
        if (attlen > 0)
        {
            memcpy((char *)buffer + pos, values[c], (size_t)nrows * attlen);
            pos += align_up(prows * (size_t)attlen, KDS_COLUMN_ALIGN);
        }
        else
            pos += align_up(prows * sizeof(cl_uint), KDS_COLUMN_ALIGN);
        if (nulls)
            for (uint32_t r = 0; r < nrows && !has_null; r++)
                has_null = (nulls[r] != 0);
        if (attlen < 0 && values && values[c])
        {
            const void *const *ptrs = (const void *const *)values[c];
            for (uint32_t r = 0; r < nrows && !has_null; r++)
                has_null = (ptrs[r] == NULL);
        }
        if (has_null)
        {
            unsigned char *bm = (unsigned char *)buffer + pos;
            cpos->nullmap_offset = (cl_uint)pos;
            for (uint32_t r = 0; r < nrows; r++)
            {
                bool isn = (nulls && nulls[r]);
                if (attlen < 0 && values && values[c] &&
                    ((const void *const *)values[c])[r] == NULL)
                    isn = true;
                if (!isn)
                    bm[r >> 3] |= (unsigned char)(1U << (r & 7));
            }
            pos += align_up(prows / 8, KDS_COLUMN_ALIGN);
        }
        else
            cpos->nullmap_offset = 0;
    }
    for (int c = 0; c < ncols; c++)
    {
        if (colmeta[c].attlen > 0 || !values || !values[c])
            continue;
        kern_colpos *cpos = KERN_DATA_STORE_COLPOS(kds, c);
        cl_uint *offs = (cl_uint *)((char *)buffer + cpos->values_offset);
        const void *const *ptrs = (const void *const *)values[c];
        const uint8_t *nulls = (isnull ? isnull[c] : NULL);
        for (uint32_t r = 0; r < nrows; r++)
        {
            if (!ptrs[r] || (nulls && nulls[r]))
            {
                offs[r] = 0;
                continue;
            }
            size_t sz = varsize_any((const unsigned char *)ptrs[r]);
            offs[r] = (cl_uint)pos;
            memcpy((char *)buffer + pos, ptrs[r], sz);
            pos += align_up(sz, 4);
        }
    }
    kds->usage = (cl_uint)pos;
    return StromError_Success;
}

size_t
pgstrom_kds_tupslot_length(int ncols, uint32_t nrooms)
{
    return STROMALIGN(KERN_DATA_STORE_HEAD_LENGTH(ncols) +
                      KERN_DATA_STORE_SLOT_STRIDE(ncols) * (size_t)nrooms);
}

int
pgstrom_kds_tupslot_init(void *buffer, size_t buflen, int ncols,
                         const kern_colmeta *colmeta, uint32_t nrooms)
{
    size_t need = pgstrom_kds_tupslot_length(ncols, nrooms);
    if (need > buflen || need > 0xffffffffULL)
    {
        pgs::last_error = "tuple-slot store does not fit the buffer";
        return StromError_DataStoreNoSpace;
    }
    init_kern_data_store((kern_data_store *)buffer, ncols, colmeta, need,
                         nrooms, KDS_FORMAT_TUPSLOT);
    return StromError_Success;
}

int
pgstrom_fetch_data_store(const kern_data_store *kds, uint32_t row,
                         Datum *values, char *isnull)
{
    if (kds->format != KDS_FORMAT_TUPSLOT)
    {
        pgs::last_error = "pgstrom_fetch_data_store: only TUPSLOT stores can be fetched here";
        return StromError_BadRequestMessage;
    }
    if (row >= kds->nitems)
        return StromError_DataStoreOutOfRange;
    const Datum *v = KERN_DATA_STORE_VALUES(kds, row);
    const cl_char *n = KERN_DATA_STORE_ISNULL(kds, row);
    memcpy(values, v, sizeof(Datum) * kds->ncols);
    memcpy(isnull, n, kds->ncols);
    return StromError_Success;
}

int
pgstrom_fixup_kernel_numeric(Datum datum, char *buf, size_t buflen)
{
    cl_ulong    numeric_value = (cl_ulong)datum;
    bool        sign = PG_NUMERIC_SIGN(numeric_value);
    int         expo = (int)PG_NUMERIC_EXPONENT(numeric_value);
    cl_ulong    mantissa = PG_NUMERIC_MANTISSA(numeric_value);
    int n = snprintf(buf, buflen, "%c%llue%d", sign ? '-' : '+',
                     (unsigned long long)mantissa, expo);
    return (n > 0 && (size_t)n < buflen) ? StromError_Success
                                         : StromError_DataStoreNoSpace;
}

}   /* extern "C" */
