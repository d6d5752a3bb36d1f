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

/* ------------------------------------------------------------------
 * KDS_FORMAT_ROW / KDS_FORMAT_ROW_FLAT: the reference's own input formats.
 *
 * pgstrom_kds_row_*   : pgstrom_create_data_store_row (datastore.c:382-435)
 *                       and pgstrom_data_store_insert_block (:556-710).  The
 *                       visibility check belongs to PostgreSQL: the caller
 *                       passes the line pointers it found visible.  Pages
 *                       are referenced, not copied (bitem->page), exactly
 *                       like shared buffers in the reference; the CUDA layer
 *                       gathers them at DMA time.
 * pgstrom_kds_flat_*  : pgstrom_create_data_store_row_flat (:437-470) and
 *                       the ROW_FLAT branch of pgstrom_data_store_insert_tuple
 *                       (:799-823): tuples are packed from the tail.
 * ------------------------------------------------------------------ */
size_t
pgstrom_kds_row_length(int ncols, uint32_t maxblocks, uint32_t nrooms)
{
    return STROMALIGN(KERN_DATA_STORE_HEAD_LENGTH(ncols) +
                      STROMALIGN(sizeof(kern_blkitem) * (size_t)maxblocks) +
                      STROMALIGN(sizeof(kern_rowitem) * (size_t)nrooms));
}

int
pgstrom_kds_row_init(void *buffer, size_t buflen, int ncols,
                     const kern_colmeta *colmeta, uint32_t maxblocks, uint32_t nrooms)
{
    size_t need = pgstrom_kds_row_length(ncols, maxblocks, nrooms);
    kern_data_store *kds = (kern_data_store *)buffer;

    if (need > buflen || (size_t)BLCKSZ * maxblocks + need > 0xffffffffULL)
    {
        pgs::last_error = "row store does not fit the buffer";
        return StromError_DataStoreNoSpace;
    }
    if (maxblocks > 65536)
    {
        /* kern_rowitem.blk_index has 16 bits (opencl_common.h:395-401) */
        pgs::last_error = "a KDS_FORMAT_ROW chunk addresses at most 65536 pages";
        return StromError_DataStoreNoSpace;
    }
    init_kern_data_store(kds, ncols, colmeta, need, nrooms, KDS_FORMAT_ROW);
    kds->maxblocks = maxblocks;
    return StromError_Success;
}

/* returns the number of rows added, or -1 if the block has to go to the
 * next store (datastore.c:604-613) */
int
pgstrom_kds_row_insert_block(kern_data_store *kds, const void *page,
                             const uint16_t *visible_offsets, int nvisible)
{
    const unsigned char *pg = (const unsigned char *)page;
    uint16_t    pd_lower;
    int         lines;

    if (kds->format != KDS_FORMAT_ROW || kds->nblocks >= kds->maxblocks)
        return -1;
    memcpy(&pd_lower, pg + 12, 2);
    lines = (pd_lower <= 24 ? 0 : (pd_lower - 24) / 4);
    if (nvisible > lines)
    {
        pgs::last_error = "more visible tuples than line pointers";
        return -1;
    }
    if ((size_t)kds->nitems + lines > kds->nrooms ||
        KERN_DATA_STORE_HEAD_LENGTH(kds->ncols) +
        STROMALIGN(sizeof(kern_blkitem) * (size_t)kds->maxblocks) +
        STROMALIGN(sizeof(kern_rowitem) * ((size_t)kds->nitems + lines)) +
        (size_t)BLCKSZ * kds->nblocks >= (size_t)BLCKSZ * kds->maxblocks)
        return -1;
    kern_rowitem *ritem = KERN_DATA_STORE_ROWITEM(kds, kds->nitems);
    kern_blkitem *bitem = KERN_DATA_STORE_BLKITEM(kds, kds->nblocks);
    for (int i = 0; i < nvisible; i++)
    {
        ritem->blk_index = (cl_ushort)kds->nblocks;
        ritem->item_offset = visible_offsets[i];
        ritem++;
    }
    kds->nitems += (cl_uint)nvisible;
    bitem->buffer = (cl_int)(kds->nblocks + 1);
    bitem->page = (hostptr_t)(uintptr_t)page;
    kds->nblocks++;
    return nvisible;
}

int
pgstrom_kds_flat_init(void *buffer, size_t buflen, int ncols,
                      const kern_colmeta *colmeta, uint32_t nrooms)
{
    if (buflen > 0xffffffffULL || buflen < pgstrom_kds_row_length(ncols, 0, nrooms))
    {
        pgs::last_error = "flat row store does not fit the buffer";
        return StromError_DataStoreNoSpace;
    }
    init_kern_data_store((kern_data_store *)buffer, ncols, colmeta,
                         buflen & ~(size_t)15, nrooms, KDS_FORMAT_ROW_FLAT);
    return StromError_Success;
}

/* htup: HeapTupleHeaderData + data, t_len bytes.  Returns 1, or 0 when the
 * store is full. */
int
pgstrom_kds_flat_insert_tuple(kern_data_store *kds, const void *htup, uint32_t t_len)
{
    if (kds->format != KDS_FORMAT_ROW_FLAT || kds->nitems >= kds->nrooms)
        return 0;
    kern_rowitem *ritem = KERN_DATA_STORE_ROWITEM(kds, kds->nitems);
    size_t usage = (size_t)((char *)(ritem + 1) - (char *)kds) +
                   kds->usage + LONGALIGN(t_len);
    if (usage > kds->length)
        return 0;
    char *dest = (char *)kds + kds->length - kds->usage - LONGALIGN(t_len);
    memcpy(dest, htup, t_len);
    ritem->htup_offset = (cl_uint)(dest - (char *)kds);
    kds->usage += (cl_uint)LONGALIGN(t_len);
    kds->nitems++;
    return 1;
}

/* attcacheoff of every column the way TupleDesc caches it: known until the
 * first variable-length attribute (heap_deform_tuple's fast path) */
void
pgstrom_colmeta_set_cacheoff(int ncols, kern_colmeta *colmeta)
{
    long off = 0;
    bool known = true;

    for (int i = 0; i < ncols; i++)
    {
        colmeta[i].attnum = (cl_short)(i + 1);
        if (known && colmeta[i].attlen > 0)
        {
            off = (long)TYPEALIGN(colmeta[i].attalign, off);
            colmeta[i].attcacheoff = (cl_short)off;
            off += colmeta[i].attlen;
        }
        else
        {
            /* the first varlena still has a known offset if it needs no padding */
            colmeta[i].attcacheoff = -1;
            known = false;
        }
    }
}

/* ------------------------------------------------------------------
 * Synthetic heap pages (benchmarks and tests; PostgreSQL itself is the real
 * producer).  Forms tuples like heap_form_tuple and adds them like
 * PageAddItem: 24-byte page header, line pointers growing up, MAXALIGNed
 * tuples growing down; 23-byte tuple header + NULL bitmap, t_hoff MAXALIGNed;
 * attributes aligned by attalign; varlena values whose payload is < 127
 * bytes get the 1-byte header.
 *   values[c]   : attlen > 0: nrows * attlen bytes; attlen < 0: nrows uint32
 *                 offsets into varlena_blob of 4-byte-header datums
 *   isnull[c]   : nrows bytes (1 = NULL) or NULL; values[c] == NULL means the
 *                 column is NULL in every row
 * Returns the number of pages written, -1 if maxpages is too small;
 * rows_per_page[p] = tuples on page p (line pointers 1..n, all normal).
 * ------------------------------------------------------------------ */
long
pgstrom_heap_form_pages(int ncols, const kern_colmeta *colmeta, uint32_t nrows,
                        const void *const *values, const unsigned char *const *isnull,
                        const unsigned char *varlena_blob,
                        unsigned char *pages, size_t maxpages, uint32_t *rows_per_page)
{
    long        npages = 0;
    unsigned char *page = NULL;
    uint32_t    lower = 0, upper = 0, nlines = 0;
    unsigned char tup[BLCKSZ];

    for (uint32_t r = 0; r < nrows; r++)
    {
        bool hasnull = false;
        for (int c = 0; c < ncols; c++)
            if (!values[c] || (isnull && isnull[c] && isnull[c][r]))
                hasnull = true;
        uint32_t hoff = (uint32_t)MAXALIGN(23 + (hasnull ? (ncols + 7) / 8 : 0));
        uint32_t off = hoff;
        memset(tup, 0, hoff);
        uint16_t infomask = (uint16_t)((hasnull ? 0x0001 : 0) | 0x0100 | 0x0800);
        uint16_t infomask2 = (uint16_t)ncols;
        for (int c = 0; c < ncols; c++)
        {
            bool isn = (!values[c] || (isnull && isnull[c] && isnull[c][r]));
            if (isn)
                continue;
            if (hasnull)
                tup[23 + (c >> 3)] |= (unsigned char)(1 << (c & 7));
            int attlen = colmeta[c].attlen;
            if (attlen > 0)
            {
                uint32_t a = (uint32_t)TYPEALIGN(colmeta[c].attalign, off);
                if (a + attlen > sizeof(tup)) return -1;
                memset(tup + off, 0, a - off);
                memcpy(tup + a, (const char *)values[c] + (size_t)r * attlen, attlen);
                off = a + attlen;
            }
            else
            {
                const unsigned char *d = varlena_blob + ((const uint32_t *)values[c])[r];
                size_t vl = varsize_any(d);
                infomask |= 0x0002;
                if (!(d[0] & 0x01) && vl - 4 + 1 <= 127)
                {
                    /* VARATT_CAN_MAKE_SHORT: 1-byte header, no alignment */
                    size_t sl = vl - 4 + 1;
                    if (off + sl > sizeof(tup)) return -1;
                    tup[off] = (unsigned char)((sl << 1) | 0x01);
                    memcpy(tup + off + 1, d + 4, vl - 4);
                    off += (uint32_t)sl;
                }
                else
                {
                    uint32_t a = (d[0] & 0x01) ? off : (uint32_t)TYPEALIGN(colmeta[c].attalign, off);
                    if (a + vl > sizeof(tup)) return -1;
                    memset(tup + off, 0, a - off);
                    memcpy(tup + a, d, vl);
                    off = a + (uint32_t)vl;
                }
            }
        }
        memcpy(tup + 18, &infomask2, 2);
        memcpy(tup + 20, &infomask, 2);
        tup[22] = (unsigned char)hoff;
        uint32_t need = (uint32_t)MAXALIGN(off);
        if (!page || lower + 4 + need > upper)
        {
            if (page)
            {
                uint16_t v = (uint16_t)lower; memcpy(page + 12, &v, 2);
                v = (uint16_t)upper; memcpy(page + 14, &v, 2);
                rows_per_page[npages - 1] = nlines;
            }
            if ((size_t)npages >= maxpages)
                return -1;
            page = pages + (size_t)BLCKSZ * npages++;
            memset(page, 0, BLCKSZ);
            uint16_t v = BLCKSZ; memcpy(page + 16, &v, 2);          /* pd_special */
            v = (uint16_t)(BLCKSZ | 4); memcpy(page + 18, &v, 2);   /* size | version */
            lower = 24; upper = BLCKSZ; nlines = 0;
        }
        upper -= need;
        memcpy(page + upper, tup, off);
        uint32_t lp = (upper & 0x7fff) | (1U << 15) | ((off & 0x7fff) << 17);   /* LP_NORMAL */
        memcpy(page + lower, &lp, 4);
        /* t_ctid = (block, line) */
        {
            uint16_t bi_hi = (uint16_t)((npages - 1) >> 16), bi_lo = (uint16_t)(npages - 1);
            uint16_t posid = (uint16_t)(nlines + 1);
            memcpy(page + upper + 12, &bi_hi, 2);
            memcpy(page + upper + 14, &bi_lo, 2);
            memcpy(page + upper + 16, &posid, 2);
        }
        lower += 4;
        nlines++;
    }
    if (page)
    {
        uint16_t v = (uint16_t)lower; memcpy(page + 12, &v, 2);
        v = (uint16_t)upper; memcpy(page + 14, &v, 2);
        rows_per_page[npages - 1] = nlines;
    }
    return npages;
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

/*
 * "kernel text" (kern_textlib.cuh) -> varlena: a text / bpchar grouping key
 * of at most 7 bytes comes back from the device by value - payload byte i in
 * bits 8i..8i+7, the length in the top byte.  bpchar keys were stripped of
 * their trailing blanks (they compare without them); `typmod` (atttypmod of
 * the column: VARHDRSZ + n for character(n), -1 if unknown / text) pads the
 * value back to n characters as PostgreSQL stores it.  Writes a varlena with
 * a 4-byte header into buf and returns its size, 0 when buf is too small.
 * Host-side fix-up like pgstrom_fixup_kernel_numeric (datastore.c:150-167);
 * the reference fixes key pointers up on the device
 * (opencl_gpupreagg.h:326-366).
 */
size_t
pgstrom_fixup_kernel_text(Datum datum, int typmod, void *buf, size_t buflen)
{
    return pgstrom_fixup_kernel_text_heap(datum, typmod, NULL, 0, buf, buflen);
}

/*
 * The same for any text / bpchar grouping key of a result row.  A key of more
 * than 7 bytes comes back as a word of the session's key heap (top byte 0x80,
 * below it the offset of [uint64 length | bytes] in the heap that
 * pgs_preagg_key_heap() returns after pgs_preagg_finish()).  Returns 0 for a
 * word that does not point at an entry inside the heap.
 */
size_t
pgstrom_fixup_kernel_text_heap(Datum datum, int typmod,
                               const void *key_heap, size_t key_heap_len,
                               void *buf, size_t buflen)
{
    cl_ulong    word = (cl_ulong)datum;
    size_t      len = (size_t)(word >> 56);
    unsigned char inline_payload[8];
    const unsigned char *payload = inline_payload;
    size_t      nchars = 0, pad = 0;

    if (len == 0x80)
    {
        size_t      off = (size_t)(word & 0x00FFFFFFFFFFFFFFULL);
        uint64_t    n;

        if (!key_heap || (off & 7) != 0 || off + 8 > key_heap_len)
            return 0;
        memcpy(&n, (const char *)key_heap + off, 8);
        if (n <= 7 || n > key_heap_len - off - 8)
            return 0;
        len = (size_t)n;
        payload = (const unsigned char *)key_heap + off + 8;
    }
    else if (len > 7)
        return 0;
    else
    {
        for (size_t i = 0; i < len; i++)
            inline_payload[i] = (unsigned char)(word >> (8 * i));
    }
    for (size_t i = 0; i < len; i++)
    {
        if ((payload[i] & 0xC0) != 0x80)    /* not a UTF-8 continuation byte */
            nchars++;
    }
    if (typmod >= 4 && (size_t)(typmod - 4) > nchars)
        pad = (size_t)(typmod - 4) - nchars;
    size_t      total = 4 + len + pad;
    if (total > buflen)
        return 0;
    uint32_t    hdr = (uint32_t)(total << 2);       /* SET_VARSIZE */
    memcpy(buf, &hdr, 4);
    memcpy((char *)buf + 4, payload, len);
    memset((char *)buf + 4 + len, ' ', pad);
    return total;
}

}   /* extern "C" */
