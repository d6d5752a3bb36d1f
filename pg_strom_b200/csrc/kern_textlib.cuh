/*
 * kern_textlib.cuh - bpchar / text comparison of the device runtime (the
 * counterpart of the reference's opencl_textlib.h; catalogue entries
 * codegen.c:611-629).
 *
 * A text / bpchar value on the device is a pointer to the varlena image
 * (1-byte or 4-byte header) inside the chunk or the kern_parambuf - the
 * staged / de-formed value of a varlena column is the offset of the datum
 * from the head of the chunk, as for numeric (kern_numeric.cuh).  Compressed
 * and out-of-line (TOAST pointer) datums cannot be read here: the value reads
 * as NULL and the row is flagged StromError_CpuReCheck.
 *
 * Comparison is bytewise on UNSIGNED bytes, i.e. PostgreSQL's result under
 * the "C" collation (varstr_cmp: memcmp, then the shorter string first);
 * equality is the same under every deterministic collation.  The planner
 * half only offloads the ordering operators when the input collation is C
 * (codegen.cpp).  bpchar ignores trailing blanks (bcTruelen).  The reference
 * compares signed cl_char, which orders bytes >= 0x80 (every multi-byte UTF-8
 * sequence) before ASCII - PostgreSQL does not, so that is not reproduced.
 */
#ifndef KERN_TEXTLIB_CUH
#define KERN_TEXTLIB_CUH
#include "kern_shared.h"

typedef struct {
    const unsigned char *value;     /* varlena header */
    bool        isnull;
} pg_varlena_t;
typedef pg_varlena_t    pg_text_t;
typedef pg_varlena_t    pg_bpchar_t;

/* VARDATA_ANY / VARSIZE_ANY_EXHDR of a little-endian varlena
 * (opencl_common.h:459-475); false for compressed / external datums */
DEVFN bool
pgs_varlena_payload(const unsigned char *p, const unsigned char **data, cl_int *len)
{
    cl_uint     b0 = p[0];

    if (b0 & 0x01)
    {
        if (b0 == 0x01)
            return false;           /* VARATT_IS_1B_E: TOAST pointer */
        *len = (cl_int)(b0 >> 1) - 1;
        *data = p + 1;
        return (*len >= 0);
    }
    if (b0 & 0x02)
        return false;               /* VARATT_IS_4B_C: compressed in-line */
    cl_uint     hdr = b0 | ((cl_uint)p[1] << 8) | ((cl_uint)p[2] << 16) | ((cl_uint)p[3] << 24);
    *len = (cl_int)(hdr >> 2) - 4;
    *data = p + 4;
    return (*len >= 0);
}

DEVFN pg_varlena_t
pgs_varlena_make(cl_int *errcode, const unsigned char *p)
{
    pg_varlena_t    r;
    const unsigned char *data;
    cl_int      len;

    r.value = p;
    r.isnull = false;
    if (!pgs_varlena_payload(p, &data, &len))
    {
        r.value = NULL;
        r.isnull = true;
        STROM_SET_ERROR(errcode, StromError_CpuReCheck);
    }
    return r;
}

template <typename KDS>
DEVFN pg_varlena_t
pg_varlena_vref(const KDS &kds, const void *ktoast, cl_int *errcode,
                cl_uint colidx, cl_uint rowidx)
{
    cl_uint     offset = 0;
    bool        ok = kds.template fetch<cl_uint>(GPUPREAGG_INCOL_SLOT(colidx), rowidx, offset);

    if (!ok || offset == 0)
    {
        pg_varlena_t r;
        r.value = NULL;
        r.isnull = true;
        return r;
    }
    return pgs_varlena_make(errcode, (const unsigned char *)ktoast + offset);
}

DEVFN pg_varlena_t
pg_varlena_param(const kern_parambuf *kparams, cl_int *errcode, cl_uint param_id)
{
    if (param_id < kparams->nparams && kparams->poffset[param_id] > 0)
        return pgs_varlena_make(errcode, (const unsigned char *)kparams +
                                kparams->poffset[param_id]);
    pg_varlena_t r;
    r.value = NULL;
    r.isnull = true;
    return r;
}

DEVFN pg_varlena_t
pg_varlena_null(void)
{
    pg_varlena_t r;
    r.value = NULL;
    r.isnull = true;
    return r;
}

#define pg_text_vref        pg_varlena_vref
#define pg_bpchar_vref      pg_varlena_vref
#define pg_text_param       pg_varlena_param
#define pg_bpchar_param     pg_varlena_param
#define pg_text_null        pg_varlena_null
#define pg_bpchar_null      pg_varlena_null

#define PGS_VARLENA_NULLTEST_TEMPLATE(NAME)                         \
    DEVFN pg_bool_t                                                 \
    pgfn_##NAME##_isnull(cl_int *errcode, pg_varlena_t arg)         \
    {                                                               \
        pg_bool_t result;                                           \
        result.isnull = false;                                      \
        result.value = arg.isnull;                                  \
        return result;                                              \
    }                                                               \
    DEVFN pg_bool_t                                                 \
    pgfn_##NAME##_isnotnull(cl_int *errcode, pg_varlena_t arg)      \
    {                                                               \
        pg_bool_t result;                                           \
        result.isnull = false;                                      \
        result.value = !arg.isnull;                                 \
        return result;                                              \
    }
PGS_VARLENA_NULLTEST_TEMPLATE(text)
PGS_VARLENA_NULLTEST_TEMPLATE(bpchar)

/* memcmp order, the shorter string first on a common prefix */
DEVFN cl_int
pgs_bytes_compare(const unsigned char *s1, cl_int len1,
                  const unsigned char *s2, cl_int len2)
{
    cl_int      len = (len1 < len2 ? len1 : len2);

    for (cl_int i = 0; i < len; i++)
    {
        cl_uint     c1 = s1[i], c2 = s2[i];

        if (c1 != c2)
            return (c1 < c2 ? -1 : 1);
    }
    return (len1 == len2 ? 0 : (len1 < len2 ? -1 : 1));
}

DEVFN cl_int
pgs_text_compare(pg_varlena_t arg1, pg_varlena_t arg2, bool ignore_trailing_blanks)
{
    const unsigned char *s1, *s2;
    cl_int      len1 = 0, len2 = 0;

    /* both were checked by pgs_varlena_make */
    pgs_varlena_payload(arg1.value, &s1, &len1);
    pgs_varlena_payload(arg2.value, &s2, &len2);
    if (ignore_trailing_blanks)
    {
        while (len1 > 0 && s1[len1 - 1] == ' ')
            len1--;
        while (len2 > 0 && s2[len2 - 1] == ' ')
            len2--;
    }
    return pgs_bytes_compare(s1, len1, s2, len2);
}

#define PGS_TEXT_COMPARE_TEMPLATE(FNAME,TYPE,BLANKS,OPER)                   \
    DEVFN pg_bool_t                                                         \
    pgfn_##FNAME(cl_int *errcode, pg_##TYPE##_t arg1, pg_##TYPE##_t arg2)   \
    {                                                                       \
        pg_bool_t   result;                                                 \
                                                                            \
        result.isnull = (arg1.isnull | arg2.isnull);                        \
        result.value = (cl_bool)(!result.isnull &&                          \
                                 pgs_text_compare(arg1, arg2, BLANKS) OPER 0); \
        return result;                                                      \
    }
PGS_TEXT_COMPARE_TEMPLATE(bpchareq, bpchar, true, ==)
PGS_TEXT_COMPARE_TEMPLATE(bpcharne, bpchar, true, !=)
PGS_TEXT_COMPARE_TEMPLATE(bpcharlt, bpchar, true, <)
PGS_TEXT_COMPARE_TEMPLATE(bpcharle, bpchar, true, <=)
PGS_TEXT_COMPARE_TEMPLATE(bpchargt, bpchar, true, >)
PGS_TEXT_COMPARE_TEMPLATE(bpcharge, bpchar, true, >=)
PGS_TEXT_COMPARE_TEMPLATE(texteq, text, false, ==)
PGS_TEXT_COMPARE_TEMPLATE(textne, text, false, !=)
PGS_TEXT_COMPARE_TEMPLATE(text_lt, text, false, <)
PGS_TEXT_COMPARE_TEMPLATE(text_le, text, false, <=)
PGS_TEXT_COMPARE_TEMPLATE(text_gt, text, false, >)
PGS_TEXT_COMPARE_TEMPLATE(text_ge, text, false, >=)

DEVFN pg_int4_t
pgfn_bpcharcmp(cl_int *errcode, pg_bpchar_t arg1, pg_bpchar_t arg2)
{
    pg_int4_t   result;

    result.isnull = (arg1.isnull | arg2.isnull);
    result.value = (result.isnull ? 0 : pgs_text_compare(arg1, arg2, true));
    return result;
}

DEVFN pg_int4_t
pgfn_text_cmp(cl_int *errcode, pg_text_t arg1, pg_text_t arg2)
{
    pg_int4_t   result;

    result.isnull = (arg1.isnull | arg2.isnull);
    result.value = (result.isnull ? 0 : pgs_text_compare(arg1, arg2, false));
    return result;
}

/* ------------------------------------------------------------------
 * text / bpchar as a GROUP BY key: "kernel text".
 *
 * A hash-table slot keeps 8 bytes per key.  A string of at most 7 bytes is
 * its own key word: payload byte i in bits 8i..8i+7, the length in the top
 * byte - equal words <=> equal strings, so the tables, the partitions and
 * the NCCL merge treat it like an int8 key.  The word is also what the
 * result row carries (by value, like the 64-bit device numeric); the host
 * turns it back into a varlena with pgstrom_fixup_kernel_text()
 * (datastore.cpp), the counterpart of the reference's varlena fix-up of
 * grouping keys (opencl_gpupreagg.h:326-366, pg_fixup_tupslot_varlena).
 * bpchar compares without its trailing blanks, so those are not part of the
 * word; the host pads the value back to the column's typmod.
 *
 * A longer key goes to the session's KEY HEAP (below): the string is stored
 * once, its key word is the heap offset under the top byte 0x80 - again
 * equal words <=> equal strings, within one session.  The reference keeps a
 * varlena key as a pointer into the chunk's toast area and compares the
 * strings in its generated keycomp (opencl_gpupreagg.h:326-366,
 * gpupreagg.c:1234-1243); here the comparison happens once per row, when
 * the string is looked up, and everything behind it works on 8-byte words.
 * A compressed / external datum, a full heap or a session without a heap is
 * a row for the host: CpuReCheck.
 * ------------------------------------------------------------------ */
#define PGS_KERNEL_TEXT_MAXLEN  7
#define PGS_KEYHEAP_TAG         0x80ULL     /* top byte of a key-heap word */

/*
 * The key heap of a session: an open-addressing table of (hash, ref) pairs
 * over an append-only string heap, both in HBM.  `ref` is 1 + the byte
 * offset of the entry [cl_ulong length | bytes, zero-padded to 8]; 0 =
 * claimed but not yet published; PGS_KEYHEAP_NOROOM = the claimer found the
 * heap full.  The control block is a module global that the CUDA layer fills
 * in when it opens the session (a session loads its own instance of the
 * program); a program that is never given a heap (nslots = 0) re-checks long
 * keys.
 */
/* (pgs_keyheap_ctl: kern_shared.h) */

#define PGS_KEYHEAP_NOROOM      0xFFFFFFFFFFFFFFFFULL
#define PGS_KEYHEAP_MAX_SPINS   20000
/* longer keys are left to the host: one thread hashes and compares the whole
 * string (and a corrupt varlena header must not send it through gigabytes) */
#define PGS_KEYHEAP_MAXLEN      65535

#ifdef __CUDACC__
extern "C" { __device__ pgs_keyheap_ctl pgs_keyheap; }
#define PGS_KEYHEAP_FENCE()         __threadfence()
/* slots and entries are written by other SMs while this one reads their
 * neighbours: every look goes to L2 */
#define PGS_KEYHEAP_LOAD64(p)       __ldcg((const unsigned long long *)(p))
#define PGS_KEYHEAP_PAUSE()         __nanosleep(100)
#else
static pgs_keyheap_ctl pgs_keyheap;     /* CPU build of the tests */
#ifdef PGS_KEYHEAP_HOST_ATOMICS         /* ... with real threads */
#define PGS_KEYHEAP_FENCE()         __atomic_thread_fence(__ATOMIC_SEQ_CST)
#define PGS_KEYHEAP_LOAD64(p)       __atomic_load_n((const unsigned long long *)(p), __ATOMIC_ACQUIRE)
#define PGS_KEYHEAP_PAUSE()         sched_yield()
#else
#define PGS_KEYHEAP_FENCE()         ((void)0)
#define PGS_KEYHEAP_LOAD64(p)       (*(const volatile unsigned long long *)(p))
#define PGS_KEYHEAP_PAUSE()         ((void)0)
#endif
#endif

DEVFN cl_ulong
pgs_keyheap_hash(const unsigned char *data, cl_int len)
{
    cl_ulong    h = 0xCBF29CE484222325ULL ^ (cl_ulong)(cl_uint)len;

    for (cl_int i = 0; i < len; i++)
        h = (h ^ data[i]) * 0x100000001B3ULL;       /* FNV-1a */
    h ^= h >> 29;
    h *= 0xBF58476D1CE4E5B9ULL;
    h ^= h >> 32;
    return h ? h : 1;
}

/* bytes i .. i+7 of the string as one little-endian word, zero past its end */
DEVFN cl_ulong
pgs_keyheap_word(const unsigned char *data, cl_int len, cl_int i)
{
    cl_ulong    w = 0;
    cl_int      n = (len - i < 8 ? len - i : 8);

    for (cl_int k = 0; k < n; k++)
        w |= (cl_ulong)data[i + k] << (8 * k);
    return w;
}

/* -> key word of the string, or 0 with *ok = false (no heap / heap full /
 * probe limit: the row is for the host) */
DEVFN cl_ulong
pgs_keyheap_intern(const unsigned char *data, cl_int len, bool *ok)
{
    const cl_uint   nslots = pgs_keyheap.nslots;
    cl_ulong       *slots = pgs_keyheap.slots;
    unsigned char  *heap = pgs_keyheap.heap;
    cl_ulong        h;
    cl_uint         pos, probes = 0, spins = 0;

    *ok = false;
    if (nslots == 0 || len > PGS_KEYHEAP_MAXLEN)
        return 0;
    h = pgs_keyheap_hash(data, len);
    pos = (cl_uint)(h >> 17) & (nslots - 1);
    /* one flat loop: a lane that finds a claimed but unpublished slot comes
     * round again instead of spinning in a nested loop, so that the claimer
     * - possibly a lane of the same warp - gets to publish.  The wait is
     * bounded: a lane that has waited for ~2 ms gives its row to the host */
    for (;;)
    {
        cl_ulong   *slot = slots + 2 * (size_t)pos;
        cl_ulong    cur = PGS_KEYHEAP_LOAD64(slot);

        if (cur == 0)
        {
            cur = atomicCAS((unsigned long long *)slot, 0ULL, (unsigned long long)h);
            if (cur == 0)
            {
                /* ours: store the string, then publish it */
                cl_ulong    need = 8 + (((cl_ulong)len + 7) & ~7ULL);
                cl_ulong    off = atomicAdd((unsigned long long *)pgs_keyheap.heap_used,
                                            (unsigned long long)need);
                if (off + need > pgs_keyheap.heap_bytes)
                {
                    atomicExch((unsigned long long *)(slot + 1), PGS_KEYHEAP_NOROOM);
                    return 0;
                }
                cl_ulong   *ent = (cl_ulong *)(heap + off);
                ent[0] = (cl_ulong)(cl_uint)len;
                for (cl_int i = 0; i < len; i += 8)
                    ent[1 + (i >> 3)] = pgs_keyheap_word(data, len, i);
                PGS_KEYHEAP_FENCE();
                atomicExch((unsigned long long *)(slot + 1), (unsigned long long)(off + 1));
                *ok = true;
                return (PGS_KEYHEAP_TAG << 56) | off;
            }
        }
        if (cur == h)
        {
            cl_ulong    ref = PGS_KEYHEAP_LOAD64(slot + 1);

            if (ref == 0)
            {
                /* not published yet: look again */
                if (++spins > PGS_KEYHEAP_MAX_SPINS)
                    return 0;
                PGS_KEYHEAP_PAUSE();
                continue;
            }
            if (ref == PGS_KEYHEAP_NOROOM)
                return 0;
            PGS_KEYHEAP_FENCE();
            const cl_ulong *ent = (const cl_ulong *)(heap + (ref - 1));
            if (PGS_KEYHEAP_LOAD64(ent) == (cl_ulong)(cl_uint)len)
            {
                cl_int  i = 0;
                while (i < len &&
                       PGS_KEYHEAP_LOAD64(ent + 1 + (i >> 3)) == pgs_keyheap_word(data, len, i))
                    i += 8;
                if (i >= len)
                {
                    *ok = true;
                    return (PGS_KEYHEAP_TAG << 56) | (ref - 1);
                }
            }
            /* same hash, another string: next slot */
        }
        if (++probes > pgs_keyheap.max_probe)
            return 0;
        pos = (pos + 1) & (nslots - 1);
    }
}

DEVFN cl_ulong
pgs_text_keybits(cl_int *errcode, pg_varlena_t arg, bool ignore_trailing_blanks,
                 bool *isnull)
{
    const unsigned char *data;
    cl_int      len = 0;
    cl_ulong    word = 0;

    *isnull = arg.isnull;
    if (arg.isnull)
        return 0;
    pgs_varlena_payload(arg.value, &data, &len);    /* checked by pgs_varlena_make */
    if (ignore_trailing_blanks)
        while (len > 0 && data[len - 1] == ' ')
            len--;
    if (len > PGS_KERNEL_TEXT_MAXLEN)
    {
        bool    ok;

        word = pgs_keyheap_intern(data, len, &ok);
        if (!ok)
        {
            *isnull = true;
            STROM_SET_ERROR(errcode, StromError_CpuReCheck);
            return 0;
        }
        return word;
    }
    for (cl_int i = 0; i < len; i++)
        word |= (cl_ulong)data[i] << (8 * i);
    return word | ((cl_ulong)len << 56);
}

#endif  /* KERN_TEXTLIB_CUH */
