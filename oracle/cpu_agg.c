/*
 * cpu_agg.c - CPU restatement of PostgreSQL's Agg-over-SeqScan for the bench
 * queries (the "pg_strom.enabled = off" plan: HashAggregate / Aggregate over
 * Seq Scan, /root/reference/expected/explain_agg.out first plan shape).
 *
 * TEST INFRASTRUCTURE (oracle) - built into oracle/_build/libcpu_agg.so; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  It is the checker and the reported CPU
 * baseline, never the product path.
 *
 * What is restated (PostgreSQL core is not vendored in the reference tree):
 * the executor's per-tuple loop - ExecQual on the scan qual, grouping hash
 * lookup (TupleHashTable), then advance_aggregates calling each strict
 * transition function only for non-NULL input:
 *   count(*)            int8inc
 *   count(x)            int8inc_any
 *   sum(int4)           int4_sum   (int8 state)
 *   avg(int4)           int4_avg_accum {count, sum}
 *   avg(int8)           int8_avg_accum (numeric state; here a 128-bit sum,
 *                       which is exact for the same inputs)
 *   min/max             int4smaller/larger, float8smaller/larger
 *   sum/avg(float8)     float8pl / float8_accum {N, sumX, sumX2}
 *   variance(float8)    float8_accum {N, sumX, sumX2}
 * It is "port" not "reference": the expression interpreter and the fmgr call
 * overhead of a real PostgreSQL are absent, so a real PostgreSQL backend is
 * slower than this on the same core.  cpu_agg_run() scans column arrays;
 * cpu_agg_run_heap() scans heap pages and de-forms every tuple like
 * slot_deform_tuple (the baseline of the heap-page workload).
 * The parallel variant cuts the rows into per-thread ranges, aggregates each
 * with its own hash table and combines the states (what parallel aggregation
 * of later PostgreSQL versions does; the reference-era 9.4 has none, so
 * 1 thread is the faithful figure).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { AGG_COUNT_STAR = 0, AGG_COUNT, AGG_SUM_INT4, AGG_AVG_INT4, AGG_AVG_INT8,
       AGG_MIN_INT4, AGG_MAX_INT4, AGG_SUM_FLOAT8, AGG_AVG_FLOAT8,
       AGG_MIN_FLOAT8, AGG_MAX_FLOAT8, AGG_VAR_FLOAT8 };
enum { COL_INT4 = 4, COL_INT8 = 8, COL_FLOAT8 = 108 };

#define MAX_COLS    8
#define MAX_AGGS    16

typedef struct {
    int32_t     ncols;
    int32_t     coltype[MAX_COLS];
    const void *values[MAX_COLS];
    const uint8_t *nulls[MAX_COLS];     /* one byte per row, or NULL */
    int64_t     nrows;
} cpu_table;

typedef struct {
    int32_t     qual_col;       /* WHERE col < qual_const ; -1 = no qual */
    int32_t     qual_const;
    int32_t     key_col;        /* GROUP BY col ; -1 = none */
    int32_t     naggs;
    int32_t     agg_kind[MAX_AGGS];
    int32_t     agg_col[MAX_AGGS];
    int32_t     partitioned;    /* GROUP BY with very many groups: see cpu_agg_run */
    int32_t     pad;
} cpu_query;

/* transition state of one aggregate of one group */
typedef struct {
    int64_t     n;
    int64_t     isum_lo;        /* 128-bit sum */
    int64_t     isum_hi;
    double      fsum;
    double      fsum2;
    int64_t     imin, imax;
    double      fmin, fmax;
    int32_t     has_value;
    int32_t     pad;
} agg_state;

typedef struct {
    int64_t     key;
    int64_t     index;          /* group number + 1; 0 = empty slot */
} group_entry;

typedef struct {
    group_entry *slots;
    uint64_t    nslots;         /* power of two */
    uint64_t    nused;
    int32_t     naggs;
    agg_state  *pool;           /* [ngroups][naggs] transition states */
    int64_t    *keys;           /* [ngroups] */
    uint64_t    pool_cap;       /* in groups */
} group_table;

static inline uint64_t
mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

static void
state_init(agg_state *s)
{
    memset(s, 0, sizeof(*s));
}

static void
table_init(group_table *t, uint64_t nslots, int naggs)
{
    t->nslots = nslots;
    t->nused = 0;
    t->naggs = naggs;
    t->slots = (group_entry *)calloc(nslots, sizeof(group_entry));
    t->pool_cap = nslots;
    t->pool = (agg_state *)malloc(sizeof(agg_state) * naggs * t->pool_cap);
    t->keys = (int64_t *)malloc(sizeof(int64_t) * t->pool_cap);
}

static void
table_free(group_table *t)
{
    free(t->slots);
    free(t->pool);
    free(t->keys);
}

static void
table_grow(group_table *t)
{
    uint64_t nn = t->nslots * 2;
    group_entry *ns = (group_entry *)calloc(nn, sizeof(group_entry));
    for (uint64_t i = 0; i < t->nslots; i++)
        if (t->slots[i].index)
        {
            uint64_t h = mix64((uint64_t)t->slots[i].key) & (nn - 1);
            while (ns[h].index)
                h = (h + 1) & (nn - 1);
            ns[h] = t->slots[i];
        }
    free(t->slots);
    t->slots = ns;
    t->nslots = nn;
}

/* returns the transition states of the group, creating it on first sight */
static agg_state *
table_lookup(group_table *t, int64_t key)
{
    for (;;)
    {
        uint64_t mask = t->nslots - 1;
        uint64_t h = mix64((uint64_t)key) & mask;
        for (;;)
        {
            group_entry *e = &t->slots[h];
            if (!e->index)
                break;
            if (e->key == key)
                return t->pool + (e->index - 1) * t->naggs;
            h = (h + 1) & mask;
        }
        if ((t->nused + 1) * 4 > t->nslots * 3)
        {
            table_grow(t);
            continue;
        }
        if (t->nused == t->pool_cap)
        {
            t->pool_cap *= 2;
            t->pool = (agg_state *)realloc(t->pool, sizeof(agg_state) * t->naggs * t->pool_cap);
            t->keys = (int64_t *)realloc(t->keys, sizeof(int64_t) * t->pool_cap);
        }
        t->slots[h].key = key;
        t->slots[h].index = (int64_t)(++t->nused);
        t->keys[t->nused - 1] = key;
        agg_state *st = t->pool + (t->nused - 1) * t->naggs;
        for (int j = 0; j < t->naggs; j++)
            state_init(&st[j]);
        return st;
    }
}

static inline void
add128(agg_state *s, int64_t v)
{
    uint64_t old = (uint64_t)s->isum_lo;
    uint64_t nw = old + (uint64_t)v;
    s->isum_lo = (int64_t)nw;
    s->isum_hi += (v < 0 ? -1 : 0) + (nw < old ? 1 : 0);
}

/* float8_cmp_internal: NaN sorts above everything */
static inline int
f8cmp(double a, double b)
{
    if (isnan(a)) return isnan(b) ? 0 : 1;
    if (isnan(b)) return -1;
    return (a > b) - (a < b);
}

static inline void
advance(agg_state *s, int kind, const cpu_table *t, int col, int64_t r)
{
    if (kind == AGG_COUNT_STAR)
    {
        s->n++;
        return;
    }
    if (t->nulls[col] && t->nulls[col][r])
        return;                         /* strict transition function */
    switch (kind)
    {
        case AGG_COUNT:
            s->n++;
            break;
        case AGG_SUM_INT4:
        case AGG_AVG_INT4:
            s->n++;
            s->isum_lo += ((const int32_t *)t->values[col])[r];
            s->has_value = 1;
            break;
        case AGG_AVG_INT8:
            s->n++;
            add128(s, ((const int64_t *)t->values[col])[r]);
            s->has_value = 1;
            break;
        case AGG_MIN_INT4:
        {
            int64_t v = ((const int32_t *)t->values[col])[r];
            if (!s->has_value || v < s->imin) s->imin = v;
            s->has_value = 1;
            break;
        }
        case AGG_MAX_INT4:
        {
            int64_t v = ((const int32_t *)t->values[col])[r];
            if (!s->has_value || v > s->imax) s->imax = v;
            s->has_value = 1;
            break;
        }
        case AGG_SUM_FLOAT8:
        case AGG_AVG_FLOAT8:
            s->n++;
            s->fsum += ((const double *)t->values[col])[r];
            s->has_value = 1;
            break;
        case AGG_VAR_FLOAT8:
        {
            double v = ((const double *)t->values[col])[r];
            s->n++;
            s->fsum += v;
            s->fsum2 += v * v;
            s->has_value = 1;
            break;
        }
        case AGG_MIN_FLOAT8:
        {
            double v = ((const double *)t->values[col])[r];
            if (!s->has_value || f8cmp(s->fmin, v) >= 0) s->fmin = v;
            s->has_value = 1;
            break;
        }
        case AGG_MAX_FLOAT8:
        {
            double v = ((const double *)t->values[col])[r];
            if (!s->has_value || f8cmp(s->fmax, v) <= 0) s->fmax = v;
            s->has_value = 1;
            break;
        }
    }
}

/* combine function (what a Gather + Finalize Aggregate does) */
static void
combine(agg_state *d, const agg_state *s, int kind)
{
    uint64_t old;
    d->n += s->n;
    old = (uint64_t)d->isum_lo;
    d->isum_lo = (int64_t)(old + (uint64_t)s->isum_lo);
    d->isum_hi += s->isum_hi + (((uint64_t)d->isum_lo) < old ? 1 : 0);
    d->fsum += s->fsum;
    d->fsum2 += s->fsum2;
    if (s->has_value)
    {
        if (kind == AGG_MIN_INT4 && (!d->has_value || s->imin < d->imin)) d->imin = s->imin;
        if (kind == AGG_MAX_INT4 && (!d->has_value || s->imax > d->imax)) d->imax = s->imax;
        if (kind == AGG_MIN_FLOAT8 && (!d->has_value || f8cmp(d->fmin, s->fmin) > 0)) d->fmin = s->fmin;
        if (kind == AGG_MAX_FLOAT8 && (!d->has_value || f8cmp(d->fmax, s->fmax) < 0)) d->fmax = s->fmax;
        d->has_value = 1;
    }
}

static void
scan_range(const cpu_table *t, const cpu_query *q, int64_t lo, int64_t hi,
           group_table *gt, agg_state *nogroup, int part, int nparts)
{
    for (int64_t r = lo; r < hi; r++)
    {
        agg_state *st;
        /* ExecQual: NULL or false => row is filtered */
        if (q->qual_col >= 0)
        {
            if (t->nulls[q->qual_col] && t->nulls[q->qual_col][r])
                continue;
            if (!(((const int32_t *)t->values[q->qual_col])[r] < q->qual_const))
                continue;
        }
        if (q->key_col >= 0)
        {
            int64_t key = (t->coltype[q->key_col] == COL_INT8)
                ? ((const int64_t *)t->values[q->key_col])[r]
                : (int64_t)((const int32_t *)t->values[q->key_col])[r];
            /* key-partitioned plan: this worker owns the groups whose key
             * hashes into its partition (high bits: the table uses the low) */
            if (nparts > 1 &&
                (int)(((mix64((uint64_t)key) >> 40) * (uint64_t)nparts) >> 24) != part)
                continue;
            st = table_lookup(gt, key);
        }
        else
            st = nogroup;
        for (int j = 0; j < q->naggs; j++)
            advance(&st[j], q->agg_kind[j], t, q->agg_col[j], r);
    }
}

typedef struct {
    const cpu_table *t;
    const cpu_query *q;
    int64_t     lo, hi;
    int         part, nparts;
    group_table gt;
    agg_state   nogroup[MAX_AGGS];
} worker_arg;

static void *
worker_main(void *p)
{
    worker_arg *w = (worker_arg *)p;
    scan_range(w->t, w->q, w->lo, w->hi, &w->gt, w->nogroup, w->part, w->nparts);
    return NULL;
}

/*
 * cpu_agg_run: returns the number of groups; results are written into
 * keys[max_groups] and states[max_groups * naggs] (group order is arbitrary).
 * If max_groups is too small the count is still returned.
 */
int64_t
cpu_agg_run(const cpu_table *t, const cpu_query *q, int nthreads,
            int64_t *keys, agg_state *states, int64_t max_groups)
{
    if (nthreads < 1)
        nthreads = 1;
    worker_arg *w = (worker_arg *)calloc(nthreads, sizeof(worker_arg));
    pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
    int64_t per = (t->nrows + nthreads - 1) / nthreads;
    int64_t ngroups = 0;

    /* GROUP BY with very many groups on several cores: per-worker tables over
     * row ranges would each hold nearly every group and the combine step - one
     * core re-inserting all of them - costs more than the scan (measured:
     * 16 cores slower than 1).  Instead every worker reads all keys and
     * aggregates the groups of its own hash partition: the tables are
     * disjoint, their union is the result, nothing is combined.  (A worker
     * per partition behind a repartitioning exchange is what a parallel
     * executor does with such a plan.) */
    const int partitioned = (q->partitioned && q->key_col >= 0 && nthreads > 1);

    for (int i = 0; i < nthreads; i++)
    {
        w[i].t = t;
        w[i].q = q;
        w[i].lo = per * i < t->nrows ? per * i : t->nrows;
        w[i].hi = per * (i + 1) < t->nrows ? per * (i + 1) : t->nrows;
        w[i].part = 0;
        w[i].nparts = 1;
        if (partitioned)
        {
            w[i].lo = 0;
            w[i].hi = t->nrows;
            w[i].part = i;
            w[i].nparts = nthreads;
        }
        if (q->key_col >= 0)
            table_init(&w[i].gt, 1024, q->naggs);
        for (int j = 0; j < q->naggs; j++)
            state_init(&w[i].nogroup[j]);
        if (nthreads > 1)
            pthread_create(&th[i], NULL, worker_main, &w[i]);
        else
            worker_main(&w[i]);
    }
    if (nthreads > 1)
        for (int i = 0; i < nthreads; i++)
            pthread_join(th[i], NULL);
    if (q->key_col < 0)
    {
        for (int i = 1; i < nthreads; i++)
            for (int j = 0; j < q->naggs; j++)
                combine(&w[0].nogroup[j], &w[i].nogroup[j], q->agg_kind[j]);
        ngroups = 1;
        if (max_groups >= 1)
        {
            keys[0] = 0;
            memcpy(states, w[0].nogroup, sizeof(agg_state) * q->naggs);
        }
    }
    else if (partitioned)
    {
        /* disjoint tables: the result is their concatenation */
        for (int i = 0; i < nthreads; i++)
        {
            int64_t room = max_groups - ngroups;
            int64_t ncopy = (int64_t)w[i].gt.nused < room ? (int64_t)w[i].gt.nused : (room > 0 ? room : 0);

            memcpy(keys + ngroups, w[i].gt.keys, sizeof(int64_t) * ncopy);
            memcpy(states + ngroups * q->naggs, w[i].gt.pool, sizeof(agg_state) * q->naggs * ncopy);
            ngroups += (int64_t)w[i].gt.nused;
            table_free(&w[i].gt);
        }
    }
    else
    {
        for (int i = 1; i < nthreads; i++)
        {
            for (uint64_t g = 0; g < w[i].gt.nused; g++)
            {
                agg_state *d = table_lookup(&w[0].gt, w[i].gt.keys[g]);
                for (int j = 0; j < q->naggs; j++)
                    combine(&d[j], &w[i].gt.pool[g * q->naggs + j], q->agg_kind[j]);
            }
            table_free(&w[i].gt);
        }
        ngroups = (int64_t)w[0].gt.nused;
        {
            int64_t ncopy = ngroups < max_groups ? ngroups : max_groups;
            memcpy(keys, w[0].gt.keys, sizeof(int64_t) * ncopy);
            memcpy(states, w[0].gt.pool, sizeof(agg_state) * q->naggs * ncopy);
        }
        table_free(&w[0].gt);
    }
    free(w);
    free(th);
    return ngroups;
}

/* ------------------------------------------------------------------
 * The same over HEAP PAGES: what a PostgreSQL backend actually scans
 * (SURVEY.md 8d asks for per-tuple deform in the CPU baseline).  The page and
 * tuple layout is PostgreSQL's (bufpage.h PageHeaderData / ItemIdData,
 * htup_details.h HeapTupleHeaderData; not vendored in the reference tree):
 *   page  : 24-byte header with pd_lower @12 / pd_upper @14, line pointers
 *           (lp_off:15 | lp_flags:2 | lp_len:15) from byte 24, tuples packed
 *           from the end of the page, each MAXALIGNed
 *   tuple : 23-byte header - t_infomask2 @18 (natts in the low 11 bits),
 *           t_infomask @20 (HEAP_HASNULL = 0x0001), t_hoff @22 - then the
 *           NULL bitmap when HEAP_HASNULL, then at MAXALIGN(t_hoff) the
 *           attributes, each aligned to its typalign, NULLs taking no room
 * The scan below is heapgettup_pagemode + slot_deform_tuple for the columns
 * a query needs, followed by the same ExecQual / hash lookup / advance as the
 * columnar loop above.
 * ------------------------------------------------------------------ */
#define CPU_BLCKSZ          8192
#define CPU_PAGE_HEADER     24
#define CPU_HTUP_HEADER     23
#define CPU_HEAP_HASNULL    0x0001
#define CPU_LP_NORMAL       1
#define CPU_MAXALIGN(x)     (((x) + 7) & ~(size_t)7)

static inline int
col_len(int coltype)
{
    return coltype == COL_INT4 ? 4 : 8;
}

/*
 * heap_form_tuple + PageAddItem for rows [row0, nrows) of a columnar table:
 * fills `pages` (maxpages * 8192 bytes, zeroed by the caller) and returns the
 * number of pages used; *rows_done = rows that found room.
 */
int64_t
cpu_heap_form_pages(const cpu_table *t, unsigned char *pages, int64_t maxpages,
                    int64_t row0, int64_t *rows_done)
{
    int64_t npages = 0;
    int64_t r = row0;

    while (r < t->nrows && npages < maxpages)
    {
        unsigned char  *page = pages + npages * CPU_BLCKSZ;
        uint16_t        lower = CPU_PAGE_HEADER, upper = CPU_BLCKSZ;

        for (; r < t->nrows; r++)
        {
            unsigned char   tup[CPU_HTUP_HEADER + 8 + 8 * MAX_COLS + 16];
            int             hasnull = 0;
            size_t          hoff, off, len;

            for (int c = 0; c < t->ncols; c++)
                if (t->nulls[c] && t->nulls[c][r])
                    hasnull = 1;
            hoff = CPU_MAXALIGN(CPU_HTUP_HEADER + (hasnull ? (size_t)(t->ncols + 7) / 8 : 0));
            memset(tup, 0, sizeof(tup));
            off = hoff;
            for (int c = 0; c < t->ncols; c++)
            {
                size_t  alen = (size_t)col_len(t->coltype[c]);

                if (t->nulls[c] && t->nulls[c][r])
                    continue;
                if (hasnull)
                    tup[CPU_HTUP_HEADER + (c >> 3)] |= (unsigned char)(1 << (c & 7));
                off = (off + alen - 1) & ~(alen - 1);
                memcpy(tup + off, (const char *)t->values[c] + (size_t)r * alen, alen);
                off += alen;
            }
            len = off;
            tup[18] = (unsigned char)(t->ncols & 0xff);
            tup[19] = (unsigned char)((t->ncols >> 8) & 0x07);
            tup[20] = (unsigned char)(hasnull ? CPU_HEAP_HASNULL : 0);
            tup[22] = (unsigned char)hoff;
            if ((size_t)lower + 4 + CPU_MAXALIGN(len) > (size_t)upper)
                break;                  /* page full */
            upper = (uint16_t)(upper - CPU_MAXALIGN(len));
            memcpy(page + upper, tup, len);
            {
                uint32_t lp = (uint32_t)upper | ((uint32_t)CPU_LP_NORMAL << 15) | ((uint32_t)len << 17);
                memcpy(page + lower, &lp, 4);
            }
            lower = (uint16_t)(lower + 4);
        }
        memcpy(page + 12, &lower, 2);
        memcpy(page + 14, &upper, 2);
        {
            uint16_t special = CPU_BLCKSZ, pagesize_version = CPU_BLCKSZ | 4;
            memcpy(page + 16, &special, 2);
            memcpy(page + 18, &pagesize_version, 2);
        }
        npages++;
    }
    if (rows_done)
        *rows_done = r - row0;
    return npages;
}

typedef struct {
    int32_t     ncols;
    int32_t     coltype[MAX_COLS];
    const unsigned char *pages;
    int64_t     npages;
} cpu_heap;

static void
scan_pages(const cpu_heap *h, const cpu_query *q, int64_t lo, int64_t hi,
           group_table *gt, agg_state *nogroup)
{
    /* the slot one tuple is de-formed into: a one-row "table" */
    cpu_table   slot;
    uint8_t     isnull[MAX_COLS];
    int         lastcol = 0;

    memset(&slot, 0, sizeof(slot));
    slot.ncols = h->ncols;
    slot.nrows = 1;
    for (int c = 0; c < h->ncols; c++)
    {
        slot.coltype[c] = h->coltype[c];
        slot.nulls[c] = &isnull[c];
    }
    /* slot_getsomeattrs: de-form up to the last attribute the query needs */
    if (q->qual_col >= lastcol) lastcol = q->qual_col + 1;
    if (q->key_col >= lastcol) lastcol = q->key_col + 1;
    for (int j = 0; j < q->naggs; j++)
        if (q->agg_kind[j] != AGG_COUNT_STAR && q->agg_col[j] >= lastcol)
            lastcol = q->agg_col[j] + 1;

    for (int64_t p = lo; p < hi; p++)
    {
        const unsigned char *page = h->pages + p * CPU_BLCKSZ;
        uint16_t    lower;
        int         nlines;

        memcpy(&lower, page + 12, 2);
        nlines = (lower <= CPU_PAGE_HEADER ? 0 : (lower - CPU_PAGE_HEADER) / 4);
        for (int i = 0; i < nlines; i++)
        {
            uint32_t    lp;
            const unsigned char *tup;
            int         hasnull, natts;
            size_t      off;
            agg_state  *st;

            memcpy(&lp, page + CPU_PAGE_HEADER + 4 * i, 4);
            if (((lp >> 15) & 3) != CPU_LP_NORMAL)
                continue;
            tup = page + (lp & 0x7fff);
            hasnull = (tup[20] & CPU_HEAP_HASNULL) != 0;
            natts = tup[18] | ((tup[19] & 0x07) << 8);
            off = tup[22];
            for (int c = 0; c < lastcol; c++)
            {
                size_t  alen = (size_t)col_len(h->coltype[c]);

                if (c >= natts || (hasnull && !((tup[CPU_HTUP_HEADER + (c >> 3)] >> (c & 7)) & 1)))
                {
                    isnull[c] = 1;
                    continue;
                }
                isnull[c] = 0;
                off = (off + alen - 1) & ~(alen - 1);
                slot.values[c] = tup + off;
                off += alen;
            }
            /* ExecQual: NULL or false => row is filtered */
            if (q->qual_col >= 0)
            {
                if (isnull[q->qual_col])
                    continue;
                if (!(*(const int32_t *)slot.values[q->qual_col] < q->qual_const))
                    continue;
            }
            if (q->key_col >= 0)
            {
                int64_t key = (h->coltype[q->key_col] == COL_INT8)
                    ? *(const int64_t *)slot.values[q->key_col]
                    : (int64_t)*(const int32_t *)slot.values[q->key_col];
                st = table_lookup(gt, key);
            }
            else
                st = nogroup;
            for (int j = 0; j < q->naggs; j++)
                advance(&st[j], q->agg_kind[j], &slot, q->agg_col[j], 0);
        }
    }
}

typedef struct {
    const cpu_heap  *h;
    const cpu_query *q;
    int64_t     lo, hi;
    group_table gt;
    agg_state   nogroup[MAX_AGGS];
} heap_worker_arg;

static void *
heap_worker_main(void *p)
{
    heap_worker_arg *w = (heap_worker_arg *)p;
    scan_pages(w->h, w->q, w->lo, w->hi, &w->gt, w->nogroup);
    return NULL;
}

/* cpu_agg_run over heap pages; same results and conventions */
int64_t
cpu_agg_run_heap(const cpu_heap *h, const cpu_query *q, int nthreads,
                 int64_t *keys, agg_state *states, int64_t max_groups)
{
    if (nthreads < 1)
        nthreads = 1;
    heap_worker_arg *w = (heap_worker_arg *)calloc(nthreads, sizeof(heap_worker_arg));
    pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
    int64_t per = (h->npages + nthreads - 1) / nthreads;
    int64_t ngroups = 0;

    for (int i = 0; i < nthreads; i++)
    {
        w[i].h = h;
        w[i].q = q;
        w[i].lo = per * i < h->npages ? per * i : h->npages;
        w[i].hi = per * (i + 1) < h->npages ? per * (i + 1) : h->npages;
        if (q->key_col >= 0)
            table_init(&w[i].gt, 1024, q->naggs);
        for (int j = 0; j < q->naggs; j++)
            state_init(&w[i].nogroup[j]);
        if (nthreads > 1)
            pthread_create(&th[i], NULL, heap_worker_main, &w[i]);
        else
            heap_worker_main(&w[i]);
    }
    if (nthreads > 1)
        for (int i = 0; i < nthreads; i++)
            pthread_join(th[i], NULL);
    if (q->key_col < 0)
    {
        for (int i = 1; i < nthreads; i++)
            for (int j = 0; j < q->naggs; j++)
                combine(&w[0].nogroup[j], &w[i].nogroup[j], q->agg_kind[j]);
        ngroups = 1;
        if (max_groups >= 1)
        {
            keys[0] = 0;
            memcpy(states, w[0].nogroup, sizeof(agg_state) * q->naggs);
        }
    }
    else
    {
        for (int i = 1; i < nthreads; i++)
        {
            for (uint64_t g = 0; g < w[i].gt.nused; g++)
            {
                agg_state *d = table_lookup(&w[0].gt, w[i].gt.keys[g]);
                for (int j = 0; j < q->naggs; j++)
                    combine(&d[j], &w[i].gt.pool[g * q->naggs + j], q->agg_kind[j]);
            }
            table_free(&w[i].gt);
        }
        ngroups = (int64_t)w[0].gt.nused;
        {
            int64_t ncopy = ngroups < max_groups ? ngroups : max_groups;
            memcpy(keys, w[0].gt.keys, sizeof(int64_t) * ncopy);
            memcpy(states, w[0].gt.pool, sizeof(agg_state) * q->naggs * ncopy);
        }
        table_free(&w[0].gt);
    }
    free(w);
    free(th);
    return ngroups;
}

int
cpu_agg_state_size(void)
{
    return (int)sizeof(agg_state);
}
