/*
 * kern_gpupreagg.cuh - hand-written kernel templates of GpuPreAgg for sm_100a.
 *
 * The reference runs six OpenCL kernels per chunk
 * (opencl_gpupreagg.h:380-856: preparation, set_rindex, bitonic_local/step/
 * merge, reduction) and materialises an intermediate TUPSLOT store.  Here one
 * persistent kernel per chunk does all of it:
 *
 *   TMA (cp.async.bulk) column slices -> shared-memory ring   [producer warp]
 *   gpupreagg_qual_eval()   generated, per row                [consumer warps]
 *   gpupreagg_projection()  generated, into registers (pagg_row)
 *   no GROUP BY : per-thread accumulators -> warp shuffle -> CTA -> one
 *                 deterministic merge by the last CTA into the state row
 *   GROUP BY    : shared-memory open-addressing table (SoA) per CTA, spill and
 *                 final merge into a global open-addressing table (AoS)
 *
 * and gpupreagg_flush writes the partial rows as a TUPSLOT kern_data_store
 * exactly as ExecStoreVirtualTuple expects them (opencl_common.h:417-427,
 * 567-580).  Merge rules follow opencl_gpupreagg.h:862-987 (PMIN/PMAX ignore
 * NULL, PSUM becomes non-NULL on the first non-NULL input) with three
 * deliberate differences, each making the device handle strictly more rows
 * without changing any result:
 *   - integer sums are 64/128-bit and counts 64-bit internally; values that
 *     do not fit the output column are emitted as several partial rows
 *     (the final aggregates add them up), never CpuReCheck;
 *   - float sums re-check a *row* whose magnitude could overflow the sum in
 *     some order (|x| > 2^960, float4: 2^88) instead of the whole chunk;
 *   - float min/max order NaN above everything like PostgreSQL
 *     (opencl_common.h:1553-1560), not OpenCL min()/max().
 * A row that the device cannot finish is flagged in a per-chunk bitmap and
 * the chunk status becomes StromError_CpuReCheck; every other row of the
 * chunk is still aggregated on the device.
 *
 * Compile-time inputs, emitted by codegen (see codegen.cpp):
 *   GPUPREAGG_NUM_INCOLS, GPUPREAGG_INCOL_SLOT(colidx), GPUPREAGG_INCOL_LIST(_)
 *   GPUPREAGG_NUM_KEYS,  GPUPREAGG_KEY_LIST(_)
 *   GPUPREAGG_NUM_AGGS,  GPUPREAGG_NUM_CELLS, GPUPREAGG_AGG_LIST(_)
 *   GPUPREAGG_NUM_OUTCOLS, GPUPREAGG_OUT_LIST(_)
 *   GPUPREAGG_FIELD_ROLE(colidx), GPUPREAGG_FIELD_INDEX(colidx)
 *   GPUPREAGG_HAS_QUAL, GPUPREAGG_NNCLASS_LIST(_)
 * and by the CUDA layer (-D): GPUPREAGG_CONSUMER_WARPS.
 */
#ifndef KERN_GPUPREAGG_CUH
#define KERN_GPUPREAGG_CUH

#include "kern_shared.h"

/* tile size and pipeline depth are launch parameters (the CUDA layer fits
 * them, together with the CTA-local hash table, into the 227 KB of shared
 * memory); only their upper bounds are compile time */
#define GPUPREAGG_MAX_STAGES        8
#ifndef GPUPREAGG_CONSUMER_WARPS
#define GPUPREAGG_CONSUMER_WARPS    16
#endif
#ifndef GPUPREAGG_MIN_CTAS
#define GPUPREAGG_MIN_CTAS          1
#endif
#define GPUPREAGG_CONSUMER_THREADS  (GPUPREAGG_CONSUMER_WARPS * 32)
#define GPUPREAGG_BLOCK_THREADS     (GPUPREAGG_CONSUMER_THREADS + 32)

/* 0 when gpupreagg_qual_eval() is the constant `true` (no WHERE clause) */
#ifndef GPUPREAGG_HAS_QUAL
#define GPUPREAGG_HAS_QUAL          1
#endif

/* 1 when the planner expects very many groups: the scan kernels can deal
 * rows into partitions (pgs_part_emit) and gpupreagg_partagg has a body */
#ifndef GPUPREAGG_PARTITIONED
#define GPUPREAGG_PARTITIONED       0
#endif
#if GPUPREAGG_PARTITIONED
#define PGS_PART_NPARTS(gs)         ((gs).part_nparts)
#else
#define PGS_PART_NPARTS(gs)         0U
#endif

/* 1 (experimental, PGSTROM_GATHER_PAYLOAD): only the columns the qual reads
 * are staged (GPUPREAGG_INCOL_STAGED(slot)); rows that pass the qual fetch
 * the others from HBM by row number */
/* many-groups deal pass: 4-row batches a lane has under way at a time
 * (measured: 1 and 2 take the same time - the global cursor atomics are bound
 * by throughput, not by latency - so the default is the one without spills) */
#ifndef GPUPREAGG_DEAL_STEPS
#define GPUPREAGG_DEAL_STEPS        1
#endif
#ifndef GPUPREAGG_GATHER_PAYLOAD
#define GPUPREAGG_GATHER_PAYLOAD    0
#endif

/* 1 when gpupreagg_qual_eval() reads varlena datums (numeric, text) */
#ifndef GPUPREAGG_QUAL_DEREFS
#define GPUPREAGG_QUAL_DEREFS       0
#endif

#ifndef GPUPREAGG_DEBUG_LEVEL
#define GPUPREAGG_DEBUG_LEVEL       0
#endif
/* GPUPREAGG_DEBUG_LEVEL 4: cycle counters of the consumer warps (lane 0),
 * summed into 8-byte words behind gs.nrows_scanned: [16] wait for the tile,
 * [17] qual + queueing, [18] hash chain, [19] number of chains */
#if GPUPREAGG_DEBUG_LEVEL == 4
#define PGS_DBG_DECL        long long __dbg_t0 = 0, __dbg_acc[4] = {0, 0, 0, 0};
#define PGS_DBG_START()     __dbg_t0 = clock64();
#define PGS_DBG_STOP(i)     __dbg_acc[i] += clock64() - __dbg_t0;
#define PGS_DBG_COUNT(i)    __dbg_acc[i] += 1;
#define PGS_DBG_FLUSH()                                                 \
    if ((threadIdx.x & 31) == 0)                                        \
        for (int __i = 0; __i < 4; __i++)                               \
            atomicAdd((unsigned long long *)gs.nrows_scanned + 14 + __i, \
                      (unsigned long long)__dbg_acc[__i]);
#else
#define PGS_DBG_DECL
#define PGS_DBG_START()
#define PGS_DBG_STOP(i)
#define PGS_DBG_COUNT(i)
#define PGS_DBG_FLUSH()
#endif

#ifndef PGS_ROWS_PER_THREAD
#define PGS_ROWS_PER_THREAD         4
#endif

#define PGS_MAX(a,b)    ((a) > (b) ? (a) : (b))

/* ------------------------------------------------------------------
 * The session's kern_parambuf (Const / Param values of the query, the same
 * image that travels in kern_gpupreagg.kparams) also lives in constant
 * memory: the generated code fetches its KPARAM_n per row
 * (pg_<type>_param), and from the kern_gpupreagg in HBM that is a chain of
 * three dependent global loads the compiler may not hoist over the atomics
 * of the loop.  From the constant bank it is three LDC.  The CUDA layer
 * fills the symbol when it opens the session and refuses parameter buffers
 * that do not fit.
 * ------------------------------------------------------------------ */
#define PGS_CONST_KPARAMS_MAX       16384
extern "C" { __constant__ __align__(16) unsigned char pgs_const_kparams[PGS_CONST_KPARAMS_MAX]; }
#define KERN_GPUPREAGG_PARAMBUF_CONST   ((const kern_parambuf *)pgs_const_kparams)

/* ------------------------------------------------------------------
 * pagg_datum / pagg_row: what gpupreagg_projection() produces for one input
 * row (the reference writes the same thing into a kds_src TUPSLOT row,
 * gpupreagg.c:1495-1748; here it lives in registers).
 * ------------------------------------------------------------------ */
typedef struct
{
    bool            isnull;
    union {
        cl_short    short_val;
        cl_int      int_val;
        cl_long     long_val;
        cl_float    float_val;
        cl_double   double_val;
        cl_ulong    ulong_val;
    };
} pagg_datum;

struct pagg_row
{
    pagg_datum  key[PGS_MAX(GPUPREAGG_NUM_KEYS, 1)];
    pagg_datum  agg[PGS_MAX(GPUPREAGG_NUM_AGGS, 1)];

    __device__ __forceinline__ void
    store(cl_uint colidx, bool isnull, cl_ulong bits)
    {
        /* colidx is a literal in generated code: both switches fold */
        if (GPUPREAGG_FIELD_ROLE(colidx) == GPUPREAGG_FIELD_IS_GROUPKEY)
        {
            key[GPUPREAGG_FIELD_INDEX(colidx)].isnull = isnull;
            key[GPUPREAGG_FIELD_INDEX(colidx)].ulong_val = bits;
        }
        else if (GPUPREAGG_FIELD_ROLE(colidx) == GPUPREAGG_FIELD_IS_AGGFUNC)
        {
            agg[GPUPREAGG_FIELD_INDEX(colidx)].isnull = isnull;
            agg[GPUPREAGG_FIELD_INDEX(colidx)].ulong_val = bits;
        }
    }
};

/* the two generated functions (defined after this header is included) */
template <typename KDS>
DEVFN bool
gpupreagg_qual_eval(cl_int *errcode, const kern_parambuf *kparams,
                    const KDS &kds, const void *ktoast, cl_uint kds_index);
template <typename KDS>
DEVFN void
gpupreagg_projection(cl_int *errcode, const kern_parambuf *kparams,
                     const KDS &kds_in, pagg_row &kds_src, const void *ktoast,
                     cl_uint rowidx_in, cl_uint rowidx_out);

/* pg_<type>_vstore(kds_src, kds_in, errcode, colidx, rowidx_out, datum):
 * same call shape as the reference's generated projection
 * (gpupreagg.c:1508-1518); kds_src is the pagg_row. */
#define STROMCL_SIMPLE_VARSTORE_TEMPLATE(NAME,BASE)                     \
    template <typename KDS>                                             \
    DEVFN void                                                          \
    pg_##NAME##_vstore(pagg_row &kds_src, const KDS &kds_in,            \
                       cl_int *errcode, cl_uint colidx,                 \
                       cl_uint rowidx_out, pg_##NAME##_t datum)         \
    {                                                                   \
        union { BASE v_base; cl_ulong v_datum; } temp;                  \
        temp.v_datum = 0;                                               \
        temp.v_base = datum.value;                                      \
        kds_src.store(colidx, datum.isnull, temp.v_datum);              \
    }
STROMCL_SIMPLE_VARSTORE_TEMPLATE(bool, cl_bool)
STROMCL_SIMPLE_VARSTORE_TEMPLATE(int2, cl_short)
STROMCL_SIMPLE_VARSTORE_TEMPLATE(int4, cl_int)
STROMCL_SIMPLE_VARSTORE_TEMPLATE(int8, cl_long)
STROMCL_SIMPLE_VARSTORE_TEMPLATE(float4, cl_float)
STROMCL_SIMPLE_VARSTORE_TEMPLATE(float8, cl_double)
STROMCL_SIMPLE_VARSTORE_TEMPLATE(date, cl_int)
STROMCL_SIMPLE_VARSTORE_TEMPLATE(time, cl_long)
STROMCL_SIMPLE_VARSTORE_TEMPLATE(timestamp, cl_long)

#ifdef KERN_NUMERIC_CUH
template <typename KDS>
DEVFN void
pg_numeric_vstore(pagg_row &kds_src, const KDS &kds_in, cl_int *errcode,
                  cl_uint colidx, cl_uint rowidx_out, pg_numeric_t datum)
{
    kds_src.store(colidx, datum.isnull, datum.value);
}
#endif

#ifdef KERN_TEXTLIB_CUH
/* text / bpchar appear in the pagg_row only as grouping keys ("kernel text",
 * kern_textlib.cuh) */
template <typename KDS>
DEVFN void
pg_text_vstore(pagg_row &kds_src, const KDS &kds_in, cl_int *errcode,
               cl_uint colidx, cl_uint rowidx_out, pg_text_t datum)
{
    bool        isnull;
    cl_ulong    word = pgs_text_keybits(errcode, datum, false, &isnull);

    kds_src.store(colidx, isnull, word);
}
template <typename KDS>
DEVFN void
pg_bpchar_vstore(pagg_row &kds_src, const KDS &kds_in, cl_int *errcode,
                 cl_uint colidx, cl_uint rowidx_out, pg_bpchar_t datum)
{
    bool        isnull;
    cl_ulong    word = pgs_text_keybits(errcode, datum, true, &isnull);

    kds_src.store(colidx, isnull, word);
}
#endif

/* NULL-const output columns need no work on the device */
template <typename KDS>
DEVFN void
pg_common_vstore(pagg_row &kds_src, const KDS &kds_in, cl_int *errcode,
                 cl_uint colidx, cl_uint rowidx_out, bool isnull)
{}

/* ------------------------------------------------------------------
 * state cells.  One group's running state = GPUPREAGG_NUM_CELLS 8-byte
 * cells + one bit per aggregate ("saw a non-NULL input").
 *
 *   PSUM  INT / LONGS : cl_long sum            (nrows; psum of int2/int4 cast)
 *   PSUM  LONG        : 128-bit sum, 2 cells   (psum of a genuine int8)
 *   PSUM  FLOAT/DOUBLE: cl_double sum
 *   PMIN/PMAX SHORT/INT/LONG : cl_long
 *   PMIN/PMAX FLOAT/DOUBLE   : order-preserving cl_ulong key, NaN highest
 * ------------------------------------------------------------------ */
#define PGS_F8_SIGNBIT      0x8000000000000000ULL
#define PGS_F8_NANKEY       0xFFFFFFFFFFFFFFFFULL
/* re-check thresholds of float sums (see the head of this file) */
#define PGS_PSUM_DOUBLE_LIMIT   9.7453140113999990e+288     /* 2^960 */
#define PGS_PSUM_FLOAT_LIMIT    3.0948500982134507e+26      /* 2^88  */

DEVFN cl_ulong
pgs_f8_sortkey(double v)
{
    cl_long     b = __double_as_longlong(v);
    /* negative: flip all bits; non-negative: set the sign bit */
    cl_ulong    k = (cl_ulong)b ^ ((cl_ulong)(b >> 63) | PGS_F8_SIGNBIT);

    return (v != v) ? PGS_F8_NANKEY : k;
}

DEVFN double
pgs_f8_from_sortkey(cl_ulong k)
{
    if (k == PGS_F8_NANKEY)
        return __longlong_as_double(0x7FF8000000000000LL);
    return __longlong_as_double((cl_long)((k & PGS_F8_SIGNBIT)
                                          ? (k & ~PGS_F8_SIGNBIT) : ~k));
}

/* identity element of each cell kind */
#define PGS_CELL_INIT_PSUM_INT(c,p)     (p)[c] = 0;
#define PGS_CELL_INIT_PSUM_LONGS(c,p)   (p)[c] = 0;
#define PGS_CELL_INIT_PSUM_LONG(c,p)    (p)[c] = 0; (p)[(c)+1] = 0;
#define PGS_CELL_INIT_PSUM_FLOAT(c,p)   (p)[c] = 0;
#define PGS_CELL_INIT_PSUM_DOUBLE(c,p)  (p)[c] = 0;
#define PGS_CELL_INIT_PMIN_SHORT(c,p)   (p)[c] = (cl_ulong)(cl_long)INT_MAX;
#define PGS_CELL_INIT_PMIN_INT(c,p)     (p)[c] = (cl_ulong)(cl_long)INT_MAX;
#define PGS_CELL_INIT_PMIN_LONG(c,p)    (p)[c] = (cl_ulong)LONG_MAX;
/* float min/max cells: with GROUP BY the table cells are updated by 64-bit
 * integer atomics, so they hold order-preserving keys (NaN highest).  Without
 * GROUP BY everything lives in registers and the cell is the raw double:
 * compares run on the FP64 pipe instead of four ALU ops per value.  The
 * identity of min is NaN ("greater than everything"), of max -Infinity. */
#if GPUPREAGG_NUM_KEYS == 0
#define PGS_F8_MIN_IDENTITY     0x7FF8000000000000ULL
#define PGS_F8_MAX_IDENTITY     0xFFF0000000000000ULL
#define PGS_F8_CELL(v)          ((cl_ulong)__double_as_longlong(v))
#define PGS_F8_UNCELL(k)        __longlong_as_double((cl_long)(k))
#else
#define PGS_F8_MIN_IDENTITY     PGS_F8_NANKEY
#define PGS_F8_MAX_IDENTITY     0ULL
#define PGS_F8_CELL(v)          pgs_f8_sortkey(v)
#define PGS_F8_UNCELL(k)        pgs_f8_from_sortkey(k)
#endif
#define PGS_CELL_INIT_PMIN_FLOAT(c,p)   (p)[c] = PGS_F8_MIN_IDENTITY;
#define PGS_CELL_INIT_PMIN_DOUBLE(c,p)  (p)[c] = PGS_F8_MIN_IDENTITY;
#define PGS_CELL_INIT_PMAX_SHORT(c,p)   (p)[c] = (cl_ulong)(cl_long)INT_MIN;
#define PGS_CELL_INIT_PMAX_INT(c,p)     (p)[c] = (cl_ulong)(cl_long)INT_MIN;
#define PGS_CELL_INIT_PMAX_LONG(c,p)    (p)[c] = (cl_ulong)LONG_MIN;
#define PGS_CELL_INIT_PMAX_FLOAT(c,p)   (p)[c] = PGS_F8_MAX_IDENTITY;
#define PGS_CELL_INIT_PMAX_DOUBLE(c,p)  (p)[c] = PGS_F8_MAX_IDENTITY;

/* NUMERIC (kern_numeric.cuh): a sum takes three cells - a 128-bit integer at
 * scale PGS_NUMERIC_SUM_SCALE (lo, hi) and the largest display scale seen -;
 * min / max keep the 64-bit device numeric, all ones = nothing seen */
#define PGS_CELL_INIT_PSUM_NUMERIC(c,p) (p)[c] = 0; (p)[(c)+1] = 0; (p)[(c)+2] = 0;
#define PGS_CELL_INIT_PMIN_NUMERIC(c,p) (p)[c] = 0xFFFFFFFFFFFFFFFFULL;
#define PGS_CELL_INIT_PMAX_NUMERIC(c,p) (p)[c] = 0xFFFFFFFFFFFFFFFFULL;

/* ---- value of one projected datum in "cell domain" ---- */
#define PGS_NEWVAL_PSUM_INT(d)      ((cl_ulong)(cl_long)(d).int_val)
#define PGS_NEWVAL_PSUM_LONGS(d)    ((cl_ulong)(d).long_val)
#define PGS_NEWVAL_PSUM_LONG(d)     ((cl_ulong)(d).long_val)
#define PGS_NEWVAL_PSUM_FLOAT(d)    ((cl_ulong)__double_as_longlong((double)(d).float_val))
#define PGS_NEWVAL_PSUM_DOUBLE(d)   ((cl_ulong)__double_as_longlong((d).double_val))
#define PGS_NEWVAL_PMIN_SHORT(d)    ((cl_ulong)(cl_long)(d).short_val)
#define PGS_NEWVAL_PMIN_INT(d)      ((cl_ulong)(cl_long)(d).int_val)
#define PGS_NEWVAL_PMIN_LONG(d)     ((cl_ulong)(d).long_val)
#define PGS_NEWVAL_PMIN_FLOAT(d)    PGS_F8_CELL((double)(d).float_val)
#define PGS_NEWVAL_PMIN_DOUBLE(d)   PGS_F8_CELL((d).double_val)
#define PGS_NEWVAL_PMAX_SHORT(d)    PGS_NEWVAL_PMIN_SHORT(d)
#define PGS_NEWVAL_PMAX_INT(d)      PGS_NEWVAL_PMIN_INT(d)
#define PGS_NEWVAL_PMAX_LONG(d)     PGS_NEWVAL_PMIN_LONG(d)
#define PGS_NEWVAL_PMAX_FLOAT(d)    PGS_NEWVAL_PMIN_FLOAT(d)
#define PGS_NEWVAL_PMAX_DOUBLE(d)   PGS_NEWVAL_PMIN_DOUBLE(d)

/* ---- merge one cell-domain value into a cell: plain (registers / one
 * owner) and atomic (shared or global table) flavours.  They take the value
 * already in cell domain so that the same code merges rows and states. ---- */
/* PLAIN flavours are branch free (`ok` selects), so that the per-row body of
 * the no-group kernel is straight-line code; ATOMIC flavours skip the
 * memory operation when `ok` is false. */
#define PGS_MERGE_PLAIN_SUM_I64(p,c,v,ok)   (p)[c] += ((ok) ? (cl_ulong)(v) : 0ULL);
#define PGS_MERGE_PLAIN_SUM_F64(p,c,v,ok)                               \
    { double __t = __longlong_as_double((cl_long)(p)[c]) +              \
                   __longlong_as_double((cl_long)(v));                  \
      (p)[c] = (ok) ? (cl_ulong)__double_as_longlong(__t) : (p)[c]; }
#define PGS_MERGE_PLAIN_MIN_I64(p,c,v,ok)                               \
    (p)[c] = ((ok) & ((cl_long)(v) < (cl_long)(p)[c])) ? (cl_ulong)(v) : (p)[c];
#define PGS_MERGE_PLAIN_MAX_I64(p,c,v,ok)                               \
    (p)[c] = ((ok) & ((cl_long)(v) > (cl_long)(p)[c])) ? (cl_ulong)(v) : (p)[c];
#define PGS_MERGE_PLAIN_MIN_I32(p,c,v,ok)                               \
    (p)[c] = (cl_ulong)(cl_uint)((ok) ? min((cl_int)(p)[c], (cl_int)(v)) : (cl_int)(p)[c]);
#define PGS_MERGE_PLAIN_MAX_I32(p,c,v,ok)                               \
    (p)[c] = (cl_ulong)(cl_uint)((ok) ? max((cl_int)(p)[c], (cl_int)(v)) : (cl_int)(p)[c]);
#define PGS_MERGE_PLAIN_MIN_U64(p,c,v,ok)                               \
    (p)[c] = ((ok) & ((cl_ulong)(v) < (cl_ulong)(p)[c])) ? (cl_ulong)(v) : (p)[c];
#define PGS_MERGE_PLAIN_MAX_U64(p,c,v,ok)                               \
    (p)[c] = ((ok) & ((cl_ulong)(v) > (cl_ulong)(p)[c])) ? (cl_ulong)(v) : (p)[c];
#define PGS_MERGE_THREAD_CNT_I32(p,c,v,ok)                              \
    (p)[c] = (cl_ulong)((cl_uint)(p)[c] + ((ok) ? (cl_uint)(v) : 0U));
#define PGS_MERGE_THREAD_SUM_I64(p,c,v,ok)  PGS_MERGE_PLAIN_SUM_I64(p,c,v,ok)
/* the addend is known to fit 32 bits (int2/int4 widened to int8): one
 * IMAD.WIDE on the FMA pipe instead of select + sign extension + 64-bit add */
#define PGS_MERGE_THREAD_SUM_I32W(p,c,v,ok)                             \
    (p)[c] = (cl_ulong)((cl_long)(p)[c] +                               \
                        (cl_long)(cl_int)(v) * (cl_long)(cl_int)(ok));
#define PGS_MERGE_PLAIN_SUM_I32W(p,c,v,ok)  PGS_MERGE_PLAIN_SUM_I64(p,c,v,ok)
#define PGS_MERGE_ATOMIC_SUM_I32W(p,c,v,ok) PGS_MERGE_ATOMIC_SUM_I64(p,c,v,ok)
#define PGS_MERGE_THREAD_SUM_F64(p,c,v,ok)  PGS_MERGE_PLAIN_SUM_F64(p,c,v,ok)
#define PGS_MERGE_THREAD_MIN_I64(p,c,v,ok)  PGS_MERGE_PLAIN_MIN_I64(p,c,v,ok)
#define PGS_MERGE_THREAD_MAX_I64(p,c,v,ok)  PGS_MERGE_PLAIN_MAX_I64(p,c,v,ok)
#define PGS_MERGE_THREAD_MIN_I32(p,c,v,ok)  PGS_MERGE_PLAIN_MIN_I32(p,c,v,ok)
#define PGS_MERGE_THREAD_MAX_I32(p,c,v,ok)  PGS_MERGE_PLAIN_MAX_I32(p,c,v,ok)
#define PGS_MERGE_THREAD_MIN_U64(p,c,v,ok)  PGS_MERGE_PLAIN_MIN_U64(p,c,v,ok)
#define PGS_MERGE_THREAD_MAX_U64(p,c,v,ok)  PGS_MERGE_PLAIN_MAX_U64(p,c,v,ok)
#define PGS_MERGE_PLAIN_CNT_I32(p,c,v,ok)   PGS_MERGE_PLAIN_SUM_I64(p,c,v,ok)
#define PGS_MERGE_ATOMIC_CNT_I32(p,c,v,ok)  PGS_MERGE_ATOMIC_SUM_I64(p,c,v,ok)

/* float min/max (see PGS_F8_CELL) */
#if GPUPREAGG_NUM_KEYS == 0
#define PGS_MERGE_PLAIN_MIN_F64(p,c,v,ok)                               \
    { double __a = PGS_F8_UNCELL((p)[c]), __x = PGS_F8_UNCELL(v);        \
      (p)[c] = ((ok) & ((__x < __a) | (__a != __a))) ? (cl_ulong)(v) : (p)[c]; }
#define PGS_MERGE_PLAIN_MAX_F64(p,c,v,ok)                               \
    { double __a = PGS_F8_UNCELL((p)[c]), __x = PGS_F8_UNCELL(v);        \
      (p)[c] = ((ok) & ((__x > __a) | (__x != __x))) ? (cl_ulong)(v) : (p)[c]; }
#else
#define PGS_MERGE_PLAIN_MIN_F64(p,c,v,ok)   PGS_MERGE_PLAIN_MIN_U64(p,c,v,ok)
#define PGS_MERGE_PLAIN_MAX_F64(p,c,v,ok)   PGS_MERGE_PLAIN_MAX_U64(p,c,v,ok)
#endif
/* per-row flavour (no GROUP BY): a running min/max changes only O(log n)
 * times, so the update sits under a rarely taken branch and the common path
 * is a single FP64 compare.  min: cell NaN = nothing but NaN seen so far. */
#if GPUPREAGG_NUM_KEYS == 0
#define PGS_MERGE_THREAD_MIN_F64(p,c,v,ok)                              \
    { double __a = PGS_F8_UNCELL((p)[c]), __x = PGS_F8_UNCELL(v);        \
      if ((ok) & !(__x >= __a))                                         \
      {                                                                 \
          asm volatile("" ::: "memory");                                \
          if (__x == __x)                                               \
              (p)[c] = (cl_ulong)(v);                                   \
      } }
#define PGS_MERGE_THREAD_MAX_F64(p,c,v,ok)                              \
    { double __a = PGS_F8_UNCELL((p)[c]), __x = PGS_F8_UNCELL(v);        \
      if ((ok) & !(__x <= __a))                                         \
      {                                                                 \
          asm volatile("" ::: "memory");                                \
          if (__a == __a)                                               \
              (p)[c] = (cl_ulong)(v);                                   \
      } }
#else
#define PGS_MERGE_THREAD_MIN_F64(p,c,v,ok)  PGS_MERGE_PLAIN_MIN_F64(p,c,v,ok)
#define PGS_MERGE_THREAD_MAX_F64(p,c,v,ok)  PGS_MERGE_PLAIN_MAX_F64(p,c,v,ok)
#endif
#define PGS_MERGE_ATOMIC_MIN_F64(p,c,v,ok)  PGS_MERGE_ATOMIC_MIN_U64(p,c,v,ok)
#define PGS_MERGE_ATOMIC_MAX_F64(p,c,v,ok)  PGS_MERGE_ATOMIC_MAX_U64(p,c,v,ok)

#define PGS_MERGE_ATOMIC_SUM_I64(p,c,v,ok)                              \
    if (ok) atomicAdd((unsigned long long *)&(p)[c], (unsigned long long)(v));
#define PGS_MERGE_ATOMIC_SUM_F64(p,c,v,ok)                              \
    if (ok) atomicAdd((double *)&(p)[c], __longlong_as_double((cl_long)(v)));
#define PGS_MERGE_ATOMIC_MIN_I64(p,c,v,ok)                              \
    if ((ok) && (cl_long)(v) < *((volatile cl_long *)&(p)[c]))          \
        atomicMin((long long *)&(p)[c], (long long)(v));
#define PGS_MERGE_ATOMIC_MAX_I64(p,c,v,ok)                              \
    if ((ok) && (cl_long)(v) > *((volatile cl_long *)&(p)[c]))          \
        atomicMax((long long *)&(p)[c], (long long)(v));
/* table cells keep int4 min/max sign-extended (64-bit atomics) */
#define PGS_MERGE_ATOMIC_MIN_I32(p,c,v,ok)  PGS_MERGE_ATOMIC_MIN_I64(p,c,(cl_long)(cl_int)(v),ok)
#define PGS_MERGE_ATOMIC_MAX_I32(p,c,v,ok)  PGS_MERGE_ATOMIC_MAX_I64(p,c,(cl_long)(cl_int)(v),ok)
#define PGS_MERGE_ATOMIC_MIN_U64(p,c,v,ok)                              \
    if ((ok) && (cl_ulong)(v) < *((volatile cl_ulong *)&(p)[c]))        \
        atomicMin((unsigned long long *)&(p)[c], (unsigned long long)(v));
#define PGS_MERGE_ATOMIC_MAX_U64(p,c,v,ok)                              \
    if ((ok) && (cl_ulong)(v) > *((volatile cl_ulong *)&(p)[c]))        \
        atomicMax((unsigned long long *)&(p)[c], (unsigned long long)(v));

/* ---- typed thread accumulators (no GROUP BY fast path).
 * The running state of a consumer thread lives in registers of the natural
 * C type of each aggregate (pgs_tacc, one member a<i> per aggregate) and is
 * folded into the 8-byte cells once, after the last tile.  Rows reach it
 * only in batches that are free of errors and of NaN / infinite / huge float
 * inputs (pgs_row_special), so nothing can overflow and every float compare
 * is one DSETP: min starts as NaN ("nothing seen") and !(x >= NaN) adopts the
 * first value, max starts at -Infinity.  Updates are predicated on `ok`. ---- */
struct pgs_i128 { cl_ulong lo, hi; };

#define PGS_TACC_TYPE_PSUM_INT      cl_uint
#define PGS_TACC_TYPE_PSUM_LONGS    cl_long
#define PGS_TACC_TYPE_PSUM_LONG     pgs_i128
#define PGS_TACC_TYPE_PSUM_FLOAT    double
#define PGS_TACC_TYPE_PSUM_DOUBLE   double
#define PGS_TACC_TYPE_PMIN_SHORT    cl_int
#define PGS_TACC_TYPE_PMIN_INT      cl_int
#define PGS_TACC_TYPE_PMIN_LONG     cl_long
#define PGS_TACC_TYPE_PMIN_FLOAT    double
#define PGS_TACC_TYPE_PMIN_DOUBLE   double
#define PGS_TACC_TYPE_PMAX_SHORT    cl_int
#define PGS_TACC_TYPE_PMAX_INT      cl_int
#define PGS_TACC_TYPE_PMAX_LONG     cl_long
#define PGS_TACC_TYPE_PMAX_FLOAT    double
#define PGS_TACC_TYPE_PMAX_DOUBLE   double

/* numeric aggregates never take the batch fast path (PGS_SPECIAL_NUMERIC) */
#define PGS_TACC_TYPE_PSUM_NUMERIC  cl_ulong
#define PGS_TACC_TYPE_PMIN_NUMERIC  cl_ulong
#define PGS_TACC_TYPE_PMAX_NUMERIC  cl_ulong
#define PGS_TACC_INIT_PSUM_NUMERIC(a)   (a) = 0;
#define PGS_TACC_INIT_PMIN_NUMERIC(a)   (a) = 0;
#define PGS_TACC_INIT_PMAX_NUMERIC(a)   (a) = 0;
#define PGS_TACC_CALC4_PSUM_NUMERIC(a,d0,d1,d2,d3,k0,k1,k2,k3)
#define PGS_TACC_CALC4_PMIN_NUMERIC(a,d0,d1,d2,d3,k0,k1,k2,k3)
#define PGS_TACC_CALC4_PMAX_NUMERIC(a,d0,d1,d2,d3,k0,k1,k2,k3)
#define PGS_TACC_CELL_PSUM_NUMERIC(a,src,c) (src)[c] = 0; (src)[(c)+1] = 0; (src)[(c)+2] = 0;
#define PGS_TACC_CELL_PMIN_NUMERIC(a,src,c) (src)[c] = 0xFFFFFFFFFFFFFFFFULL;
#define PGS_TACC_CELL_PMAX_NUMERIC(a,src,c) (src)[c] = 0xFFFFFFFFFFFFFFFFULL;

#define PGS_F8_NAN      __longlong_as_double(0x7FF8000000000000LL)
#define PGS_F8_NEGINF   __longlong_as_double((cl_long)0xFFF0000000000000ULL)

#define PGS_TACC_INIT_PSUM_INT(a)       (a) = 0;
#define PGS_TACC_INIT_PSUM_LONGS(a)     (a) = 0;
#define PGS_TACC_INIT_PSUM_LONG(a)      (a).lo = 0; (a).hi = 0;
#define PGS_TACC_INIT_PSUM_FLOAT(a)     (a) = 0.0;
#define PGS_TACC_INIT_PSUM_DOUBLE(a)    (a) = 0.0;
#define PGS_TACC_INIT_PMIN_SHORT(a)     (a) = INT_MAX;
#define PGS_TACC_INIT_PMIN_INT(a)       (a) = INT_MAX;
#define PGS_TACC_INIT_PMIN_LONG(a)      (a) = LONG_MAX;
#define PGS_TACC_INIT_PMIN_FLOAT(a)     (a) = PGS_F8_NAN;
#define PGS_TACC_INIT_PMIN_DOUBLE(a)    (a) = PGS_F8_NAN;
#define PGS_TACC_INIT_PMAX_SHORT(a)     (a) = INT_MIN;
#define PGS_TACC_INIT_PMAX_INT(a)       (a) = INT_MIN;
#define PGS_TACC_INIT_PMAX_LONG(a)      (a) = LONG_MIN;
#define PGS_TACC_INIT_PMAX_FLOAT(a)     (a) = PGS_F8_NEGINF;
#define PGS_TACC_INIT_PMAX_DOUBLE(a)    (a) = PGS_F8_NEGINF;

/* Updates take a batch of 4 rows (d0..d3 projected datums, k0..k3 "row is
 * valid and the datum is not NULL").  The consumer loop is bound by the ALU
 * pipe (LOP3 / SEL / IADD3 / VIMNMX, one warp instruction per two cycles and
 * SM sub-partition), so the formulations below push work to the FMA and FP64
 * pipes where they can:
 *   - int4 -> int8 sums: zero the NULL addends (1 SEL each), then
 *     IMAD.WIDE acc = x * one + acc; `one` is a run-time 1 the compiler cannot
 *     fold, otherwise it emits IADD3 + LEA.HI.X.SX32 (two more ALU ops)
 *   - float8 sums: acc = fma(x, k ? 1.0 : 0.0, acc): one SEL for the high word
 *     of the factor and one DFMA, instead of DADD + two FSEL.  x * 0.0 is an
 *     exact zero because special floats never reach this path.
 *   - counts: one 3-input add per two rows. */
#define PGS_K01(k)      ((k) ? 1U : 0U)
#define PGS_KF8(k)      __hiloint2double((k) ? 0x3FF00000 : 0, 0)

DEVFN cl_long
pgs_madwide(cl_int x, cl_int y, cl_long acc)
{
    cl_long r;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(x), "r"(y), "l"(acc));
    return r;
}

/* nrows(): the addend is 0 or 1 (gpupreagg.c:1559-1600) - count bits */
#define PGS_TACC_CALC4_PSUM_INT(a,d0,d1,d2,d3,k0,k1,k2,k3)              \
    (a) += __popc((((k0) & ((d0).int_val != 0)) ? 1U : 0U) |            \
                  (((k1) & ((d1).int_val != 0)) ? 2U : 0U) |            \
                  (((k2) & ((d2).int_val != 0)) ? 4U : 0U) |            \
                  (((k3) & ((d3).int_val != 0)) ? 8U : 0U));
#define PGS_TACC_CALC4_PSUM_LONGS(a,d0,d1,d2,d3,k0,k1,k2,k3)            \
    (a) = pgs_madwide((k0) ? (cl_int)(d0).long_val : 0, one, (a));      \
    (a) = pgs_madwide((k1) ? (cl_int)(d1).long_val : 0, one, (a));      \
    (a) = pgs_madwide((k2) ? (cl_int)(d2).long_val : 0, one, (a));      \
    (a) = pgs_madwide((k3) ? (cl_int)(d3).long_val : 0, one, (a));
#define PGS_TACC_ADD128(a,d,k)                                          \
    pgs_add128_PLAIN(&(a).lo, &(a).hi, (cl_ulong)(d).long_val,          \
                     (d).long_val < 0 ? ~0ULL : 0ULL, (k));
#define PGS_TACC_CALC4_PSUM_LONG(a,d0,d1,d2,d3,k0,k1,k2,k3)             \
    PGS_TACC_ADD128(a,d0,k0) PGS_TACC_ADD128(a,d1,k1)                   \
    PGS_TACC_ADD128(a,d2,k2) PGS_TACC_ADD128(a,d3,k3)
#define PGS_TACC_CALC4_PSUM_FLOAT(a,d0,d1,d2,d3,k0,k1,k2,k3)            \
    (a) = fma((double)(d0).float_val, PGS_KF8(k0), (a));                \
    (a) = fma((double)(d1).float_val, PGS_KF8(k1), (a));                \
    (a) = fma((double)(d2).float_val, PGS_KF8(k2), (a));                \
    (a) = fma((double)(d3).float_val, PGS_KF8(k3), (a));
#define PGS_TACC_CALC4_PSUM_DOUBLE(a,d0,d1,d2,d3,k0,k1,k2,k3)           \
    (a) = fma((d0).double_val, PGS_KF8(k0), (a));                       \
    (a) = fma((d1).double_val, PGS_KF8(k1), (a));                       \
    (a) = fma((d2).double_val, PGS_KF8(k2), (a));                       \
    (a) = fma((d3).double_val, PGS_KF8(k3), (a));
#define PGS_TACC_MIN1(a,v,k)    if (k) (a) = min((a), (v));
#define PGS_TACC_MAX1(a,v,k)    if (k) (a) = max((a), (v));
#define PGS_TACC_CALC4_PMIN_SHORT(a,d0,d1,d2,d3,k0,k1,k2,k3)            \
    PGS_TACC_MIN1(a,(cl_int)(d0).short_val,k0) PGS_TACC_MIN1(a,(cl_int)(d1).short_val,k1) \
    PGS_TACC_MIN1(a,(cl_int)(d2).short_val,k2) PGS_TACC_MIN1(a,(cl_int)(d3).short_val,k3)
#define PGS_TACC_CALC4_PMIN_INT(a,d0,d1,d2,d3,k0,k1,k2,k3)              \
    PGS_TACC_MIN1(a,(d0).int_val,k0) PGS_TACC_MIN1(a,(d1).int_val,k1)   \
    PGS_TACC_MIN1(a,(d2).int_val,k2) PGS_TACC_MIN1(a,(d3).int_val,k3)
#define PGS_TACC_CALC4_PMIN_LONG(a,d0,d1,d2,d3,k0,k1,k2,k3)             \
    PGS_TACC_MIN1(a,(d0).long_val,k0) PGS_TACC_MIN1(a,(d1).long_val,k1) \
    PGS_TACC_MIN1(a,(d2).long_val,k2) PGS_TACC_MIN1(a,(d3).long_val,k3)
#define PGS_TACC_CALC4_PMAX_SHORT(a,d0,d1,d2,d3,k0,k1,k2,k3)            \
    PGS_TACC_MAX1(a,(cl_int)(d0).short_val,k0) PGS_TACC_MAX1(a,(cl_int)(d1).short_val,k1) \
    PGS_TACC_MAX1(a,(cl_int)(d2).short_val,k2) PGS_TACC_MAX1(a,(cl_int)(d3).short_val,k3)
#define PGS_TACC_CALC4_PMAX_INT(a,d0,d1,d2,d3,k0,k1,k2,k3)              \
    PGS_TACC_MAX1(a,(d0).int_val,k0) PGS_TACC_MAX1(a,(d1).int_val,k1)   \
    PGS_TACC_MAX1(a,(d2).int_val,k2) PGS_TACC_MAX1(a,(d3).int_val,k3)
#define PGS_TACC_CALC4_PMAX_LONG(a,d0,d1,d2,d3,k0,k1,k2,k3)             \
    PGS_TACC_MAX1(a,(d0).long_val,k0) PGS_TACC_MAX1(a,(d1).long_val,k1) \
    PGS_TACC_MAX1(a,(d2).long_val,k2) PGS_TACC_MAX1(a,(d3).long_val,k3)
/* x is never NaN here: the unordered compare also adopts x while the cell is
 * still NaN.  Written in PTX so that the predicate stays an input of the one
 * DSETP instead of a second level of selects. */
DEVFN void
pgs_pred_dmin(double &acc, double x, bool ok)
{
    asm("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %2, 0;\n\t"
        "setp.ltu.and.f64 q, %1, %0, p;\n\tselp.f64 %0, %1, %0, q;\n\t}"
        : "+d"(acc) : "d"(x), "r"((cl_uint)ok));
}
DEVFN void
pgs_pred_dmax(double &acc, double x, bool ok)
{
    asm("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %2, 0;\n\t"
        "setp.gt.and.f64 q, %1, %0, p;\n\tselp.f64 %0, %1, %0, q;\n\t}"
        : "+d"(acc) : "d"(x), "r"((cl_uint)ok));
}
#define PGS_TACC_FMIN1(a,x,k)   pgs_pred_dmin((a), (x), (k));
#define PGS_TACC_FMAX1(a,x,k)   pgs_pred_dmax((a), (x), (k));
#define PGS_TACC_CALC4_PMIN_FLOAT(a,d0,d1,d2,d3,k0,k1,k2,k3)            \
    PGS_TACC_FMIN1(a,(double)(d0).float_val,k0) PGS_TACC_FMIN1(a,(double)(d1).float_val,k1) \
    PGS_TACC_FMIN1(a,(double)(d2).float_val,k2) PGS_TACC_FMIN1(a,(double)(d3).float_val,k3)
#define PGS_TACC_CALC4_PMIN_DOUBLE(a,d0,d1,d2,d3,k0,k1,k2,k3)           \
    PGS_TACC_FMIN1(a,(d0).double_val,k0) PGS_TACC_FMIN1(a,(d1).double_val,k1) \
    PGS_TACC_FMIN1(a,(d2).double_val,k2) PGS_TACC_FMIN1(a,(d3).double_val,k3)
#define PGS_TACC_CALC4_PMAX_FLOAT(a,d0,d1,d2,d3,k0,k1,k2,k3)            \
    PGS_TACC_FMAX1(a,(double)(d0).float_val,k0) PGS_TACC_FMAX1(a,(double)(d1).float_val,k1) \
    PGS_TACC_FMAX1(a,(double)(d2).float_val,k2) PGS_TACC_FMAX1(a,(double)(d3).float_val,k3)
#define PGS_TACC_CALC4_PMAX_DOUBLE(a,d0,d1,d2,d3,k0,k1,k2,k3)           \
    PGS_TACC_FMAX1(a,(d0).double_val,k0) PGS_TACC_FMAX1(a,(d1).double_val,k1) \
    PGS_TACC_FMAX1(a,(d2).double_val,k2) PGS_TACC_FMAX1(a,(d3).double_val,k3)

/* accumulator -> cell-domain value(s) at src[c] */
#define PGS_TACC_CELL_PSUM_INT(a,src,c)     (src)[c] = (cl_ulong)(a);
#define PGS_TACC_CELL_PSUM_LONGS(a,src,c)   (src)[c] = (cl_ulong)(a);
#define PGS_TACC_CELL_PSUM_LONG(a,src,c)    (src)[c] = (a).lo; (src)[(c)+1] = (a).hi;
#define PGS_TACC_CELL_PSUM_FLOAT(a,src,c)   (src)[c] = (cl_ulong)__double_as_longlong(a);
#define PGS_TACC_CELL_PSUM_DOUBLE(a,src,c)  (src)[c] = (cl_ulong)__double_as_longlong(a);
#define PGS_TACC_CELL_PMIN_SHORT(a,src,c)   (src)[c] = (cl_ulong)(cl_long)(a);
#define PGS_TACC_CELL_PMIN_INT(a,src,c)     (src)[c] = (cl_ulong)(cl_long)(a);
#define PGS_TACC_CELL_PMIN_LONG(a,src,c)    (src)[c] = (cl_ulong)(a);
#define PGS_TACC_CELL_PMIN_FLOAT(a,src,c)   (src)[c] = PGS_F8_CELL(a);
#define PGS_TACC_CELL_PMIN_DOUBLE(a,src,c)  (src)[c] = PGS_F8_CELL(a);
#define PGS_TACC_CELL_PMAX_SHORT(a,src,c)   PGS_TACC_CELL_PMIN_SHORT(a,src,c)
#define PGS_TACC_CELL_PMAX_INT(a,src,c)     PGS_TACC_CELL_PMIN_INT(a,src,c)
#define PGS_TACC_CELL_PMAX_LONG(a,src,c)    PGS_TACC_CELL_PMIN_LONG(a,src,c)
#define PGS_TACC_CELL_PMAX_FLOAT(a,src,c)   PGS_TACC_CELL_PMIN_FLOAT(a,src,c)
#define PGS_TACC_CELL_PMAX_DOUBLE(a,src,c)  PGS_TACC_CELL_PMIN_DOUBLE(a,src,c)

/* ---- SHARED flavours (CTA-local table): shared memory has native 32-bit
 * atomics only, every 64-bit add / min / max is a compare-and-swap loop.
 * Counts of one launch fit 32 bits (the low half of the cell); a sum of
 * 32-bit addends is carried by hand from the low into the high half. ---- */
#define PGS_MERGE_SHARED_CNT_I32(p,c,v,ok)                              \
    if (ok) atomicAdd((cl_uint *)&(p)[c], (cl_uint)(v));
#define PGS_MERGE_SHARED_SUM_I32W(p,c,v,ok)                             \
    if (ok)                                                             \
    {                                                                   \
        cl_uint __lo = (cl_uint)(v);                                    \
        cl_uint __old = atomicAdd((cl_uint *)&(p)[c], __lo);            \
        cl_uint __hi = (cl_uint)((cl_int)__lo >> 31) +                  \
                       ((cl_uint)(__old + __lo) < __old ? 1U : 0U);     \
        if (__hi != 0)                                                  \
            atomicAdd((cl_uint *)&(p)[c] + 1, __hi);                    \
    }
#define PGS_MERGE_SHARED_SUM_I64(p,c,v,ok)  PGS_MERGE_ATOMIC_SUM_I64(p,c,v,ok)
#define PGS_MERGE_SHARED_SUM_F64(p,c,v,ok)  PGS_MERGE_ATOMIC_SUM_F64(p,c,v,ok)
#define PGS_MERGE_SHARED_MIN_I32(p,c,v,ok)  PGS_MERGE_ATOMIC_MIN_I32(p,c,v,ok)
#define PGS_MERGE_SHARED_MAX_I32(p,c,v,ok)  PGS_MERGE_ATOMIC_MAX_I32(p,c,v,ok)
#define PGS_MERGE_SHARED_MIN_I64(p,c,v,ok)  PGS_MERGE_ATOMIC_MIN_I64(p,c,v,ok)
#define PGS_MERGE_SHARED_MAX_I64(p,c,v,ok)  PGS_MERGE_ATOMIC_MAX_I64(p,c,v,ok)
#define PGS_MERGE_SHARED_MIN_F64(p,c,v,ok)  PGS_MERGE_ATOMIC_MIN_F64(p,c,v,ok)
#define PGS_MERGE_SHARED_MAX_F64(p,c,v,ok)  PGS_MERGE_ATOMIC_MAX_F64(p,c,v,ok)

/* 128-bit sums: (lo,hi) two's complement.  The carry out of `lo` is known
 * from the value atomicAdd returns, and additions commute, so two 64-bit
 * atomics give an exact 128-bit sum without a lock. */
DEVFN void
pgs_add128_PLAIN(cl_ulong *plo, cl_ulong *phi, cl_ulong vlo, cl_ulong vhi, bool ok)
{
    cl_ulong    old = *plo;

    vlo = ok ? vlo : 0;
    vhi = ok ? vhi : 0;
    *plo = old + vlo;
    *phi += vhi + (*plo < old ? 1 : 0);
}
DEVFN void
pgs_add128_THREAD(cl_ulong *plo, cl_ulong *phi, cl_ulong vlo, cl_ulong vhi, bool ok)
{
    pgs_add128_PLAIN(plo, phi, vlo, vhi, ok);
}
DEVFN void
pgs_add128_FAST(cl_ulong *plo, cl_ulong *phi, cl_ulong vlo, cl_ulong vhi, bool ok)
{
    pgs_add128_PLAIN(plo, phi, vlo, vhi, ok);
}
DEVFN void
pgs_add128_ATOMIC(cl_ulong *plo, cl_ulong *phi, cl_ulong vlo, cl_ulong vhi, bool ok)
{
    if (ok)
    {
        cl_ulong    old = atomicAdd((unsigned long long *)plo,
                                    (unsigned long long)vlo);
        cl_ulong    inc = vhi + ((old + vlo) < old ? 1 : 0);

        if (inc != 0)
            atomicAdd((unsigned long long *)phi, (unsigned long long)inc);
    }
}

DEVFN void
pgs_add128_SHARED(cl_ulong *plo, cl_ulong *phi, cl_ulong vlo, cl_ulong vhi, bool ok)
{
    pgs_add128_ATOMIC(plo, phi, vlo, vhi, ok);
}

#ifdef KERN_NUMERIC_CUH
/* numeric sum: the addend at the fixed scale, as two 64-bit halves */
#define PGS_NUMSUM_TEMPLATE(MODE)                                       \
    template <typename CELLS>                                           \
    DEVFN void                                                          \
    pgs_numsum_##MODE(CELLS p, int c, cl_ulong packed, bool ok)         \
    {                                                                   \
        int         ds = (ok ? pgs_numeric_dscale(packed) : 0);         \
        pgs_s128    add = (ok ? pgs_numeric_scaled(packed, PGS_NUMERIC_SUM_SCALE) : 0); \
        pgs_add128_##MODE(&p[c], &p[c + 1], (cl_ulong)add,              \
                          (cl_ulong)((pgs_u128)add >> 64), ok);         \
        PGS_MERGE_##MODE##_MAX_I64(p, c + 2, (cl_ulong)(cl_long)ds, ok) \
    }                                                                   \
    template <typename CELLS, typename SRC>                             \
    DEVFN void                                                          \
    pgs_numsum_merge_##MODE(CELLS p, int c, SRC q, bool ok)             \
    {                                                                   \
        pgs_add128_##MODE(&p[c], &p[c + 1], q[c], q[c + 1], ok);        \
        PGS_MERGE_##MODE##_MAX_I64(p, c + 2, q[c + 2], ok)              \
    }
PGS_NUMSUM_TEMPLATE(PLAIN)
PGS_NUMSUM_TEMPLATE(THREAD)
PGS_NUMSUM_TEMPLATE(ATOMIC)
PGS_NUMSUM_TEMPLATE(SHARED)

/* numeric min / max; WANT = -1 keeps the smaller, +1 the larger */
DEVFN void
pgs_numext_PLAIN(cl_ulong *p, cl_ulong v, bool ok, int want)
{
    if (ok && v != PGS_NUMERIC_EMPTY &&
        (*p == PGS_NUMERIC_EMPTY || pgs_numeric_cmp(v, *p) == want))
        *p = v;
}
DEVFN void
pgs_numext_THREAD(cl_ulong *p, cl_ulong v, bool ok, int want)
{
    pgs_numext_PLAIN(p, v, ok, want);
}
DEVFN void
pgs_numext_ATOMIC(cl_ulong *p, cl_ulong v, bool ok, int want)
{
    if (ok && v != PGS_NUMERIC_EMPTY)
    {
        cl_ulong    old = *((volatile cl_ulong *)p);

        while (old == PGS_NUMERIC_EMPTY || pgs_numeric_cmp(v, old) == want)
        {
            cl_ulong seen = atomicCAS((unsigned long long *)p, (unsigned long long)old,
                                      (unsigned long long)v);
            if (seen == old)
                break;
            old = seen;
        }
    }
}
DEVFN void
pgs_numext_SHARED(cl_ulong *p, cl_ulong v, bool ok, int want)
{
    pgs_numext_ATOMIC(p, v, ok, want);
}
#endif  /* KERN_NUMERIC_CUH */
#define PGS_AGGCALC_PSUM_NUMERIC(MODE,p,c,d,ok) pgs_numsum_##MODE(p, c, (d).ulong_val, (ok));
#define PGS_AGGCALC_PMIN_NUMERIC(MODE,p,c,d,ok) pgs_numext_##MODE(&(p)[c], (d).ulong_val, (ok), -1);
#define PGS_AGGCALC_PMAX_NUMERIC(MODE,p,c,d,ok) pgs_numext_##MODE(&(p)[c], (d).ulong_val, (ok), 1);
#define PGS_AGGMERGE_PSUM_NUMERIC(MODE,p,c,q,ok) pgs_numsum_merge_##MODE(p, c, q, (ok));
#define PGS_AGGMERGE_PMIN_NUMERIC(MODE,p,c,q,ok) pgs_numext_##MODE(&(p)[c], (q)[c], (ok), -1);
#define PGS_AGGMERGE_PMAX_NUMERIC(MODE,p,c,q,ok) pgs_numext_##MODE(&(p)[c], (q)[c], (ok), 1);
/* a sum takes values up to the fixed scale whose scaled mantissa stays below
 * 2^96: 2^31 of them cannot overflow the 128-bit cell */
#define PGS_AGGCHECK_PSUM_NUMERIC(d)                                    \
    {                                                                   \
        int __ds = pgs_numeric_dscale((d).ulong_val);                   \
        if (__ds < 0 || __ds > PGS_NUMERIC_SUM_SCALE ||                 \
            (((pgs_u128)PG_NUMERIC_MANTISSA((d).ulong_val) *            \
              pgs_pow10_u128(PGS_NUMERIC_SUM_SCALE - (__ds < 0 ? 0 : (__ds > PGS_NUMERIC_SUM_SCALE ? PGS_NUMERIC_SUM_SCALE : __ds)))) >> 96) != 0) \
            STROM_SET_ERROR(errcode, StromError_CpuReCheck);            \
    }
#define PGS_AGGCHECK_PMIN_NUMERIC(d)
#define PGS_AGGCHECK_PMAX_NUMERIC(d)
#define PGS_SPECIAL_NUMERIC(d)      md = 0xffffffffU;

/* row -> state (MODE = PLAIN | THREAD | FAST | ATOMIC | SHARED); `d` is a
 * pagg_datum */
#define PGS_AGGCALC_PSUM_INT(MODE,p,c,d,ok)    PGS_MERGE_##MODE##_CNT_I32(p,c,PGS_NEWVAL_PSUM_INT(d),ok)
#define PGS_AGGCALC_PSUM_LONGS(MODE,p,c,d,ok)  PGS_MERGE_##MODE##_SUM_I32W(p,c,PGS_NEWVAL_PSUM_LONGS(d),ok)
#define PGS_AGGCALC_PSUM_LONG(MODE,p,c,d,ok)                            \
    pgs_add128_##MODE(&(p)[c], &(p)[(c)+1], (cl_ulong)(d).long_val,     \
                      (d).long_val < 0 ? ~0ULL : 0ULL, (ok));
#define PGS_AGGCALC_PSUM_FLOAT(MODE,p,c,d,ok)  PGS_MERGE_##MODE##_SUM_F64(p,c,PGS_NEWVAL_PSUM_FLOAT(d),ok)
#define PGS_AGGCALC_PSUM_DOUBLE(MODE,p,c,d,ok) PGS_MERGE_##MODE##_SUM_F64(p,c,PGS_NEWVAL_PSUM_DOUBLE(d),ok)
#define PGS_AGGCALC_PMIN_SHORT(MODE,p,c,d,ok)  PGS_MERGE_##MODE##_MIN_I32(p,c,PGS_NEWVAL_PMIN_SHORT(d),ok)
#define PGS_AGGCALC_PMIN_INT(MODE,p,c,d,ok)    PGS_MERGE_##MODE##_MIN_I32(p,c,PGS_NEWVAL_PMIN_INT(d),ok)
#define PGS_AGGCALC_PMIN_LONG(MODE,p,c,d,ok)   PGS_MERGE_##MODE##_MIN_I64(p,c,PGS_NEWVAL_PMIN_LONG(d),ok)
#define PGS_AGGCALC_PMIN_FLOAT(MODE,p,c,d,ok)  PGS_MERGE_##MODE##_MIN_F64(p,c,PGS_NEWVAL_PMIN_FLOAT(d),ok)
#define PGS_AGGCALC_PMIN_DOUBLE(MODE,p,c,d,ok) PGS_MERGE_##MODE##_MIN_F64(p,c,PGS_NEWVAL_PMIN_DOUBLE(d),ok)
#define PGS_AGGCALC_PMAX_SHORT(MODE,p,c,d,ok)  PGS_MERGE_##MODE##_MAX_I32(p,c,PGS_NEWVAL_PMAX_SHORT(d),ok)
#define PGS_AGGCALC_PMAX_INT(MODE,p,c,d,ok)    PGS_MERGE_##MODE##_MAX_I32(p,c,PGS_NEWVAL_PMAX_INT(d),ok)
#define PGS_AGGCALC_PMAX_LONG(MODE,p,c,d,ok)   PGS_MERGE_##MODE##_MAX_I64(p,c,PGS_NEWVAL_PMAX_LONG(d),ok)
#define PGS_AGGCALC_PMAX_FLOAT(MODE,p,c,d,ok)  PGS_MERGE_##MODE##_MAX_F64(p,c,PGS_NEWVAL_PMAX_FLOAT(d),ok)
#define PGS_AGGCALC_PMAX_DOUBLE(MODE,p,c,d,ok) PGS_MERGE_##MODE##_MAX_F64(p,c,PGS_NEWVAL_PMAX_DOUBLE(d),ok)

/* state -> state (q = source cells) */
#define PGS_AGGMERGE_PSUM_INT(MODE,p,c,q,ok)    PGS_MERGE_##MODE##_SUM_I64(p,c,(q)[c],ok)
#define PGS_AGGMERGE_PSUM_LONGS(MODE,p,c,q,ok)  PGS_MERGE_##MODE##_SUM_I64(p,c,(q)[c],ok)
#define PGS_AGGMERGE_PSUM_LONG(MODE,p,c,q,ok) pgs_add128_##MODE(&(p)[c], &(p)[(c)+1], (q)[c], (q)[(c)+1], (ok));
#define PGS_AGGMERGE_PSUM_FLOAT(MODE,p,c,q,ok)  PGS_MERGE_##MODE##_SUM_F64(p,c,(q)[c],ok)
#define PGS_AGGMERGE_PSUM_DOUBLE(MODE,p,c,q,ok) PGS_MERGE_##MODE##_SUM_F64(p,c,(q)[c],ok)
#define PGS_AGGMERGE_PMIN_SHORT(MODE,p,c,q,ok)  PGS_MERGE_##MODE##_MIN_I32(p,c,(q)[c],ok)
#define PGS_AGGMERGE_PMIN_INT(MODE,p,c,q,ok)    PGS_MERGE_##MODE##_MIN_I32(p,c,(q)[c],ok)
#define PGS_AGGMERGE_PMIN_LONG(MODE,p,c,q,ok)   PGS_MERGE_##MODE##_MIN_I64(p,c,(q)[c],ok)
#define PGS_AGGMERGE_PMIN_FLOAT(MODE,p,c,q,ok)  PGS_MERGE_##MODE##_MIN_F64(p,c,(q)[c],ok)
#define PGS_AGGMERGE_PMIN_DOUBLE(MODE,p,c,q,ok) PGS_MERGE_##MODE##_MIN_F64(p,c,(q)[c],ok)
#define PGS_AGGMERGE_PMAX_SHORT(MODE,p,c,q,ok)  PGS_MERGE_##MODE##_MAX_I32(p,c,(q)[c],ok)
#define PGS_AGGMERGE_PMAX_INT(MODE,p,c,q,ok)    PGS_MERGE_##MODE##_MAX_I32(p,c,(q)[c],ok)
#define PGS_AGGMERGE_PMAX_LONG(MODE,p,c,q,ok)   PGS_MERGE_##MODE##_MAX_I64(p,c,(q)[c],ok)
#define PGS_AGGMERGE_PMAX_FLOAT(MODE,p,c,q,ok)  PGS_MERGE_##MODE##_MAX_F64(p,c,(q)[c],ok)
#define PGS_AGGMERGE_PMAX_DOUBLE(MODE,p,c,q,ok) PGS_MERGE_##MODE##_MAX_F64(p,c,(q)[c],ok)

/* per-row admission check: rows whose float inputs could overflow a sum in
 * some summation order are left to the CPU (row level) */
#define PGS_AGGCHECK_PSUM_INT(d)
#define PGS_AGGCHECK_PSUM_LONGS(d)
#define PGS_AGGCHECK_PSUM_LONG(d)
#define PGS_AGGCHECK_PSUM_FLOAT(d)                                      \
    if (fabsf((d).float_val) > (float)PGS_PSUM_FLOAT_LIMIT)             \
        STROM_SET_ERROR(errcode, StromError_CpuReCheck);
#define PGS_AGGCHECK_PSUM_DOUBLE(d)                                     \
    if (fabs((d).double_val) > PGS_PSUM_DOUBLE_LIMIT)                   \
        STROM_SET_ERROR(errcode, StromError_CpuReCheck);
#define PGS_AGGCHECK_PMIN_SHORT(d)
#define PGS_AGGCHECK_PMIN_INT(d)
#define PGS_AGGCHECK_PMIN_LONG(d)
#define PGS_AGGCHECK_PMIN_FLOAT(d)
#define PGS_AGGCHECK_PMIN_DOUBLE(d)
#define PGS_AGGCHECK_PMAX_SHORT(d)
#define PGS_AGGCHECK_PMAX_INT(d)
#define PGS_AGGCHECK_PMAX_LONG(d)
#define PGS_AGGCHECK_PMAX_FLOAT(d)
#define PGS_AGGCHECK_PMAX_DOUBLE(d)

/* per-row "special float" probe of the no-group fast path: the largest
 * exponent field among the float inputs of a batch of rows, as an integer
 * (sign stripped).  At or above the limit the batch holds NaN, +-Infinity or
 * a magnitude the sum re-checks (PGS_PSUM_*_LIMIT) and takes the careful path. */
#define PGS_SPECIAL_F8_LIMIT    0x7BF00000U     /* high word of 2^960 */
#define PGS_SPECIAL_F4_LIMIT    0x6B800000U     /* bits of 2^88 */
#define PGS_SPECIAL_SHORT(d)
#define PGS_SPECIAL_INT(d)
#define PGS_SPECIAL_LONG(d)
#define PGS_SPECIAL_LONGS(d)
#define PGS_SPECIAL_FLOAT(d)                                            \
    mf = max(mf, __float_as_uint((d).float_val) & 0x7fffffffU);
#define PGS_SPECIAL_DOUBLE(d)                                           \
    md = max(md, (cl_uint)((d).ulong_val >> 32) & 0x7fffffffU);
#define PGS_X_SPECIAL(i,c,OP,TYPE)      PGS_SPECIAL_##TYPE(row.agg[i])

/* X-macro adaptors over GPUPREAGG_AGG_LIST(_) = _(aggidx, cellidx, OP, TYPE) */
#define PGS_X_INIT(i,c,OP,TYPE)         PGS_CELL_INIT_##OP##_##TYPE(c,cells)
#define PGS_X_CHECK(i,c,OP,TYPE)                                        \
    if (!row.agg[i].isnull) { PGS_AGGCHECK_##OP##_##TYPE(row.agg[i]) }
#define PGS_X_CALC_PLAIN(i,c,OP,TYPE)                                   \
    { bool __ok = valid && !row.agg[i].isnull;                          \
      PGS_AGGCALC_##OP##_##TYPE(PLAIN,cells,c,row.agg[i],__ok)          \
      nnmask |= (__ok ? (1U << (i)) : 0U); }
#define PGS_X_CALC_THREAD(i,c,OP,TYPE)                                  \
    { bool __ok = valid & !row.agg[i].isnull;                           \
      PGS_AGGCALC_##OP##_##TYPE(THREAD,cells,c,row.agg[i],__ok) }
#define PGS_X_TACC_DECL(i,c,OP,TYPE)    PGS_TACC_TYPE_##OP##_##TYPE a##i;
#define PGS_X_TACC_INIT(i,c,OP,TYPE)    PGS_TACC_INIT_##OP##_##TYPE(a##i)
#define PGS_X_TACC_CALC4(i,c,OP,TYPE)                                   \
    PGS_TACC_CALC4_##OP##_##TYPE(a##i,                                  \
        rows[0].agg[i], rows[1].agg[i], rows[2].agg[i], rows[3].agg[i], \
        (valids[0] & !rows[0].agg[i].isnull),                           \
        (valids[1] & !rows[1].agg[i].isnull),                           \
        (valids[2] & !rows[2].agg[i].isnull),                           \
        (valids[3] & !rows[3].agg[i].isnull))
#define PGS_X_TACC_CELL(i,c,OP,TYPE)    PGS_TACC_CELL_##OP##_##TYPE(a##i,src,c)
#define PGS_X_CALC_SHARED(i,c,OP,TYPE)                                  \
    { bool __ok = !row.agg[i].isnull;                                   \
      PGS_AGGCALC_##OP##_##TYPE(SHARED,cells,c,row.agg[i],__ok)         \
      nnmask |= (__ok ? (1U << (i)) : 0U); }
/* "saw a non-NULL input" per class of aggregates that share the NULL-ness of
 * one argument (GPUPREAGG_NNCLASS_LIST(_) = _(representative agg, bit mask)) */
#define PGS_X_NNCLASS(rep,mask)                                         \
    nnmask |= ((valid & !row.agg[rep].isnull) ? (mask) : 0U);
#define PGS_X_NNCLASS_COUNT(rep,mask)   + 1
#define PGS_NUM_NNCLASSES   (0 GPUPREAGG_NNCLASS_LIST(PGS_X_NNCLASS_COUNT))
#define PGS_X_NNCLASS_FLAG(rep,mask)                                    \
    nnflag[__k] = nnflag[__k] | (valid & !row.agg[rep].isnull); __k++;
#define PGS_X_NNCLASS_FLAG4(rep,mask)                                   \
    nnflag[__k] = nnflag[__k] |                                         \
        ((valids[0] & !rows[0].agg[rep].isnull) |                       \
         (valids[1] & !rows[1].agg[rep].isnull) |                       \
         (valids[2] & !rows[2].agg[rep].isnull) |                       \
         (valids[3] & !rows[3].agg[rep].isnull)); __k++;
#define PGS_X_NNCLASS_MASK(rep,mask)                                    \
    nnmask |= (nnflag[__k] ? (mask) : 0U); __k++;
#define PGS_X_CALC_ATOMIC(i,c,OP,TYPE)                                  \
    { bool __ok = valid && !row.agg[i].isnull;                          \
      PGS_AGGCALC_##OP##_##TYPE(ATOMIC,cells,c,row.agg[i],__ok)         \
      nnmask |= (__ok ? (1U << (i)) : 0U); }
#define PGS_X_MERGE_PLAIN(i,c,OP,TYPE)                                  \
    { bool __ok = ((src_nn >> (i)) & 1U) != 0;                          \
      PGS_AGGMERGE_##OP##_##TYPE(PLAIN,cells,c,src,__ok) }
#define PGS_X_MERGE_ATOMIC(i,c,OP,TYPE)                                 \
    { bool __ok = ((src_nn >> (i)) & 1U) != 0;                          \
      PGS_AGGMERGE_##OP##_##TYPE(ATOMIC,cells,c,src,__ok) }

template <typename CELLS>
DEVFN void
pgs_cells_init(CELLS cells)
{
    GPUPREAGG_AGG_LIST(PGS_X_INIT)
}

DEVFN void
gpupreagg_aggcheck(cl_int *errcode, const pagg_row &row)
{
    GPUPREAGG_AGG_LIST(PGS_X_CHECK)
}

/* gpupreagg_aggcalc: merge one projected row into a state
 * (the reference's generated switch(resno), gpupreagg.c:1319-1440) */
template <typename CELLS>
DEVFN cl_uint
gpupreagg_aggcalc_plain(CELLS cells, const pagg_row &row, bool valid)
{
    cl_uint nnmask = 0;
    GPUPREAGG_AGG_LIST(PGS_X_CALC_PLAIN)
    return nnmask;
}
/* per-row accumulation into thread registers (no-group kernel); the
 * "non-NULL seen" flags are kept as predicates over the rows of one loop
 * iteration and folded into the bit mask by pgs_nnflags_to_mask() */
DEVFN void
gpupreagg_aggcalc_thread(cl_ulong *cells, const pagg_row &row, bool valid,
                         bool *nnflag)
{
    int __k = 0;
    GPUPREAGG_AGG_LIST(PGS_X_CALC_THREAD)
    GPUPREAGG_NNCLASS_LIST(PGS_X_NNCLASS_FLAG)
    (void)__k;
}
DEVFN cl_uint
pgs_nnflags_to_mask(const bool *nnflag)
{
    cl_uint nnmask = 0;
    int __k = 0;
    GPUPREAGG_NNCLASS_LIST(PGS_X_NNCLASS_MASK)
    (void)__k;
    return nnmask;
}
/* typed per-thread state of the no-group fast path (see PGS_TACC_*) */
struct pgs_tacc
{
    GPUPREAGG_AGG_LIST(PGS_X_TACC_DECL)
    bool    nnflag[PGS_MAX(PGS_NUM_NNCLASSES, 1)];

    __device__ __forceinline__ void
    init(void)
    {
        GPUPREAGG_AGG_LIST(PGS_X_TACC_INIT)
#pragma unroll
        for (int k = 0; k < PGS_MAX(PGS_NUM_NNCLASSES, 1); k++)
            nnflag[k] = false;
    }
    /* a batch of 4 rows without errors and special floats; `one` is 1 */
    __device__ __forceinline__ void
    calc4(const pagg_row *rows, const bool *valids, cl_int one)
    {
        GPUPREAGG_AGG_LIST(PGS_X_TACC_CALC4)
        (void)one;
    }
    /* "saw a non-NULL input" of the whole batch: OR over the rows first, so
     * that tests of adjacent validity bits fold into one */
    __device__ __forceinline__ void
    note_nonnull(const pagg_row *rows, const bool *valids)
    {
        int __k = 0;
        GPUPREAGG_NNCLASS_LIST(PGS_X_NNCLASS_FLAG4)
        (void)__k;
    }
    /* fold into the cells of the thread */
    __device__ __forceinline__ cl_uint
    fold(cl_ulong *cells) const
    {
        cl_ulong    src[PGS_MAX(GPUPREAGG_NUM_CELLS, 1)];
        cl_uint     src_nn = pgs_nnflags_to_mask(nnflag);

        GPUPREAGG_AGG_LIST(PGS_X_TACC_CELL)
        GPUPREAGG_AGG_LIST(PGS_X_MERGE_PLAIN)
        return src_nn;
    }
};

DEVFN void
pgs_row_special(const pagg_row &row, cl_uint &md, cl_uint &mf)
{
    GPUPREAGG_AGG_LIST(PGS_X_SPECIAL)
}
template <typename CELLS>
DEVFN cl_uint
gpupreagg_aggcalc_atomic(CELLS cells, const pagg_row &row)
{
    const bool valid = true;
    cl_uint nnmask = 0;
    GPUPREAGG_AGG_LIST(PGS_X_CALC_ATOMIC)
    return nnmask;
}
template <typename CELLS>
DEVFN cl_uint
gpupreagg_aggcalc_shared(CELLS cells, const pagg_row &row)
{
    cl_uint nnmask = 0;
    GPUPREAGG_AGG_LIST(PGS_X_CALC_SHARED)
    return nnmask;
}
template <typename CELLS, typename SRC>
DEVFN void
gpupreagg_aggmerge_plain(CELLS cells, SRC src, cl_uint src_nn)
{
    GPUPREAGG_AGG_LIST(PGS_X_MERGE_PLAIN)
}
template <typename CELLS, typename SRC>
DEVFN void
gpupreagg_aggmerge_atomic(CELLS cells, SRC src, cl_uint src_nn)
{
    GPUPREAGG_AGG_LIST(PGS_X_MERGE_ATOMIC)
}

/* ------------------------------------------------------------------
 * device-resident session state (allocated by the CUDA layer)
 * ------------------------------------------------------------------ */
#define PGS_SLOT_EMPTY      0U
#define PGS_SLOT_BUSY       1U
#define PGS_SLOT_READY      2U

/* a slot of the global (AoS) table:
 *   ctrl_lo : PGS_SLOT_* in bits 0-1, key-is-NULL bits from bit 8
 *   ctrl_hi : "saw non-NULL" bit per aggregate
 *   key[NKEYS], cell[NCELLS]
 */
#define PGS_SLOT_WORDS  (1 + GPUPREAGG_NUM_KEYS + GPUPREAGG_NUM_CELLS)
#define PGS_SLOT_BYTES  (8 * PGS_SLOT_WORDS)
/* in the table a slot starts on a 32-byte boundary: it then covers whole
 * sectors (the unit of L2 atomics and DRAM traffic) and never more of them
 * than it has to; exported records stay packed */
#define PGS_SLOT_STRIDE ((PGS_SLOT_WORDS + 3) & ~3)

/* pgs_gstate / pgs_kern_desc: see kern_shared.h (shared with the host) */

/* ------------------------------------------------------------------
 * stage layout: [col0 values | col1 values | ... | col0 bitmap | ...]
 * every piece 128-byte aligned
 * ------------------------------------------------------------------ */
#define PGS_ALIGN128(x)     (((x) + 127U) & ~127U)

/* tile_rows is a multiple of 1024, so every piece is 128-byte aligned */
/* columns that are not staged (GPUPREAGG_GATHER_PAYLOAD) take no room */
#define PGS_STAGE_ATTLEN(s) \
    (GPUPREAGG_INCOL_STAGED(s) ? GPUPREAGG_INCOL_ATTLEN(s) : 0U)
DEVFN cl_uint
pgs_stage_val_off(int slot, cl_uint tile_rows)
{
    cl_uint off = 0;
#pragma unroll
    for (int s = 0; s < slot; s++)
        off += tile_rows * PGS_STAGE_ATTLEN(s);
    return off;
}
DEVFN cl_uint
pgs_stage_nul_off(int slot, cl_uint tile_rows)
{
#if GPUPREAGG_GATHER_PAYLOAD
    cl_uint nstaged = 0;
#pragma unroll
    for (int s = 0; s < slot; s++)
        nstaged += (GPUPREAGG_INCOL_STAGED(s) ? 1U : 0U);
    return pgs_stage_val_off(GPUPREAGG_NUM_INCOLS, tile_rows) +
        nstaged * (tile_rows / 8);
#else
    return pgs_stage_val_off(GPUPREAGG_NUM_INCOLS, tile_rows) +
        (cl_uint)slot * (tile_rows / 8);
#endif
}
#define PGS_STAGE_BYTES(tile_rows)  pgs_stage_nul_off(GPUPREAGG_NUM_INCOLS, (tile_rows))
/* Per consumer warp (GROUP BY with a WHERE clause only, see the consumer
 * loop): a circular queue of the rows that passed the qual and a buffer that
 * keeps the values of the few queued rows a tile leaves behind.
 *
 *   queue   : PGS_ROWQ_ENTRIES x u16.  An entry is the row's position in the
 *             staged tile, or PGS_ROWQ_LEFT | n for row n of the leftover buffer
 *   leftover: [col0 values | col1 values | ... | validity mask (u32) | row
 *             number (u32)], PGS_ROWQ_LEFT_ENTRIES entries per array.  Queue
 *             position p always maps to entry p mod 64, so an entry that
 *             survives several tiles never moves.
 */
#define PGS_ROWQ_ENTRIES        256         /* power of two, > 31 + 128 */
/* gather variant: steps of 128 rows a warp's share of one tile may have
 * (one bit per row in a 32-bit mask); the host sizes the tiles accordingly */
#define PGS_GATHER_MAX_STEPS    8
#if GPUPREAGG_GATHER_PAYLOAD && GPUPREAGG_CONSUMER_WARPS < 8
#error "the gather variant needs at least 8 consumer warps (tiles of 8192 rows, 8 steps per warp)"
#endif
#define PGS_ROWQ_LEFT           0x8000U
#define PGS_ROWQ_LEFT_ENTRIES   64
DEVFN cl_uint
pgs_rowq_val_off(int slot)
{
    cl_uint off = 2 * PGS_ROWQ_ENTRIES;
#pragma unroll
    for (int s = 0; s < slot; s++)
        off += PGS_ROWQ_LEFT_ENTRIES * GPUPREAGG_INCOL_ATTLEN(s);
    return (off + 7U) & ~7U;
}
#define PGS_ROWQ_MASK_OFF   ((pgs_rowq_val_off(GPUPREAGG_NUM_INCOLS) + 7U) & ~7U)
#define PGS_ROWQ_ROW_OFF    (PGS_ROWQ_MASK_OFF + 4 * PGS_ROWQ_LEFT_ENTRIES)
#if GPUPREAGG_GATHER_PAYLOAD
/* gather variant: the queue holds 32-bit row numbers and nothing else */
#define PGS_ROWQ_WARP_BYTES (4U * PGS_ROWQ_ENTRIES)
#else
#define PGS_ROWQ_WARP_BYTES ((PGS_ROWQ_ROW_OFF + 4 * PGS_ROWQ_LEFT_ENTRIES + 15U) & ~15U)
#endif
#if GPUPREAGG_NUM_KEYS > 0 && GPUPREAGG_HAS_QUAL
#define PGS_ROWQ_BYTES      (PGS_ROWQ_WARP_BYTES * GPUPREAGG_CONSUMER_WARPS)
#else
#define PGS_ROWQ_BYTES      0
#endif
/* head of dynamic smem: 2 x MAX_STAGES mbarriers, column positions, misc
 * (pgs_smem_head, 128-byte aligned), then the row queues */
#define PGS_SMEM_HEAD_FIXED \
    PGS_ALIGN128(16 * GPUPREAGG_MAX_STAGES + 8 * PGS_MAX(GPUPREAGG_NUM_INCOLS,1) + 64)
#define PGS_SMEM_HEAD_BYTES     PGS_ALIGN128(PGS_SMEM_HEAD_FIXED + PGS_ROWQ_BYTES)

/* view of ONE queued row for the generated functions: the row lives either
 * in the staged tile or in the warp's leftover buffer (per lane). */
struct kern_qrow_smem
{
    cl_uint     val_off[PGS_MAX(GPUPREAGG_NUM_INCOLS, 1)];  /* array of the column */
    cl_uint     nul_off[PGS_MAX(GPUPREAGG_NUM_INCOLS, 1)];  /* staged bitmap or NO_NULLMAP */
    cl_uint     idx;            /* position in those arrays */
    cl_uint     lmask;          /* leftover row: validity bit per slot */
    bool        isleft;

    template <typename T>
    __device__ __forceinline__ bool
    fetch(int slot, cl_uint rowidx, T &out) const
    {
        cl_uint vb = 1U;

        out = *((const T *)(__pgs_smem + val_off[slot]) + idx);
        if (nul_off[slot] != KERN_TILE_NO_NULLMAP)
            vb = (*((const cl_uint *)(__pgs_smem + nul_off[slot]) + (idx >> 5))
                  >> (idx & 31)) & 1U;
        if (isleft)
            vb = (lmask >> slot) & 1U;
        return vb != 0;
    }
};
template <int ATTLEN> struct pgs_rowq_type;
template <> struct pgs_rowq_type<8> { typedef cl_ulong T; };
template <> struct pgs_rowq_type<4> { typedef cl_uint T; };
template <> struct pgs_rowq_type<2> { typedef cl_ushort T; };
template <> struct pgs_rowq_type<1> { typedef unsigned char T; };
/* where the view of queue entry __e finds column `slot` */
#define PGS_X_INCOL_QVIEW(slot,colidx,attlen)                           \
    qrow.val_off[slot] = (qrow.isleft ? __qbase + pgs_rowq_val_off(slot) \
                                      : tile.val_off[slot]);            \
    qrow.nul_off[slot] = (qrow.isleft ? KERN_TILE_NO_NULLMAP : tile.nul_off[slot]);
/* copy staged row __ri into leftover entry __pos */
#define PGS_X_INCOL_QKEEP(slot,colidx,attlen)                           \
    ((pgs_rowq_type<attlen>::T *)(__pgs_smem + __qbase + pgs_rowq_val_off(slot)))[__pos] = \
        ((const pgs_rowq_type<attlen>::T *)(__pgs_smem + tile.val_off[slot]))[__ri]; \
    {                                                                   \
        cl_uint __vb = 1U;                                              \
        if (tile.nul_off[slot] != KERN_TILE_NO_NULLMAP)                 \
            __vb = (*((const cl_uint *)(__pgs_smem + tile.nul_off[slot]) + (__ri >> 5)) \
                    >> (__ri & 31)) & 1U;                               \
        __mask |= (__vb << (slot));                                     \
    }

/* ------------------------------------------------------------------
 * mbarrier + bulk async copy (TMA engine, 1-D) wrappers
 * ------------------------------------------------------------------ */
DEVFN cl_uint
pgs_smem_addr(const void *p)
{
    return (cl_uint)__cvta_generic_to_shared(p);
}
DEVFN void
pgs_mbar_init(cl_ulong *bar, cl_uint count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;"
                 :: "r"(pgs_smem_addr(bar)), "r"(count) : "memory");
}
DEVFN void
pgs_mbar_arrive_expect_tx(cl_ulong *bar, cl_uint bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(pgs_smem_addr(bar)), "r"(bytes) : "memory");
}
DEVFN void
pgs_mbar_arrive(cl_ulong *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];"
                 :: "r"(pgs_smem_addr(bar)) : "memory");
}
DEVFN void
pgs_mbar_wait(cl_ulong *bar, cl_uint parity)
{
    cl_uint addr = pgs_smem_addr(bar);
    cl_uint done;

    do {
        asm volatile("{\n\t"
                     ".reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t"
                     "}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
/* the producer lane's wait for a free stage: it has nothing else to do, but
 * its polling must not take issue slots from the consumer warps that share
 * its scheduler (ncu, WHERE + GROUP BY: 12% of all executed instructions were
 * this loop) */
DEVFN void
pgs_mbar_wait_relaxed(cl_ulong *bar, cl_uint parity)
{
    cl_uint addr = pgs_smem_addr(bar);
    cl_uint done;

    for (;;)
    {
        asm volatile("{\n\t"
                     ".reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t"
                     "}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done)
            break;
        __nanosleep(64);
    }
}
DEVFN void
pgs_bulk_g2s(void *dst_smem, const void *src_gmem, cl_uint bytes, cl_ulong *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes"
                 " [%0], [%1], %2, [%3];"
                 :: "r"(pgs_smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes),
                    "r"(pgs_smem_addr(bar)) : "memory");
}

/* ------------------------------------------------------------------
 * hashing of the group key
 * ------------------------------------------------------------------ */
DEVFN cl_ulong
pgs_mix64(cl_ulong x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

/* hash of (key values with NULL -> 0, NULL bits): 32-bit multiply-add fold of
 * the key halves, then the murmur3 finaliser - a dozen 32-bit instructions
 * instead of three 64-bit multiplies per key.  The CTA-local table indexes
 * with the high bits, the global table with the low bits. */
DEVFN cl_uint
pgs_fmix32(cl_uint h)
{
    h ^= h >> 16;
    h *= 0x85ebca6bU;
    h ^= h >> 13;
    h *= 0xc2b2ae35U;
    h ^= h >> 16;
    return h;
}

DEVFN cl_ulong
pgs_hash_keyvals(const cl_ulong *keyvals, cl_uint knull)
{
    cl_uint     h = 0x9e3779b9U + knull;

#pragma unroll
    for (int k = 0; k < GPUPREAGG_NUM_KEYS; k++)
    {
        h = h * 0x9e3779b1U + (cl_uint)keyvals[k];
        h = h * 0x85ebca77U + (cl_uint)(keyvals[k] >> 32);
    }
    h = pgs_fmix32(h);
    return ((cl_ulong)h << 32) | h;
}

DEVFN cl_ulong
pgs_hash_keys(const pagg_row &row, cl_uint &knull)
{
    cl_ulong    keyvals[PGS_MAX(GPUPREAGG_NUM_KEYS, 1)];

    knull = 0;
#pragma unroll
    for (int k = 0; k < GPUPREAGG_NUM_KEYS; k++)
    {
        /* NULL keys form one group (gpupreagg.c:1234-1243) */
        keyvals[k] = row.key[k].isnull ? 0 : row.key[k].ulong_val;
        if (row.key[k].isnull)
            knull |= (1U << k);
    }
    return pgs_hash_keyvals(keyvals, knull);
}

/* key normalisation: float keys must compare like PostgreSQL's btree
 * operators (-0 = +0, NaN = NaN), so they are canonicalised at projection
 * time by codegen (pgs_f8_canon); other keys compare bitwise. */
DEVFN double
pgs_f8_canon(double v)
{
    if (isnan(v))
        return __longlong_as_double(0x7FF8000000000000LL);
    return (v == 0.0) ? 0.0 : v;
}

/* ------------------------------------------------------------------
 * CTA-local table (shared memory, SoA):
 *   ctrl_lo[n] (u32) | ctrl_hi[n] (u32) | key[NKEYS][n] (u64) | cell[NCELLS][n]
 * ctrl_lo = tag (see pgs_sh_find_slot): PGS_SLOT_* in bits 0-1, key-is-NULL
 * bits from bit 2, fingerprint of the hash above them; ctrl_hi = "saw a
 * non-NULL input" bit per aggregate.  nslots is a multiple of 32.
 * ------------------------------------------------------------------ */
#define PGS_SH_SLOT_BYTES   (8 + 8 * GPUPREAGG_NUM_KEYS + 8 * GPUPREAGG_NUM_CELLS)

struct pgs_sh_table
{
    cl_uint     base;       /* offset in __pgs_smem */
    cl_uint     nslots;     /* multiple of 32, 0 = disabled */
    cl_uint     salt;       /* odd multiplier of the hash before the bucket is
                             * taken from its high bits; 1 = as is.  A table
                             * image of a partition sees only keys whose hash
                             * shares the high bits (that is what made them
                             * land in the partition), so it mixes first. */

    __device__ __forceinline__ cl_uint *ctrl_lo(cl_uint s) const
    { return (cl_uint *)(__pgs_smem + base) + s; }
    __device__ __forceinline__ cl_uint *ctrl_hi(cl_uint s) const
    { return (cl_uint *)(__pgs_smem + base + 4 * nslots) + s; }
    __device__ __forceinline__ cl_ulong *key(int k, cl_uint s) const
    { return (cl_ulong *)(__pgs_smem + base + 8 * nslots) + (cl_ulong)k * nslots + s; }
    __device__ __forceinline__ cl_ulong *cell(int c, cl_uint s) const
    { return (cl_ulong *)(__pgs_smem + base + 8 * nslots * (1 + GPUPREAGG_NUM_KEYS))
            + (cl_ulong)c * nslots + s; }
};

/* strided view of a slot's cells so that the AGG macros can index p[c] */
struct pgs_sh_cells
{
    cl_ulong   *p0;
    cl_uint     stride;
    __device__ __forceinline__ cl_ulong &operator[](int c) const
    { return p0[(cl_ulong)c * stride]; }
};

/* ------------------------------------------------------------------
 * find-or-insert in the global table.  Returns the slot's word pointer or
 * NULL when the probe limit is hit (table full => StromError_DataStoreNoSpace)
 * ------------------------------------------------------------------ */
/* Both find-or-insert loops below leave only through the loop condition:
 * a lane that won a slot publishes it inside its own iteration, and a lane
 * that found the slot BUSY just goes round again.  (With an early `return`
 * the compiler may park the winner at the reconvergence point behind the
 * loop while a lane of the same warp spins on the slot it has yet to
 * publish.) */
DEVFN cl_ulong *
pgs_gh_find_slot(const pgs_gstate &gs, const cl_ulong *keyvals,
                 cl_uint knull, cl_ulong hash, cl_uint &ninserted)
{
    const cl_uint mask = gs.gh_nslots - 1;
    cl_uint     h = (cl_uint)hash & mask;
    cl_uint     probe = 0;
    cl_ulong   *found = NULL;
    bool        done = false;

    while (!done)
    {
        cl_ulong   *slot = gs.gh_slots + (cl_ulong)h * PGS_SLOT_STRIDE;
        cl_uint     st = *((volatile cl_uint *)slot);

        if ((st & 3U) == PGS_SLOT_READY)
        {
            bool    same = ((st >> 8) == knull);
#pragma unroll
            for (int k = 0; k < GPUPREAGG_NUM_KEYS; k++)
                same = same && (__ldcg(slot + 1 + k) == keyvals[k]);
            if (same)
            {
                found = slot;
                done = true;
            }
            else
            {
                h = (h + 1) & mask;
                if (++probe >= gs.gh_max_probe)
                    done = true;        /* table full */
            }
        }
        else if ((st & 3U) == PGS_SLOT_EMPTY)
        {
            if (atomicCAS((cl_uint *)slot, PGS_SLOT_EMPTY, PGS_SLOT_BUSY) == PGS_SLOT_EMPTY)
            {
#pragma unroll
                for (int k = 0; k < GPUPREAGG_NUM_KEYS; k++)
                    slot[1 + k] = keyvals[k];
                /* publish: the keys are visible before the state is (a
                 * release store, not the sequentially consistent fence of
                 * __threadfence()); the number of groups is counted per
                 * thread and added up once per warp and launch - one
                 * counter bumped by every insert serialises in L2 */
                asm volatile("st.release.gpu.global.u32 [%0], %1;"
                             :: "l"(slot), "r"(PGS_SLOT_READY | (knull << 8)) : "memory");
                ninserted++;
                found = slot;
                done = true;
            }
            /* lost the race: look at the same slot again */
        }
        /* BUSY: the owner publishes shortly; look at the same slot again */
    }
    return found;
}

/* a state record that found no room in the global table goes to the overflow
 * log (record = ctrl | keys | cells, the exported format); false = the log is
 * full as well */
DEVFN bool
pgs_ovf_append(const pgs_gstate &gs, const cl_ulong *keyvals, cl_uint knull,
               const cl_ulong *src, cl_uint src_nn)
{
    cl_uint     pos = atomicAdd(gs.ovf_count, 1U);
    cl_ulong   *rec;

    if (pos >= gs.ovf_cap)
        return false;
    rec = gs.ovf_recs + (cl_ulong)pos * PGS_SLOT_WORDS;
    rec[0] = ((cl_ulong)src_nn << 32) | (knull << 8) | PGS_SLOT_READY;
#pragma unroll
    for (int k = 0; k < GPUPREAGG_NUM_KEYS; k++)
        rec[1 + k] = keyvals[k];
#pragma unroll
    for (int c = 0; c < GPUPREAGG_NUM_CELLS; c++)
        rec[1 + GPUPREAGG_NUM_KEYS + c] = src[c];
    return true;
}

/* merge a state (cells + nn bits) into the global table; a full table sends
 * it to the overflow log, so that a state is never lost */
DEVFN bool
pgs_gh_merge_state(const pgs_gstate &gs, const cl_ulong *keyvals, cl_uint knull,
                   cl_ulong hash, const cl_ulong *src, cl_uint src_nn,
                   cl_uint &ninserted)
{
    cl_ulong   *slot = pgs_gh_find_slot(gs, keyvals, knull, hash, ninserted);
    cl_ulong   *cells;
    cl_uint    *p_nn;

    if (!slot)
        return pgs_ovf_append(gs, keyvals, knull, src, src_nn);
    cells = slot + 1 + GPUPREAGG_NUM_KEYS;
    gpupreagg_aggmerge_atomic(cells, src, src_nn);
    p_nn = (cl_uint *)slot + 1;
    if ((*((volatile cl_uint *)p_nn) & src_nn) != src_nn)
        atomicOr(p_nn, src_nn);
    return true;
}

/*
 * find-or-insert in the CTA-local table; returns the slot index or ~0U if the
 * table is (nearly) full and the row has to go to the global table.
 *
 * The table is probed a bucket of four slots at a time: one 128-bit load
 * brings four tags, a tag = fingerprint of the hash | key-is-NULL bits |
 * PGS_SLOT_* state, and the keys of a slot are compared only when its tag
 * matches.  With linear probing of single slots the longest probe sequence
 * among the 32 lanes of a warp decided how often the warp went round the loop
 * (measured: 7 rounds with 7 lanes active on average at 58% load); with
 * buckets nearly every lane is done after the first round.
 * A key is only ever inserted into the first EMPTY slot of the first bucket
 * that has one, slots never become EMPTY again, and a slot that is BUSY with
 * the fingerprint looked for is waited for - so a key cannot be inserted twice.
 */
#define PGS_TAG_SHIFT       (2 + GPUPREAGG_NUM_KEYS)

DEVFN uint4
pgs_lds_volatile_v4(const void *p)
{
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(pgs_smem_addr(p)) : "memory");
    return v;
}

DEVFN cl_uint
pgs_sh_find_slot(const pgs_sh_table &sh, cl_uint *sh_nused,
                 const cl_ulong *keyvals, cl_uint knull, cl_ulong hash)
{
    const cl_uint nbuckets = sh.nslots >> 2;
    const cl_uint limit = sh.nslots - (sh.nslots >> 2);     /* 75% */
    const cl_uint want = ((cl_uint)hash << PGS_TAG_SHIFT) | (knull << 2) | PGS_SLOT_READY;
    const cl_uint busy = want ^ (PGS_SLOT_READY ^ PGS_SLOT_BUSY);
    cl_uint     b = __umulhi((cl_uint)(hash >> 32) * sh.salt, nbuckets);   /* high bits */
    cl_uint     probe = 0;
    cl_uint     found = ~0U;
    bool        done = false;

    while (!done)
    {
        uint4       t = pgs_lds_volatile_v4(sh.ctrl_lo(4 * b));
        cl_uint     hits = (t.x == want ? 1U : 0U) | (t.y == want ? 2U : 0U) |
                           (t.z == want ? 4U : 0U) | (t.w == want ? 8U : 0U);

        /* fingerprints collide once in 2^(30 - NKEYS) slots: nearly always
         * the first hit is the group */
        while (hits != 0 && found == ~0U)
        {
            cl_uint s = 4 * b + (__ffs(hits) - 1);
            bool    same = true;
#pragma unroll
            for (int k = 0; k < GPUPREAGG_NUM_KEYS; k++)
                same = same && (*((volatile cl_ulong *)sh.key(k, s)) == keyvals[k]);
            if (same)
                found = s;
            hits &= hits - 1;
        }
        if (found != ~0U)
            done = true;
        else if (t.x == busy || t.y == busy || t.z == busy || t.w == busy)
            ;   /* perhaps our key, published shortly: look at the bucket again */
        else
        {
            cl_uint empties = (t.x == PGS_SLOT_EMPTY ? 1U : 0U) | (t.y == PGS_SLOT_EMPTY ? 2U : 0U) |
                              (t.z == PGS_SLOT_EMPTY ? 4U : 0U) | (t.w == PGS_SLOT_EMPTY ? 8U : 0U);
            if (empties != 0)
            {
                cl_uint s = 4 * b + (__ffs(empties) - 1);

                if (*((volatile cl_uint *)sh_nused) >= limit)
                    done = true;
                else if (atomicCAS(sh.ctrl_lo(s), PGS_SLOT_EMPTY, busy) == PGS_SLOT_EMPTY)
                {
                    atomicAdd(sh_nused, 1U);
#pragma unroll
                    for (int k = 0; k < GPUPREAGG_NUM_KEYS; k++)
                        *sh.key(k, s) = keyvals[k];
                    __threadfence_block();
                    atomicExch(sh.ctrl_lo(s), want);
                    found = s;
                    done = true;
                }
                /* lost the race: look at the bucket again */
            }
            else
            {
                b = (b + 1 < nbuckets ? b + 1 : 0);         /* any table size */
                if (++probe >= 16)
                    done = true;
            }
        }
    }
    return found;
}

/* ------------------------------------------------------------------
 * Partition records (GROUP BY with very many groups, see gpupreagg_partagg):
 * the referenced columns of one row in slot order, each on its natural
 * alignment, then the validity mask (bit per slot) and the row number.
 * ------------------------------------------------------------------ */
__host__ __device__ constexpr cl_uint
pgs_rec_val_off(int slot)
{
    cl_uint off = 0;
    for (int s = 0; s <= slot && s < GPUPREAGG_NUM_INCOLS; s++)
    {
        cl_uint a = GPUPREAGG_INCOL_ATTLEN(s);
        off = (off + a - 1U) & ~(a - 1U);
        if (s < slot)
            off += a;
    }
    return off;
}
__host__ __device__ constexpr cl_uint
pgs_rec_end_off(void)
{
    cl_uint off = 0;
    for (int s = 0; s < GPUPREAGG_NUM_INCOLS; s++)
    {
        cl_uint a = GPUPREAGG_INCOL_ATTLEN(s);
        off = ((off + a - 1U) & ~(a - 1U)) + a;
    }
    return (off + 3U) & ~3U;
}
#define PGS_REC_MASK_OFF    pgs_rec_end_off()
#define PGS_REC_ROW_OFF     (PGS_REC_MASK_OFF + 4)
/* a record is a whole number of 16-byte units: it leaves the SM as 128-bit
 * stores (one or two per 32-byte sector), never as one store per column */
#define PGS_REC_BYTES       ((PGS_REC_ROW_OFF + 4 + 15U) & ~15U)
/* cursor of partition p: the cursors may be spread out (1 << part_pad words
 * apart) so that neighbours do not share the 32-byte sector L2 atomics work on */
#define PGS_PART_CURSOR(gs,p)   ((gs).part_cursor + ((cl_ulong)(p) << (gs).part_pad))

struct kern_rec_gmem
{
    const unsigned char *rec;

    template <typename T>
    __device__ __forceinline__ bool
    fetch(int slot, cl_uint rowidx, T &out) const
    {
        out = *((const T *)(rec + pgs_rec_val_off(slot)));
        return ((*((const cl_uint *)(rec + PGS_REC_MASK_OFF)) >> slot) & 1U) != 0;
    }
};
/* the same record held in registers (gpupreagg_partagg reads the records one
 * batch ahead of the one it works on) */
template <typename T, int SIZE> struct pgs_rec_bits;
template <typename T> struct pgs_rec_bits<T, 8>
{
    static __device__ __forceinline__ T
    get(const cl_uint *w, cl_uint off)
    {
        union { cl_ulong u; T t; } x;
        x.u = ((cl_ulong)w[off / 4 + 1] << 32) | w[off / 4];
        return x.t;
    }
};
template <typename T> struct pgs_rec_bits<T, 4>
{
    static __device__ __forceinline__ T
    get(const cl_uint *w, cl_uint off)
    {
        union { cl_uint u; T t; } x;
        x.u = w[off / 4];
        return x.t;
    }
};
template <typename T> struct pgs_rec_bits<T, 2>
{
    static __device__ __forceinline__ T
    get(const cl_uint *w, cl_uint off)
    {
        union { cl_ushort u; T t; } x;
        x.u = (cl_ushort)(w[off / 4] >> (8 * (off & 3U)));
        return x.t;
    }
};
template <typename T> struct pgs_rec_bits<T, 1>
{
    static __device__ __forceinline__ T
    get(const cl_uint *w, cl_uint off)
    {
        union { unsigned char u; T t; } x;
        x.u = (unsigned char)(w[off / 4] >> (8 * (off & 3U)));
        return x.t;
    }
};
struct kern_rec_regs
{
    cl_uint     w[PGS_REC_BYTES / 4];

    __device__ __forceinline__ void
    load(const unsigned char *rec)
    {
#pragma unroll
        for (int i = 0; i < (int)(PGS_REC_BYTES / 16); i++)
        {
            uint4   q = __ldcs((const uint4 *)rec + i);     /* read once: streaming */

            w[4 * i] = q.x; w[4 * i + 1] = q.y; w[4 * i + 2] = q.z; w[4 * i + 3] = q.w;
        }
    }
    __device__ __forceinline__ void
    clear(void)
    {
#pragma unroll
        for (int i = 0; i < (int)(PGS_REC_BYTES / 4); i++)
            w[i] = 0;
    }
    __device__ __forceinline__ cl_uint rownum(void) const { return w[PGS_REC_ROW_OFF / 4]; }

    template <typename T>
    __device__ __forceinline__ bool
    fetch(int slot, cl_uint rowidx, T &out) const
    {
        out = pgs_rec_bits<T, sizeof(T)>::get(w, pgs_rec_val_off(slot));
        return ((w[PGS_REC_MASK_OFF / 4] >> slot) & 1U) != 0;
    }
};
#define PGS_X_INCOL_RECPUT(slot,colidx,attlen)                          \
    {                                                                   \
        pgs_rowq_type<attlen>::T __v;                                   \
        bool __ok = kds.template fetch<pgs_rowq_type<attlen>::T>(slot, rowidx, __v); \
        pgs_rec_put<attlen>(__w, pgs_rec_val_off(slot), (cl_ulong)__v); \
        __mask |= (__ok ? (1U << (slot)) : 0U);                         \
    }
/* the record is put together in registers (every offset is a compile time
 * constant, so __w[] never leaves them) */
template <int ATTLEN>
DEVFN void
pgs_rec_put(cl_uint *w, cl_uint off, cl_ulong v)
{
    if (ATTLEN == 8)
    {
        w[off / 4] = (cl_uint)v;
        w[off / 4 + 1] = (cl_uint)(v >> 32);
    }
    else if (ATTLEN == 4)
        w[off / 4] = (cl_uint)v;
    else if (ATTLEN == 2)
        w[off / 4] |= ((cl_uint)v & 0xffffU) << (8 * (off & 3U));
    else
        w[off / 4] |= ((cl_uint)v & 0xffU) << (8 * (off & 3U));
}

/* deal one row into its partition: reserve a record (one atomic on the
 * partition's cursor), then write it.  A position at or beyond part_cap means
 * the partition takes no more records this chunk (the caller sends the row
 * to the global table). */
DEVFN cl_uint
pgs_part_reserve(const pgs_gstate &gs, cl_ulong hash, cl_uint &part)
{
    part = __umulhi((cl_uint)(hash >> 32), gs.part_nparts);
    return atomicAdd(PGS_PART_CURSOR(gs, part), 1U);
}
/* Segment mode: the cursors of this CTA's segments are 16-bit counters in
 * shared memory, two per word.  A counter is read before it is bumped and
 * left alone once it has reached the capacity, so it never runs into its
 * neighbour: at most one add per thread and row in flight can follow a read
 * that still saw room (capacity + 8 x block size < 65536, the host checks).
 * Returns the position in the segment; >= part_seg_cap: no room. */
DEVFN cl_uint
pgs_part_reserve_seg(const pgs_gstate &gs, cl_uint *segcur, cl_ulong hash, cl_uint &part)
{
    part = __umulhi((cl_uint)(hash >> 32), gs.part_nparts);
    const cl_uint   sh = (part & 1U) * 16U;
    cl_uint         c = (*((volatile cl_uint *)&segcur[part >> 1]) >> sh) & 0xffffU;

    if (c >= gs.part_seg_cap)
        return gs.part_seg_cap;
    return (atomicAdd(&segcur[part >> 1], 1U << sh) >> sh) & 0xffffU;
}
/* position of record `pos` of segment `seg` of a partition inside its area */
DEVFN cl_uint
pgs_part_seg_pos(const pgs_gstate &gs, cl_uint seg, cl_uint pos)
{
    return seg * gs.part_seg_cap + pos;
}
template <typename KDS>
DEVFN void
pgs_part_write(const pgs_gstate &gs, const KDS &kds, cl_uint rowidx,
               cl_uint rownum, cl_uint part, cl_uint pos)
{
    uint4          *__rec = (uint4 *)(gs.part_recs +
                                      ((cl_ulong)part * gs.part_cap + pos) * PGS_REC_BYTES);
    cl_uint         __w[PGS_REC_BYTES / 4];
    cl_uint         __mask = 0;

#pragma unroll
    for (int i = 0; i < (int)(PGS_REC_BYTES / 4); i++)
        __w[i] = 0;
    GPUPREAGG_INCOL_LIST(PGS_X_INCOL_RECPUT)
    __w[PGS_REC_MASK_OFF / 4] = __mask;
    __w[PGS_REC_ROW_OFF / 4] = rownum;
#if __CUDACC_VER_MAJOR__ > 12 || (__CUDACC_VER_MAJOR__ == 12 && __CUDACC_VER_MINOR__ >= 9)
    if ((PGS_REC_BYTES & 31U) == 0)
    {
        /* whole 32-byte sectors: one 256-bit store each (STG.256, sm_100;
         * PTX ISA 8.8, i.e. NVRTC 12.9 or later).
         * L2 works on requests, not bytes: four 8-byte stores per record made
         * the deal pass twice as slow as two 16-byte ones (measured) */
#pragma unroll
        for (int i = 0; i < (int)(PGS_REC_BYTES / 32); i++)
            asm volatile("st.global.v4.u64 [%0], {%1, %2, %3, %4};"
                         :: "l"((unsigned char *)__rec + 32 * i),
                            "l"(((cl_ulong)__w[8 * i + 1] << 32) | __w[8 * i]),
                            "l"(((cl_ulong)__w[8 * i + 3] << 32) | __w[8 * i + 2]),
                            "l"(((cl_ulong)__w[8 * i + 5] << 32) | __w[8 * i + 4]),
                            "l"(((cl_ulong)__w[8 * i + 7] << 32) | __w[8 * i + 6])
                         : "memory");
    }
    else
#endif
    {
#pragma unroll
        for (int i = 0; i < (int)(PGS_REC_BYTES / 16); i++)
            __rec[i] = make_uint4(__w[4 * i], __w[4 * i + 1], __w[4 * i + 2], __w[4 * i + 3]);
    }
}
template <typename KDS>
DEVFN bool
pgs_part_emit(const pgs_gstate &gs, const KDS &kds, cl_uint rowidx,
              cl_uint rownum, cl_ulong hash)
{
    cl_uint     part;
    cl_uint     pos = pgs_part_reserve(gs, hash, part);

    if (pos >= gs.part_cap)
        return false;
    pgs_part_write(gs, kds, rowidx, rownum, part, pos);
    return true;
}

/* ------------------------------------------------------------------
 * per-row body shared by the staged and the gather path
 * ------------------------------------------------------------------ */
struct pgs_row_ctx
{
    cl_uint     nfiltered;
    cl_uint     nrecheck;
    cl_uint     ninserted;      /* groups this thread added to the global table */
    cl_int      errcode;        /* first significant error seen */
};

template <typename KDS>
DEVFN bool
pgs_eval_row(const kern_parambuf *kparams, const KDS &kds, const void *ktoast,
             cl_uint row, cl_uint *recheck_map, pgs_row_ctx &ctx, pagg_row &prow)
{
    cl_int      errcode = StromError_Success;
    bool        valid;

    valid = gpupreagg_qual_eval(&errcode, kparams, kds, ktoast, row);
    if (valid)
    {
        gpupreagg_projection(&errcode, kparams, kds, prow, ktoast, row, 0);
        gpupreagg_aggcheck(&errcode, prow);
    }
    if (errcode != StromError_Success)
    {
        if (errcode == StromError_CpuReCheck)
        {
            atomicOr(&recheck_map[row >> 5], 1U << (row & 31));
            ctx.nrecheck++;
        }
        else if (ctx.errcode == StromError_Success)
            ctx.errcode = errcode;
        return false;
    }
    if (!valid)
        ctx.nfiltered++;
    return valid;
}

/* ------------------------------------------------------------------
 * block-level epilogue helpers
 * ------------------------------------------------------------------ */
DEVFN cl_uint
pgs_warp_sum(cl_uint v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
        v += __shfl_xor_sync(0xffffffffU, v, d);
    return v;
}

DEVFN void
pgs_writeback_status(kern_gpupreagg *kgpreagg, const pgs_gstate &gs,
                     const pgs_row_ctx &ctx)
{
    /* first significant error wins, else CpuReCheck
     * (kern_writeback_error_status, opencl_common.h:1481-1527) */
    cl_uint     nf = pgs_warp_sum(ctx.nfiltered);
    cl_uint     nr = pgs_warp_sum(ctx.nrecheck);
    cl_uint     ni = pgs_warp_sum(ctx.ninserted);
    cl_int      ec = ctx.errcode;

#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        cl_int o = __shfl_xor_sync(0xffffffffU, ec, d);
        if (ec == StromError_Success)
            ec = o;
    }
    if ((threadIdx.x & 31) == 0)
    {
        if (nf)
            atomicAdd((unsigned long long *)gs.nrows_filtered, (unsigned long long)nf);
        if (ni)
        {
            atomicAdd(gs.gh_ngroups, ni);
            atomicAdd(gs.gh_nused, ni);
        }
        if (ec != StromError_Success)
        {
            cl_int cur = atomicCAS(&kgpreagg->status, StromError_Success, ec);
            if (cur == StromError_CpuReCheck)
                atomicCAS(&kgpreagg->status, StromError_CpuReCheck, ec);
        }
        else if (nr)
            atomicCAS(&kgpreagg->status, StromError_Success, StromError_CpuReCheck);
    }
}


/* ------------------------------------------------------------------
 * the main kernel
 * ------------------------------------------------------------------ */
struct pgs_smem_head
{
    cl_ulong    full_bar[GPUPREAGG_MAX_STAGES];
    cl_ulong    empty_bar[GPUPREAGG_MAX_STAGES];
    cl_uint     val_pos[PGS_MAX(GPUPREAGG_NUM_INCOLS, 1)];  /* kern_colpos copy */
    cl_uint     nul_pos[PGS_MAX(GPUPREAGG_NUM_INCOLS, 1)];
    cl_uint     sh_nused;
    cl_uint     is_last_cta;
};

#define PGS_X_INCOL_LOADPOS(slot,colidx,attlen)                         \
    head->val_pos[slot] = KERN_DATA_STORE_COLPOS(kds_in, colidx)->values_offset; \
    head->nul_pos[slot] = KERN_DATA_STORE_COLPOS(kds_in, colidx)->nullmap_offset;

#define PGS_X_INCOL_ISSUE(slot,colidx,attlen)                           \
    if (GPUPREAGG_INCOL_STAGED(slot))                                   \
    {                                                                   \
        cl_uint nb = ((rows * (attlen)) + 15U) & ~15U;                  \
        pgs_bulk_g2s(stage_base + pgs_stage_val_off(slot, tile_rows),              \
                     (const char *)kds_in + head->val_pos[slot] +       \
                     (cl_ulong)row0 * (attlen), nb, &head->full_bar[stage]); \
        if (head->nul_pos[slot] != 0)                                   \
        {                                                               \
            cl_uint mb = (((rows + 7U) >> 3) + 15U) & ~15U;             \
            pgs_bulk_g2s(stage_base + pgs_stage_nul_off(slot, tile_rows),          \
                         (const char *)kds_in + head->nul_pos[slot] +   \
                         (row0 >> 3), mb, &head->full_bar[stage]);      \
        }                                                               \
    }

#define PGS_X_INCOL_TXBYTES(slot,colidx,attlen)                         \
    if (GPUPREAGG_INCOL_STAGED(slot))                                   \
    {                                                                   \
        txbytes += ((rows * (attlen)) + 15U) & ~15U;                    \
        if (head->nul_pos[slot] != 0)                                   \
            txbytes += (((rows + 7U) >> 3) + 15U) & ~15U;               \
    }

#define PGS_X_INCOL_VIEW(slot,colidx,attlen)                            \
    tile.val_off[slot] = stage_off + pgs_stage_val_off(slot, tile_rows);           \
    tile.nul_off[slot] = (head->nul_pos[slot] != 0 && GPUPREAGG_INCOL_STAGED(slot) \
                          ? stage_off + pgs_stage_nul_off(slot, tile_rows)         \
                          : KERN_TILE_NO_NULLMAP);

/* careful per-row path of the no-group kernel */
#define PGS_CONSUME_ROW(j)                                              \
    {                                                                   \
        pagg_row    prow;                                               \
        bool valid = pgs_eval_row(kparams, rr[j], kds_in, row0 + r + (j), \
                                  recheck_map, ctx, prow);              \
        gpupreagg_aggcalc_thread(acc, prow, valid, nnflag);             \
    }
/* GROUP BY without a WHERE clause: every lane of the warp comes here for
 * row j of its batch (pgs_group_add_row is warp-collective) */
#define PGS_CONSUME_ROW_GROUPED(j)                                      \
    {                                                                   \
        pagg_row    prow;                                               \
        bool valid = false;                                             \
        if (r + (j) < rows)                                             \
            valid = pgs_eval_row(kparams, rr[j], kds_in, row0 + r + (j), \
                                 recheck_map, ctx, prow);               \
        pgs_group_add_row(gs, sh, &head->sh_nused, prow, ctx, valid,    \
                          row0 + r + (j), recheck_map);                 \
    }

/* phase 1 of the staged consumer loop: 4 adjacent rows of one column */
#define PGS_X_INCOL_LOAD4(slot,colidx,attlen)                           \
    if (GPUPREAGG_INCOL_STAGED(slot))                                   \
        pgs_rowload<attlen>::load4(__pgs_smem + tile.val_off[slot] +    \
                                   r * (attlen),                        \
                                   rr[0].v[slot], rr[1].v[slot],        \
                                   rr[2].v[slot], rr[3].v[slot]);       \
    else                                                                \
        rr[0].v[slot] = rr[1].v[slot] = rr[2].v[slot] = rr[3].v[slot] = 0; \
    {                                                                   \
        cl_uint __vb = 0xFU;                                            \
        if (tile.nul_off[slot] != KERN_TILE_NO_NULLMAP)                 \
            __vb = *((const cl_uint *)(__pgs_smem + tile.nul_off[slot]) + (r >> 5)) >> (r & 31); \
        rr[0].vbits[slot] = rr[1].vbits[slot] = __vb;                   \
        rr[2].vbits[slot] = rr[3].vbits[slot] = __vb;                   \
    }

#define PGS_X_INCOL_GVIEW(slot,colidx,attlen)                           \
    gtile.val_ptr[slot] = (const char *)kds_in +                        \
        KERN_DATA_STORE_COLPOS(kds_in, colidx)->values_offset;          \
    gtile.nul_ptr[slot] =                                               \
        (KERN_DATA_STORE_COLPOS(kds_in, colidx)->nullmap_offset != 0    \
         ? (const cl_uint *)((const char *)kds_in +                     \
               KERN_DATA_STORE_COLPOS(kds_in, colidx)->nullmap_offset)  \
         : (const cl_uint *)NULL);

/* one row of every column, by row number, raw into registers (the validity
 * word is kept whole: bit `shift` is this row's) */
#define PGS_X_INCOL_GLOAD(slot,colidx,attlen)                           \
    nx.v[slot] = (cl_ulong)__ldg((const pgs_rowq_type<attlen>::T *)gtile.val_ptr[slot] + nrow); \
    nx.vbits[slot] = (gtile.nul_ptr[slot]                               \
                      ? __ldg(gtile.nul_ptr[slot] + (nrow >> 5)) : 0xffffffffU);

/*
 * Add one projected row per lane to the group state (CTA-local table first,
 * global table when that is full or absent).
 *
 * Warp-collective: all 32 lanes call it, `active` says whether the lane has
 * a row.  The probe loops end at a different iteration for every lane; the
 * __syncwarp() behind each of them brings the lanes back together so that
 * the chain of cell updates runs once for the warp, not once per group of
 * lanes that happened to leave the loop together.
 */
DEVFN void
pgs_group_add_row(const pgs_gstate &gs, const pgs_sh_table &sh,
                  cl_uint *sh_nused, const pagg_row &prow, pgs_row_ctx &ctx,
                  bool active, cl_uint rownum, cl_uint *recheck_map)
{
    cl_ulong    keyvals[PGS_MAX(GPUPREAGG_NUM_KEYS, 1)];
    cl_uint     knull = 0;
    cl_ulong    hash;
    cl_uint     s = ~0U;
    cl_ulong   *gslot = NULL;

#pragma unroll
    for (int k = 0; k < GPUPREAGG_NUM_KEYS; k++)
    {
        /* NULL keys form one group (gpupreagg.c:1234-1243) */
        keyvals[k] = (!active || prow.key[k].isnull) ? 0 : prow.key[k].ulong_val;
        if (active && prow.key[k].isnull)
            knull |= (1U << k);
    }
    hash = pgs_hash_keyvals(keyvals, knull);
#if GPUPREAGG_DEBUG_LEVEL == 1      /* timing experiment: stop after the hash */
    ctx.nfiltered += (cl_uint)(hash & 1);
    return;
#endif
    if (sh.nslots > 0)
    {
        if (active)
            s = pgs_sh_find_slot(sh, sh_nused, keyvals, knull, hash);
        __syncwarp();
    }
#if GPUPREAGG_DEBUG_LEVEL == 2      /* timing experiment: stop after the probe */
    ctx.nfiltered += (s & 1);
    return;
#endif
    if (__any_sync(0xffffffffU, active && s == ~0U))
    {
        if (active && s == ~0U)
            gslot = pgs_gh_find_slot(gs, keyvals, knull, hash, ctx.ninserted);
        __syncwarp();
    }
    if (active)
    {
        if (s != ~0U)
        {
            /* (a per-slot lock with plain loads and stores under it was
             * tried instead of these atomics: lanes of a warp that meet in
             * one group then take turns through the whole update, measured
             * 13% slower) */
            pgs_sh_cells cells;
            cl_uint      nn;

            cells.p0 = sh.cell(0, s);
            cells.stride = sh.nslots;
            nn = gpupreagg_aggcalc_shared(cells, prow);
            if ((*((volatile cl_uint *)sh.ctrl_hi(s)) & nn) != nn)
                atomicOr(sh.ctrl_hi(s), nn);
        }
        else if (gslot)
        {
            cl_ulong   *cells = gslot + 1 + GPUPREAGG_NUM_KEYS;
            cl_uint    *p_nn = (cl_uint *)gslot + 1;
            cl_uint     nn = gpupreagg_aggcalc_atomic(cells, prow);

            if ((*((volatile cl_uint *)p_nn) & nn) != nn)
                atomicOr(p_nn, nn);
        }
        else
        {
            /* no room in the tables (the planner's estimate of the number of
             * groups was too low and the table has not been enlarged yet):
             * the row becomes a state record of its own in the overflow log
             * (the host folds the log into a larger table after the chunk);
             * with the log full as well it is left to the host like any
             * other row the device cannot finish - the reference reduces
             * every chunk on its own and never fails on a wrong ndistinct,
             * neither may this */
            cl_ulong    one[PGS_MAX(GPUPREAGG_NUM_CELLS, 1)];
            cl_uint     nn;

            pgs_cells_init(one);
            nn = gpupreagg_aggcalc_plain(one, prow, true);
            if (!pgs_ovf_append(gs, keyvals, knull, one, nn))
            {
                atomicOr(&recheck_map[rownum >> 5], 1U << (rownum & 31));
                ctx.nrecheck++;
            }
        }
    }
    __syncwarp();
}

/* a row whose qual / projection raised an error: re-check rows are flagged in
 * the chunk's bitmap, the first significant error is kept */
DEVFN void
pgs_note_error(cl_int errcode, cl_uint row, cl_uint *recheck_map, pgs_row_ctx &ctx)
{
    if (errcode == StromError_CpuReCheck)
    {
        atomicOr(&recheck_map[row >> 5], 1U << (row & 31));
        ctx.nrecheck++;
    }
    else if (ctx.errcode == StromError_Success)
        ctx.errcode = errcode;
}

/* threads of the CTA -> thread 0, always combined in the same order: lanes by
 * a butterfly where the lower lane merges, then warps in index order.
 * `scratch`: shared memory nobody else uses any more, 8*(1+NCELLS)*nwarps. */
DEVFN void
pgs_block_reduce(cl_ulong *acc, cl_uint &acc_nn, cl_ulong *scratch)
{
    const cl_uint   warp_id = threadIdx.x >> 5;
    const cl_uint   lane_id = threadIdx.x & 31;
    const cl_uint   nwarps = blockDim.x >> 5;
    const int       W = 1 + GPUPREAGG_NUM_CELLS;

#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        cl_ulong    src[PGS_MAX(GPUPREAGG_NUM_CELLS, 1)];
        cl_uint     src_nn = __shfl_xor_sync(0xffffffffU, acc_nn, d);
#pragma unroll
        for (int c = 0; c < GPUPREAGG_NUM_CELLS; c++)
            src[c] = __shfl_xor_sync(0xffffffffU, acc[c], d);
        if ((lane_id & d) == 0)
        {
            gpupreagg_aggmerge_plain(acc, src, src_nn);
            acc_nn |= src_nn;
        }
    }
    __syncthreads();        /* scratch is free: every tile was consumed */
    if (lane_id == 0)
    {
        scratch[warp_id * W] = acc_nn;
#pragma unroll
        for (int c = 0; c < GPUPREAGG_NUM_CELLS; c++)
            scratch[warp_id * W + 1 + c] = acc[c];
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (cl_uint w = 1; w < nwarps; w++)
        {
            cl_uint src_nn = (cl_uint)scratch[w * W];
            gpupreagg_aggmerge_plain(acc, scratch + w * W + 1, src_nn);
            acc_nn |= src_nn;
        }
    }
    __syncthreads();
}

/* epilogue of both main kernels.  `scratch` is shared memory nobody else
 * uses any more, at least 8 * (1 + NCELLS) * nwarps bytes. */
DEVFN void
pgs_main_epilogue(kern_gpupreagg *kgpreagg, const pgs_gstate &gs,
                  const pgs_sh_table &sh, cl_ulong *scratch,
                  cl_uint *p_is_last, cl_ulong *acc, cl_uint acc_nn,
                  pgs_row_ctx &ctx)
{
    const int       W = 1 + GPUPREAGG_NUM_CELLS;

    if (GPUPREAGG_NUM_KEYS == 0)
    {
        cl_ulong   *mine = gs.ng_partial + (cl_ulong)blockIdx.x * W;

        /* threads -> CTA state (thread 0) */
        pgs_block_reduce(acc, acc_nn, scratch);
        if (threadIdx.x == 0)
        {
            cl_uint     ticket;

            mine[0] = acc_nn;
#pragma unroll
            for (int c = 0; c < GPUPREAGG_NUM_CELLS; c++)
                mine[1 + c] = acc[c];
            __threadfence();
            ticket = atomicAdd(gs.ng_ticket, 1U);
            *p_is_last = (ticket == gridDim.x - 1 ? 1U : 0U);
        }
        __syncthreads();
        if (*p_is_last)
        {
            /* last CTA: fold the CTA partials of this launch into the
             * persistent state row.  Thread t takes partials t, t+B, ... in
             * order and the block reduction has a fixed shape, so the result
             * does not depend on which CTA happened to finish last. */
            __threadfence();
            pgs_cells_init(acc);
            acc_nn = 0;
            for (cl_uint b = threadIdx.x; b < gridDim.x; b += blockDim.x)
            {
                const cl_ulong *part = gs.ng_partial + (cl_ulong)b * W;
                cl_uint     src_nn = (cl_uint)__ldcg(part);
                cl_ulong    src[PGS_MAX(GPUPREAGG_NUM_CELLS, 1)];
#pragma unroll
                for (int c = 0; c < GPUPREAGG_NUM_CELLS; c++)
                    src[c] = __ldcg(part + 1 + c);
                gpupreagg_aggmerge_plain(acc, src, src_nn);
                acc_nn |= src_nn;
            }
            __syncthreads();
            pgs_block_reduce(acc, acc_nn, scratch);
            if (threadIdx.x == 0)
            {
                cl_ulong   *state = gs.ng_state;

                gpupreagg_aggmerge_plain(state + 1, acc, acc_nn);
                state[0] = (cl_uint)state[0] | acc_nn;
                *gs.ng_ticket = 0;
            }
        }
    }
    else if (sh.nslots > 0)
    {
        /* spill the CTA-local table into the global one */
        /* every CTA starts somewhere else in its table: the tables hold the
         * same groups, and CTAs that finish together would otherwise queue
         * up on the same global slots in the same order */
        __syncthreads();
        for (cl_uint s0 = threadIdx.x; s0 < sh.nslots; s0 += blockDim.x)
        {
            cl_uint     s = (s0 + blockIdx.x * 37U) % sh.nslots;
            cl_uint     st = *sh.ctrl_lo(s);
            cl_ulong    keyvals[PGS_MAX(GPUPREAGG_NUM_KEYS, 1)];
            cl_ulong    src[PGS_MAX(GPUPREAGG_NUM_CELLS, 1)];
            cl_uint     knull;

            if ((st & 3U) != PGS_SLOT_READY)
                continue;
            knull = (st >> 2) & ((1U << GPUPREAGG_NUM_KEYS) - 1U);
#pragma unroll
            for (int k = 0; k < GPUPREAGG_NUM_KEYS; k++)
                keyvals[k] = *sh.key(k, s);
#pragma unroll
            for (int c = 0; c < GPUPREAGG_NUM_CELLS; c++)
                src[c] = *sh.cell(c, s);
            if (!pgs_gh_merge_state(gs, keyvals, knull,
                                    pgs_hash_keyvals(keyvals, knull),
                                    src, *sh.ctrl_hi(s), ctx.ninserted))
            {
                if (ctx.errcode == StromError_Success)
                    ctx.errcode = StromError_DataStoreNoSpace;
            }
        }
    }
    pgs_writeback_status(kgpreagg, gs, ctx);
}

#define PGS_SH_TABLE_INIT()                                             \
    if (sh.nslots > 0)                                                  \
    {                                                                   \
        for (cl_uint s = threadIdx.x; s < sh.nslots; s += blockDim.x)   \
        {                                                               \
            pgs_sh_cells cells;                                         \
            *sh.ctrl_lo(s) = PGS_SLOT_EMPTY;                            \
            *sh.ctrl_hi(s) = 0;                                         \
            cells.p0 = sh.cell(0, s);                                   \
            cells.stride = sh.nslots;                                   \
            pgs_cells_init(cells);                                      \
        }                                                               \
    }

/*
 * gpupreagg_main - staged (TMA) path over a KDS_FORMAT_COLUMN chunk.
 * grid = persistent CTAs (multiple of the SM count), block = 1 producer
 * warp + GPUPREAGG_CONSUMER_WARPS consumer warps.
 */
extern "C" __global__ void
__launch_bounds__(GPUPREAGG_BLOCK_THREADS, GPUPREAGG_MIN_CTAS)
gpupreagg_main(kern_gpupreagg *kgpreagg,
               const kern_data_store *kds_in,
               pgs_gstate gs,
               cl_uint *recheck_map,
               cl_uint sh_nslots,
               cl_uint tile_rows,
               cl_uint nstages)
{
    pgs_smem_head  *head = (pgs_smem_head *)__pgs_smem;
    unsigned char  *stages = __pgs_smem + PGS_SMEM_HEAD_BYTES;
    const kern_parambuf *kparams = KERN_GPUPREAGG_PARAMBUF_CONST;
    const cl_uint   nrows = kds_in->nitems;
    const cl_uint   ntiles = (nrows + tile_rows - 1) / tile_rows;
    const cl_uint   warp_id = threadIdx.x >> 5;
    const cl_uint   lane_id = threadIdx.x & 31;
    cl_ulong        acc[PGS_MAX(GPUPREAGG_NUM_CELLS, 1)];
    cl_uint         acc_nn = 0;
    pgs_row_ctx     ctx;
    pgs_sh_table    sh;

    ctx.nfiltered = 0;
    ctx.nrecheck = 0;
    ctx.ninserted = 0;
    ctx.errcode = StromError_Success;
    sh.base = PGS_SMEM_HEAD_BYTES + nstages * PGS_STAGE_BYTES(tile_rows);
    sh.nslots = (GPUPREAGG_NUM_KEYS > 0 ? sh_nslots : 0);
    sh.salt = 1;
    pgs_cells_init(acc);

    if (threadIdx.x == 0)
    {
        for (cl_uint s = 0; s < nstages; s++)
        {
            pgs_mbar_init(&head->full_bar[s], 1);
            pgs_mbar_init(&head->empty_bar[s], GPUPREAGG_CONSUMER_WARPS);
        }
        GPUPREAGG_INCOL_LIST(PGS_X_INCOL_LOADPOS)
        head->sh_nused = 0;
        head->is_last_cta = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    PGS_SH_TABLE_INIT()
#if GPUPREAGG_PARTITIONED
    /* segment mode of the deal pass: this CTA's cursors, behind the ring */
    if (gs.part_nparts != 0 && gs.part_seg_cap != 0)
        for (cl_uint i = threadIdx.x; i < (gs.part_nparts + 1) / 2; i += blockDim.x)
            ((cl_uint *)(__pgs_smem + sh.base))[i] = 0;
#endif
    __syncthreads();

    if (warp_id == 0)
    {
        /* ===== producer warp: lane 0 feeds the ring ===== */
        cl_uint it = 0;
        for (cl_uint t = blockIdx.x; t < ntiles; t += gridDim.x, it++)
        {
            if (lane_id == 0)
            {
                cl_uint stage = it % nstages;
                cl_uint phase = (it / nstages) & 1;
                cl_uint row0 = t * tile_rows;
                cl_uint rows = min(tile_rows, nrows - row0);
                unsigned char *stage_base = stages + stage * PGS_STAGE_BYTES(tile_rows);
                cl_uint txbytes = 0;

                pgs_mbar_wait_relaxed(&head->empty_bar[stage], phase ^ 1);
                GPUPREAGG_INCOL_LIST(PGS_X_INCOL_TXBYTES)
                pgs_mbar_arrive_expect_tx(&head->full_bar[stage], txbytes);
                GPUPREAGG_INCOL_LIST(PGS_X_INCOL_ISSUE)
                (void)stage_base;
            }
            __syncwarp();
        }
    }
    else
    {
        /* ===== consumer warps ===== */
        const cl_uint   ctid = threadIdx.x - 32;
        cl_uint         it = 0;
#if GPUPREAGG_NUM_KEYS == 0
        pgs_tacc        tacc;
        const cl_int    one = (cl_int)(nstages != 0);   /* 1, opaque to the compiler */

        tacc.init();
#elif GPUPREAGG_HAS_QUAL
        /* the warp's row queue (see PGS_ROWQ_*): free-running positions,
         * warp-uniform; npassed counts the rows its qual let through */
        const cl_uint   __qbase = PGS_SMEM_HEAD_FIXED + (warp_id - 1) * PGS_ROWQ_WARP_BYTES;
        cl_ushort      *rowq = (cl_ushort *)(__pgs_smem + __qbase);
        const cl_uint   lanes_lt = (1U << lane_id) - 1U;
        cl_uint         qhead = 0, qtail = 0;
        cl_uint         nscanned = 0, npassed = 0;
#if GPUPREAGG_GATHER_PAYLOAD
        kern_tile_gmem  gtile;
        /* the batch of queued rows whose columns are on their way from HBM */
        kern_row_regs   g_row;
        cl_uint         g_rownum = 0;
        bool            g_have = false, g_act = false;

        GPUPREAGG_INCOL_LIST(PGS_X_INCOL_GVIEW)
#endif
#endif
        cl_uint         stage = 0, phase = 0;
        PGS_DBG_DECL

        for (cl_uint t = blockIdx.x; t < ntiles;
             t += gridDim.x, it++, stage++)
        {
            if (stage == nstages)
            {
                stage = 0;
                phase ^= 1;
            }
            cl_uint row0 = t * tile_rows;
            cl_uint rows = min(tile_rows, nrows - row0);
            cl_uint stage_off = PGS_SMEM_HEAD_BYTES + stage * PGS_STAGE_BYTES(tile_rows);
            kern_tile_smem tile;

            tile.row0 = row0;
            GPUPREAGG_INCOL_LIST(PGS_X_INCOL_VIEW)
            (void)stage_off;
            PGS_DBG_START()
            pgs_mbar_wait(&head->full_bar[stage], phase);
            PGS_DBG_STOP(0)
            PGS_DBG_START()

            /* each thread takes 4 consecutive rows: phase 1 pulls them out of
             * the stage with 128-bit shared memory loads, phase 2 evaluates
             * them from registers */
#if GPUPREAGG_NUM_KEYS == 0
            for (cl_uint r = ctid * 4; r < rows; r += GPUPREAGG_CONSUMER_THREADS * 4)
            {
                kern_row_regs rr[4];
                pagg_row    prow4[4];
                bool        valid4[4];
                bool        fast = (r + 4 <= rows);

                rr[0].shift = 0; rr[1].shift = 1; rr[2].shift = 2; rr[3].shift = 3;
                GPUPREAGG_INCOL_LIST(PGS_X_INCOL_LOAD4)
                if (fast)
                {
                    /* qual and projection of all 4 rows, straight-line; any
                     * error or special float sends the batch down the
                     * careful path below, which evaluates it again */
                    cl_int      berr = StromError_Success;
                    cl_uint     md = 0, mf = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++)
                    {
                        cl_int  e1 = StromError_Success;
                        cl_int  e2 = StromError_Success;

                        valid4[j] = gpupreagg_qual_eval(&e1, kparams, rr[j], kds_in,
                                                        row0 + r + j);
                        gpupreagg_projection(&e2, kparams, rr[j], prow4[j], kds_in,
                                             row0 + r + j, 0);
                        berr |= e1 | (valid4[j] ? e2 : StromError_Success);
                        pgs_row_special(prow4[j], md, mf);
                    }
                    fast = (berr == StromError_Success) &
                           (md < PGS_SPECIAL_F8_LIMIT) & (mf < PGS_SPECIAL_F4_LIMIT);
                }
                if (fast)
                {
                    tacc.calc4(prow4, valid4, one);
                    tacc.note_nonnull(prow4, valid4);
#if GPUPREAGG_HAS_QUAL
                    ctx.nfiltered += PGS_K01(!valid4[0]) + PGS_K01(!valid4[1]) +
                                     PGS_K01(!valid4[2]) + PGS_K01(!valid4[3]);
#endif
                }
                else
                {
                    bool    nnflag[PGS_MAX(PGS_NUM_NNCLASSES, 1)];
#pragma unroll
                    for (int k = 0; k < PGS_MAX(PGS_NUM_NNCLASSES, 1); k++)
                        nnflag[k] = false;
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (r + j < rows)
                            PGS_CONSUME_ROW(j)
                    acc_nn |= pgs_nnflags_to_mask(nnflag);
                }
            }
#elif !GPUPREAGG_HAS_QUAL
            /* GROUP BY, every row takes part.
             * Without a CTA-local table (many groups) every row goes to its
             * slot of the global table in HBM: a chain of dependent accesses
             * (state word, keys, atomics) whose first one misses L2 - the
             * warp would sit out a DRAM round trip per row (measured: 4.7%
             * issue slots used, 13 ms per 50 M rows).  So the warp first asks
             * L2 for the home slot of every row it owns in this tile
             * (prefetch.global.L2, nothing waits for it) and only then works
             * through the rows. */
            if (sh.nslots == 0 && PGS_PART_NPARTS(gs) == 0)
            {
                for (cl_uint rb = (ctid & ~31U) * 4; rb < rows;
                     rb += GPUPREAGG_CONSUMER_THREADS * 4)
                {
                    const cl_uint r = rb + lane_id * 4;
                    kern_row_regs rr[4];

                    rr[0].shift = 0; rr[1].shift = 1; rr[2].shift = 2; rr[3].shift = 3;
                    GPUPREAGG_INCOL_LIST(PGS_X_INCOL_LOAD4)
#pragma unroll
                    for (int j = 0; j < 4; j++)
                    {
                        if (r + j < rows)
                        {
                            pagg_row    prow;
                            cl_int      e = StromError_Success;
                            cl_uint     knull;
                            cl_ulong    hash;
                            const cl_ulong *home;

                            gpupreagg_projection(&e, kparams, rr[j], prow, kds_in,
                                                 row0 + r + j, 0);
                            hash = pgs_hash_keys(prow, knull);
                            home = gs.gh_slots +
                                (cl_ulong)((cl_uint)hash & (gs.gh_nslots - 1)) * PGS_SLOT_STRIDE;
                            asm volatile("prefetch.global.L2 [%0];" :: "l"(home));
                            if (8 * PGS_SLOT_STRIDE > 64)
                                asm volatile("prefetch.global.L2 [%0];" :: "l"(home + 8));
                        }
                    }
                }
            }
#if GPUPREAGG_PARTITIONED
            if (gs.part_nparts != 0)
            {
                /* very many groups: deal the rows into their partitions.
                 * What a row costs here is the round trip of the atomic on
                 * its partition's cursor (half of them to the far L2: ncu
                 * shows 50% lookup misses on 100 M atomics, 60% of all warp
                 * samples waiting for one), so a lane takes
                 * GPUPREAGG_DEAL_STEPS x 4 rows at a time and has all their
                 * atomics under way before the first record is written; the
                 * global table only takes what a full partition refuses */
                const cl_uint   step_rows = GPUPREAGG_CONSUMER_THREADS * 4;
                const bool      seg_mode = (gs.part_seg_cap != 0);
                cl_uint        *segcur = (cl_uint *)(__pgs_smem + sh.base);

                for (cl_uint rb = (ctid & ~31U) * 4; rb < rows;
                     rb += GPUPREAGG_DEAL_STEPS * step_rows)
                {
                    bool            validd[GPUPREAGG_DEAL_STEPS][4];
                    cl_uint         partd[GPUPREAGG_DEAL_STEPS][4];
                    cl_uint         posd[GPUPREAGG_DEAL_STEPS][4];

#pragma unroll
                    for (int h = 0; h < GPUPREAGG_DEAL_STEPS; h++)
                    {
                        /* (warp-uniform: rows past the end of the tile are
                         * not even read - the stage ends there) */
                        const bool      have = (h == 0 || rb + h * step_rows < rows);
                        const cl_uint   r = rb + h * step_rows + lane_id * 4;
                        kern_row_regs   rr[4];

                        rr[0].shift = 0; rr[1].shift = 1; rr[2].shift = 2; rr[3].shift = 3;
                        if (have)
                        {
                            GPUPREAGG_INCOL_LIST(PGS_X_INCOL_LOAD4)
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++)
                        {
                            pagg_row    prow;
                            cl_uint     knull;

                            validd[h][j] = false;
                            partd[h][j] = 0;
                            posd[h][j] = 0;
                            if (have && r + j < rows)
                                validd[h][j] = pgs_eval_row(kparams, rr[j], kds_in, row0 + r + j,
                                                            recheck_map, ctx, prow);
                            if (validd[h][j])
                            {
                                if (seg_mode)
                                {
                                    /* this CTA's own segment of the partition,
                                     * cursor in shared memory */
                                    const cl_uint   q = pgs_part_reserve_seg(
                                        gs, segcur, pgs_hash_keys(prow, knull), partd[h][j]);

                                    posd[h][j] = (q >= gs.part_seg_cap ? gs.part_cap
                                                  : pgs_part_seg_pos(gs, blockIdx.x, q));
                                }
                                else
                                    posd[h][j] = pgs_part_reserve(gs, pgs_hash_keys(prow, knull),
                                                                  partd[h][j]);
                            }
                        }
                    }
                    /* the rows are read again from the stage (it is still
                     * ours) instead of being kept in registers while the
                     * atomics are under way: 8 rows x 3 columns would not fit */
                    asm volatile("" ::: "memory");
#pragma unroll
                    for (int h = 0; h < GPUPREAGG_DEAL_STEPS; h++)
                    {
                        const bool      have = (h == 0 || rb + h * step_rows < rows);
                        const cl_uint   r = rb + h * step_rows + lane_id * 4;
                        kern_row_regs   rr[4];

                        rr[0].shift = 0; rr[1].shift = 1; rr[2].shift = 2; rr[3].shift = 3;
                        if (have)
                        {
                            GPUPREAGG_INCOL_LIST(PGS_X_INCOL_LOAD4)
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++)
                        {
                            bool    refused = (validd[h][j] && posd[h][j] >= gs.part_cap);

                            if (validd[h][j] && !refused)
                                pgs_part_write(gs, rr[j], row0 + r + j, row0 + r + j,
                                               partd[h][j], posd[h][j]);
                            if (__any_sync(0xffffffffU, refused))
                            {
                                pagg_row    prow;

                                if (refused)
                                {
                                    cl_int  e = StromError_Success;     /* evaluated above */
                                    gpupreagg_projection(&e, kparams, rr[j], prow, kds_in,
                                                         row0 + r + j, 0);
                                }
                                pgs_group_add_row(gs, sh, &head->sh_nused, prow, ctx, refused,
                                                  row0 + r + j, recheck_map);
                            }
                        }
                    }
                }
            }
            else
#endif
            /* The trip count is the same for every lane of the warp
             * (pgs_group_add_row is collective). */
            for (cl_uint rb = (ctid & ~31U) * 4; rb < rows;
                 rb += GPUPREAGG_CONSUMER_THREADS * 4)
            {
                const cl_uint r = rb + lane_id * 4;
                kern_row_regs rr[4];

                rr[0].shift = 0; rr[1].shift = 1; rr[2].shift = 2; rr[3].shift = 3;
                GPUPREAGG_INCOL_LIST(PGS_X_INCOL_LOAD4)
#pragma unroll
                for (int j = 0; j < 4; j++)
                    PGS_CONSUME_ROW_GROUPED(j)
            }
#elif GPUPREAGG_GATHER_PAYLOAD
            /* GROUP BY under a WHERE clause: the ring only carries the columns
             * the qual reads (GPUPREAGG_GATHER_PAYLOAD, the default), so a
             * tile of the same byte size holds several times the rows and
             * more of the chunk is in flight per SM.  Per tile a warp
             *   A. evaluates the qual of its whole share of the tile (up to
             *      PGS_GATHER_MAX_STEPS steps of 128 rows, 4 adjacent rows
             *      per lane and step): independent 128-bit shared-memory
             *      loads and compares, one bit per row in `vm`; nothing of
             *      the stage is needed after that, it goes back to the
             *      producer at once;
             *   B. step by step turns the bits into queue entries - the ROW
             *      NUMBERS of the survivors, compacted with ballots - and,
             *      whenever 32 rows are queued, every lane takes one through
             *      projection - its columns gathered from HBM by row number
             *      (kern_tile_gmem, the view of the row-map kernel) - and
             *      find-or-insert.
             * (A and B used to be one loop per step: load -> qual -> ballot
             * -> push, one dependent chain of ~1400 cycles per 128 rows with
             * four warps per scheduler to hide it - at 1% selectivity, with
             * next to nothing reaching the table, the kernel still needed
             * 0.30 ms per 125 M rows.)  With a selective qual most 32-byte
             * sectors of the unstaged columns are never read. */
            {
                cl_uint        *rowq32 = (cl_uint *)(__pgs_smem + __qbase);
                const cl_uint   step_rows = GPUPREAGG_CONSUMER_THREADS * 4;
                const cl_uint   rows_up = ((rows + step_rows - 1) / step_rows) * step_rows;
                const bool      last_tile = (t + gridDim.x >= ntiles);
                const cl_uint   rb0 = (ctid & ~31U) * 4;
                cl_uint         rb = rb0;
                cl_uint         vm = 0;     /* bit 4 k + j: row j of step k passed */
                bool            scanning = true;

                /* ---- A ---- */
#define PGS_GATHER_SCAN_QUALS(NV)                                               \
                        _Pragma("unroll")                                       \
                        for (int j = 0; j < 4; j++)                             \
                        {                                                       \
                            cl_int      e = StromError_Success;                 \
                            bool        v = false;                              \
                                                                                \
                            if ((cl_uint)j < (NV))                              \
                                v = gpupreagg_qual_eval(&e, kparams, rr[j], kds_in, \
                                                        row0 + r + j);          \
                            if (e != StromError_Success)                        \
                            {                                                   \
                                pgs_note_error(e, row0 + r + j, recheck_map, ctx); \
                                ctx.nfiltered--;    /* neither passed nor filtered */ \
                                v = false;                                      \
                            }                                                   \
                            vmask |= (v ? (1U << j) : 0U);                      \
                        }
                if (rows == tile_rows && tile_rows == 4 * step_rows)
                {
                    /* a whole tile of four steps (the default shape): one
                     * straight line, the four loads and sixteen compares
                     * are independent of each other */
#pragma unroll
                    for (int k = 0; k < 4; k++)
                    {
                        const cl_uint r = rb0 + (cl_uint)k * step_rows + lane_id * 4;
                        kern_row_regs rr[4];
                        cl_uint     vmask = 0;

                        rr[0].shift = 0; rr[1].shift = 1; rr[2].shift = 2; rr[3].shift = 3;
                        GPUPREAGG_INCOL_LIST(PGS_X_INCOL_LOAD4)
                        PGS_GATHER_SCAN_QUALS(4U)
                        vm |= vmask << (4 * k);
                    }
                    nscanned += 4 * 128U;
                }
                else
                {
#pragma unroll 1
                    for (int k = 0; k < PGS_GATHER_MAX_STEPS; k++)
                    {
                        const cl_uint   rbk = rb0 + (cl_uint)k * step_rows;
                        const cl_uint   r = rbk + lane_id * 4;
                        const cl_uint   nv = (r < rows ? min(rows - r, 4U) : 0U);
                        kern_row_regs   rr[4];
                        cl_uint         vmask = 0;

                        if (rbk >= rows)        /* warp-uniform */
                            break;
                        rr[0].shift = 0; rr[1].shift = 1; rr[2].shift = 2; rr[3].shift = 3;
                        GPUPREAGG_INCOL_LIST(PGS_X_INCOL_LOAD4)
                        PGS_GATHER_SCAN_QUALS(nv)
                        vm |= vmask << (4 * k);
                        nscanned += min(rows - rbk, 128U);
                    }
                }
#undef PGS_GATHER_SCAN_QUALS
                /* nothing of the stage is needed any more */
                __syncwarp();
                if (lane_id == 0)
                    pgs_mbar_arrive(&head->empty_bar[stage]);
                PGS_DBG_STOP(1)
                PGS_DBG_START()

                /* ---- B ---- */
                {
                    /* all survivors of the warp's share at once when the queue
                     * has room for them (it has, unless most rows pass): one
                     * prefix sum over the lanes' counts instead of four ballots
                     * and eight population counts per step (the per-step loop
                     * below stays for the rest: 0.54 k cycles per step and
                     * warp, 18% of the kernel at 10% selectivity) */
                    const cl_uint   cnt = __popc(vm);
                    cl_uint         incl = cnt;

#pragma unroll
                    for (int d = 1; d < 32; d <<= 1)
                    {
                        const cl_uint v = __shfl_up_sync(0xffffffffU, incl, d);

                        if (lane_id >= (cl_uint)d)
                            incl += v;
                    }
                    const cl_uint   total = __shfl_sync(0xffffffffU, incl, 31);

                    if (total + (qtail - qhead) <= PGS_ROWQ_ENTRIES)
                    {
                        cl_uint     pos = qtail + incl - cnt;
                        cl_uint     m = vm;

                        while (m != 0)
                        {
                            const cl_uint b = __ffs(m) - 1;

                            m &= m - 1;
                            rowq32[pos & (PGS_ROWQ_ENTRIES - 1)] =
                                row0 + rb0 + (b >> 2) * step_rows + lane_id * 4 + (b & 3U);
                            pos++;
                        }
                        qtail += total;
                        rb = rows_up;
                        scanning = false;
                        __syncwarp();
                    }
                }
                for (;;)
                {
                    const cl_uint   qn = qtail - qhead;

                    if (scanning && rb < rows_up && qn <= PGS_ROWQ_ENTRIES - 128)
                    {
                        const cl_uint r = rb + lane_id * 4;
                        const cl_uint vmask = vm & 0xfU;
                        cl_uint     votes4[4];

                        vm >>= 4;
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            votes4[j] = __ballot_sync(0xffffffffU, (vmask >> j) & 1U);
#pragma unroll
                        for (int j = 0; j < 4; j++)
                        {
                            if ((vmask >> j) & 1U)
                                rowq32[(qtail + __popc(votes4[j] & lanes_lt)) & (PGS_ROWQ_ENTRIES - 1)] =
                                    row0 + r + j;
                            qtail += __popc(votes4[j]);
                        }
                        rb += step_rows;
                        __syncwarp();
                        continue;
                    }
                    if (scanning && rb >= rows_up)
                        scanning = false;
                    {
                        /* Two batches of 32 queued rows are under way at any
                         * time: the columns of the batch popped now are asked
                         * for (plain loads into registers, nothing waits for
                         * them yet), then the batch popped last time - whose
                         * loads have had a whole round of scanning or a whole
                         * chain to arrive - goes through projection +
                         * find-or-insert + the cell updates.  (Taken one batch
                         * at a time the warp sat out a DRAM round trip per 32
                         * rows: 18% of all samples on the first use of the
                         * gathered key.) */
                        const bool  pop = (qn >= 32 || (!scanning && last_tile && qn != 0));

                        if (pop || g_have)
                        {
                            kern_row_regs   nx;
                            cl_uint         nrow = row0;
                            bool            nact = false;

                            if (pop)
                            {
                                const cl_uint   n = min(qn, 32U);

                                nact = (lane_id < n);
                                if (nact)
                                    nrow = rowq32[(qhead + lane_id) & (PGS_ROWQ_ENTRIES - 1)];
                                nx.shift = (int)(nrow & 31U);
                                GPUPREAGG_INCOL_LIST(PGS_X_INCOL_GLOAD)
                                qhead += n;
                                npassed += n;
                            }
                            if (g_have)
                            {
                                bool        active = g_act;
                                pagg_row    prow;

                                PGS_DBG_COUNT(3)
                                if (active)
                                {
                                    cl_int  e = StromError_Success;

                                    gpupreagg_projection(&e, kparams, g_row, prow, kds_in, g_rownum, 0);
                                    gpupreagg_aggcheck(&e, prow);
                                    if (e != StromError_Success)
                                    {
                                        pgs_note_error(e, g_rownum, recheck_map, ctx);
                                        active = false;
                                    }
                                }
                                pgs_group_add_row(gs, sh, &head->sh_nused, prow, ctx, active,
                                                  g_rownum, recheck_map);
                                __syncwarp();
                            }
                            g_have = pop;
                            if (pop)
                            {
                                g_row = nx;
                                g_rownum = nrow;
                                g_act = nact;
                            }
                            continue;
                        }
                    }
                    if (!scanning)
                        break;
                }
            }
            PGS_DBG_STOP(2)
            continue;       /* the stage was handed back above */
#else
            /* GROUP BY under a WHERE clause.  Only a fraction of the rows
             * reaches the hash table; taking that path per row would run it
             * with a few lanes of each warp.  So a warp evaluates the qual
             * of 128 rows (4 per lane, from registers), and the survivors'
             * positions in the tile go - compacted with ballots - into its
             * queue: one 16-bit store per surviving row.  When the warp is
             * done with the tile it copies the (< 64) queued rows by value
             * into its row buffer - one row per lane - and hands the stage
             * back to the producer; only then, whenever 32 rows are queued,
             * every lane takes one through projection + find-or-insert +
             * the cell updates.  That chain of dependent shared-memory
             * operations is the slow part; run behind the release it
             * overlaps the other warps and the TMA refill instead of holding
             * up the CTA's ring (measured: 33% of the warp cycles were spent
             * waiting for tiles when the chain ran before the release).
             * The chain exists once in the program: the loop below is a
             * small state machine around it.  Every warp makes the same
             * number of steps per tile (bounds are checked per row), and the
             * last tile of the CTA is where each warp drains what is left. */
            {
                const cl_uint   step_rows = GPUPREAGG_CONSUMER_THREADS * 4;
                const cl_uint   rows_up = ((rows + step_rows - 1) / step_rows) * step_rows;
                const bool      last_tile = (t + gridDim.x >= ntiles);
                cl_uint         rb = (ctid & ~31U) * 4;
                bool            scanning = true;

                for (;;)
                {
                    if (scanning && rb < rows_up)
                    {
                        const cl_uint r = rb + lane_id * 4;
                        const cl_uint nv = (r < rows ? min(rows - r, 4U) : 0U);
                        kern_row_regs rr[4];
                        bool        valid4[4];
                        cl_uint     votes4[4];

                        rr[0].shift = 0; rr[1].shift = 1; rr[2].shift = 2; rr[3].shift = 3;
                        GPUPREAGG_INCOL_LIST(PGS_X_INCOL_LOAD4)
                        /* the 4 rows are independent: all four quals, then
                         * the four ballots, so that their latencies overlap */
#pragma unroll
                        for (int j = 0; j < 4; j++)
                        {
                            cl_int      e = StromError_Success;

#if GPUPREAGG_QUAL_DEREFS
                            /* the qual follows varlena offsets: rows past the
                             * end of the tile hold stale ones */
                            valid4[j] = false;
                            if ((cl_uint)j < nv)
                                valid4[j] = gpupreagg_qual_eval(&e, kparams, rr[j], kds_in,
                                                                row0 + r + j);
#else
                            valid4[j] = gpupreagg_qual_eval(&e, kparams, rr[j], kds_in,
                                                            row0 + r + j) & ((cl_uint)j < nv);
#endif
                            if (e != StromError_Success)
                            {
                                if ((cl_uint)j < nv)
                                {
                                    pgs_note_error(e, row0 + r + j, recheck_map, ctx);
                                    ctx.nfiltered--;    /* neither passed nor filtered */
                                }
                                valid4[j] = false;
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            votes4[j] = __ballot_sync(0xffffffffU, valid4[j]);
#pragma unroll
                        for (int j = 0; j < 4; j++)
                        {
                            if (valid4[j])
                                rowq[(qtail + __popc(votes4[j] & lanes_lt)) & (PGS_ROWQ_ENTRIES - 1)] =
                                    (cl_ushort)(r + j);
                            qtail += __popc(votes4[j]);
                        }
                        nscanned += min(rows - min(rb, rows), 128U);
                        rb += step_rows;
                        __syncwarp();
                    }
                    else if (scanning && qtail - qhead < PGS_ROWQ_LEFT_ENTRIES)
                    {
                        /* done with the tile: keep what is queued by value,
                         * queue position p in buffer entry p mod 64 */
                        for (cl_uint k = lane_id; k < qtail - qhead; k += 32)
                        {
                            const cl_uint   __qp = (qhead + k) & (PGS_ROWQ_ENTRIES - 1);
                            const cl_uint   __ri = rowq[__qp];

                            if ((__ri & PGS_ROWQ_LEFT) == 0)
                            {
                                const cl_uint   __pos = __qp & (PGS_ROWQ_LEFT_ENTRIES - 1);
                                cl_uint         __mask = 0;

                                GPUPREAGG_INCOL_LIST(PGS_X_INCOL_QKEEP)
                                *((cl_uint *)(__pgs_smem + __qbase + PGS_ROWQ_MASK_OFF) + __pos) = __mask;
                                *((cl_uint *)(__pgs_smem + __qbase + PGS_ROWQ_ROW_OFF) + __pos) = row0 + __ri;
                                rowq[__qp] = (cl_ushort)(PGS_ROWQ_LEFT | __pos);
                            }
                        }
                        __syncwarp();
                        if (lane_id == 0)
                            pgs_mbar_arrive(&head->empty_bar[stage]);
                        scanning = false;
                        PGS_DBG_STOP(1)
                    }
                    {
                        const cl_uint   qn = qtail - qhead;

                        if (scanning ? (qn >= PGS_ROWQ_LEFT_ENTRIES)
                                     : (qn >= 32 || (last_tile && qn != 0)))
                        {
                            const cl_uint   n = min(qn, 32U);
                            const cl_uint   __e = rowq[(qhead + lane_id) & (PGS_ROWQ_ENTRIES - 1)];
                            bool            active = (lane_id < n);
                            kern_qrow_smem  qrow;
                            pagg_row        prow;

                            PGS_DBG_COUNT(3)
                            qrow.isleft = ((__e & PGS_ROWQ_LEFT) != 0);
                            qrow.idx = (__e & (PGS_ROWQ_LEFT - 1));
                            qrow.lmask = *((const cl_uint *)(__pgs_smem + __qbase + PGS_ROWQ_MASK_OFF) +
                                           (qrow.idx & (PGS_ROWQ_LEFT_ENTRIES - 1)));
                            GPUPREAGG_INCOL_LIST(PGS_X_INCOL_QVIEW)
                            if (active)
                            {
                                cl_int  e = StromError_Success;

                                gpupreagg_projection(&e, kparams, qrow, prow, kds_in, 0, 0);
                                gpupreagg_aggcheck(&e, prow);
                                if (e != StromError_Success)
                                {
                                    cl_uint rownum = (qrow.isleft
                                        ? *((const cl_uint *)(__pgs_smem + __qbase + PGS_ROWQ_ROW_OFF) + qrow.idx)
                                        : row0 + qrow.idx);
                                    pgs_note_error(e, rownum, recheck_map, ctx);
                                    active = false;
                                }
                            }
                            const cl_uint   rownum = (qrow.isleft
                                ? *((const cl_uint *)(__pgs_smem + __qbase + PGS_ROWQ_ROW_OFF) + qrow.idx)
                                : row0 + qrow.idx);
                            if (PGS_PART_NPARTS(gs) != 0)
                            {
                                if (active)
                                {
                                    cl_uint     knull;
                                    cl_ulong    hash = pgs_hash_keys(prow, knull);
                                    active = !pgs_part_emit(gs, qrow, 0, rownum, hash);
                                }
                            }
                            if (PGS_PART_NPARTS(gs) == 0 || __any_sync(0xffffffffU, active))
                                pgs_group_add_row(gs, sh, &head->sh_nused, prow, ctx, active,
                                                  rownum, recheck_map);
                            qhead += n;
                            npassed += n;
                        }
                        else if (!scanning)
                            break;
                    }
                }
            }
            continue;       /* the stage was handed back above */
            PGS_DBG_STOP(1)
#endif
            __syncwarp();
            if (lane_id == 0)
                pgs_mbar_arrive(&head->empty_bar[stage]);
        }
#if GPUPREAGG_NUM_KEYS == 0
        acc_nn |= tacc.fold(acc);
#elif GPUPREAGG_HAS_QUAL
        /* rows the qual removed, counted per warp */
        if (lane_id == 0)
            ctx.nfiltered += nscanned - npassed;
        PGS_DBG_FLUSH()
#endif
    }
#if GPUPREAGG_PARTITIONED
    if (gs.part_nparts != 0 && gs.part_seg_cap != 0)
    {
        /* segment mode: how many records this CTA left in each partition */
        const cl_uint  *segcur = (const cl_uint *)(__pgs_smem + sh.base);
        cl_ushort      *counts = gs.part_seg_counts + (cl_ulong)blockIdx.x * gs.part_nparts;

        __syncthreads();
        for (cl_uint p = threadIdx.x; p < gs.part_nparts; p += blockDim.x)
            counts[p] = (cl_ushort)min((segcur[p >> 1] >> ((p & 1U) * 16U)) & 0xffffU,
                                       gs.part_seg_cap);
        __syncthreads();
    }
#endif
    pgs_main_epilogue(kgpreagg, gs, sh, (cl_ulong *)stages,
                      &head->is_last_cta, acc, acc_nn, ctx);
}

/*
 * gpupreagg_main_rowmap - gather path: the chunk comes with a kern_row_map
 * (rows a bulk-load child found visible, gpuscan.c:1425-1427), so columns are
 * read straight from HBM by row index.  Same per-row body and epilogue.
 */
extern "C" __global__ void
__launch_bounds__(GPUPREAGG_BLOCK_THREADS)
gpupreagg_main_rowmap(kern_gpupreagg *kgpreagg,
                      const kern_data_store *kds_in,
                      pgs_gstate gs,
                      cl_uint *recheck_map,
                      cl_uint sh_nslots,
                      cl_uint tile_rows,
                      cl_uint nstages)
{
    pgs_smem_head  *head = (pgs_smem_head *)__pgs_smem;
    unsigned char  *stages = __pgs_smem + PGS_SMEM_HEAD_BYTES;
    const kern_parambuf *kparams = KERN_GPUPREAGG_PARAMBUF_CONST;
    const kern_row_map  *krowmap = KERN_GPUPREAGG_KROWMAP(kgpreagg);
    const cl_uint   nrows = kds_in->nitems;
    const cl_uint   nvalids = (krowmap->nvalids < 0
                               ? nrows : (cl_uint)krowmap->nvalids);
    cl_ulong        acc[PGS_MAX(GPUPREAGG_NUM_CELLS, 1)];
    cl_uint         acc_nn = 0;
    pgs_row_ctx     ctx;
    pgs_sh_table    sh;
    kern_tile_gmem  gtile;

    ctx.nfiltered = 0;
    ctx.nrecheck = 0;
    ctx.ninserted = 0;
    ctx.errcode = StromError_Success;
    sh.base = PGS_SMEM_HEAD_BYTES + nstages * PGS_STAGE_BYTES(tile_rows);
    sh.nslots = (GPUPREAGG_NUM_KEYS > 0 ? sh_nslots : 0);
    sh.salt = 1;
    pgs_cells_init(acc);
    GPUPREAGG_INCOL_LIST(PGS_X_INCOL_GVIEW)
    if (threadIdx.x == 0)
    {
        head->sh_nused = 0;
        head->is_last_cta = 0;
    }
    PGS_SH_TABLE_INIT()
    __syncthreads();

    /* warp-uniform trip count: the group path is warp-collective */
    for (cl_ulong base = (cl_ulong)blockIdx.x * blockDim.x + (threadIdx.x & ~31U);
         base < nvalids;
         base += (cl_ulong)gridDim.x * blockDim.x)
    {
        cl_ulong    i = base + (threadIdx.x & 31U);
        pagg_row    prow;
        bool        valid = false;
        cl_uint     row = 0;

        if (i < nvalids)
        {
            row = (krowmap->nvalids < 0
                   ? (cl_uint)i : (cl_uint)krowmap->rindex[i]);
            if (row >= nrows)
            {
                if (ctx.errcode == StromError_Success)
                    ctx.errcode = StromError_DataStoreOutOfRange;
            }
            else
                valid = pgs_eval_row(kparams, gtile, kds_in, row, recheck_map, ctx, prow);
        }
#if GPUPREAGG_NUM_KEYS == 0
        acc_nn |= gpupreagg_aggcalc_plain(acc, prow, valid);
#else
        pgs_group_add_row(gs, sh, &head->sh_nused, prow, ctx, valid, row, recheck_map);
#endif
    }
    pgs_main_epilogue(kgpreagg, gs, sh, (cl_ulong *)stages,
                      &head->is_last_cta, acc, acc_nn, ctx);
}

/* ------------------------------------------------------------------
 * Heap-page input: KDS_FORMAT_ROW / KDS_FORMAT_ROW_FLAT chunks exactly as
 * pgstrom_data_store_insert_block / _insert_tuple fill them
 * (datastore.c:556-823) and clserv_dmasend_data_store lays them out on the
 * device (datastore.c:837-973):
 *   ROW      : head | kern_blkitem[maxblocks] | kern_rowitem[nitems] | pad to
 *              BLCKSZ | raw heap pages[nblocks]; a row item names (block,
 *              line pointer)
 *   ROW_FLAT : head | kern_rowitem[nitems] | ... | heap tuples from the tail;
 *              a row item is the offset of a HeapTupleHeaderData
 * One thread de-forms one tuple: a single walk over the attributes up to the
 * last referenced column (alignment padding, NULL bitmap, 1- and 4-byte
 * varlena headers; the rules of kern_get_datum_tuple, opencl_common.h:
 * 817-864) fills the register row the generated code reads, so the tuple is
 * walked once per row and not once per referenced column.  The sanity checks
 * of kern_get_tuple_rs (opencl_common.h:866-899) are kept: a page that does
 * not look like a heap page yields StromError_DataStoreCorruption, never a
 * wild address.
 * ------------------------------------------------------------------ */
#define PGS_HEAP_HASNULL        0x0001
#define PGS_HEAP_NATTS_MASK     0x07FF
#define PGS_HTUP_INFOMASK2_OFF  18
#define PGS_HTUP_INFOMASK_OFF   20
#define PGS_HTUP_HOFF_OFF       22
#define PGS_HTUP_BITS_OFF       23
#define PGS_PAGE_HEADER_SIZE    24      /* offsetof(PageHeaderData, pd_linp) */
#define PGS_PAGE_LOWER_OFF      12
/* ItemIdData on a little-endian host: lp_off:15, lp_flags:2, lp_len:15 */
#define PGS_ITEMID_OFFSET(x)    ((x) & 0x7fffU)
#define PGS_ITEMID_FLAGS(x)     (((x) >> 15) & 0x3U)
#define PGS_ITEMID_LENGTH(x)    (((x) >> 17) & 0x7fffU)
#define PGS_LP_NORMAL           1

struct pgs_heap_chunk
{
    const unsigned char *base;      /* the kern_data_store */
    const kern_rowitem  *rowitems;
    const unsigned char *blocks;    /* ROW: first page */
    cl_uint     nblocks;
    cl_uint     length;             /* ROW_FLAT: bytes of the chunk */
    bool        flat;
};

DEVFN void
pgs_heap_chunk_init(pgs_heap_chunk &hc, const kern_data_store *kds)
{
    hc.base = (const unsigned char *)kds;
    hc.flat = (kds->format == KDS_FORMAT_ROW_FLAT);
    hc.rowitems = KERN_DATA_STORE_ROWITEM(kds, 0);
    hc.blocks = (const unsigned char *)KERN_DATA_STORE_ROWBLOCK(kds, 0);
    hc.nblocks = kds->nblocks;
    hc.length = kds->length;
}

/* tuple behind line pointer `item_offset` (1-based) of a heap page, or NULL
 * (bytes available behind it in *p_avail); the page may sit in HBM or in a
 * staged copy in shared memory */
DEVFN const unsigned char *
pgs_heap_page_tuple(const unsigned char *page, cl_uint item_offset, cl_uint *p_avail)
{
    cl_uint     lower, nlines, lp, off;

    lower = *((const cl_ushort *)(page + PGS_PAGE_LOWER_OFF));
    nlines = (lower <= PGS_PAGE_HEADER_SIZE ? 0 : (lower - PGS_PAGE_HEADER_SIZE) / 4);
    if (PGS_PAGE_HEADER_SIZE + 4 * (nlines + 1) >= BLCKSZ ||
        item_offset == 0 || item_offset > nlines)
        return NULL;
    lp = *((const cl_uint *)(page + PGS_PAGE_HEADER_SIZE) + (item_offset - 1));
    off = PGS_ITEMID_OFFSET(lp);
    if (PGS_ITEMID_FLAGS(lp) != PGS_LP_NORMAL || (off & 7U) != 0 ||
        off + PGS_HTUP_BITS_OFF >= BLCKSZ)
        return NULL;
    *p_avail = BLCKSZ - off;
    return page + off;
}

/* tuple of row `rowidx`, or NULL (bytes available behind it in *p_avail) */
DEVFN const unsigned char *
pgs_heap_tuple(const pgs_heap_chunk &hc, cl_uint rowidx, cl_uint *p_avail)
{
    kern_rowitem    ri = hc.rowitems[rowidx];

    if (hc.flat)
    {
        if (ri.htup_offset >= hc.length || (ri.htup_offset & 7U) != 0 ||
            ri.htup_offset + PGS_HTUP_BITS_OFF >= hc.length)
            return NULL;
        *p_avail = hc.length - ri.htup_offset;
        return hc.base + ri.htup_offset;
    }
    if (ri.blk_index >= hc.nblocks)
        return NULL;
    return pgs_heap_page_tuple(hc.blocks + (cl_ulong)BLCKSZ * ri.blk_index,
                               ri.item_offset, p_avail);
}

/* VARSIZE_ANY (opencl_common.h:459-463) of a little-endian varlena */
DEVFN cl_uint
pgs_varsize_any(const unsigned char *p)
{
    cl_uint b = p[0];

    if (b == 0x01)                      /* 1B_E: external TOAST pointer */
        return 2 + (p[1] == 18 ? 16 : (p[1] == 1 ? 8 : 16));
    if (b & 0x01)                       /* 1B: short header */
        return (b >> 1) & 0x7fU;
    /* 4B header: may sit unaligned only when it is a pad-free short one,
     * and that case was handled above */
    return (*((const cl_uint *)p) >> 2) & 0x3fffffffU;
}

/* value of a by-value attribute of `attlen` bytes, zero-extended */
template <int ATTLEN> struct pgs_heap_load;
template <> struct pgs_heap_load<8>
{ static __device__ __forceinline__ cl_ulong get(const unsigned char *p) { return *((const cl_ulong *)p); } };
template <> struct pgs_heap_load<4>
{ static __device__ __forceinline__ cl_ulong get(const unsigned char *p) { return *((const cl_uint *)p); } };
template <> struct pgs_heap_load<2>
{ static __device__ __forceinline__ cl_ulong get(const unsigned char *p) { return *((const cl_ushort *)p); } };
template <> struct pgs_heap_load<1>
{ static __device__ __forceinline__ cl_ulong get(const unsigned char *p) { return p[0]; } };

#define PGS_X_INCOL_HEAPMAX(slot,colidx,attlen)                         \
    if ((cl_uint)(colidx) + 1 > lastcol) lastcol = (cl_uint)(colidx) + 1;
#define PGS_X_INCOL_HEAPCLEAR(slot,colidx,attlen)                       \
    rr.v[slot] = 0; rr.vbits[slot] = 0;
#define PGS_X_INCOL_HEAPTAKE(slot,colidx,attlen)                        \
    if (i == (cl_uint)(colidx))                                         \
    {                                                                   \
        if ((cl_int)(attlen) != alen)                                   \
            return false;                                               \
        rr.v[slot] = pgs_heap_load<attlen>::get(addr);                  \
        rr.vbits[slot] = 1U;                                            \
    }

#define PGS_X_INCOL_HEAPTAKEV(slot,colidx,attlen)                       \
    if (i == (cl_uint)(colidx))                                         \
    {                                                                   \
        rr.v[slot] = htup_off + (cl_ulong)(addr - htup);                \
        rr.vbits[slot] = 1U;                                            \
    }

/* columns up to the last referenced one: what a tuple walk has to step over */
__device__ constexpr cl_uint
pgs_heap_lastcol_const(void)
{
    cl_uint lastcol = 0;

    GPUPREAGG_INCOL_LIST(PGS_X_INCOL_HEAPMAX)
    return lastcol;
}
#define PGS_HEAP_LASTCOL    (pgs_heap_lastcol_const())
/*
 * attlen / attalign of those columns, read from colmeta[] ONCE per thread and
 * kept in registers (the walk is unrolled over PGS_HEAP_LASTCOL, so every
 * index is a constant).  Reading kds->colmeta[i] per attribute and tuple -
 * what the reference does (opencl_common.h:817-864) - put a global load in
 * front of every step of a chain that is already serial in `offset`
 * (ncu, heap scan of the bench table: LDG.U16 / LDG.U8 of colmeta in the
 * loop body, 16 of 32 lanes active).  Wider prefixes than
 * PGS_HEAP_META_MAX columns keep the loads.
 */
#define PGS_HEAP_META_MAX   24
struct pgs_heap_meta
{
    cl_uint     m[PGS_HEAP_META_MAX];   /* (ushort)attlen | attalign << 16 */
    cl_uint     ncols;
};

DEVFN void
pgs_heap_meta_load(pgs_heap_meta &hm, const kern_data_store *kds)
{
    hm.ncols = kds->ncols;
#pragma unroll
    for (cl_uint i = 0; i < PGS_HEAP_META_MAX; i++)
    {
        hm.m[i] = 0;
        if (i < PGS_HEAP_LASTCOL && i < hm.ncols)
        {
            kern_colmeta    cmeta = kds->colmeta[i];

            hm.m[i] = (cl_uint)(cl_ushort)cmeta.attlen |
                      ((cl_uint)(unsigned char)cmeta.attalign << 16);
        }
    }
}

/* false: the tuple does not fit what colmeta[] says (corruption).
 * `htup` may be a staged copy; `htup_off` is where the tuple sits in the
 * chunk (varlena values travel as chunk offsets and are read from HBM) */
DEVFN bool
pgs_heap_deform(const kern_data_store *kds, const unsigned char *htup,
                cl_uint avail, kern_row_regs &rr, cl_ulong htup_off,
                const pgs_heap_meta &hm)
{
    const cl_uint   infomask = *((const cl_ushort *)(htup + PGS_HTUP_INFOMASK_OFF));
    const cl_uint   natts = *((const cl_ushort *)(htup + PGS_HTUP_INFOMASK2_OFF)) & PGS_HEAP_NATTS_MASK;
    const bool      hasnull = (infomask & PGS_HEAP_HASNULL) != 0;
    cl_uint         offset = htup[PGS_HTUP_HOFF_OFF];
    cl_uint         lastcol = PGS_HEAP_LASTCOL;

    rr.shift = 0;
    GPUPREAGG_INCOL_LIST(PGS_X_INCOL_HEAPCLEAR)
    if (lastcol > hm.ncols)
        return false;
    /* attributes beyond natts are NULL (tuple older than ADD COLUMN) */
    if (lastcol > natts)
        lastcol = natts;
    if (offset < PGS_HTUP_BITS_OFF + (hasnull ? (natts + 7) / 8 : 0))
        return false;
    if (PGS_HEAP_LASTCOL <= PGS_HEAP_META_MAX)
    {
#pragma unroll
        for (cl_uint i = 0; i < PGS_HEAP_META_MAX; i++)
        {
            if (i < PGS_HEAP_LASTCOL && i < lastcol)
            {
                const unsigned char *addr;
                const cl_int    alen = (cl_int)(cl_short)(hm.m[i] & 0xffffU);
                const cl_uint   aalign = hm.m[i] >> 16;

                if (hasnull && !((htup[PGS_HTUP_BITS_OFF + (i >> 3)] >> (i & 7)) & 1))
                    continue;                   /* NULL: takes no space */
                if (alen > 0)
                    offset = TYPEALIGN(aalign, offset);
                else if (offset < avail && htup[offset] == 0)   /* !VARATT_NOT_PAD_BYTE */
                    offset = TYPEALIGN(aalign, offset);
                if (offset + (alen > 0 ? (cl_uint)alen : 1U) > avail)
                    return false;
                addr = htup + offset;
                if (alen > 0)
                {
                    GPUPREAGG_INCOL_LIST(PGS_X_INCOL_HEAPTAKE)
                    offset += (cl_uint)alen;
                }
                else
                {
                    /* varlena: the row carries the offset of the datum from
                     * the head of the chunk, like a KDS_FORMAT_COLUMN value
                     * (pg_numeric_vref) */
                    GPUPREAGG_INCOL_LIST(PGS_X_INCOL_HEAPTAKEV)
                    offset += pgs_varsize_any(addr);
                }
            }
        }
        return true;
    }
    for (cl_uint i = 0; i < lastcol; i++)
    {
        kern_colmeta    cmeta;
        const unsigned char *addr;
        cl_int          alen;

        if (hasnull && !((htup[PGS_HTUP_BITS_OFF + (i >> 3)] >> (i & 7)) & 1))
            continue;                   /* NULL: takes no space */
        cmeta = kds->colmeta[i];
        alen = cmeta.attlen;
        if (alen > 0)
            offset = TYPEALIGN((cl_uint)cmeta.attalign, offset);
        else if (offset < avail && htup[offset] == 0)   /* !VARATT_NOT_PAD_BYTE */
            offset = TYPEALIGN((cl_uint)cmeta.attalign, offset);
        if (offset + (alen > 0 ? (cl_uint)alen : 1U) > avail)
            return false;
        addr = htup + offset;
        if (alen > 0)
        {
            GPUPREAGG_INCOL_LIST(PGS_X_INCOL_HEAPTAKE)
            offset += (cl_uint)alen;
        }
        else
        {
            GPUPREAGG_INCOL_LIST(PGS_X_INCOL_HEAPTAKEV)
            offset += pgs_varsize_any(addr);
        }
    }
    return true;
}

DEVFN bool
pgs_heap_deform(const kern_data_store *kds, const unsigned char *htup,
                cl_uint avail, kern_row_regs &rr, const pgs_heap_meta &hm)
{
    return pgs_heap_deform(kds, htup, avail, rr,
                           (cl_ulong)(htup - (const unsigned char *)kds), hm);
}

/* (one tuple on its own: the CPU simulation of the walk, tests/test_heap_deform_hostsim.py) */
DEVFN bool
pgs_heap_deform(const kern_data_store *kds, const unsigned char *htup,
                cl_uint avail, kern_row_regs &rr)
{
    pgs_heap_meta   hm;

    pgs_heap_meta_load(hm, kds);
    return pgs_heap_deform(kds, htup, avail, rr, hm);
}

/*
 * gpupreagg_main_heap - heap-page chunks (with or without a kern_row_map).
 * Same per-row body and epilogue as the other two main kernels; no staging:
 * rows of one page are neighbours in memory, so the lanes of a warp read
 * neighbouring sectors.
 */
extern "C" __global__ void
__launch_bounds__(GPUPREAGG_BLOCK_THREADS)
gpupreagg_main_heap(kern_gpupreagg *kgpreagg,
                    const kern_data_store *kds_in,
                    pgs_gstate gs,
                    cl_uint *recheck_map,
                    cl_uint sh_nslots,
                    cl_uint tile_rows,
                    cl_uint nstages)
{
    pgs_smem_head  *head = (pgs_smem_head *)__pgs_smem;
    unsigned char  *stages = __pgs_smem + PGS_SMEM_HEAD_BYTES;
    const kern_parambuf *kparams = KERN_GPUPREAGG_PARAMBUF_CONST;
    const kern_row_map  *krowmap = KERN_GPUPREAGG_KROWMAP(kgpreagg);
    const cl_uint   nrows = kds_in->nitems;
    const cl_uint   nvalids = (krowmap->nvalids < 0
                               ? nrows : (cl_uint)krowmap->nvalids);
    cl_ulong        acc[PGS_MAX(GPUPREAGG_NUM_CELLS, 1)];
    cl_uint         acc_nn = 0;
    pgs_row_ctx     ctx;
    pgs_sh_table    sh;
    pgs_heap_chunk  hc;

    ctx.nfiltered = 0;
    ctx.nrecheck = 0;
    ctx.ninserted = 0;
    ctx.errcode = StromError_Success;
    sh.base = PGS_SMEM_HEAD_BYTES + nstages * PGS_STAGE_BYTES(tile_rows);
    sh.nslots = (GPUPREAGG_NUM_KEYS > 0 ? sh_nslots : 0);
    sh.salt = 1;
    pgs_cells_init(acc);
    pgs_heap_chunk_init(hc, kds_in);
    pgs_heap_meta   hmeta;
    pgs_heap_meta_load(hmeta, kds_in);
    if (threadIdx.x == 0)
    {
        head->sh_nused = 0;
        head->is_last_cta = 0;
    }
    PGS_SH_TABLE_INIT()
    __syncthreads();

    /* warp-uniform trip count: the group path is warp-collective */
    for (cl_ulong base = (cl_ulong)blockIdx.x * blockDim.x + (threadIdx.x & ~31U);
         base < nvalids;
         base += (cl_ulong)gridDim.x * blockDim.x)
    {
        cl_ulong    i = base + (threadIdx.x & 31U);
        pagg_row    prow;
        bool        valid = false;
        cl_uint     row = 0;

        if (i < nvalids)
        {
            row = (krowmap->nvalids < 0
                   ? (cl_uint)i : (cl_uint)krowmap->rindex[i]);
            if (row >= nrows)
            {
                if (ctx.errcode == StromError_Success)
                    ctx.errcode = StromError_DataStoreOutOfRange;
            }
            else
            {
                kern_row_regs   rr;
                cl_uint         avail = 0;
                const unsigned char *htup = pgs_heap_tuple(hc, row, &avail);

                if (!htup || !pgs_heap_deform(kds_in, htup, avail, rr, hmeta))
                {
                    if (ctx.errcode == StromError_Success)
                        ctx.errcode = StromError_DataStoreCorruption;
                }
                else
                    valid = pgs_eval_row(kparams, rr, kds_in, row, recheck_map, ctx, prow);
            }
        }
#if GPUPREAGG_NUM_KEYS == 0
        acc_nn |= gpupreagg_aggcalc_plain(acc, prow, valid);
#else
        pgs_group_add_row(gs, sh, &head->sh_nused, prow, ctx, valid, row, recheck_map);
#endif
    }
    pgs_main_epilogue(kgpreagg, gs, sh, (cl_ulong *)stages,
                      &head->is_last_cta, acc, acc_nn, ctx);
}

/* ------------------------------------------------------------------
 * Heap pages through the staging ring (KDS_FORMAT_ROW without a row map).
 *
 * gpupreagg_main_heap reads every tuple where it lies in HBM: row item ->
 * line pointer -> tuple header -> attributes, four dependent round trips per
 * tuple, hidden only by the number of resident warps (measured 1.4 TB/s of
 * physical bytes).  Here the pages themselves travel: the producer lane
 * copies a run of whole pages and the slice of row items that points into
 * them into a stage with two bulk copies (TMA), and the consumer threads
 * de-form from shared memory.  Row items are written page by page
 * (pgstrom_data_store_insert_block, datastore.c:556-710), so the rows of a
 * run of pages are a contiguous run of row items; where it starts is what
 * gpupreagg_heap_index finds in one pass over the row items:
 *   first_row[b]       = first row item of page b (0xffffffff: none visible)
 *   first_row[nblocks] = nitems
 *   first_row[PGS_HEAP_INDEX_FLAG] != 0: the row items are not ordered by page
 *       - the kernel then falls back to reading in place, like
 *       gpupreagg_main_heap
 * (the host presets the entries to 0xffffffff and the flag to 0)
 * ------------------------------------------------------------------ */
#define PGS_HEAP_INDEX_FLAG     65537       /* a chunk has at most 65536 pages */

extern "C" __global__ void
gpupreagg_heap_index(const kern_data_store *kds, cl_uint *first_row)
{
    const kern_rowitem *items = KERN_DATA_STORE_ROWITEM(kds, 0);
    const cl_uint   nitems = kds->nitems;
    const cl_uint   nblocks = kds->nblocks;

    for (cl_ulong i = (cl_ulong)blockIdx.x * blockDim.x + threadIdx.x; i < nitems;
         i += (cl_ulong)gridDim.x * blockDim.x)
    {
        cl_uint b = items[i].blk_index;
        cl_uint prev = (i > 0 ? (cl_uint)items[i - 1].blk_index : 0xffffffffU);

        if (b >= nblocks || (i > 0 && b < prev))
            first_row[PGS_HEAP_INDEX_FLAG] = 1;
        else if (b != prev)
            first_row[b] = (cl_uint)i;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        first_row[nblocks] = nitems;
}

/* rows per page are bounded by the line pointer array: (8192 - 24) / (4 + 24) */
#define PGS_HEAP_PAGE_MAXROWS   292
/* stage: [pages | row items (+3 for the 16-byte alignment of the copy) | fr0, n] */
#define PGS_HEAP_STAGE_BYTES(pps) \
    ((pps) * BLCKSZ + PGS_ALIGN128(4 * ((pps) * PGS_HEAP_PAGE_MAXROWS + 4)) + 128)

extern "C" __global__ void
__launch_bounds__(GPUPREAGG_BLOCK_THREADS)
gpupreagg_main_heap_staged(kern_gpupreagg *kgpreagg,
                           const kern_data_store *kds_in,
                           pgs_gstate gs,
                           cl_uint *recheck_map,
                           cl_uint sh_nslots,
                           cl_uint pages_per_stage,
                           cl_uint nstages,
                           const cl_uint *first_row)
{
    pgs_smem_head  *head = (pgs_smem_head *)__pgs_smem;
    unsigned char  *stages = __pgs_smem + PGS_SMEM_HEAD_BYTES;
    const kern_parambuf *kparams = KERN_GPUPREAGG_PARAMBUF_CONST;
    const cl_uint   nrows = kds_in->nitems;
    const cl_uint   nblocks = kds_in->nblocks;
    const cl_uint   stage_bytes = PGS_HEAP_STAGE_BYTES(pages_per_stage);
    const cl_uint   ntiles = (nblocks + pages_per_stage - 1) / pages_per_stage;
    const cl_uint   warp_id = threadIdx.x >> 5;
    const cl_uint   lane_id = threadIdx.x & 31;
    const unsigned char *blocks = (const unsigned char *)KERN_DATA_STORE_ROWBLOCK(kds_in, 0);
    const kern_rowitem *items = KERN_DATA_STORE_ROWITEM(kds_in, 0);
    const bool      ordered = (first_row[PGS_HEAP_INDEX_FLAG] == 0);
    cl_ulong        acc[PGS_MAX(GPUPREAGG_NUM_CELLS, 1)];
    cl_uint         acc_nn = 0;
    pgs_row_ctx     ctx;
    pgs_sh_table    sh;

    ctx.nfiltered = 0;
    ctx.nrecheck = 0;
    ctx.ninserted = 0;
    ctx.errcode = StromError_Success;
    sh.base = PGS_SMEM_HEAD_BYTES + nstages * stage_bytes;
    sh.nslots = (GPUPREAGG_NUM_KEYS > 0 ? sh_nslots : 0);
    sh.salt = 1;
    pgs_cells_init(acc);
    if (threadIdx.x == 0)
    {
        for (cl_uint s = 0; s < nstages; s++)
        {
            pgs_mbar_init(&head->full_bar[s], 1);
            pgs_mbar_init(&head->empty_bar[s], GPUPREAGG_CONSUMER_WARPS);
        }
        head->sh_nused = 0;
        head->is_last_cta = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    PGS_SH_TABLE_INIT()
    __syncthreads();
    pgs_heap_meta   hmeta;
    pgs_heap_meta_load(hmeta, kds_in);

    if (!ordered)
    {
        /* row items in some other order: every thread reads its tuple in place */
        pgs_heap_chunk  hc;

        pgs_heap_chunk_init(hc, kds_in);
        for (cl_ulong base = (cl_ulong)blockIdx.x * blockDim.x + (threadIdx.x & ~31U);
             base < nrows;
             base += (cl_ulong)gridDim.x * blockDim.x)
        {
            cl_uint     row = (cl_uint)base + lane_id;
            pagg_row    prow;
            bool        valid = false;

            if (row < nrows)
            {
                kern_row_regs   rr;
                cl_uint         avail = 0;
                const unsigned char *htup = pgs_heap_tuple(hc, row, &avail);

                if (!htup || !pgs_heap_deform(kds_in, htup, avail, rr, hmeta))
                {
                    if (ctx.errcode == StromError_Success)
                        ctx.errcode = StromError_DataStoreCorruption;
                }
                else
                    valid = pgs_eval_row(kparams, rr, kds_in, row, recheck_map, ctx, prow);
            }
#if GPUPREAGG_NUM_KEYS == 0
            acc_nn |= gpupreagg_aggcalc_plain(acc, prow, valid);
#else
            pgs_group_add_row(gs, sh, &head->sh_nused, prow, ctx, valid, row, recheck_map);
#endif
        }
    }
    else if (warp_id == 0)
    {
        /* ===== producer warp: lane 0 feeds the ring ===== */
        cl_uint it = 0;
        for (cl_uint t = blockIdx.x; t < ntiles; t += gridDim.x, it++)
        {
            if (lane_id == 0)
            {
                const cl_uint   stage = it % nstages;
                const cl_uint   phase = (it / nstages) & 1;
                const cl_uint   b0 = t * pages_per_stage;
                const cl_uint   b1 = min(b0 + pages_per_stage, nblocks);
                unsigned char  *stage_base = stages + stage * stage_bytes;
                cl_uint        *meta = (cl_uint *)(stage_base + stage_bytes - 128);
                cl_uint         fr0, fr1, bb, n, lead, ibytes;

                /* first row item at or behind page b0 / b1 (pages without a
                 * visible row carry no entry) */
                for (bb = b0; bb < nblocks && __ldg(first_row + bb) == 0xffffffffU; bb++)
                    ;
                fr0 = __ldg(first_row + bb);
                for (bb = b1; bb < nblocks && __ldg(first_row + bb) == 0xffffffffU; bb++)
                    ;
                fr1 = __ldg(first_row + bb);
                n = (fr1 > fr0 ? fr1 - fr0 : 0);
                if (n > pages_per_stage * PGS_HEAP_PAGE_MAXROWS)
                    n = 0xffffffffU;        /* more rows than line pointers: not a heap chunk */
                lead = fr0 & 3U;
                ibytes = (n == 0xffffffffU ? 0 : ((lead + n) * 4 + 15U) & ~15U);
                pgs_mbar_wait_relaxed(&head->empty_bar[stage], phase ^ 1);
                meta[0] = fr0;
                meta[1] = n;
                pgs_mbar_arrive_expect_tx(&head->full_bar[stage], (b1 - b0) * BLCKSZ + ibytes);
                pgs_bulk_g2s(stage_base, blocks + (cl_ulong)b0 * BLCKSZ, (b1 - b0) * BLCKSZ,
                             &head->full_bar[stage]);
                if (ibytes != 0)
                    pgs_bulk_g2s(stage_base + pages_per_stage * BLCKSZ,
                                 (const unsigned char *)(items + (fr0 - lead)), ibytes,
                                 &head->full_bar[stage]);
            }
            __syncwarp();
        }
    }
    else
    {
        /* ===== consumer warps ===== */
        const cl_uint   ctid = threadIdx.x - 32;
        cl_uint         stage = 0, phase = 0;

        for (cl_uint t = blockIdx.x; t < ntiles; t += gridDim.x, stage++)
        {
            if (stage == nstages)
            {
                stage = 0;
                phase ^= 1;
            }
            const unsigned char *stage_base = stages + stage * stage_bytes;
            const cl_uint  *meta = (const cl_uint *)(stage_base + stage_bytes - 128);
            const cl_uint   b0 = t * pages_per_stage;
            const cl_uint   npages = min(pages_per_stage, nblocks - b0);

            pgs_mbar_wait(&head->full_bar[stage], phase);
            const cl_uint   fr0 = meta[0];
            const cl_uint   n = meta[1];
            const kern_rowitem *sitems = (const kern_rowitem *)(stage_base + pages_per_stage * BLCKSZ) +
                (fr0 & 3U);

            if (n == 0xffffffffU)
            {
                if (ctx.errcode == StromError_Success)
                    ctx.errcode = StromError_DataStoreCorruption;
            }
            else
            {
                /* warp-uniform trip count: the group path is warp-collective */
                for (cl_uint i0 = (ctid & ~31U); i0 < n; i0 += GPUPREAGG_CONSUMER_THREADS)
                {
                    const cl_uint   i = i0 + lane_id;
                    const cl_uint   row = fr0 + i;
                    pagg_row        prow;
                    bool            valid = false;

                    if (i < n)
                    {
                        kern_rowitem    ri = sitems[i];
                        cl_uint         pg = (cl_uint)ri.blk_index - b0;
                        kern_row_regs   rr;
                        cl_uint         avail = 0;
                        const unsigned char *htup = NULL;

                        if (pg < npages)
                            htup = pgs_heap_page_tuple(stage_base + pg * BLCKSZ, ri.item_offset, &avail);
                        if (!htup ||
                            !pgs_heap_deform(kds_in, htup, avail, rr,
                                             (cl_ulong)(blocks - (const unsigned char *)kds_in) +
                                             (cl_ulong)ri.blk_index * BLCKSZ +
                                             (cl_ulong)(htup - (stage_base + pg * BLCKSZ)),
                                             hmeta))
                        {
                            if (ctx.errcode == StromError_Success)
                                ctx.errcode = StromError_DataStoreCorruption;
                        }
                        else
                            valid = pgs_eval_row(kparams, rr, kds_in, row, recheck_map, ctx, prow);
                    }
#if GPUPREAGG_NUM_KEYS == 0
                    acc_nn |= gpupreagg_aggcalc_plain(acc, prow, valid);
#else
                    pgs_group_add_row(gs, sh, &head->sh_nused, prow, ctx, valid, row, recheck_map);
#endif
                }
            }
            __syncwarp();
            if (lane_id == 0)
                pgs_mbar_arrive(&head->empty_bar[stage]);
        }
    }
    pgs_main_epilogue(kgpreagg, gs, sh, (cl_ulong *)stages,
                      &head->is_last_cta, acc, acc_nn, ctx);
}

/* ------------------------------------------------------------------
 * gpupreagg_partagg - second pass of the partitioned GROUP BY.
 *
 * With ~10 M groups every row used to cost a handful of atomics on a random
 * 96-byte slot of a multi-GB table: 276 bytes of random DRAM traffic per row,
 * 1 TB/s at best (ncu: 13.6 ms per 50 M rows).  Instead the scan deals the
 * rows into partitions by the high bits of the key hash (pgs_part_emit: one
 * 32-byte record, one cursor atomic), and here one CTA at a time takes a
 * partition: its persistent table image (a few hundred groups, the CTA-local
 * table layout) comes into shared memory with coalesced loads, the records
 * stream through the same find-or-insert + shared-memory atomics the
 * low-cardinality path uses, and the image goes back.  All HBM traffic is
 * sequential.  Rows that do not fit (a full partition, a full image) take the
 * global table as before; PostgreSQL's final Agg merges partial rows of the
 * same key, so a group may live in both.
 * block = PGS_PARTAGG_THREADS, dynamic smem = 128 + image bytes.
 * ------------------------------------------------------------------ */
#define PGS_PARTAGG_THREADS     256

/* shared memory of gpupreagg_partagg: [0] used slots of the image, [64...]
 * prefix sums of the segment lengths (segment mode), then the image */
#define PGS_PARTAGG_MAX_SEGS    256
#define PGS_PARTAGG_HEAD_BYTES  (64 + 4 * (PGS_PARTAGG_MAX_SEGS + 1) + 60)

extern "C" __global__ void
__launch_bounds__(PGS_PARTAGG_THREADS)
gpupreagg_partagg(kern_gpupreagg *kgpreagg,
                  const kern_data_store *kds_in,
                  pgs_gstate gs,
                  cl_uint *recheck_map,
                  cl_uint nseg)
{
#if GPUPREAGG_PARTITIONED
    const kern_parambuf *kparams = KERN_GPUPREAGG_PARAMBUF_CONST;
    cl_uint        *p_nused = (cl_uint *)__pgs_smem;
    cl_uint        *prefix = (cl_uint *)(__pgs_smem + 64);  /* [nseg + 1] */
    const cl_uint   image_bytes = gs.part_slots * PGS_SH_SLOT_BYTES;
    const cl_uint   lane_id = threadIdx.x & 31;
    pgs_row_ctx     ctx;
    pgs_sh_table    sh;
    cl_uint         ngrown = 0;

    ctx.nfiltered = 0;
    ctx.nrecheck = 0;
    ctx.ninserted = 0;
    ctx.errcode = StromError_Success;
    sh.base = PGS_PARTAGG_HEAD_BYTES;
    sh.nslots = gs.part_slots;
    sh.salt = 0x9E3779B1U;

    for (cl_uint part = blockIdx.x; part < gs.part_nparts; part += gridDim.x)
    {
        uint4          *image = (uint4 *)(gs.part_images + (cl_ulong)part * image_bytes);
        uint4          *local = (uint4 *)(__pgs_smem + sh.base);
        const unsigned char *recs = gs.part_recs +
            (cl_ulong)part * gs.part_cap * PGS_REC_BYTES;
        cl_uint         n;

        if (nseg == 0)
            n = min(*PGS_PART_CURSOR(gs, part), gs.part_cap);
        else
        {
            /* segment mode: the records lie in one segment per CTA of the scan
             * kernel.  Warp 0 turns the segment lengths into prefix sums (8
             * segments per lane); record i of the partition is then found by
             * a binary search over them */
            if (threadIdx.x < 32)
            {
                cl_uint c[8], sum = 0, excl;

#pragma unroll
                for (int k = 0; k < 8; k++)
                {
                    const cl_uint seg = threadIdx.x * 8 + k;

                    c[k] = (seg < nseg
                            ? (cl_uint)gs.part_seg_counts[(cl_ulong)seg * gs.part_nparts + part]
                            : 0U);
                    sum += c[k];
                }
                excl = sum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1)
                {
                    cl_uint v = __shfl_up_sync(0xffffffffU, excl, d);
                    if (lane_id >= (cl_uint)d)
                        excl += v;
                }
                excl -= sum;
#pragma unroll
                for (int k = 0; k < 8; k++)
                {
                    prefix[threadIdx.x * 8 + k] = excl;
                    excl += c[k];
                }
                if (threadIdx.x == 31)
                    prefix[PGS_PARTAGG_MAX_SEGS] = excl;
            }
            __syncthreads();
            n = prefix[PGS_PARTAGG_MAX_SEGS];
            __syncthreads();        /* (prefix[] is rewritten for the next partition) */
        }
        if (n == 0)
            continue;           /* nothing for this image in this chunk */
        for (cl_uint i = threadIdx.x; i < image_bytes / 16; i += blockDim.x)
            local[i] = image[i];
        if (threadIdx.x == 0)
            *p_nused = gs.part_nused[part];
        __syncthreads();
        /* where record i of the partition lies */
#define PGS_PARTAGG_REC(i, out)                                                 \
        {                                                                       \
            cl_uint __pos = (i);                                                \
            if (nseg != 0)                                                      \
            {                                                                   \
                cl_uint __lo = 0, __hi = PGS_PARTAGG_MAX_SEGS;                  \
                _Pragma("unroll")                                               \
                for (int __s = 0; __s < 8; __s++)                               \
                {                                                               \
                    const cl_uint __mid = (__lo + __hi) >> 1;                   \
                    if (prefix[__mid] <= (i)) __lo = __mid; else __hi = __mid;  \
                }                                                               \
                __pos = pgs_part_seg_pos(gs, __lo, (i) - prefix[__lo]);         \
            }                                                                   \
            (out) = recs + (cl_ulong)__pos * PGS_REC_BYTES;                     \
        }
        /* the records of the batch after this one are asked for before this
         * one goes through the chain (ncu: 18% of the samples sat on the
         * first use of a record read in place) */
        kern_rec_regs   cur, nxt;

        cur.clear();
        if ((threadIdx.x & ~31U) + lane_id < n)
        {
            const unsigned char *rp;

            PGS_PARTAGG_REC((threadIdx.x & ~31U) + lane_id, rp)
            cur.load(rp);
        }
        for (cl_uint i0 = (threadIdx.x & ~31U); i0 < n; i0 += blockDim.x)
        {
            const cl_uint   i = i0 + lane_id;
            bool            active = (i < n);
            pagg_row        prow;

            nxt.clear();
            if (i + blockDim.x < n)
            {
                const unsigned char *rp;

                PGS_PARTAGG_REC(i + blockDim.x, rp)
                nxt.load(rp);
            }
            if (active)
            {
                cl_int  e = StromError_Success;

                gpupreagg_projection(&e, kparams, cur, prow, kds_in, 0, 0);
                gpupreagg_aggcheck(&e, prow);
                if (e != StromError_Success)
                {
                    pgs_note_error(e, cur.rownum(), recheck_map, ctx);
                    active = false;
                }
            }
            pgs_group_add_row(gs, sh, p_nused, prow, ctx, active, cur.rownum(), recheck_map);
            cur = nxt;
        }
#undef PGS_PARTAGG_REC
        __syncthreads();
        for (cl_uint i = threadIdx.x; i < image_bytes / 16; i += blockDim.x)
            image[i] = local[i];
        if (threadIdx.x == 0)
        {
            ngrown += *p_nused - gs.part_nused[part];
            gs.part_nused[part] = *p_nused;
            if (nseg == 0)
                *PGS_PART_CURSOR(gs, part) = 0;
        }
        __syncthreads();
    }
    if (ngrown)                 /* thread 0: groups the images gained */
        atomicAdd(gs.gh_ngroups, ngrown);
    pgs_writeback_status(kgpreagg, gs, ctx);
#endif
}

/*
 * One slot of the persistent GROUP BY state, wherever it lives: index
 * [0, gh_nslots) is the global table, the rest are the slots of the table
 * images.  Returns whether the slot holds a group.
 */
DEVFN cl_ulong
pgs_state_nslots(const pgs_gstate &gs)
{
    return (cl_ulong)gs.gh_nslots + (cl_ulong)gs.part_nparts * gs.part_slots +
        (cl_ulong)min(*((volatile cl_uint *)gs.ovf_count), gs.ovf_cap);
}
DEVFN bool
pgs_state_slot(const pgs_gstate &gs, cl_ulong i, cl_ulong *keys, cl_ulong *cells,
               cl_uint &knull, cl_uint &nn)
{
    if (i < gs.gh_nslots)
    {
        const cl_ulong *slot = gs.gh_slots + i * PGS_SLOT_STRIDE;
        cl_uint     st = (cl_uint)slot[0];

        if ((st & 3U) != PGS_SLOT_READY)
            return false;
        knull = st >> 8;
        nn = (cl_uint)(slot[0] >> 32);
#pragma unroll
        for (int k = 0; k < GPUPREAGG_NUM_KEYS; k++)
            keys[k] = slot[1 + k];
#pragma unroll
        for (int c = 0; c < GPUPREAGG_NUM_CELLS; c++)
            cells[c] = slot[1 + GPUPREAGG_NUM_KEYS + c];
        return true;
    }
    else if (i >= (cl_ulong)gs.gh_nslots + (cl_ulong)gs.part_nparts * gs.part_slots)
    {
        /* a record of the overflow log */
        const cl_ulong *rec = gs.ovf_recs +
            (i - gs.gh_nslots - (cl_ulong)gs.part_nparts * gs.part_slots) * PGS_SLOT_WORDS;

        knull = (cl_uint)rec[0] >> 8;
        nn = (cl_uint)(rec[0] >> 32);
#pragma unroll
        for (int k = 0; k < GPUPREAGG_NUM_KEYS; k++)
            keys[k] = rec[1 + k];
#pragma unroll
        for (int c = 0; c < GPUPREAGG_NUM_CELLS; c++)
            cells[c] = rec[1 + GPUPREAGG_NUM_KEYS + c];
        return true;
    }
    else
    {
        const cl_ulong  j = i - gs.gh_nslots;
        const cl_uint   S = gs.part_slots;
        const cl_uint   s = (cl_uint)(j % S);
        const unsigned char *image = gs.part_images + (j / S) * ((cl_ulong)S * PGS_SH_SLOT_BYTES);
        cl_uint     tag = ((const cl_uint *)image)[s];

        if ((tag & 3U) != PGS_SLOT_READY)
            return false;
        knull = (tag >> 2) & ((1U << GPUPREAGG_NUM_KEYS) - 1U);
        nn = ((const cl_uint *)(image + 4 * (cl_ulong)S))[s];
#pragma unroll
        for (int k = 0; k < GPUPREAGG_NUM_KEYS; k++)
            keys[k] = ((const cl_ulong *)(image + 8 * (cl_ulong)S))[(cl_ulong)k * S + s];
#pragma unroll
        for (int c = 0; c < GPUPREAGG_NUM_CELLS; c++)
            cells[c] = ((const cl_ulong *)(image + 8 * (cl_ulong)S * (1 + GPUPREAGG_NUM_KEYS)))
                [(cl_ulong)c * S + s];
        return true;
    }
}

/* ------------------------------------------------------------------
 * gpupreagg_init_state - identity state everywhere
 * ------------------------------------------------------------------ */
extern "C" __global__ void
gpupreagg_init_state(pgs_gstate gs)
{
    cl_ulong    i = (cl_ulong)blockIdx.x * blockDim.x + threadIdx.x;

    if (i == 0)
    {
        gs.ng_state[0] = 0;
        pgs_cells_init(gs.ng_state + 1);
        *gs.ng_ticket = 0;
        *gs.gh_ngroups = 0;
        *gs.gh_nused = 0;
        *gs.ovf_count = 0;
    }
    for (cl_ulong j = i; j < gs.gh_nslots; j += (cl_ulong)gridDim.x * blockDim.x)
    {
        cl_ulong   *slot = gs.gh_slots + j * PGS_SLOT_STRIDE;

        slot[0] = 0;
#pragma unroll
        for (int k = 0; k < GPUPREAGG_NUM_KEYS; k++)
            slot[1 + k] = 0;
        pgs_cells_init(slot + 1 + GPUPREAGG_NUM_KEYS);
    }
    /* table images of the partitions (CTA-local table layout) */
    for (cl_ulong j = i; j < (cl_ulong)gs.part_nparts * gs.part_slots;
         j += (cl_ulong)gridDim.x * blockDim.x)
    {
        const cl_uint   S = gs.part_slots;
        const cl_uint   s = (cl_uint)(j % S);
        unsigned char  *image = gs.part_images + (j / S) * ((cl_ulong)S * PGS_SH_SLOT_BYTES);
        pgs_sh_cells    cells;

        ((cl_uint *)image)[s] = PGS_SLOT_EMPTY;
        ((cl_uint *)(image + 4 * (cl_ulong)S))[s] = 0;
#pragma unroll
        for (int k = 0; k < GPUPREAGG_NUM_KEYS; k++)
            ((cl_ulong *)(image + 8 * (cl_ulong)S))[(cl_ulong)k * S + s] = 0;
        cells.p0 = (cl_ulong *)(image + 8 * (cl_ulong)S * (1 + GPUPREAGG_NUM_KEYS)) + s;
        cells.stride = S;
        pgs_cells_init(cells);
    }
    for (cl_ulong j = i; j < gs.part_nparts; j += (cl_ulong)gridDim.x * blockDim.x)
    {
        *PGS_PART_CURSOR(gs, j) = 0;
        gs.part_nused[j] = 0;
    }
}

/* ------------------------------------------------------------------
 * gpupreagg_rehash - the global table was too small for the number of
 * groups the scan meets (the planner's numGroups is an estimate; the
 * reference never depends on it: it reduces chunk by chunk and lets the CPU
 * Agg merge, gpupreagg.c:2169-2186).  The host allocates a larger table and
 * overflow log (`to`, initialised) and this kernel moves every group of the
 * old table and of the old log over.  Table images are shared by both.
 * ------------------------------------------------------------------ */
extern "C" __global__ void
gpupreagg_rehash(pgs_gstate from, pgs_gstate to, kern_gpupreagg *kgpreagg)
{
#if GPUPREAGG_NUM_KEYS > 0
    const cl_ulong  nlog = min(*from.ovf_count, from.ovf_cap);
    const cl_ulong  n = (cl_ulong)from.gh_nslots + nlog;
    cl_uint         ninserted = 0;

    for (cl_ulong i = (cl_ulong)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (cl_ulong)gridDim.x * blockDim.x)
    {
        const cl_ulong *rec;
        cl_uint         knull, nn;

        if (i < from.gh_nslots)
        {
            rec = from.gh_slots + i * PGS_SLOT_STRIDE;
            if (((cl_uint)rec[0] & 3U) != PGS_SLOT_READY)
                continue;
        }
        else
            rec = from.ovf_recs + (i - from.gh_nslots) * PGS_SLOT_WORDS;
        knull = (cl_uint)rec[0] >> 8;
        nn = (cl_uint)(rec[0] >> 32);
        if (!pgs_gh_merge_state(to, rec + 1, knull, pgs_hash_keyvals(rec + 1, knull),
                                rec + 1 + GPUPREAGG_NUM_KEYS, nn, ninserted))
            atomicCAS(&kgpreagg->status, StromError_Success, StromError_DataStoreNoSpace);
    }
    /* `to` counts in its own words (the host swaps the pointers) */
    if (ninserted)
        atomicAdd(to.gh_nused, ninserted);
#endif
}

/* ------------------------------------------------------------------
 * gpupreagg_flush - state -> TUPSLOT partial rows
 *
 * Datum encoding as pg_<type>_vstore (opencl_common.h:567-580): by-value
 * types zero-extended into 8 bytes; isnull[] one char per column after the
 * ncols Datums; row stride LONGALIGN(9 * ncols).
 * A group whose count exceeds int4 or whose 128-bit sum exceeds int8 is
 * written as several rows (the final aggregates add partial rows up).
 * ------------------------------------------------------------------ */
#define PGS_INT4_MAX    2147483647LL

DEVFN cl_uint
pgs_nsplit_i128(cl_ulong lo, cl_ulong hi)
{
    __int128    v = (__int128)(((unsigned __int128)hi << 64) | lo);
    __int128    lim = (v >= 0 ? (__int128)LONG_MAX : -(__int128)LONG_MIN);
    __int128    a = (v >= 0 ? v : -v);
    __int128    n = (a + lim - 1) / lim;

    if (n < 1)
        n = 1;
    return (n > 0x7fffffff ? 0x7fffffffU : (cl_uint)n);
}
DEVFN cl_long
pgs_piece_i128(cl_ulong lo, cl_ulong hi, cl_uint r)
{
    __int128    v = (__int128)(((unsigned __int128)hi << 64) | lo);

    if (v >= 0)
    {
        __int128 rest = v - (__int128)r * (__int128)LONG_MAX;
        if (rest <= 0)
            return 0;
        return (rest > (__int128)LONG_MAX ? LONG_MAX : (cl_long)rest);
    }
    else
    {
        __int128 rest = v - (__int128)r * (__int128)LONG_MIN;
        if (rest >= 0)
            return 0;
        return (rest < (__int128)LONG_MIN ? LONG_MIN : (cl_long)rest);
    }
}

/* rows needed by one aggregate of one group */
#define PGS_NSPLIT_PSUM_INT(c)                                          \
    { cl_long v = (cl_long)cells[c];                                    \
      cl_long n = (v + PGS_INT4_MAX - 1) / PGS_INT4_MAX;                \
      if (n > (cl_long)nsplit) nsplit = (cl_uint)n; }
#define PGS_NSPLIT_PSUM_LONG(c)                                         \
    { cl_uint n = pgs_nsplit_i128(cells[c], cells[(c)+1]);              \
      if (n > nsplit) nsplit = n; }
#define PGS_NSPLIT_OTHER(c)
#define PGS_NSPLIT_PSUM_LONGS(c)    PGS_NSPLIT_OTHER(c)
#define PGS_NSPLIT_PSUM_FLOAT(c)    PGS_NSPLIT_OTHER(c)
#define PGS_NSPLIT_PSUM_DOUBLE(c)   PGS_NSPLIT_OTHER(c)
#define PGS_NSPLIT_PMIN_SHORT(c)    PGS_NSPLIT_OTHER(c)
#define PGS_NSPLIT_PMIN_INT(c)      PGS_NSPLIT_OTHER(c)
#define PGS_NSPLIT_PMIN_LONG(c)     PGS_NSPLIT_OTHER(c)
#define PGS_NSPLIT_PMIN_FLOAT(c)    PGS_NSPLIT_OTHER(c)
#define PGS_NSPLIT_PMIN_DOUBLE(c)   PGS_NSPLIT_OTHER(c)
#define PGS_NSPLIT_PMAX_SHORT(c)    PGS_NSPLIT_OTHER(c)
#define PGS_NSPLIT_PMAX_INT(c)      PGS_NSPLIT_OTHER(c)
#define PGS_NSPLIT_PMAX_LONG(c)     PGS_NSPLIT_OTHER(c)
#define PGS_NSPLIT_PMAX_FLOAT(c)    PGS_NSPLIT_OTHER(c)
#define PGS_NSPLIT_PMAX_DOUBLE(c)   PGS_NSPLIT_OTHER(c)
#ifdef KERN_NUMERIC_CUH
/* numeric sum cell(s) -> value at its display scale (exact: every addend had
 * at most that scale) */
DEVFN pgs_s128
pgs_numsum_value(const cl_ulong *cells, int c, int *p_dscale)
{
    pgs_s128    v = (pgs_s128)(((pgs_u128)cells[c + 1] << 64) | cells[c]);
    int         ds = (int)(cl_long)cells[c + 2];

    if (ds < 0) ds = 0;
    if (ds > PGS_NUMERIC_SUM_SCALE) ds = PGS_NUMERIC_SUM_SCALE;
    *p_dscale = ds;
    return v / (pgs_s128)pgs_pow10_u128(PGS_NUMERIC_SUM_SCALE - ds);
}
/* A sum is emitted positionally in base 10^17 (< 2^57): at most three device
 * numerics  lo * 10^-ds  +  mid * 10^(17-ds)  +  hi * 10^(34-ds)  hold any
 * 128-bit value exactly, and PostgreSQL's final sum adds them up.  Piece 0
 * always carries the display scale, also when it is zero.  (Cutting the value
 * into pieces of one full mantissa each needs |v| / 2^57 rows: a sum of 10^12
 * and 10^-16 would have asked for 10^11 of them.) */
#define PGS_NUMSUM_BASE     100000000000000000ULL       /* 10^17 */
DEVFN cl_uint
pgs_numsum_nsplit(const cl_ulong *cells, int c)
{
    int         ds;
    pgs_s128    v = pgs_numsum_value(cells, c, &ds);
    pgs_u128    a = (pgs_u128)(v < 0 ? -v : v);

    if (a < (pgs_u128)PGS_NUMSUM_BASE)
        return 1;
    return (a / PGS_NUMSUM_BASE < (pgs_u128)PGS_NUMSUM_BASE) ? 2 : 3;
}
/* piece r of the sum as a device numeric */
DEVFN cl_ulong
pgs_numsum_piece(const cl_ulong *cells, int c, cl_uint r)
{
    int         ds;
    pgs_s128    v = pgs_numsum_value(cells, c, &ds);
    pgs_u128    a = (pgs_u128)(v < 0 ? -v : v);
    cl_ulong    mant = 0;
    int         expo = -ds;

    if (r == 0)
        mant = (cl_ulong)(a % PGS_NUMSUM_BASE);
    else if (r == 1)
    {
        mant = (cl_ulong)((a / PGS_NUMSUM_BASE) % PGS_NUMSUM_BASE);
        expo = 17 - ds;
    }
    else if (r == 2)
    {
        /* a < 2^127: what is left is below 1.8 * 10^4 */
        mant = (cl_ulong)((a / PGS_NUMSUM_BASE) / PGS_NUMSUM_BASE);
        expo = 34 - ds;
        if (expo > PG_NUMERIC_EXPONENT_MAX)
        {
            mant *= pgs_pow10_u64(expo - PG_NUMERIC_EXPONENT_MAX);
            expo = PG_NUMERIC_EXPONENT_MAX;
        }
    }
    return PG_NUMERIC_SET(expo, (v < 0) && mant != 0, mant);
}
#endif
#define PGS_NSPLIT_PSUM_NUMERIC(c)                                      \
    { cl_uint n = pgs_numsum_nsplit(cells, c);                          \
      if (n > nsplit) nsplit = n; }
#define PGS_NSPLIT_PMIN_NUMERIC(c)  PGS_NSPLIT_OTHER(c)
#define PGS_NSPLIT_PMAX_NUMERIC(c)  PGS_NSPLIT_OTHER(c)
#define PGS_OUT_PSUM_NUMERIC(i,c)                                       \
    { isnull = !(nn & (1U << (i)));                                     \
      datum = pgs_numsum_piece(cells, c, r); }
#define PGS_OUT_PMIN_NUMERIC(i,c)                                       \
    { isnull = !(nn & (1U << (i))) || cells[c] == 0xFFFFFFFFFFFFFFFFULL; \
      datum = cells[c]; }
#define PGS_OUT_PMAX_NUMERIC(i,c)   PGS_OUT_PMIN_NUMERIC(i,c)
#define PGS_X_NSPLIT(i,c,OP,TYPE)   PGS_NSPLIT_##OP##_##TYPE(c)

/* Datum of aggregate cell(s) for split row r; sets `isnull` */
#define PGS_OUT_PSUM_INT(i,c)                                           \
    { cl_long v = (cl_long)cells[c] - (cl_long)r * PGS_INT4_MAX;        \
      if (v < 0) v = 0;                                                 \
      if (v > PGS_INT4_MAX) v = PGS_INT4_MAX;                           \
      datum = (cl_ulong)(cl_uint)v; isnull = false; }
#define PGS_OUT_PSUM_LONGS(i,c)                                         \
    { isnull = !(nn & (1U << (i)));                                     \
      datum = (r == 0 ? cells[c] : 0); }
#define PGS_OUT_PSUM_LONG(i,c)                                          \
    { isnull = !(nn & (1U << (i)));                                     \
      datum = (cl_ulong)pgs_piece_i128(cells[c], cells[(c)+1], r); }
#define PGS_OUT_PSUM_FLOAT(i,c)                                         \
    { isnull = !(nn & (1U << (i)));                                     \
      datum = (r == 0 ? (cl_ulong)__float_as_uint(                      \
                   (float)__longlong_as_double((cl_long)cells[c])) : 0); }
#define PGS_OUT_PSUM_DOUBLE(i,c)                                        \
    { isnull = !(nn & (1U << (i)));                                     \
      datum = (r == 0 ? cells[c] : 0); }
#define PGS_OUT_MINMAX_SHORT(i,c)                                       \
    { isnull = !(nn & (1U << (i)));                                     \
      datum = (cl_ulong)(cl_ushort)(cl_short)(cl_long)cells[c]; }
#define PGS_OUT_MINMAX_INT(i,c)                                         \
    { isnull = !(nn & (1U << (i)));                                     \
      datum = (cl_ulong)(cl_uint)(cl_int)(cl_long)cells[c]; }
#define PGS_OUT_MINMAX_LONG(i,c)                                        \
    { isnull = !(nn & (1U << (i)));                                     \
      datum = cells[c]; }
#define PGS_OUT_MINMAX_FLOAT(i,c)                                       \
    { isnull = !(nn & (1U << (i)));                                     \
      datum = (cl_ulong)__float_as_uint((float)PGS_F8_UNCELL(cells[c])); }
#define PGS_OUT_MINMAX_DOUBLE(i,c)                                      \
    { isnull = !(nn & (1U << (i)));                                     \
      datum = (cl_ulong)__double_as_longlong(PGS_F8_UNCELL(cells[c])); }
#define PGS_OUT_PMIN_SHORT(i,c)     PGS_OUT_MINMAX_SHORT(i,c)
#define PGS_OUT_PMIN_INT(i,c)       PGS_OUT_MINMAX_INT(i,c)
#define PGS_OUT_PMIN_LONG(i,c)      PGS_OUT_MINMAX_LONG(i,c)
#define PGS_OUT_PMIN_FLOAT(i,c)     PGS_OUT_MINMAX_FLOAT(i,c)
#define PGS_OUT_PMIN_DOUBLE(i,c)    PGS_OUT_MINMAX_DOUBLE(i,c)
#define PGS_OUT_PMAX_SHORT(i,c)     PGS_OUT_MINMAX_SHORT(i,c)
#define PGS_OUT_PMAX_INT(i,c)       PGS_OUT_MINMAX_INT(i,c)
#define PGS_OUT_PMAX_LONG(i,c)      PGS_OUT_MINMAX_LONG(i,c)
#define PGS_OUT_PMAX_FLOAT(i,c)     PGS_OUT_MINMAX_FLOAT(i,c)
#define PGS_OUT_PMAX_DOUBLE(i,c)    PGS_OUT_MINMAX_DOUBLE(i,c)

/* GPUPREAGG_OUT_LIST(_) = _(colidx, ROLE, idx, cellidx, OP, TYPE) with
 * ROLE in {NUL, KEY, AGG} */
#define PGS_OUTCOL_NUL(col,idx,c,OP,TYPE)                              \
    values[col] = 0; isnulls[col] = 1;
#define PGS_OUTCOL_KEY(col,idx,c,OP,TYPE)                               \
    values[col] = ((knull >> (idx)) & 1U) ? 0 : keys[idx];              \
    isnulls[col] = (cl_char)((knull >> (idx)) & 1U);
#define PGS_OUTCOL_AGG(col,idx,c,OP,TYPE)                               \
    { cl_ulong datum; bool isnull;                                      \
      PGS_OUT_##OP##_##TYPE(idx,c)                                      \
      values[col] = isnull ? 0 : datum;                                 \
      isnulls[col] = (cl_char)(isnull ? 1 : 0); }
#define PGS_X_OUTCOL(col,ROLE,idx,c,OP,TYPE)    PGS_OUTCOL_##ROLE(col,idx,c,OP,TYPE)

/* rows one group needs in the result */
DEVFN cl_uint
pgs_flush_nsplit(const cl_ulong *cells)
{
    cl_uint     nsplit = 1;

    GPUPREAGG_AGG_LIST(PGS_X_NSPLIT)
    return nsplit;
}

/* write the `nsplit` rows of one group at row `base` of the result */
DEVFN void
pgs_flush_rows(kern_data_store *kds_dst, kern_gpupreagg *kgpreagg,
               cl_uint base, cl_uint nsplit,
               const cl_ulong *keys, cl_uint knull,
               const cl_ulong *cells, cl_uint nn)
{
    if (base + nsplit > kds_dst->nrooms)
    {
        atomicCAS(&kgpreagg->status, StromError_Success,
                  StromError_DataStoreNoSpace);
        return;
    }
    for (cl_uint r = 0; r < nsplit; r++)
    {
        Datum      *values = KERN_DATA_STORE_VALUES(kds_dst, base + r);
        cl_char    *isnulls = KERN_DATA_STORE_ISNULL(kds_dst, base + r);

        GPUPREAGG_OUT_LIST(PGS_X_OUTCOL)
    }
}

/* launched with whole warps.  Result rows are reserved once per warp (an
 * exclusive scan of the rows each lane needs, one atomicAdd on nitems by
 * lane 0): one atomicAdd per group on that single word serialises in L2 and
 * took 26 ms for 10 M groups. */
extern "C" __global__ void
gpupreagg_flush(pgs_gstate gs, kern_data_store *kds_dst,
                kern_gpupreagg *kgpreagg)
{
    const cl_uint   lane_id = threadIdx.x & 31;

    if (GPUPREAGG_NUM_KEYS == 0)
    {
        if (blockIdx.x == 0 && threadIdx.x == 0)
        {
            cl_uint nsplit = pgs_flush_nsplit(gs.ng_state + 1);
            cl_uint base = atomicAdd(&kds_dst->nitems, nsplit);

            pgs_flush_rows(kds_dst, kgpreagg, base, nsplit, (const cl_ulong *)NULL, 0,
                           gs.ng_state + 1, (cl_uint)gs.ng_state[0]);
        }
        return;
    }
    const cl_ulong  nstate = pgs_state_nslots(gs);

    for (cl_ulong i0 = (cl_ulong)blockIdx.x * blockDim.x + (threadIdx.x & ~31U);
         i0 < nstate;
         i0 += (cl_ulong)gridDim.x * blockDim.x)
    {
        const cl_ulong  i = i0 + lane_id;
        cl_ulong    keys[PGS_MAX(GPUPREAGG_NUM_KEYS, 1)];
        cl_ulong    cells[PGS_MAX(GPUPREAGG_NUM_CELLS, 1)];
        cl_uint     knull = 0, nn = 0;
        bool        ready = (i < nstate && pgs_state_slot(gs, i, keys, cells, knull, nn));
        cl_uint     nsplit = (ready ? pgs_flush_nsplit(cells) : 0);
        cl_uint     incl = nsplit;
        cl_uint     base = 0;

#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
            cl_uint v = __shfl_up_sync(0xffffffffU, incl, d);
            if (lane_id >= (cl_uint)d)
                incl += v;
        }
        if (lane_id == 31 && incl > 0)
            base = atomicAdd(&kds_dst->nitems, incl);
        base = __shfl_sync(0xffffffffU, base, 31);
        if (ready)
            pgs_flush_rows(kds_dst, kgpreagg, base + incl - nsplit, nsplit,
                           keys, knull, cells, nn);
    }
}

/* ------------------------------------------------------------------
 * export / import of raw states: the unit that crosses NVLink when the
 * per-GPU states are merged (NCCL gather of these records, then import on
 * the receiving GPU).  record = PGS_SLOT_WORDS words, same as a slot.
 * ------------------------------------------------------------------ */
extern "C" __global__ void
gpupreagg_export(pgs_gstate gs, cl_ulong *records, cl_uint *nrecords,
                 cl_uint max_records)
{
    if (GPUPREAGG_NUM_KEYS == 0)
    {
        if (blockIdx.x == 0 && threadIdx.x == 0)
        {
            records[0] = (gs.ng_state[0] << 32) | PGS_SLOT_READY;
            for (int c = 0; c < GPUPREAGG_NUM_CELLS; c++)
                records[1 + c] = gs.ng_state[1 + c];
            *nrecords = 1;
        }
        return;
    }
    /* whole warps; record positions are handed out once per warp */
    const cl_ulong  nstate = pgs_state_nslots(gs);

    for (cl_ulong i0 = (cl_ulong)blockIdx.x * blockDim.x + (threadIdx.x & ~31U);
         i0 < nstate;
         i0 += (cl_ulong)gridDim.x * blockDim.x)
    {
        const cl_uint   lane_id = threadIdx.x & 31;
        const cl_ulong  i = i0 + lane_id;
        cl_ulong    keys[PGS_MAX(GPUPREAGG_NUM_KEYS, 1)];
        cl_ulong    cells[PGS_MAX(GPUPREAGG_NUM_CELLS, 1)];
        cl_uint     knull = 0, nn = 0;
        bool        ready = (i < nstate && pgs_state_slot(gs, i, keys, cells, knull, nn));
        cl_uint     votes = __ballot_sync(0xffffffffU, ready);
        cl_uint     base = 0;

        if (lane_id == 0 && votes != 0)
            base = atomicAdd(nrecords, (cl_uint)__popc(votes));
        base = __shfl_sync(0xffffffffU, base, 0);
        if (ready)
        {
            cl_uint pos = base + __popc(votes & ((1U << lane_id) - 1U));
            if (pos < max_records)
            {
                cl_ulong *rec = records + (cl_ulong)pos * PGS_SLOT_WORDS;

                rec[0] = ((cl_ulong)nn << 32) | (knull << 8) | PGS_SLOT_READY;
#pragma unroll
                for (int k = 0; k < GPUPREAGG_NUM_KEYS; k++)
                    rec[1 + k] = keys[k];
#pragma unroll
                for (int c = 0; c < GPUPREAGG_NUM_CELLS; c++)
                    rec[1 + GPUPREAGG_NUM_KEYS + c] = cells[c];
            }
        }
    }
}

extern "C" __global__ void
gpupreagg_import(pgs_gstate gs, const cl_ulong *records, cl_uint nrecords,
                 kern_gpupreagg *kgpreagg)
{
    if (GPUPREAGG_NUM_KEYS == 0)
    {
        /* one thread, records in rank order: deterministic */
        if (blockIdx.x == 0 && threadIdx.x == 0)
        {
            cl_uint nn = (cl_uint)gs.ng_state[0];
            for (cl_uint i = 0; i < nrecords; i++)
            {
                const cl_ulong *rec = records + (cl_ulong)i * PGS_SLOT_WORDS;
                cl_uint src_nn = (cl_uint)(rec[0] >> 32);
                gpupreagg_aggmerge_plain(gs.ng_state + 1, rec + 1, src_nn);
                nn |= src_nn;
            }
            gs.ng_state[0] = nn;
        }
        return;
    }
    cl_uint     ninserted = 0;

    for (cl_ulong i = (cl_ulong)blockIdx.x * blockDim.x + threadIdx.x;
         i < nrecords;
         i += (cl_ulong)gridDim.x * blockDim.x)
    {
        const cl_ulong *rec = records + i * PGS_SLOT_WORDS;
        cl_uint     knull = (cl_uint)rec[0] >> 8;

        if (!pgs_gh_merge_state(gs, rec + 1, knull,
                                pgs_hash_keyvals(rec + 1, knull),
                                rec + 1 + GPUPREAGG_NUM_KEYS,
                                (cl_uint)(rec[0] >> 32), ninserted))
            atomicCAS(&kgpreagg->status, StromError_Success,
                      StromError_DataStoreNoSpace);
    }
    if (ninserted)
    {
        atomicAdd(gs.gh_ngroups, ninserted);
        atomicAdd(gs.gh_nused, ninserted);
    }
}

/*
 * gpupreagg_import_blocks - import what ncclAllGather delivered: one block of
 * (1 + cap) records per rank, record 0 = header whose first word is the
 * number of records that follow.  The root merges every block but its own.
 */
extern "C" __global__ void
gpupreagg_import_blocks(pgs_gstate gs, const cl_ulong *blocks, cl_uint nranks,
                        cl_uint root, cl_uint cap, kern_gpupreagg *kgpreagg)
{
    const cl_ulong  block_words = (cl_ulong)(1 + cap) * PGS_SLOT_WORDS;

    if (GPUPREAGG_NUM_KEYS == 0)
    {
        /* one thread, rank order: deterministic */
        if (blockIdx.x == 0 && threadIdx.x == 0)
        {
            cl_uint nn = (cl_uint)gs.ng_state[0];
            for (cl_uint r = 0; r < nranks; r++)
            {
                const cl_ulong *rec = blocks + r * block_words + PGS_SLOT_WORDS;
                cl_uint src_nn = (cl_uint)(rec[0] >> 32);
                if (r == root || (cl_uint)blocks[r * block_words] == 0)
                    continue;
                gpupreagg_aggmerge_plain(gs.ng_state + 1, rec + 1, src_nn);
                nn |= src_nn;
            }
            gs.ng_state[0] = nn;
        }
        return;
    }
    cl_uint     ninserted = 0;

    for (cl_uint r = 0; r < nranks; r++)
    {
        const cl_ulong *recs = blocks + r * block_words + PGS_SLOT_WORDS;
        cl_uint     n = (cl_uint)blocks[r * block_words];

        if (r == root)
            continue;
        if (n > cap)
        {
            atomicCAS(&kgpreagg->status, StromError_Success, StromError_DataStoreNoSpace);
            n = cap;
        }
        for (cl_ulong i = (cl_ulong)blockIdx.x * blockDim.x + threadIdx.x;
             i < n;
             i += (cl_ulong)gridDim.x * blockDim.x)
        {
            const cl_ulong *rec = recs + i * PGS_SLOT_WORDS;
            cl_uint     knull = (cl_uint)rec[0] >> 8;

            if (!pgs_gh_merge_state(gs, rec + 1, knull,
                                    pgs_hash_keyvals(rec + 1, knull),
                                    rec + 1 + GPUPREAGG_NUM_KEYS,
                                    (cl_uint)(rec[0] >> 32), ninserted))
                atomicCAS(&kgpreagg->status, StromError_Success,
                          StromError_DataStoreNoSpace);
        }
    }
    if (ninserted)
    {
        atomicAdd(gs.gh_ngroups, ninserted);
        atomicAdd(gs.gh_nused, ninserted);
    }
}

/* ------------------------------------------------------------------
 * Partitioned exchange of large states (~10 M groups): rank
 * r ends up owning the groups whose key hash maps to r, so that after the
 * exchange the ranks hold disjoint sets of groups and every rank flushes its
 * own share.  Two passes over the state: pass 0 counts the records per
 * destination, the host turns the counts into offsets, pass 1 writes the
 * records bucket by bucket (one buffer, G runs).  The bucket comes from a
 * re-mixed hash: the low bits pick the slot of the global table and the high
 * bits the partition, neither may correlate with the rank.
 * ------------------------------------------------------------------ */
DEVFN cl_uint
pgs_rank_of_hash(cl_ulong hash, cl_uint nranks)
{
    return pgs_fmix32((cl_uint)hash ^ 0x5bd1e995U) % nranks;
}

#define PGS_EXCHANGE_MAX_RANKS  16

extern "C" __global__ void
gpupreagg_export_parts(pgs_gstate gs, cl_ulong *records, cl_uint *counts,
                       const cl_uint *offsets, cl_uint *cursors,
                       cl_uint nranks, cl_uint pass)
{
#if GPUPREAGG_NUM_KEYS > 0
    const cl_ulong  nstate = pgs_state_nslots(gs);
    const cl_uint   lane_id = threadIdx.x & 31;

    for (cl_ulong i0 = (cl_ulong)blockIdx.x * blockDim.x + (threadIdx.x & ~31U);
         i0 < nstate;
         i0 += (cl_ulong)gridDim.x * blockDim.x)
    {
        const cl_ulong  i = i0 + lane_id;
        cl_ulong    keys[PGS_MAX(GPUPREAGG_NUM_KEYS, 1)];
        cl_ulong    cells[PGS_MAX(GPUPREAGG_NUM_CELLS, 1)];
        cl_uint     knull = 0, nn = 0;
        bool        ready = (i < nstate && pgs_state_slot(gs, i, keys, cells, knull, nn));
        cl_uint     dest = 0;
        cl_uint     pos = 0;

        if (ready)
        {
            cl_ulong kv[PGS_MAX(GPUPREAGG_NUM_KEYS, 1)];
#pragma unroll
            for (int k = 0; k < GPUPREAGG_NUM_KEYS; k++)
                kv[k] = ((knull >> k) & 1U) ? 0 : keys[k];
            dest = pgs_rank_of_hash(pgs_hash_keyvals(kv, knull), nranks);
        }
        /* positions are handed out once per warp and destination */
        for (cl_uint d = 0; d < nranks; d++)
        {
            cl_uint votes = __ballot_sync(0xffffffffU, ready && dest == d);
            cl_uint base = 0;

            if (votes == 0)
                continue;
            if (lane_id == 0)
                base = atomicAdd((pass == 0 ? counts : cursors) + d, (cl_uint)__popc(votes));
            base = __shfl_sync(0xffffffffU, base, 0);
            if (ready && dest == d)
                pos = base + __popc(votes & ((1U << lane_id) - 1U));
        }
        if (ready && pass != 0)
        {
            cl_ulong *rec = records + (cl_ulong)(offsets[dest] + pos) * PGS_SLOT_WORDS;

            rec[0] = ((cl_ulong)nn << 32) | (knull << 8) | PGS_SLOT_READY;
#pragma unroll
            for (int k = 0; k < GPUPREAGG_NUM_KEYS; k++)
                rec[1 + k] = keys[k];
#pragma unroll
            for (int c = 0; c < GPUPREAGG_NUM_CELLS; c++)
                rec[1 + GPUPREAGG_NUM_KEYS + c] = cells[c];
        }
    }
#endif
}

/* ------------------------------------------------------------------
 * Merge of SMALL states over NVLink peer memory (no GROUP BY: one record;
 * GROUP BY with a few thousand groups).  No collective library in the path:
 *
 *   every rank but the root : gpupreagg_peer_push writes its state records
 *       straight into the root's exchange area (plain stores through the
 *       NVLink-mapped pointer), resets its own state, and publishes
 *       flag[rank] = epoch with a system-scope release;
 *   the root : gpupreagg_peer_pull waits for the flags (system-scope
 *       acquire), merges the records into its own state with the usual merge
 *       rules and publishes done = epoch.
 *
 * Nothing waits for a collective rendezvous: a rank that is done scanning
 * pushes and goes on with its next chunk; only the root waits, and only for
 * data.  The area holds two buffers per rank (epoch parity); a rank re-uses
 * a buffer only after the root has imported the epoch before last (`done`,
 * read over NVLink).  Exchange area (in the root's HBM, 8-byte words):
 *   [0]              done epoch
 *   [16 + 16 r]      flag of rank r
 *   [PGS_PEER_HEAD_WORDS ...]  2 x nranks blocks of (1 + cap) records; word 0
 *                    of a block's first record = number of records
 * ------------------------------------------------------------------ */
#define PGS_PEER_HEAD_WORDS     (16 + 16 * PGS_EXCHANGE_MAX_RANKS)

DEVFN cl_ulong
pgs_ld_acquire_sys(const cl_ulong *p)
{
    cl_ulong v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
/* polling: a relaxed load per look, ONE acquire when the value is there.
 * A system-scope acquire / fence costs ~10 us on a GPU whose memory its peers
 * have mapped (measured: gpupreagg_peer_pull with one acquire load per rank
 * took 18 us with 2 ranks and 84 us with 8, the flags already set) */
DEVFN cl_ulong
pgs_ld_relaxed_sys(const cl_ulong *p)
{
    cl_ulong v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
DEVFN void
pgs_fence_acquire_sys(void)
{
    asm volatile("fence.acq_rel.sys;" ::: "memory");
}
DEVFN void
pgs_st_release_sys(cl_ulong *p, cl_ulong v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
DEVFN cl_ulong *
pgs_peer_block(cl_ulong *area, cl_uint nranks, cl_uint cap, cl_ulong epoch, cl_uint rank)
{
    return area + PGS_PEER_HEAD_WORDS +
        ((epoch & 1) * nranks + rank) * ((cl_ulong)(1 + cap) * PGS_SLOT_WORDS);
}

/* `local`: two words of this rank's own memory: [0] records written so far,
 * [1] CTA ticket.  `moved` (mapped host memory) is set to 1 when the whole
 * state went over and was reset here - the rank then has nothing to flush.
 * Launched with whole warps. */
extern "C" __global__ void
gpupreagg_peer_push(pgs_gstate gs, cl_ulong *area, cl_uint rank, cl_uint nranks,
                    cl_uint cap, cl_ulong epoch, cl_uint *local, cl_uint *moved)
{
    cl_ulong   *block = pgs_peer_block(area, nranks, cap, epoch, rank);
    cl_ulong   *recs = block + PGS_SLOT_WORDS;
    __shared__ cl_uint is_last;

    /* the buffer was last used two epochs ago: has the root imported that?
     * (the word lives in the root's HBM: a look every microsecond, so that
     * seven waiting ranks do not keep the root's memory busy) */
    if (threadIdx.x == 0)
    {
        while (pgs_ld_relaxed_sys(area) + 2 < epoch)
            __nanosleep(1000);
        pgs_fence_acquire_sys();
    }
    __syncthreads();
    if (GPUPREAGG_NUM_KEYS == 0)
    {
        if (blockIdx.x == 0 && threadIdx.x == 0)
        {
            recs[0] = (gs.ng_state[0] << 32) | PGS_SLOT_READY;
            for (int c = 0; c < GPUPREAGG_NUM_CELLS; c++)
                recs[1 + c] = gs.ng_state[1 + c];
            gs.ng_state[0] = 0;
            pgs_cells_init(gs.ng_state + 1);
            block[0] = 1;
            *moved = 1;         /* (host memory: read after the stream is idle) */
            /* (the release store orders this thread's writes before it) */
            pgs_st_release_sys(area + 16 + 16 * rank, epoch);
        }
        return;
    }
#if GPUPREAGG_NUM_KEYS > 0
    /* more records than a block takes (the estimate of the number of groups
     * was far off): nothing moves, this rank flushes its own partial rows and
     * PostgreSQL's final Agg merges them */
    const cl_uint   nlog = min(*gs.ovf_count, gs.ovf_cap);
    const bool      fits = ((cl_ulong)*gs.gh_nused + nlog <= cap) && gs.part_nparts == 0;
    const cl_ulong  n = (fits ? (cl_ulong)gs.gh_nslots + nlog : 0);
    const cl_uint   lane_id = threadIdx.x & 31;

    for (cl_ulong i0 = (cl_ulong)blockIdx.x * blockDim.x + (threadIdx.x & ~31U);
         i0 < n;
         i0 += (cl_ulong)gridDim.x * blockDim.x)
    {
        const cl_ulong  i = i0 + lane_id;
        const cl_ulong *src = NULL;
        cl_ulong        ctrl = 0;
        bool            ready = false;
        cl_uint         votes, base = 0;

        if (i < gs.gh_nslots)
        {
            cl_ulong *slot = gs.gh_slots + i * PGS_SLOT_STRIDE;

            ctrl = slot[0];
            ready = (((cl_uint)ctrl & 3U) == PGS_SLOT_READY);
            src = slot;
        }
        else if (i < n)
        {
            src = gs.ovf_recs + (i - gs.gh_nslots) * PGS_SLOT_WORDS;
            ctrl = src[0];
            ready = true;
        }
        votes = __ballot_sync(0xffffffffU, ready);
        if (lane_id == 0 && votes != 0)
            base = atomicAdd(local, (cl_uint)__popc(votes));
        base = __shfl_sync(0xffffffffU, base, 0);
        if (ready)
        {
            cl_ulong *rec = recs + (cl_ulong)(base + __popc(votes & ((1U << lane_id) - 1U))) * PGS_SLOT_WORDS;

            /* (a table slot keeps its NULL-key bits from bit 8 like a record) */
            rec[0] = ctrl;
#pragma unroll
            for (int w = 1; w < PGS_SLOT_WORDS; w++)
                rec[w] = src[w];
            if (i < gs.gh_nslots)
            {
                /* the group now lives on the root */
                cl_ulong *slot = gs.gh_slots + i * PGS_SLOT_STRIDE;

                slot[0] = 0;
#pragma unroll
                for (int k = 0; k < GPUPREAGG_NUM_KEYS; k++)
                    slot[1 + k] = 0;
                pgs_cells_init(slot + 1 + GPUPREAGG_NUM_KEYS);
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
        is_last = (atomicAdd(local + 1, 1U) == gridDim.x - 1 ? 1U : 0U);
    __syncthreads();
    if (is_last && threadIdx.x == 0)
    {
        __threadfence();
        block[0] = *((volatile cl_uint *)local);
        if (fits)
        {
            *gs.gh_ngroups = 0;
            *gs.gh_nused = 0;
            *gs.ovf_count = 0;
        }
        local[0] = 0;
        local[1] = 0;
        *moved = (fits ? 1U : 0U);
        /* (every thread fenced its records before the ticket; the release
         * store is cumulative over what this thread has observed) */
        pgs_st_release_sys(area + 16 + 16 * rank, epoch);
    }
#endif
}

extern "C" __global__ void
gpupreagg_peer_pull(pgs_gstate gs, cl_ulong *area, cl_uint root, cl_uint nranks,
                    cl_uint cap, cl_ulong epoch, cl_uint *local,
                    kern_gpupreagg *kgpreagg)
{
    __shared__ cl_uint is_last;
    cl_uint     ninserted = 0;

    /* wait for every rank's flag (local memory, relaxed looks), then ONE
     * system-scope acquire for all of them */
    if (threadIdx.x == 0)
    {
        for (cl_uint r = 0; r < nranks; r++)
            if (r != root)
                while (pgs_ld_relaxed_sys(area + 16 + 16 * r) < epoch)
                    __nanosleep(100);
        pgs_fence_acquire_sys();
    }
    __syncthreads();
    for (cl_uint r = 0; r < nranks; r++)
    {
        const cl_ulong *block = pgs_peer_block(area, nranks, cap, epoch, r);
        const cl_ulong *recs = block + PGS_SLOT_WORDS;
        cl_uint         n;

        if (r == root)
            continue;
        /* written by another GPU: read at the L2, never from this SM's L1 */
        n = min((cl_uint)__ldcg(block), cap);
        if (GPUPREAGG_NUM_KEYS == 0)
        {
            /* one thread, rank order: deterministic */
            if (blockIdx.x == 0 && threadIdx.x == 0 && n > 0)
            {
                cl_ulong    src[PGS_MAX(GPUPREAGG_NUM_CELLS, 1)];
                cl_uint     src_nn = (cl_uint)(__ldcg(recs) >> 32);
#pragma unroll
                for (int c = 0; c < GPUPREAGG_NUM_CELLS; c++)
                    src[c] = __ldcg(recs + 1 + c);
                gpupreagg_aggmerge_plain(gs.ng_state + 1, src, src_nn);
                gs.ng_state[0] = (cl_uint)gs.ng_state[0] | src_nn;
            }
            continue;
        }
#if GPUPREAGG_NUM_KEYS > 0
        for (cl_ulong i = (cl_ulong)blockIdx.x * blockDim.x + threadIdx.x; i < n;
             i += (cl_ulong)gridDim.x * blockDim.x)
        {
            cl_ulong    rec[PGS_SLOT_WORDS];
            cl_uint     knull;

#pragma unroll
            for (int w = 0; w < PGS_SLOT_WORDS; w++)
                rec[w] = __ldcg(recs + i * PGS_SLOT_WORDS + w);
            knull = ((cl_uint)rec[0] >> 8) & ((1U << GPUPREAGG_NUM_KEYS) - 1U);
            if (!pgs_gh_merge_state(gs, rec + 1, knull, pgs_hash_keyvals(rec + 1, knull),
                                    rec + 1 + GPUPREAGG_NUM_KEYS,
                                    (cl_uint)(rec[0] >> 32), ninserted))
                atomicCAS(&kgpreagg->status, StromError_Success, StromError_DataStoreNoSpace);
        }
#endif
    }
    if (ninserted)
    {
        atomicAdd(gs.gh_ngroups, ninserted);
        atomicAdd(gs.gh_nused, ninserted);
    }
    __syncthreads();
    if (threadIdx.x == 0)
        is_last = (atomicAdd(local + 1, 1U) == gridDim.x - 1 ? 1U : 0U);
    __syncthreads();
    if (is_last && threadIdx.x == 0)
    {
        local[1] = 0;
        pgs_st_release_sys(area, epoch);
    }
}

/* ------------------------------------------------------------------
 * gpupreagg_describe - layout constants for the host
 * ------------------------------------------------------------------ */
#define PGS_X_INCOL_ROWBYTES(slot,colidx,attlen)    rb += (attlen);
/* text / bpchar grouping keys (their long values live in the key heap) */
__host__ __device__ constexpr cl_uint pgs_is_text_key(const void *) { return 0; }
#ifdef KERN_TEXTLIB_CUH
__host__ __device__ constexpr cl_uint pgs_is_text_key(const pg_varlena_t *) { return 1; }
#endif
#define PGS_X_KEY_ISTEXT(keyidx,colidx,NAME)    ntk += pgs_is_text_key((const pg_##NAME##_t *)0);

extern "C" __global__ void
gpupreagg_describe(pgs_kern_desc *desc)
{
    cl_uint rb = 0, ntk = 0;

    GPUPREAGG_INCOL_LIST(PGS_X_INCOL_ROWBYTES)
    GPUPREAGG_KEY_LIST(PGS_X_KEY_ISTEXT)
    desc->num_incols = GPUPREAGG_NUM_INCOLS;
    desc->num_keys = GPUPREAGG_NUM_KEYS;
    desc->num_aggs = GPUPREAGG_NUM_AGGS;
    desc->num_cells = GPUPREAGG_NUM_CELLS;
    desc->num_outcols = GPUPREAGG_NUM_OUTCOLS;
    desc->slot_bytes = PGS_SLOT_BYTES;
    desc->slot_stride_bytes = 8 * PGS_SLOT_STRIDE;
    desc->part_rec_bytes = (GPUPREAGG_PARTITIONED ? PGS_REC_BYTES : 0);
    desc->tile_rows = 1024;             /* granule of the tile size */
    desc->num_stages = GPUPREAGG_MAX_STAGES;
    desc->stage_bytes = PGS_STAGE_BYTES(1024);  /* per 1024 rows */
    desc->static_smem_bytes = PGS_SMEM_HEAD_BYTES;
    desc->block_threads = GPUPREAGG_BLOCK_THREADS;
    desc->sh_slot_bytes = PGS_SH_SLOT_BYTES;
    desc->row_bytes = rb;
    /* largest tile worth using: with only the qual's columns staged a tile
     * of the usual byte size holds many more rows */
    desc->max_tile_rows = (GPUPREAGG_GATHER_PAYLOAD ? 8192 : 4096);
    desc->has_qual = GPUPREAGG_HAS_QUAL;
    desc->partagg_head_bytes = PGS_PARTAGG_HEAD_BYTES;
    desc->num_text_keys = ntk;
}

#endif  /* KERN_GPUPREAGG_CUH */
