/*
 * pgstrom_kds.h - chunk layouts shared by host and device code.
 *
 * Byte-for-byte restatement of the structures the reference moves between the
 * PostgreSQL backend and the device (citations are /root/reference/<file>:<line>):
 *
 *   error codes          opencl_common.h:108-123
 *   kern_colmeta         opencl_common.h:335-346
 *   kern_rowitem         opencl_common.h:353-359
 *   kern_blkitem         opencl_common.h:361-369
 *   kern_data_store      opencl_common.h:375-389   (+ access macros :392-434)
 *   kern_parambuf        opencl_common.h:443-457
 *   kern_row_map         opencl_common.h:483-486
 *   kern_gpupreagg       opencl_gpupreagg.h:67-106
 *   field role flags     opencl_gpupreagg.h:135-137
 *
 * One layout is new: KDS_FORMAT_COLUMN (4).  The header is unchanged; after
 * colmeta[] comes one kern_colpos per column that locates a 128-byte aligned
 * value array (attlen > 0: packed values; attlen < 0: uint32 offsets from the
 * head of the chunk to an in-line varlena, 0 = NULL) and an optional
 * validity bitmap (1 bit per row, bit set = NOT NULL, the same sense as
 * HeapTupleHeaderData.t_bits).  It exists because 128-bit coalesced / TMA
 * column loads cannot be done on heap pages.
 *
 * This file is compiled three ways: by gcc/g++ for the host library, by nvcc
 * for the static device code and by NVRTC (no libc headers) for the generated
 * kernels, hence the self-contained typedefs.
 */
#ifndef PGSTROM_KDS_H
#define PGSTROM_KDS_H

#ifdef __CUDACC_RTC__
typedef signed char         cl_char;
typedef unsigned char       cl_uchar;
typedef short               cl_short;
typedef unsigned short      cl_ushort;
typedef int                 cl_int;
typedef unsigned int        cl_uint;
typedef long long           cl_long;
typedef unsigned long long  cl_ulong;
typedef float               cl_float;
typedef double              cl_double;
typedef unsigned long long  hostptr_t;
typedef unsigned long long  Datum;
#ifndef offsetof
#define offsetof(T, f)      ((unsigned long)&(((T *)0)->f))
#endif
#else
#include <stdint.h>
#include <stddef.h>
typedef int8_t      cl_char;
typedef uint8_t     cl_uchar;
typedef int16_t     cl_short;
typedef uint16_t    cl_ushort;
typedef int32_t     cl_int;
typedef uint32_t    cl_uint;
typedef int64_t     cl_long;
typedef uint64_t    cl_ulong;
typedef float       cl_float;
typedef double      cl_double;
typedef uint64_t    hostptr_t;
typedef uint64_t    Datum;
#endif
typedef cl_char     cl_bool;

/* C++ has no flexible array members; one element keeps every offsetof()
 * identical and nothing here depends on sizeof() of these structs */
#ifndef FLEXIBLE_ARRAY_MEMBER
#ifdef __cplusplus
#define FLEXIBLE_ARRAY_MEMBER   1
#else
#define FLEXIBLE_ARRAY_MEMBER
#endif
#endif

/* ---- error codes (opencl_common.h:108-123) ---- */
#define StromError_Success              0   /* nothing to report */
#define StromError_RowFiltered          1   /* the qual rejected the row */
#define StromError_CpuReCheck           2   /* the host has to evaluate this itself */
#define StromError_ServerNotReady       100 /* device layer is not ready */
#define StromError_BadRequestMessage    101 /* malformed request / arguments */
#define StromError_OpenCLInternal       102 /* internal error of the device runtime */
#define StromError_OutOfSharedMemory    105 /* out of pinned host memory */
#define StromError_OutOfMemory          106 /* out of host memory */
#define StromError_DataStoreCorruption  300 /* a chunk fails its consistency checks */
#define StromError_DataStoreNoSpace     301 /* destination store / table is full */
#define StromError_DataStoreOutOfRange  302 /* row or column index beyond the chunk */
#define StromError_DataStoreReCheck     303 /* whole chunk goes back to the host */
#define StromError_SanityCheckViolation 999 /* an internal invariant does not hold */
/* new in the CUDA layer: device program build failure (the reference reports
 * CL_BUILD_PROGRAM_FAILURE = -11 here, gpupreagg.c:2751-2764) */
#define StromError_ProgramBuildFailure  (-11)
#define StromError_CudaInternal         (-9999)

/* codes the host must turn into an ERROR (everything but 0, 1, 2) */
#define StromErrorIsSignificant(errcode)    ((errcode) >= 100 || (errcode) < 0)

/* ---- alignment (opencl_common.h:272-274) ---- */
#define STROMALIGN_LEN          16
/* (PostgreSQL's c.h has its own TYPEALIGN family: inside a backend those win) */
#ifndef TYPEALIGN
#define TYPEALIGN(ALIGNVAL,LEN) \
    (((cl_ulong)(LEN) + ((ALIGNVAL) - 1)) & ~((cl_ulong)((ALIGNVAL) - 1)))
#endif
#ifndef TYPEALIGN_DOWN
#define TYPEALIGN_DOWN(ALIGNVAL,LEN) \
    (((cl_ulong)(LEN)) & ~((cl_ulong)((ALIGNVAL) - 1)))
#endif
#define STROMALIGN(LEN)         TYPEALIGN(STROMALIGN_LEN,(LEN))
#define STROMALIGN_DOWN(LEN)    TYPEALIGN_DOWN(STROMALIGN_LEN,(LEN))
#ifndef LONGALIGN
#define LONGALIGN(LEN)          TYPEALIGN(8,(LEN))
#endif
#ifndef INTALIGN
#define INTALIGN(LEN)           TYPEALIGN(4,(LEN))
#endif
#ifndef MAXALIGN
#define MAXALIGN(LEN)           TYPEALIGN(8,(LEN))
#endif
#ifndef BLCKSZ
#define BLCKSZ                  8192
#endif
/* alignment of every array inside a KDS_FORMAT_COLUMN chunk: one L2 line,
 * which also satisfies the 16-byte rule of cp.async.bulk */
#define KDS_COLUMN_ALIGN        128
/* rows per staged tile must keep every per-tile slice 16-byte aligned, also
 * for 1-bit-per-row bitmaps: a multiple of 128 rows */
#define KDS_COLUMN_ROW_QUANTUM  128

typedef struct {
    cl_char         attbyval;   /* != 0: the datum is the value; 0: it points to it */
    cl_char         attalign;   /* in bytes (1, 2, 4, 8) - not pg_attribute's letter */
    cl_short        attlen;     /* bytes, or < 0 for varlena as in pg_attribute */
    cl_short        attnum;     /* 1-based column number of the source relation */
    cl_short        attcacheoff;/* fixed offset inside a NULL-free tuple, or -1 */
} kern_colmeta;

typedef union {
    struct {
        cl_ushort   blk_index;      /* KDS_FORMAT_ROW: which page ... */
        cl_ushort   item_offset;    /* ... and where in it the tuple starts */
    };
    cl_uint         htup_offset;    /* KDS_FORMAT_ROW_FLAT: from the head of the chunk */
} kern_rowitem;

typedef struct {
    cl_int          buffer;         /* shared-buffer id the backend pinned */
    hostptr_t       page;           /* where that page lives in host memory */
} kern_blkitem;

#define KDS_FORMAT_ROW          1
#define KDS_FORMAT_ROW_FLAT     2
#define KDS_FORMAT_TUPSLOT      3
#define KDS_FORMAT_COLUMN       4   /* new: see the head of this file */

typedef struct {
    hostptr_t       hostptr;    /* the chunk's own host address (for the way back) */
    cl_uint         length;     /* bytes, head included */
    cl_uint         usage;      /* bytes taken so far (ROW_FLAT grows from the tail) */
    cl_uint         ncols;      /* entries of colmeta[] */
    cl_uint         nitems;     /* rows present */
    cl_uint         nrooms;     /* rows the chunk was sized for */
    cl_uint         nblocks;    /* pages present (KDS_FORMAT_ROW) */
    cl_uint         maxblocks;  /* pages the chunk was sized for */
    cl_char         format;     /* KDS_FORMAT_* */
    cl_char         tdhasoid;   /* the three TupleDesc fields a heap tuple's */
    cl_uint         tdtypeid;   /*   header depends on, taken over from the */
    cl_int          tdtypmod;   /*   relation's descriptor */
    kern_colmeta    colmeta[FLEXIBLE_ARRAY_MEMBER];
} kern_data_store;

#define KERN_DATA_STORE_HEAD_LENGTH(ncols)                      \
    STROMALIGN(offsetof(kern_data_store, colmeta) +             \
               sizeof(kern_colmeta) * (ncols))

/* KDS_FORMAT_ROW: block items, row items, then the pages on a BLCKSZ boundary */
#define KERN_DATA_STORE_BLKITEM(kds,blk_index)                  \
    (((kern_blkitem *)                                          \
      ((char *)(kds) + KERN_DATA_STORE_HEAD_LENGTH((kds)->ncols))) + (blk_index))
#define KERN_DATA_STORE_ROWITEM(kds,row_index)                  \
    (((kern_rowitem *)                                          \
      ((char *)(kds) + KERN_DATA_STORE_HEAD_LENGTH((kds)->ncols) + \
       STROMALIGN(sizeof(kern_blkitem) * (kds)->maxblocks))) + (row_index))
#define KERN_DATA_STORE_ROWBLOCK(kds,blk_index)                 \
    ((char *)(kds) +                                            \
     (TYPEALIGN(BLCKSZ,                                         \
                KERN_DATA_STORE_HEAD_LENGTH((kds)->ncols) +     \
                STROMALIGN(sizeof(kern_blkitem) * (kds)->maxblocks) + \
                STROMALIGN(sizeof(kern_rowitem) * (kds)->nitems)) \
      + (cl_ulong)BLCKSZ * (blk_index)))

/* KDS_FORMAT_TUPSLOT: per row ncols Datums followed by ncols isnull bytes */
#define KERN_DATA_STORE_SLOT_STRIDE(ncols)                      \
    LONGALIGN((sizeof(Datum) + sizeof(cl_char)) * (ncols))
#define KERN_DATA_STORE_VALUES(kds,row_index)                   \
    ((Datum *)((char *)(kds) +                                  \
               KERN_DATA_STORE_HEAD_LENGTH((kds)->ncols) +      \
               KERN_DATA_STORE_SLOT_STRIDE((kds)->ncols) * (cl_ulong)(row_index)))
#define KERN_DATA_STORE_ISNULL(kds,row_index)                   \
    ((cl_char *)(KERN_DATA_STORE_VALUES((kds),(row_index)) + (kds)->ncols))

/* access for the column format */
typedef struct {
    cl_uint         values_offset;  /* from the head of kds; KDS_COLUMN_ALIGN-ed */
    cl_uint         nullmap_offset; /* 0, if the column has no NULL in this chunk */
} kern_colpos;

#define KERN_DATA_STORE_COLPOS(kds,colidx)                      \
    (((kern_colpos *)                                           \
      ((char *)(kds) + KERN_DATA_STORE_HEAD_LENGTH((kds)->ncols))) + (colidx))
#define KERN_DATA_STORE_COLUMN_HEAD_LENGTH(ncols)               \
    TYPEALIGN(KDS_COLUMN_ALIGN,                                 \
              KERN_DATA_STORE_HEAD_LENGTH(ncols) + sizeof(kern_colpos) * (ncols))

/* ---- kern_parambuf (opencl_common.h:443-457) ---- */
typedef struct {
    cl_uint     length;     /* bytes, this head included */
    cl_uint     nparams;    /* entries of poffset[] */
    cl_uint     poffset[FLEXIBLE_ARRAY_MEMBER]; /* where each value starts; 0 = NULL */
} kern_parambuf;

/* ---- kern_row_map (opencl_common.h:483-486) ---- */
typedef struct {
    cl_int      nvalids;    /* entries of rindex[]; < 0: every row of the chunk counts */
    cl_int      rindex[FLEXIBLE_ARRAY_MEMBER];
} kern_row_map;

/* ---- kern_gpupreagg (opencl_gpupreagg.h:67-106) ---- */
typedef struct {
    cl_int          status;     /* StromError_* the kernels leave behind */
    cl_int          sortbuf_len;/* unused by the hash based kernels; kept for layout */
    char            __padding[8];   /* kparams starts on a 16-byte boundary */
    kern_parambuf   kparams;
    /* kern_row_map follows at STROMALIGN(offsetof(kparams) + kparams.length) */
} kern_gpupreagg;

#define KERN_GPUPREAGG_PARAMBUF(kgpreagg)   (&(kgpreagg)->kparams)
#define KERN_GPUPREAGG_KROWMAP(kgpreagg)                        \
    ((kern_row_map *)((char *)(kgpreagg) +                      \
                      STROMALIGN(offsetof(kern_gpupreagg, kparams) + \
                                 (kgpreagg)->kparams.length)))

/* KPARAM_0 of GpuPreAgg: one role byte per output column
 * (opencl_gpupreagg.h:135-137) */
#define GPUPREAGG_FIELD_IS_NULL         0
#define GPUPREAGG_FIELD_IS_GROUPKEY     1
#define GPUPREAGG_FIELD_IS_AGGFUNC      2

/* ---- device NUMERIC (opencl_numeric.h:141-162): 6-bit exp10 / sign / 57-bit mantissa */
#define PG_NUMERIC_EXPONENT_BITS    6
#define PG_NUMERIC_EXPONENT_POS     58
#define PG_NUMERIC_EXPONENT_MAX     ((1 << ((PG_NUMERIC_EXPONENT_BITS) - 1)) - 1)
#define PG_NUMERIC_EXPONENT_MIN     (0 - (1 << ((PG_NUMERIC_EXPONENT_BITS) - 1)))
#define PG_NUMERIC_SIGN_POS         57
#define PG_NUMERIC_SIGN_MASK        (1ULL << PG_NUMERIC_SIGN_POS)
#define PG_NUMERIC_MANTISSA_BITS    57
#define PG_NUMERIC_MANTISSA_MASK    ((1ULL << PG_NUMERIC_MANTISSA_BITS) - 1)
#define PG_NUMERIC_MANTISSA_MAX     PG_NUMERIC_MANTISSA_MASK
#define PG_NUMERIC_EXPONENT(num)    ((cl_long)(num) >> 58)
#define PG_NUMERIC_SIGN(num)        (((num) & PG_NUMERIC_SIGN_MASK) != 0)
#define PG_NUMERIC_MANTISSA(num)    ((num) & PG_NUMERIC_MANTISSA_MASK)
#define PG_NUMERIC_SET(expo,sign,mant)                          \
    ((cl_ulong)((cl_ulong)((cl_long)(expo)) << 58) |            \
     ((sign) != 0 ? PG_NUMERIC_SIGN_MASK : 0ULL) |              \
     ((cl_ulong)(mant) & PG_NUMERIC_MANTISSA_MASK))

/* PostgreSQL type OIDs the device code knows (codegen.c:46-78) */
#define BOOLOID         16
#define BYTEAOID        17
#define INT8OID         20
#define INT2OID         21
#define INT4OID         23
#define TEXTOID         25
#define FLOAT4OID       700
#define FLOAT8OID       701
#define BPCHAROID       1042
#define DATEOID         1082
#define TIMEOID         1083
#define TIMESTAMPOID    1114
#define NUMERICOID      1700

#endif  /* PGSTROM_KDS_H */
