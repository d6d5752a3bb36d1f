/*
 * kern_shared.h - structures passed by value between the CUDA layer (host)
 * and the generated device programs.  Included by both sides; plain C.
 */
#ifndef KERN_SHARED_H
#define KERN_SHARED_H

typedef struct
{
    /* no-group: the persistent state row + per-CTA partials of one launch */
    cl_ulong   *ng_state;       /* [1 + NCELLS]: word0 = nn bits, then cells */
    cl_ulong   *ng_partial;     /* [max_ctas][1 + NCELLS] */
    cl_uint    *ng_ticket;      /* CTA completion counter of the launch */
    /* group-by: global open-addressing table */
    cl_ulong   *gh_slots;       /* [gh_nslots][PGS_SLOT_WORDS] */
    cl_uint     gh_nslots;      /* power of 2 */
    cl_uint     gh_max_probe;
    cl_uint    *gh_ngroups;     /* groups inserted anywhere (table, images) */
    cl_uint    *gh_nused;       /* READY slots of the global table alone */
    /* overflow log: state records (the exported record format) that found no
     * room in the global table - a CTA-local table spilled at the end of a
     * launch, or an imported record, when the planner's estimate of the
     * number of groups was too low.  They are part of the state (flush and
     * export walk them; PostgreSQL's final Agg merges partial rows of one
     * key), and the host folds them into a larger table as soon as it sees
     * the counter move (session_grow_table). */
    cl_ulong   *ovf_recs;       /* [ovf_cap][PGS_SLOT_WORDS] */
    cl_uint    *ovf_count;
    cl_uint     ovf_cap;
    cl_uint     ovf_pad;
    /* group-by with very many groups: rows are first dealt into partitions
     * (by the high bits of the key hash), then every partition is
     * aggregated in shared memory into its own persistent table image */
    cl_uint     part_nparts;    /* 0 = off */
    cl_uint     part_cap;       /* records one partition takes per chunk */
    cl_uint     part_slots;     /* slots of a table image (multiple of 32) */
    cl_uint     part_pad;
    cl_uint    *part_cursor;    /* [nparts] records of the current chunk */
    cl_uint    *part_nused;     /* [nparts] used slots of the image */
    unsigned char *part_recs;   /* [nparts][cap] records */
    unsigned char *part_images; /* [nparts] images, PGS_SH_SLOT_BYTES * slots each */
    /* segment mode of the deal pass (gpupreagg_main on column chunks): the
     * record area of a partition is cut into one segment per CTA of the scan
     * kernel, whose cursors live in that CTA's shared memory - no global
     * atomic per row.  part_cap = part_seg_max * part_seg_cap. */
    cl_uint     part_seg_cap;   /* records per segment; 0 = off */
    cl_uint     part_seg_max;   /* segments per partition the area is cut into */
    cl_ushort  *part_seg_counts;/* [part_seg_max][nparts] records of the current chunk */
    /* bookkeeping */
    cl_ulong   *nrows_scanned;  /* rows that passed visibility (row-map) */
    cl_ulong   *nrows_filtered; /* rows removed by the device qual */
} pgs_gstate;

/* what the host needs to know about a built program; filled by the
 * gpupreagg_describe kernel so that the device code is the only source of
 * truth for the layouts */
typedef struct
{
    cl_uint     num_incols;
    cl_uint     num_keys;
    cl_uint     num_aggs;
    cl_uint     num_cells;
    cl_uint     num_outcols;
    cl_uint     slot_bytes;
    cl_uint     tile_rows;
    cl_uint     num_stages;
    cl_uint     stage_bytes;        /* shared memory of one pipeline stage */
    cl_uint     static_smem_bytes;  /* barriers etc. at the head of dynamic smem */
    cl_uint     block_threads;
    cl_uint     sh_slot_bytes;      /* bytes per slot of the CTA-local table */
    cl_uint     row_bytes;          /* algorithmic bytes per row (attlen sum) */
    cl_uint     slot_stride_bytes;  /* distance of two slots of the global table */
    cl_uint     part_rec_bytes;     /* bytes of a partition record */
    cl_uint     max_tile_rows;      /* upper bound of the tile size (multiple of 1024) */
    cl_uint     has_qual;           /* the scan evaluates a WHERE clause */
    cl_uint     partagg_head_bytes; /* shared memory of gpupreagg_partagg in front of the image */
    cl_uint     num_text_keys;      /* text / bpchar grouping keys: the session gets a key heap */
} pgs_kern_desc;

/* key heap of a session whose program groups by text / bpchar columns
 * (kern_textlib.cuh): the control block is the module global `pgs_keyheap` */
typedef struct
{
    cl_ulong   *slots;          /* [nslots][2]: hash (0 = free), ref */
    unsigned char *heap;
    cl_ulong   *heap_used;      /* allocation cursor, bytes */
    cl_ulong    heap_bytes;
    cl_uint     nslots;         /* power of 2; 0 = no key heap */
    cl_uint     max_probe;
} pgs_keyheap_ctl;

#endif  /* KERN_SHARED_H */
