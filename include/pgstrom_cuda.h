/*
 * pgstrom_cuda.h - C ABI of libpgstrom_cuda.so
 *
 * The drop-in boundary of the GpuPreAgg path.  Plain C, no PostgreSQL and no
 * torch types.  Every entry point names the reference interface it replaces
 * (file:line under /root/reference).  INTEGRATION.md shows the glue a
 * PostgreSQL extension writes on top of it.
 *
 * Layers:
 *   1. GUCs / error strings            main.c:104-234, main.c:288-330
 *   2. planner half (host only)        grafter.c, gpupreagg.c:134-2187, codegen.c
 *   3. chunk builders                  datastore.c:41-148, 312-529, 718-828
 *   4. device programs (NVRTC)         opencl_devprog.c:270-659
 *   5. GpuPreAgg sessions (CUDA)       gpupreagg.c:2329-2416 (create message),
 *                                      :3849-4240 (clserv_process_gpupreagg),
 *                                      :3009-3194 (completion), mqueue.c:140-430,
 *                                      opencl_serv.c:76-215
 *   6. executor half (host + CUDA)     gpupreagg.c:2189-2941
 *   7. SQL-side functions (host only)  gpupreagg.c:4251-4773, pg_strom--1.0.sql:99-401
 *
 * Convention: functions return a StromError_* code (pgstrom_kds.h), 0 = OK;
 * pgs_last_error() gives a thread-local detail message.  CpuReCheck (2) is a
 * *minor* status: the call succeeded but some rows have to be evaluated by
 * the host (see pgs_preagg_recheck_rows).
 */
#ifndef PGSTROM_CUDA_H
#define PGSTROM_CUDA_H

#include <stddef.h>
#include <stdint.h>
#include "pgstrom_kds.h"

#ifdef __cplusplus
extern "C" {
#endif

#define PGSTROM_CUDA_ABI_VERSION    1

/* ------------------------------------------------------------------ 1 --- */
/* main.c:288 pgstrom_strerror() */
const char *pgstrom_strerror(int errcode);
const char *pgs_last_error(void);
int         pgstrom_abi_version(void);

/* GUC table kept verbatim (main.c:104-234, gpupreagg.c:2946-2967,
 * gpuscan.c:1704, mqueue.c:740, opencl_devprog.c:929-948, shmem.c:1432-1452,
 * opencl_serv.c:408, opencl_devinfo.c:1024-1070).  The glue registers each
 * with DefineCustom*Variable and forwards assignments here. */
int         pgstrom_guc_set(const char *name, const char *value);
const char *pgstrom_guc_get(const char *name);      /* NULL if unknown */
const char *pgstrom_guc_list_json(void);            /* [{name,kind,boot,...}] */
void        pgstrom_guc_reset_all(void);

/* ------------------------------------------------------------------ 2 --- */
typedef struct pgs_plan pgs_plan;

/* grafter.c:119-149 pgstrom_grafter_entrypoint(): takes the finished plan
 * tree (JSON), tries pgstrom_try_insert_gpupreagg() on
 * every Agg node (gpupreagg.c:1987) and returns the possibly rewritten tree. */
pgs_plan   *pgstrom_grafter_json(const char *plan_tree_json);
void        pgs_plan_free(pgs_plan *plan);
const char *pgs_plan_tree_json(pgs_plan *plan);
/* main.c:336-439 + gpupreagg.c:2859-2877: EXPLAIN text, lines joined by \n */
const char *pgs_plan_explain(pgs_plan *plan, int verbose);
int         pgs_plan_num_gpupreagg(pgs_plan *plan);
const char *pgs_plan_reject_reason(pgs_plan *plan);
/* per GpuPreAgg node `idx`: */
const char *pgs_plan_kernel_source(pgs_plan *plan, int idx);
int         pgs_plan_extra_flags(pgs_plan *plan, int idx);
const void *pgs_plan_kparams(pgs_plan *plan, int idx, size_t *length);
int         pgs_plan_needs_grouping(pgs_plan *plan, int idx);
double      pgs_plan_num_groups(pgs_plan *plan, int idx);
/* JSON catalogue of the GpuPreAgg target list: [{resno, role, func, type,
 * op, cell_type, text}] and of the rewritten Agg target list */
const char *pgs_plan_describe_json(pgs_plan *plan, int idx);
/* colmeta of the GpuPreAgg result store (TUPSLOT, internal_format=true:
 * numeric is attlen=8 byval, datastore.c:355-363) */
int         pgs_plan_result_colmeta(pgs_plan *plan, int idx,
                                    kern_colmeta *colmeta, int max_cols);

/* codegen.c:1631 pgstrom_codegen_available_expression() */
int         pgstrom_codegen_available_expression_json(const char *expr_json);

/* ------------------------------------------------------------------ 3 --- */
/* datastore.c:312-380 init_kern_data_store(): fills the header + colmeta */
size_t      pgstrom_kds_head_length(int ncols);
/* KDS_FORMAT_COLUMN builder.  values[c]: attlen>0 -> packed array of
 * nrows * attlen bytes; attlen<0 -> array of nrows pointers to varlena datums
 * (NULL pointer = SQL NULL).  isnull[c]: one byte per row or NULL. */
size_t      pgstrom_kds_column_length(int ncols, const kern_colmeta *colmeta,
                                      uint32_t nrows,
                                      const void *const *values,
                                      const uint8_t *const *isnull);
int         pgstrom_kds_column_build(void *buffer, size_t buflen,
                                     int ncols, const kern_colmeta *colmeta,
                                     uint32_t nrows,
                                     const void *const *values,
                                     const uint8_t *const *isnull);
/* datastore.c:501-529 pgstrom_create_data_store_tupslot() */
/* datastore.c:382-435 pgstrom_create_data_store_row(): the host form of a
 * KDS_FORMAT_ROW chunk = head + kern_blkitem[maxblocks] + kern_rowitem[nrooms];
 * the pages stay where they are (bitem->page) */
size_t      pgstrom_kds_row_length(int ncols, uint32_t maxblocks, uint32_t nrooms);
int         pgstrom_kds_row_init(void *buffer, size_t buflen, int ncols,
                                 const kern_colmeta *colmeta,
                                 uint32_t maxblocks, uint32_t nrooms);
/* datastore.c:556-710 pgstrom_data_store_insert_block(); visibility is the
 * caller's (PostgreSQL's) job: visible_offsets[] = OffsetNumbers of the
 * tuples the snapshot sees.  Returns rows added, -1 = store full. */
int         pgstrom_kds_row_insert_block(kern_data_store *kds, const void *page,
                                         const uint16_t *visible_offsets,
                                         int nvisible);
/* datastore.c:437-470, :799-823: KDS_FORMAT_ROW_FLAT */
int         pgstrom_kds_flat_init(void *buffer, size_t buflen, int ncols,
                                  const kern_colmeta *colmeta, uint32_t nrooms);
int         pgstrom_kds_flat_insert_tuple(kern_data_store *kds, const void *htup,
                                          uint32_t t_len);
/* TupleDesc's attcacheoff / attnum for a colmeta[] built by hand */
void        pgstrom_colmeta_set_cacheoff(int ncols, kern_colmeta *colmeta);
/* synthetic heap pages for benchmarks and tests (heap_form_tuple +
 * PageAddItem); see datastore.cpp */
long        pgstrom_heap_form_pages(int ncols, const kern_colmeta *colmeta,
                                    uint32_t nrows, const void *const *values,
                                    const unsigned char *const *isnull,
                                    const unsigned char *varlena_blob,
                                    unsigned char *pages, size_t maxpages,
                                    uint32_t *rows_per_page);
size_t      pgstrom_kds_tupslot_length(int ncols, uint32_t nrooms);
int         pgstrom_kds_tupslot_init(void *buffer, size_t buflen, int ncols,
                                     const kern_colmeta *colmeta,
                                     uint32_t nrooms);
/* datastore.c:169-242 pgstrom_fetch_data_store() for TUPSLOT */
int         pgstrom_fetch_data_store(const kern_data_store *kds, uint32_t row,
                                     Datum *values, char *isnull);
/* datastore.c:150-167 pgstrom_fixup_kernel_numeric(): 64-bit device numeric
 * -> decimal text "<sign><mantissa>e<exp>" for numeric_in() */
int         pgstrom_fixup_kernel_numeric(Datum datum, char *buf, size_t buflen);
/* opencl_gpupreagg.h:326-366 (varlena grouping keys of the result):
 * a text / bpchar grouping key comes back by value as "kernel text" (<= 7
 * bytes + length in one Datum); -> varlena with a 4-byte header, bpchar padded
 * to `typmod` (atttypmod, -1 = none).  Returns the size, 0 if buf is too small */
size_t      pgstrom_fixup_kernel_text(Datum datum, int typmod, void *buf, size_t buflen);
/* the same for any key: a text / bpchar key of more than 7 bytes comes back
 * as a word of the session's key heap (pgs_preagg_key_heap(), section 5) -
 * top byte 0x80, below it the offset of the string in that heap.  This is the
 * reference's fix-up of varlena key pointers to host addresses
 * (opencl_gpupreagg.h:326-366, pg_fixup_tupslot_varlena opencl_common.h:1060-1100) */
size_t      pgstrom_fixup_kernel_text_heap(Datum datum, int typmod,
                                           const void *key_heap, size_t key_heap_len,
                                           void *buf, size_t buflen);
/* PostgreSQL numeric varlena <-> decimal text (for numeric Consts and for
 * harnesses without a PostgreSQL to make datums); return the length, 0 on
 * error */
size_t      pgstrom_numeric_from_text(const char *text, void *buf, size_t buflen);
size_t      pgstrom_numeric_to_text(const void *varlena, char *buf, size_t buflen);

/* ------------------------------------------------------------------ 4 --- */
typedef struct pgs_program pgs_program;

/* opencl_devprog.c:580-659 pgstrom_get_devprog_key() +
 * :270-569 clserv_lookup_device_program(): programs are cached by
 * CRC32(source, extra_flags); building needs NVRTC only (no GPU).  On a
 * build failure returns StromError_ProgramBuildFailure and *build_log (owned
 * by the library, valid until the next build on this thread) carries the
 * compiler output together with the source, like gpupreagg.c:2751-2764. */
int         pgs_program_build(const char *kern_source, int extra_flags,
                              pgs_program **program, const char **build_log);
void        pgs_program_release(pgs_program *program);
const void *pgs_program_cubin(pgs_program *program, size_t *length);
/* admin view, pg_strom--1.0.sql:62-72 pgstrom_opencl_program_info() */
const char *pgs_program_info_json(void);

/* ------------------------------------------------------------------ 5 --- */
/* opencl_devinfo.c / opencl_serv.c startup: pick the CUDA devices this
 * process may use (GUC pg_strom.opencl_devices is re-mapped to this list).
 * devices == NULL: all visible devices. */
int         pgs_cuda_init(const int *devices, int ndevices);
int         pgs_cuda_device_count(void);
/* pg_strom--1.0.sql:47-60 pgstrom_opencl_device_info() */
const char *pgs_cuda_device_info_json(void);
void        pgs_cuda_shutdown(void);

/* pinned host memory for chunks (replaces shmem.c zones that the OpenCL
 * server registered with CL_MEM_USE_HOST_PTR, opencl_serv.c:115-215) */
void       *pgs_chunk_alloc(size_t length);
void        pgs_chunk_free(void *chunk);

typedef struct pgs_session pgs_session;

typedef struct {
    int         device;             /* index into the pgs_cuda_init() list */
    int         needs_grouping;     /* pgstrom_gpupreagg.needs_grouping */
    double      num_groups;         /* pgstrom_gpupreagg.num_groups (estimate) */
    int         max_async_chunks;   /* pg_strom.max_async_chunks, 0 = GUC */
    uint32_t    max_chunk_rows;     /* upper bound of kds->nitems, 0 = 64M */
    size_t      max_chunk_bytes;    /* upper bound of kds->length, 0 = from GUC */
    int         result_ncols;
    const kern_colmeta *result_colmeta;
} pgs_session_config;

/* gpupreagg.c:2189-2308 gpupreagg_begin(): device program, kparams and the
 * persistent per-device aggregation state */
int         pgs_preagg_open(pgs_program *program,
                            const kern_parambuf *kparams,
                            const pgs_session_config *config,
                            pgs_session **session);
/* gpupreagg.c:2329-2416 pgstrom_create_gpupreagg() + mqueue.c:140
 * pgstrom_enqueue_message(): asynchronous; the chunk must stay valid until
 * the ticket completes.  krowmap may be NULL (all rows visible). */
typedef int64_t pgs_ticket;
int         pgs_preagg_submit(pgs_session *session,
                              const kern_data_store *kds_in,
                              const kern_row_map *krowmap,
                              pgs_ticket *ticket);
/* the chunk is already in device memory (device-resident measurements,
 * chunks produced by another GPU operator) */
int         pgs_preagg_submit_device(pgs_session *session,
                                     const void *kds_in_device, size_t length,
                                     uint32_t nitems,
                                     const kern_row_map *krowmap,
                                     pgs_ticket *ticket);
/* same for a device chunk of any input format (KDS_FORMAT_ROW, ROW_FLAT or
 * COLUMN); the plain call above means KDS_FORMAT_COLUMN */
int         pgs_preagg_submit_device_format(pgs_session *session,
                                            const void *kds_in_device, size_t length,
                                            uint32_t nitems, int format,
                                            const kern_row_map *krowmap,
                                            pgs_ticket *ticket);
/* mqueue.c:331-410 pgstrom_dequeue_message / try_dequeue: returns the chunk
 * status (0, StromError_CpuReCheck, or a significant error).
 * timeout_ms < 0 waits for ever, 0 polls (returns -1 if still running). */
int         pgs_preagg_wait(pgs_session *session, pgs_ticket ticket,
                            int timeout_ms, int32_t *status);
/* rows of a CpuReCheck chunk the host has to evaluate itself
 * (gpupreagg.c:2507-2607 gpupreagg_next_tuple_fallback).  Returns the number
 * of rows, copies up to max_rows indices. */
int64_t     pgs_preagg_recheck_rows(pgs_session *session, pgs_ticket ticket,
                                    uint32_t *rows, int64_t max_rows);
/* end of scan (or whenever the caller wants the partial rows so far):
 * drains outstanding chunks, writes the state as a TUPSLOT store.
 * kds_dst: caller buffer initialised by pgstrom_kds_tupslot_init(); if it is
 * too small the call returns StromError_DataStoreNoSpace and *nrows_needed
 * tells how many rooms are required.  reset != 0 clears the state. */
int         pgs_preagg_finish(pgs_session *session, kern_data_store *kds_dst,
                              int reset, uint32_t *nrows_needed,
                              int32_t *status);
/* Long text / bpchar grouping keys (opencl_gpupreagg.h:326-366: the
 * reference keeps a varlena key as an offset into the chunk's toast area and
 * fixes the pointers up for the host): a program that groups by text columns
 * gets a key heap in HBM (GUC pg_strom.key_heap_size, MB; lookup table sized
 * from the planner's group estimate).  A key of more than 7 bytes is stored
 * there once and travels as an 8-byte word.  After pgs_preagg_finish() this
 * returns the host copy of the heap the words of the returned rows point
 * into (valid until the next finish / close; NULL, 0 when no long key was
 * seen) - pass it to pgstrom_fixup_kernel_text_heap().  A full heap turns
 * further rows with unseen long keys into CpuReCheck rows.  States with a key
 * heap are session-local: the merge / export / import calls below refuse
 * them (StromError_BadRequestMessage) and every device returns its own
 * partial rows to PostgreSQL's final Agg. */
int         pgs_preagg_key_heap(pgs_session *session, const void **heap,
                                size_t *heap_len);
/* NCCL communicator for one-process-per-GPU deployments: rank 0 makes the
 * id, the launcher (MPI, torch.distributed, PostgreSQL shared memory ...)
 * hands its 128 bytes to every rank, each rank joins.  libnccl.so.2 is
 * opened lazily; single-GPU use does not need it. */
int         pgs_nccl_get_unique_id(void *unique_id_128);
int         pgs_nccl_comm_init_rank(int device, int nranks,
                                    const void *unique_id_128, int rank,
                                    void **comm);
void        pgs_nccl_comm_destroy(void *comm);
/* multi-GPU: merge the states of all ranks into rank `root` over NCCL
 * (ncclReduce-like for no-group, gather + re-hash for GROUP BY).  Collective:
 * every rank of the communicator calls it.  comm is an ncclComm_t. */
int         pgs_preagg_merge_nccl(pgs_session *session, void *nccl_comm,
                                  int rank, int nranks, int root);
/* The same merge without a collective library in the data path, for states
 * of at most 64K groups (and the single record of a no-group aggregation):
 * the other ranks write their state records straight into an exchange area
 * in the root's HBM over NVLink peer memory and raise a flag, the root's
 * merge kernel waits for the flags on the device.  No rendezvous, no host
 * synchronisation: a rank pushes as soon as its own scan is done.
 *   setup  : every rank, once; the root's ipc_handle_64 (cudaIpcMemHandle_t)
 *            is carried to the other ranks by the launcher like the NCCL id
 *   attach : every other rank maps the root's area (another process: by
 *            handle; the same process: by session)
 *   merge  : every rank, once per scan, before pgs_preagg_finish()
 * A rank whose state does not fit pushes nothing and keeps its groups: its
 * own flush returns them and PostgreSQL's final Agg merges the partial rows,
 * as the reference does for every chunk (gpupreagg.c:2169-2186).
 * A rank whose whole state went over has nothing left to flush: its
 * pgs_preagg_finish() waits for its own chunks and returns zero rows without
 * launching anything (the push kernel has already reset the state), so the
 * other ranks run ahead of the root instead of holding it up. */
int         pgs_preagg_peer_setup(pgs_session *session, int rank, int nranks,
                                  int root, void *ipc_handle_64);
int         pgs_preagg_peer_attach(pgs_session *session,
                                   const void *root_ipc_handle_64);
int         pgs_preagg_peer_attach_session(pgs_session *session,
                                           pgs_session *root_session);
int         pgs_preagg_merge_peer(pgs_session *session);
/* Large GROUP BY states (millions of groups) are not gathered on one rank:
 * the groups are partitioned over the ranks by key hash (rank r receives
 * the records of its groups from every rank over NCCL send/recv and merges
 * them), after which every rank flushes its own, disjoint share of the
 * groups.  Collective.  pgs_preagg_merge_nccl() takes this path by itself
 * when the state is too large for its gather. */
int         pgs_preagg_merge_exchange(pgs_session *session, void *nccl_comm,
                                      int rank, int nranks);
/* raw state export / import for callers that move it themselves */
int         pgs_preagg_state_export(pgs_session *session, void *device_buf,
                                    size_t buflen, uint32_t *nrecords,
                                    size_t *record_bytes);
int         pgs_preagg_state_import(pgs_session *session, const void *device_buf,
                                    uint32_t nrecords);
int         pgs_preagg_state_reset(pgs_session *session);
/* perfmon counters of pgstrom_perfmon (pg_strom.h:177-213) as JSON */
const char *pgs_preagg_perfmon_json(pgs_session *session);
/* restrack.c:180-254 contract: abort must not leak in-flight chunks */
void        pgs_preagg_abort(pgs_session *session);
void        pgs_preagg_close(pgs_session *session);
/* CUDA stream the session launches on (for event timing by the caller) */
void       *pgs_preagg_stream(pgs_session *session);
/* kernel launch counter (bench evidence) */
uint64_t    pgs_preagg_launch_count(pgs_session *session);
/* device copy helpers for tests / benches that keep chunks resident */
void       *pgs_device_alloc(int device, size_t length);
void        pgs_device_free(int device, void *ptr);
int         pgs_device_upload(int device, void *dst_device, const void *src_host,
                              size_t length);
int         pgs_device_l2_flush(int device);

/* ------------------------------------------------------------------ 6 --- */
/* Executor half, CustomPlanMethods of "GpuPreAgg" (gpupreagg.c:2969-2978).
 * The child is a callback that hands over chunks (bulk-load protocol,
 * pg_strom.h:323-329 pgstrom_bulkslot). */
typedef struct pgs_gpupreagg_state pgs_gpupreagg_state;
typedef struct {
    const kern_data_store *kds;     /* NULL = end of scan */
    const kern_row_map    *krowmap; /* may be NULL */
    void  (*release)(void *arg, const kern_data_store *kds);
    void   *release_arg;
} pgs_bulkslot;
typedef int (*pgs_bulk_exec_fn)(void *child_state, pgs_bulkslot *slot);

/* BeginCustomPlan: gpupreagg_begin (gpupreagg.c:2189) */
int         gpupreagg_begin(pgs_plan *plan, int idx, int device,
                            pgs_bulk_exec_fn child_exec, void *child_state,
                            pgs_gpupreagg_state **state);
/* ExecCustomPlan: gpupreagg_exec (gpupreagg.c:2665): returns 1 and fills one
 * partial row, 0 at end of data, <0 on error (-errcode) */
int         gpupreagg_exec(pgs_gpupreagg_state *state, Datum *values, char *isnull);
/* text / bpchar grouping keys of those rows are "kernel text" words; the
 * strings of the long ones live in the key heap this returns once the first
 * row has come back (pgs_preagg_key_heap; pgstrom_fixup_kernel_text_heap
 * makes the varlena - the glue does that where the reference's
 * pg_fixup_tupslot_varlena made host pointers) */
int         gpupreagg_key_heap(pgs_gpupreagg_state *state, const void **heap,
                               size_t *heap_len);
/* rows the device left to the host: (chunk sequence number, row index) */
int64_t     gpupreagg_recheck_rows(pgs_gpupreagg_state *state,
                                   uint32_t *chunk_seq, uint32_t *rows,
                                   int64_t max_rows);
/* Ownership of a chunk with re-check rows: its release callback is NOT
 * called when the device is done with it - the host still has to walk the
 * rows (the reference keeps the pgstrom_data_store in curr_recheck until
 * gpupreagg_next_tuple_fallback has gone through it, gpupreagg.c:2507-2607,
 * 2746).  gpupreagg_recheck_chunk() returns the retained chunk (and its row
 * map) of a sequence number gpupreagg_recheck_rows() reported,
 * gpupreagg_recheck_done() hands it back through the release callback;
 * ReScan and EndCustomPlan release whatever is still held. */
const kern_data_store *gpupreagg_recheck_chunk(pgs_gpupreagg_state *state,
                                               uint32_t chunk_seq,
                                               const kern_row_map **krowmap);
int         gpupreagg_recheck_done(pgs_gpupreagg_state *state, uint32_t chunk_seq);
/* EndCustomPlan: gpupreagg_end (gpupreagg.c:2778); returns the NOTICE text
 * "GpuPreAgg: %u chunks were re-checked by CPU" or NULL */
const char *gpupreagg_end(pgs_gpupreagg_state *state);
/* ReScanCustomPlan: gpupreagg_rescan (gpupreagg.c:2825) */
int         gpupreagg_rescan(pgs_gpupreagg_state *state);
/* ExplainCustomPlan: gpupreagg_explain (gpupreagg.c:2859) */
const char *gpupreagg_explain(pgs_gpupreagg_state *state, int verbose, int analyze);

/* ------------------------------------------------------------------ 7 --- */
/* SQL-side half (host only): the arithmetic under the fmgr V1 functions of
 * pg_strom--1.0.sql:99-401 - partial placeholders, evaluated on the host for
 * the rows of gpupreagg_recheck_rows(), and the accumulators of the final
 * aggregates, called by PostgreSQL's Agg node for every partial row.
 * INTEGRATION.md section 5 shows the one-line wrappers. */
#define PGS_FINALFN_OVERFLOW    1   /* ereport: "value out of range: overflow" */
#define PGS_FINALFN_BAD_NROWS   2   /* elog: "Bug? NULL or negative nrows was given" */
#define PGS_FINALFN_BAD_NUMERIC 3   /* psum text is not a decimal number */
#define PGS_PCOV_X   0
#define PGS_PCOV_Y   1
#define PGS_PCOV_X2  2
#define PGS_PCOV_Y2  3
#define PGS_PCOV_XY  4
/* gpupreagg.c:4251 gpupreagg_partial_nrows(): pgstrom.nrows(bool, ...) */
int32_t     pgs_partial_nrows(int nargs, const char *values, const char *isnull);
/* gpupreagg.c:4308 gpupreagg_psum_x2_float(): pgstrom.psum_x2(float8);
 * returns the isnull flag (pgstrom.psum / pmin / pmax return their argument,
 * gpupreagg.c:4267-4306, and need no arithmetic) */
int         pgs_psum_x2_float8(double x, int x_isnull, double *result);
/* gpupreagg.c:4344-4417 gpupreagg_corr_psum_*(): pgstrom.pcov_x/y/x2/y2/xy
 * (bool, float8, float8); kind = PGS_PCOV_*; returns the isnull flag */
int         pgs_pcov_float8(int kind, int filter, int filter_isnull,
                            double x, int x_isnull, double y, int y_isnull,
                            double *result);
/* gpupreagg.c:4434 pgstrom_avg_int8_accum(int8[2], int4 nrows, int8 psum) */
int         pgs_avg_int8_accum(int64_t *trans, int32_t nrows, int64_t psum);
/* gpupreagg.c:4470 pgstrom_sum_int8_accum(int8[2], int8 psum) */
int         pgs_sum_int8_accum(int64_t *trans, int64_t psum);
/* gpupreagg.c:4508 pgstrom_sum_int8_final(int8[2]); returns the isnull flag */
int         pgs_sum_int8_final(const int64_t *trans, int64_t *result);
/* gpupreagg.c:4622 pgstrom_sum_float8_accum(float8[3], int4, float8) */
int         pgs_sum_float8_accum(double *trans, int32_t nrows, double psum);
/* gpupreagg.c:4668 pgstrom_variance_float8_accum(float8[3], int4, float8, float8) */
int         pgs_variance_float8_accum(double *trans, int32_t nrows,
                                      double psum, double psum_x2);
/* gpupreagg.c:4719 pgstrom_covariance_float8_accum(float8[6], int4, 5 x float8);
 * psum[] = {pcov_x, pcov_x2, pcov_y, pcov_y2, pcov_xy} */
int         pgs_covariance_float8_accum(double *trans, int32_t nrows,
                                        const double *psum);
/* gpupreagg.c:4540,4565 pgstrom_int8_avg_accum / pgstrom_numeric_avg_accum
 * (internal, int4 nrows, numeric psum): N and an exact decimal sum; psum as
 * decimal text (numeric_out / pgstrom_fixup_kernel_numeric), NULL = SQL NULL */
typedef struct pgs_numeric_avg_state pgs_numeric_avg_state;
pgs_numeric_avg_state *pgs_numeric_avg_init(void);
void        pgs_numeric_avg_free(pgs_numeric_avg_state *state);
int         pgs_numeric_avg_accum(pgs_numeric_avg_state *state, int32_t nrows,
                                  int nrows_isnull, const char *psum_text);
int64_t     pgs_numeric_avg_count(const pgs_numeric_avg_state *state);
size_t      pgs_numeric_avg_sum_text(const pgs_numeric_avg_state *state,
                                     char *buf, size_t buflen);

#ifdef __cplusplus
}
#endif
#endif  /* PGSTROM_CUDA_H */
