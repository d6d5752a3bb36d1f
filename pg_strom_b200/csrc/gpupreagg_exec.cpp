/*
 * gpupreagg_exec.cpp - executor half of the GpuPreAgg node.
 *
 * Mirrors gpupreagg.c:2189-2941 of the reference (gpupreagg_begin,
 * gpupreagg_load_next_outer, pgstrom_create_gpupreagg, gpupreagg_exec,
 * gpupreagg_next_tuple, gpupreagg_end, gpupreagg_rescan, gpupreagg_explain)
 * on top of the session API of cuda_layer.cpp.  Differences, all inside the
 * reference's own contract:
 *   - partial rows come out once per scan (the state is persistent in HBM),
 *     not once per chunk, so the final Agg sees #groups rows, not
 *     #groups x #chunks;
 *   - StromError_CpuReCheck is row level: the rows listed by
 *     gpupreagg_recheck_rows() are the only ones the host evaluates itself
 *     (the reference re-does the whole chunk, gpupreagg.c:2507-2607); the
 *     NOTICE still counts chunks.
 */
#include <cstring>
#include <deque>
#include <string>
#include <vector>
#include "pgs_plan.h"
#include "../../include/pgstrom_cuda.h"

namespace pgs { extern thread_local std::string last_error; }
using namespace pgs;

struct pgs_plan;    /* capi_plan.cpp */

struct RunningChunk
{
    pgs_ticket      ticket;
    uint32_t        seq;
    pgs_bulkslot    slot;
};

struct pgs_gpupreagg_state
{
    pgs_plan       *plan = NULL;
    int             idx = 0;
    int             device = 0;
    pgs_bulk_exec_fn child_exec = NULL;
    void           *child_state = NULL;
    pgs_program    *program = NULL;
    pgs_session    *session = NULL;
    std::vector<kern_colmeta> result_colmeta;
    bool            outer_done = false;
    bool            flushed = false;
    std::deque<RunningChunk> running;
    uint32_t        next_seq = 0;
    uint32_t        num_rechecked = 0;
    std::vector<std::pair<uint32_t, uint32_t> > recheck;   /* (chunk seq, row) */
    /* chunks with re-check rows stay with this node until the host has walked
     * them (gpupreagg.c:2746 keeps curr_recheck the same way):
     * gpupreagg_recheck_done(), ReScan or EndCustomPlan hand them back */
    std::vector<RunningChunk> held;
    std::vector<char> result_buf;       /* TUPSLOT store */
    uint32_t        curr_index = 0;
    std::string     notice, explain_buf;
    bool            needs_grouping = false;
    double          num_groups = 1.0;
};

static int
retire_chunk(pgs_gpupreagg_state *st, bool wait)
{
    if (st->running.empty())
        return 0;
    RunningChunk rc = st->running.front();
    int32_t status = 0;
    int ret = pgs_preagg_wait(st->session, rc.ticket, wait ? -1 : 0, &status);

    if (ret == -1)
        return 0;       /* still running */
    if (ret != StromError_Success)
        return -ret;
    st->running.pop_front();
    if (status == StromError_CpuReCheck)
    {
        int64_t n = pgs_preagg_recheck_rows(st->session, rc.ticket, NULL, 0);
        std::vector<uint32_t> rows((size_t)std::max<int64_t>(n, 0));
        pgs_preagg_recheck_rows(st->session, rc.ticket, rows.data(), n);
        for (uint32_t r : rows)
            st->recheck.push_back(std::make_pair(rc.seq, r));
        st->num_rechecked++;
        if (!rows.empty())
        {
            st->held.push_back(rc);
            rc.slot.release = NULL;     /* not yet */
        }
    }
    if (rc.slot.release)
        rc.slot.release(rc.slot.release_arg, rc.slot.kds);
    if (StromErrorIsSignificant(status))
    {
        last_error = std::string("GpuPreAgg: device error: ") + pgstrom_strerror(status);
        return -status;
    }
    return 1;
}

static void
release_held(pgs_gpupreagg_state *st)
{
    for (auto &h : st->held)
        if (h.slot.release)
            h.slot.release(h.slot.release_arg, h.slot.kds);
    st->held.clear();
}

extern "C" {

int
gpupreagg_begin(pgs_plan *plan, int idx, int device,
                pgs_bulk_exec_fn child_exec, void *child_state,
                pgs_gpupreagg_state **state)
{
    const char *source = pgs_plan_kernel_source(plan, idx);
    const char *build_log = NULL;
    size_t      kplen = 0;
    const void *kparams = pgs_plan_kparams(plan, idx, &kplen);
    int         rc;

    if (!source || !kparams || !child_exec)
    {
        last_error = "gpupreagg_begin: not a GpuPreAgg plan";
        return StromError_BadRequestMessage;
    }
    pgs_gpupreagg_state *st = new pgs_gpupreagg_state;
    st->plan = plan;
    st->idx = idx;
    st->device = device;
    st->child_exec = child_exec;
    st->child_state = child_state;
    st->needs_grouping = pgs_plan_needs_grouping(plan, idx) != 0;
    st->num_groups = pgs_plan_num_groups(plan, idx);
    st->result_colmeta.resize(256);
    int ncols = pgs_plan_result_colmeta(plan, idx, st->result_colmeta.data(), 256);
    st->result_colmeta.resize(std::max(0, ncols));

    rc = pgs_program_build(source, pgs_plan_extra_flags(plan, idx), &st->program, &build_log);
    if (rc != StromError_Success)
    {
        /* gpupreagg.c:2751-2764: ERROR with source and build log */
        if (build_log)
            last_error = std::string("device kernel build failure:\n") + build_log;
        delete st;
        return rc;
    }
    pgs_session_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.device = device;
    cfg.needs_grouping = st->needs_grouping ? 1 : 0;
    cfg.num_groups = st->num_groups;
    cfg.max_async_chunks = (int)guc_int("pg_strom.max_async_chunks");
    cfg.result_ncols = ncols;
    cfg.result_colmeta = st->result_colmeta.data();
    rc = pgs_preagg_open(st->program, (const kern_parambuf *)kparams, &cfg, &st->session);
    if (rc != StromError_Success)
    {
        pgs_program_release(st->program);
        delete st;
        return rc;
    }
    *state = st;
    return StromError_Success;
}

int
gpupreagg_exec(pgs_gpupreagg_state *st, Datum *values, char *isnull)
{
    if (!st->flushed)
    {
        int max_async = (int)guc_int("pg_strom.max_async_chunks");
        /* gpupreagg.c:2696-2717: keep the window of in-flight chunks full */
        while (!st->outer_done)
        {
            while ((int)st->running.size() >= max_async)
            {
                int r = retire_chunk(st, true);
                if (r < 0)
                    return r;
            }
            pgs_bulkslot slot;
            memset(&slot, 0, sizeof(slot));
            int rc = st->child_exec(st->child_state, &slot);
            if (rc != 0)
            {
                last_error = "GpuPreAgg: outer plan failed";
                return -StromError_BadRequestMessage;
            }
            if (!slot.kds)
            {
                st->outer_done = true;
                break;
            }
            RunningChunk run;
            run.seq = st->next_seq++;
            run.slot = slot;
            rc = pgs_preagg_submit(st->session, slot.kds, slot.krowmap, &run.ticket);
            if (rc != StromError_Success)
                return -rc;
            st->running.push_back(run);
            /* opportunistically retire what has already completed */
            for (;;)
            {
                int r = retire_chunk(st, false);
                if (r < 0)
                    return r;
                if (r == 0)
                    break;
            }
        }
        while (!st->running.empty())
        {
            int r = retire_chunk(st, true);
            if (r < 0)
                return r;
        }
        /* end of scan: partial rows of the whole scan */
        uint32_t nrooms = (uint32_t)std::max(16.0, st->needs_grouping ? st->num_groups * 1.25 + 64 : 16.0);
        int ncols = (int)st->result_colmeta.size();
        for (int attempt = 0; attempt < 4; attempt++)
        {
            size_t len = pgstrom_kds_tupslot_length(ncols, nrooms);
            uint32_t needed = 0;
            int32_t status = 0;
            st->result_buf.assign(len, 0);
            int rc = pgstrom_kds_tupslot_init(st->result_buf.data(), len, ncols,
                                              st->result_colmeta.data(), nrooms);
            if (rc != StromError_Success)
                return -rc;
            rc = pgs_preagg_finish(st->session, (kern_data_store *)st->result_buf.data(),
                                   1, &needed, &status);
            if (rc == StromError_DataStoreNoSpace && needed > nrooms)
            {
                nrooms = needed + 16;
                continue;
            }
            if (rc != StromError_Success)
                return -rc;
            break;
        }
        st->flushed = true;
        st->curr_index = 0;
    }
    /* gpupreagg_next_tuple (gpupreagg.c:2609-2663) */
    const kern_data_store *kds = (const kern_data_store *)st->result_buf.data();
    if (st->curr_index >= kds->nitems)
        return 0;
    int rc = pgstrom_fetch_data_store(kds, st->curr_index++, values, isnull);
    return rc == StromError_Success ? 1 : -rc;
}

/* the strings behind long text / bpchar keys of the rows gpupreagg_exec()
 * returns (pgs_preagg_key_heap): valid until ReScan / EndCustomPlan */
int
gpupreagg_key_heap(pgs_gpupreagg_state *st, const void **heap, size_t *heap_len)
{
    if (!st || !st->session)
        return StromError_BadRequestMessage;
    return pgs_preagg_key_heap(st->session, heap, heap_len);
}

int64_t
gpupreagg_recheck_rows(pgs_gpupreagg_state *st, uint32_t *chunk_seq, uint32_t *rows,
                       int64_t max_rows)
{
    int64_t n = (int64_t)st->recheck.size();
    for (int64_t i = 0; i < n && i < max_rows; i++)
    {
        if (chunk_seq) chunk_seq[i] = st->recheck[(size_t)i].first;
        if (rows) rows[i] = st->recheck[(size_t)i].second;
    }
    return n;
}

const kern_data_store *
gpupreagg_recheck_chunk(pgs_gpupreagg_state *st, uint32_t chunk_seq,
                        const kern_row_map **krowmap)
{
    for (auto &h : st->held)
        if (h.seq == chunk_seq)
        {
            if (krowmap)
                *krowmap = h.slot.krowmap;
            return h.slot.kds;
        }
    return NULL;
}

int
gpupreagg_recheck_done(pgs_gpupreagg_state *st, uint32_t chunk_seq)
{
    for (size_t i = 0; i < st->held.size(); i++)
        if (st->held[i].seq == chunk_seq)
        {
            RunningChunk h = st->held[i];
            st->held.erase(st->held.begin() + (long)i);
            if (h.slot.release)
                h.slot.release(h.slot.release_arg, h.slot.kds);
            return StromError_Success;
        }
    last_error = "gpupreagg_recheck_done: no chunk " + std::to_string(chunk_seq) +
        " is held for re-check";
    return StromError_BadRequestMessage;
}

const char *
gpupreagg_end(pgs_gpupreagg_state *st)
{
    static thread_local std::string notice;
    /* gpupreagg.c:2785-2787 */
    notice.clear();
    if (st->num_rechecked > 0)
        notice = "GpuPreAgg: " + std::to_string(st->num_rechecked) +
            " chunks were re-checked by CPU";
    while (!st->running.empty())
    {
        if (retire_chunk(st, true) < 0)
        {
            st->running.pop_front();
        }
    }
    release_held(st);
    if (st->session)
        pgs_preagg_close(st->session);
    if (st->program)
        pgs_program_release(st->program);
    delete st;
    return notice.empty() ? NULL : notice.c_str();
}

int
gpupreagg_rescan(pgs_gpupreagg_state *st)
{
    /* gpupreagg.c:2825-2857: drain in-flight chunks and rewind */
    while (!st->running.empty())
    {
        int r = retire_chunk(st, true);
        if (r < 0)
            return -r;
    }
    release_held(st);
    int rc = pgs_preagg_state_reset(st->session);
    if (rc != StromError_Success)
        return rc;
    st->outer_done = false;
    st->flushed = false;
    st->curr_index = 0;
    st->recheck.clear();
    st->num_rechecked = 0;
    st->next_seq = 0;
    return StromError_Success;
}

const char *
gpupreagg_explain(pgs_gpupreagg_state *st, int verbose, int analyze)
{
    /* gpupreagg.c:2859-2877 + main.c:399-439, 504-660 */
    std::string s;
    JsonPtr desc = JsonParser::parse(pgs_plan_describe_json(st->plan, st->idx));
    s += std::string("Bulkload: ") + (desc->flag("outer_bulkload") ? "On" : "Off") + "\n";
    if (verbose && guc_bool("pg_strom.show_device_kernel"))
    {
        s += "Kernel Source: ";
        s += pgs_plan_kernel_source(st->plan, st->idx);
        s += "\n";
    }
    if (analyze && guc_bool("pg_strom.perfmon") && st->session)
    {
        s += "Perfmon: ";
        s += pgs_preagg_perfmon_json(st->session);
        s += "\n";
    }
    st->explain_buf = s;
    return st->explain_buf.c_str();
}

}   /* extern "C" */
