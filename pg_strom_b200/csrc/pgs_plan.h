/*
 * pgs_plan.h - planner half of GpuPreAgg (host side, no CUDA).
 *
 * Mirrors gpupreagg.c:134-2187 of the reference: the catalogue of
 * aggregates that can be pre-aggregated, the Agg target-list rewrite into
 * alternative (final) aggregates over partial columns, code generation of
 * the device program, and injection of the GpuPreAgg node under the Agg
 * node; plus grafter.c (walk of the finished plan tree) and the EXPLAIN
 * text of main.c:336-439 / gpupreagg.c:2859-2877.
 */
#ifndef PGS_PLAN_H
#define PGS_PLAN_H

#include <string>
#include <vector>
#include "pgs_json.h"
#include "pgs_codegen.h"

namespace pgs {

/* ---- GUCs (main.c:104-234, gpupreagg.c:2946-2967, ...) ---- */
struct GucEntry
{
    const char *name;
    const char *kind;       /* bool | int | real | string */
    std::string value;      /* current value as text */
    const char *boot;       /* default */
    const char *minval;
    const char *maxval;
    const char *context;    /* USERSET | SUSET | POSTMASTER */
    const char *desc;
};
std::vector<GucEntry> &guc_table();
bool        guc_set(const std::string &name, const std::string &value, std::string *err);
std::string guc_get(const std::string &name);
bool        guc_bool(const std::string &name);
long long   guc_int(const std::string &name);
double      guc_real(const std::string &name);
void        guc_reset_all();
bool        pgstrom_enabled();      /* main.c:64-76 */

/* ---- one partial column of the GpuPreAgg target list ---- */
struct PartialColumn
{
    int         resno;          /* 1-based */
    int         role;           /* GPUPREAGG_FIELD_IS_* */
    std::string func;           /* nrows | psum | pmin | pmax | psum_x2 | pcov_* | "" */
    std::string type;           /* result type name */
    std::string op;             /* PSUM | PMIN | PMAX */
    std::string cell_type;      /* INT | LONGS | LONG | FLOAT | DOUBLE | SHORT | NUMERIC */
    int         agg_index;      /* index among aggregate fields, or key index */
    int         cell_index;
    JsonPtr     expr;           /* the target entry expression */
};

struct GpuPreAggPlan
{
    bool        valid = false;
    std::string reject_reason;
    JsonPtr     plan;               /* new plan tree (Agg on top) */
    JsonPtr     gpreagg;            /* the GpuPreAgg node inside `plan` */
    std::vector<PartialColumn> columns;     /* GpuPreAgg target list */
    std::vector<int> grp_col_idx;
    bool        needs_grouping = false;
    bool        outer_bulkload = false;
    double      num_groups = 0;
    std::string kern_source;        /* generated part of the device program */
    int         extra_flags = 0;
    std::vector<unsigned char> kparams;     /* kern_parambuf image */
    std::vector<JsonPtr> used_params;
    std::vector<std::string> outer_colnames;
    std::vector<std::string> outer_coltypes;
    std::vector<int> incol_index;   /* referenced outer columns (0-based) */
    int         num_cells = 0;
    int         row_bytes = 0;      /* algorithmic bytes per row */
};

/* pgstrom_try_insert_gpupreagg (gpupreagg.c:1987): `agg` is an Agg plan
 * node (JSON); returns a plan with valid=false and the reason when the
 * aggregate cannot be pre-processed on the device. */
GpuPreAggPlan pgstrom_try_insert_gpupreagg(const JsonPtr &agg);

/* grafter.c:24-117: walks a finished plan tree and tries to inject GpuPreAgg
 * under every Agg node.  Returns the new tree; `plans` receives the
 * GpuPreAgg plans created (at most one per Agg). */
JsonPtr pgstrom_grafter(const JsonPtr &plan_tree, std::vector<GpuPreAggPlan> *plans);

/* EXPLAIN (VERBOSE, COSTS OFF) text of a plan tree, one line per entry */
std::vector<std::string> explain_plan(const JsonPtr &plan_tree, bool verbose);

/* full device program = static headers + generated text */
std::string assemble_device_program(const std::string &kern_source, int extra_flags);

}   /* namespace pgs */
#endif  /* PGS_PLAN_H */
