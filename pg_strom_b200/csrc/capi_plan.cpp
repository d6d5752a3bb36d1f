/*
 * capi_plan.cpp - C ABI of the planner half (host only; no CUDA).
 */
#include <cstring>
#include "pgs_plan.h"
#include "../../include/pgstrom_cuda.h"

namespace pgs { extern thread_local std::string last_error; }
using namespace pgs;

struct pgs_plan
{
    JsonPtr     tree;
    std::vector<GpuPreAggPlan> nodes;
    std::string reject_reason;
    std::string buf_tree, buf_explain, buf_describe;
};

extern "C" {

pgs_plan *
pgstrom_grafter_json(const char *plan_tree_json)
{
    try
    {
        JsonPtr tree = JsonParser::parse(plan_tree_json ? plan_tree_json : "");
        pgs_plan *plan = new pgs_plan;

        plan->tree = pgstrom_grafter(tree, &plan->nodes);
        if (plan->nodes.empty())
        {
            /* report why the (first) Agg node was left alone */
            JsonPtr n = tree;
            while (n && !n->is_null() && n->s("node") != "Agg")
                n = n->getp("lefttree");
            if (n && !n->is_null())
                plan->reject_reason = pgstrom_try_insert_gpupreagg(n).reject_reason;
        }
        return plan;
    }
    catch (const std::exception &e)
    {
        last_error = e.what();
        return NULL;
    }
}

void
pgs_plan_free(pgs_plan *plan)
{
    delete plan;
}

const char *
pgs_plan_tree_json(pgs_plan *plan)
{
    plan->buf_tree = plan->tree->dump();
    return plan->buf_tree.c_str();
}

const char *
pgs_plan_explain(pgs_plan *plan, int verbose)
{
    std::vector<std::string> lines = explain_plan(plan->tree, verbose != 0);
    plan->buf_explain.clear();
    for (size_t i = 0; i < lines.size(); i++)
        plan->buf_explain += (i ? "\n" : "") + lines[i];
    return plan->buf_explain.c_str();
}

int
pgs_plan_num_gpupreagg(pgs_plan *plan)
{
    return (int)plan->nodes.size();
}

const char *
pgs_plan_reject_reason(pgs_plan *plan)
{
    return plan->reject_reason.c_str();
}

static GpuPreAggPlan *
plan_node(pgs_plan *plan, int idx)
{
    if (!plan || idx < 0 || (size_t)idx >= plan->nodes.size())
        return NULL;
    return &plan->nodes[idx];
}

const char *
pgs_plan_kernel_source(pgs_plan *plan, int idx)
{
    GpuPreAggPlan *gp = plan_node(plan, idx);
    return gp ? gp->kern_source.c_str() : NULL;
}

int
pgs_plan_extra_flags(pgs_plan *plan, int idx)
{
    GpuPreAggPlan *gp = plan_node(plan, idx);
    return gp ? gp->extra_flags : 0;
}

const void *
pgs_plan_kparams(pgs_plan *plan, int idx, size_t *length)
{
    GpuPreAggPlan *gp = plan_node(plan, idx);
    if (!gp)
        return NULL;
    if (length)
        *length = gp->kparams.size();
    return gp->kparams.data();
}

int
pgs_plan_needs_grouping(pgs_plan *plan, int idx)
{
    GpuPreAggPlan *gp = plan_node(plan, idx);
    return gp ? (gp->needs_grouping ? 1 : 0) : 0;
}

double
pgs_plan_num_groups(pgs_plan *plan, int idx)
{
    GpuPreAggPlan *gp = plan_node(plan, idx);
    return gp ? gp->num_groups : 0.0;
}

const char *
pgs_plan_describe_json(pgs_plan *plan, int idx)
{
    GpuPreAggPlan *gp = plan_node(plan, idx);
    if (!gp)
        return NULL;
    JsonPtr o = Json::object();
    JsonPtr cols = Json::array();
    for (auto &pc : gp->columns)
    {
        JsonPtr c = Json::object();
        c->set("resno", pc.resno);
        c->set("role", pc.role);
        c->set("func", pc.func);
        c->set("type", pc.type);
        c->set("op", pc.op);
        c->set("cell_type", pc.cell_type);
        c->set("agg_index", pc.agg_index);
        c->set("cell_index", pc.cell_index);
        c->set("text", deparse_expression(pc.expr, gp->outer_colnames, false));
        if (pc.expr && pc.expr->s("node") == "Var" && pc.expr->has("vartypmod"))
            c->set("typmod", (int)pc.expr->i("vartypmod"));
        cols->push(c);
    }
    o->set("columns", cols);
    o->set("agg_targetlist", gp->plan->getp("targetlist"));
    o->set("agg_qual", gp->plan->getp("qual") ? gp->plan->getp("qual") : Json::array());
    JsonPtr g = Json::array();
    for (int x : gp->grp_col_idx) g->push(Json::number(x));
    o->set("grpColIdx", g);
    JsonPtr ic = Json::array();
    for (int x : gp->incol_index) ic->push(Json::number(x));
    o->set("incol_index", ic);
    o->setb("needs_grouping", gp->needs_grouping);
    o->setb("outer_bulkload", gp->outer_bulkload);
    o->set("num_groups", Json::number(gp->num_groups));
    o->set("num_cells", gp->num_cells);
    o->set("row_bytes", gp->row_bytes);
    o->set("extra_flags", gp->extra_flags);
    JsonPtr on = Json::array(), ot = Json::array();
    for (auto &n : gp->outer_colnames) on->push(Json::string(n));
    for (auto &t : gp->outer_coltypes) ot->push(Json::string(t));
    o->set("outer_colnames", on);
    o->set("outer_coltypes", ot);
    plan->buf_describe = o->dump();
    return plan->buf_describe.c_str();
}

int
pgs_plan_result_colmeta(pgs_plan *plan, int idx, kern_colmeta *colmeta, int max_cols)
{
    GpuPreAggPlan *gp = plan_node(plan, idx);
    if (!gp)
        return -1;
    int n = (int)gp->columns.size();
    for (int i = 0; i < n && i < max_cols; i++)
    {
        const DevType *dtype = devtype_lookup(gp->columns[i].type);
        memset(&colmeta[i], 0, sizeof(kern_colmeta));
        if (dtype && dtype->type_length > 0)
        {
            colmeta[i].attbyval = 1;
            colmeta[i].attalign = (cl_char)dtype->type_align;
            colmeta[i].attlen = (cl_short)dtype->type_length;
        }
        else if (dtype && (std::string(dtype->type_name) == "numeric" ||
                           ((std::string(dtype->type_name) == "text" ||
                             std::string(dtype->type_name) == "bpchar") &&
                            gp->columns[i].role == GPUPREAGG_FIELD_IS_GROUPKEY)))
        {
            /* numeric; "kernel text" of a text / bpchar grouping key */
            /* internal_format=true: 64-bit device numeric, by value
             * (datastore.c:355-363) */
            colmeta[i].attbyval = 1;
            colmeta[i].attalign = 8;
            colmeta[i].attlen = 8;
        }
        else
        {
            colmeta[i].attbyval = 0;
            colmeta[i].attalign = (cl_char)(dtype ? dtype->type_align : 4);
            colmeta[i].attlen = -1;
        }
        colmeta[i].attnum = (cl_short)(i + 1);
        colmeta[i].attcacheoff = -1;
    }
    return n;
}

int
pgstrom_codegen_available_expression_json(const char *expr_json)
{
    try
    {
        return codegen_available_expression(JsonParser::parse(expr_json)) ? 1 : 0;
    }
    catch (const std::exception &e)
    {
        last_error = e.what();
        return -1;
    }
}

}   /* extern "C" */
