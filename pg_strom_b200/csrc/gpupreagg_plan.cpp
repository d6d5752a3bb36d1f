/*
 * gpupreagg_plan.cpp - planner half of GpuPreAgg.  See pgs_plan.h.
 */
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <sstream>
#include "pgs_plan.h"
#include "../../include/pgstrom_kds.h"

namespace pgs {

/*
 * Arguments of alternative functions (gpupreagg.c:104-113)
 */
enum {
    ALTFUNC_EXPR_NROWS = 101,   /* NROWS(X) */
    ALTFUNC_EXPR_PMIN,          /* PMIN(X) */
    ALTFUNC_EXPR_PMAX,          /* PMAX(X) */
    ALTFUNC_EXPR_PSUM,          /* PSUM(X) */
    ALTFUNC_EXPR_PSUM_X2,       /* PSUM_X2(X) = PSUM(X^2) */
    ALTFUNC_EXPR_PCOV_X,        /* PCOV_X(X,Y) */
    ALTFUNC_EXPR_PCOV_Y,        /* PCOV_Y(X,Y) */
    ALTFUNC_EXPR_PCOV_X2,       /* PCOV_X2(X,Y) */
    ALTFUNC_EXPR_PCOV_Y2,       /* PCOV_Y2(X,Y) */
    ALTFUNC_EXPR_PCOV_XY,       /* PCOV_XY(X,Y) */
};

/*
 * List of supported aggregate functions (gpupreagg.c:134-333).
 * altfn prefix: "c:" pg_catalog, "s:" pgstrom schema.
 */
struct aggfunc_catalog_t {
    const char *aggfn_name;
    int         aggfn_nargs;
    const char *aggfn_argtypes[4];
    const char *altfn_name;
    int         altfn_nargs;
    const char *altfn_argtypes[8];
    int         altfn_argexprs[8];
    int         altfn_flags;
};

#define NROWS   ALTFUNC_EXPR_NROWS
#define PSUM    ALTFUNC_EXPR_PSUM
#define PSUM_X2 ALTFUNC_EXPR_PSUM_X2
#define PCOV    ALTFUNC_EXPR_PCOV_X, ALTFUNC_EXPR_PCOV_X2, ALTFUNC_EXPR_PCOV_Y, \
                ALTFUNC_EXPR_PCOV_Y2, ALTFUNC_EXPR_PCOV_XY
#define F8x5    "float8", "float8", "float8", "float8", "float8"

static const aggfunc_catalog_t aggfunc_catalog[] = {
    /* AVG(X) = EX_AVG(NROWS(), PSUM(X)) */
    { "avg", 1, {"int2"},   "s:avg", 2, {"int4", "int8"}, {NROWS, PSUM}, 0 },
    { "avg", 1, {"int4"},   "s:avg", 2, {"int4", "int8"}, {NROWS, PSUM}, 0 },
    { "avg", 1, {"int8"},   "s:avg_numeric", 2, {"int4", "int8"}, {NROWS, PSUM}, 0 },
    { "avg", 1, {"float4"}, "s:avg", 2, {"int4", "float8"}, {NROWS, PSUM}, 0 },
    { "avg", 1, {"float8"}, "s:avg", 2, {"int4", "float8"}, {NROWS, PSUM}, 0 },
    { "avg", 1, {"numeric"},"s:avg", 2, {"int4", "numeric"}, {NROWS, PSUM}, DEVFUNC_NEEDS_NUMERIC },
    /* COUNT(*) = SUM(NROWS(*|X)) */
    { "count", 0, {},      "c:sum", 1, {"int4"}, {NROWS}, 0 },
    { "count", 1, {"any"}, "c:sum", 1, {"int4"}, {NROWS}, 0 },
    /* MAX(X) = MAX(PMAX(X)) */
    { "max", 1, {"int2"},   "c:max", 1, {"int2"},   {ALTFUNC_EXPR_PMAX}, 0 },
    { "max", 1, {"int4"},   "c:max", 1, {"int4"},   {ALTFUNC_EXPR_PMAX}, 0 },
    { "max", 1, {"int8"},   "c:max", 1, {"int8"},   {ALTFUNC_EXPR_PMAX}, 0 },
    { "max", 1, {"float4"}, "c:max", 1, {"float4"}, {ALTFUNC_EXPR_PMAX}, 0 },
    { "max", 1, {"float8"}, "c:max", 1, {"float8"}, {ALTFUNC_EXPR_PMAX}, 0 },
    { "max", 1, {"numeric"},"c:max", 1, {"numeric"},{ALTFUNC_EXPR_PMAX}, DEVFUNC_NEEDS_NUMERIC },
    /* MIN(X) = MIN(PMIN(X)) */
    { "min", 1, {"int2"},   "c:min", 1, {"int2"},   {ALTFUNC_EXPR_PMIN}, 0 },
    { "min", 1, {"int4"},   "c:min", 1, {"int4"},   {ALTFUNC_EXPR_PMIN}, 0 },
    { "min", 1, {"int8"},   "c:min", 1, {"int8"},   {ALTFUNC_EXPR_PMIN}, 0 },
    { "min", 1, {"float4"}, "c:min", 1, {"float4"}, {ALTFUNC_EXPR_PMIN}, 0 },
    { "min", 1, {"float8"}, "c:min", 1, {"float8"}, {ALTFUNC_EXPR_PMIN}, 0 },
    { "min", 1, {"numeric"},"c:min", 1, {"numeric"},{ALTFUNC_EXPR_PMIN}, DEVFUNC_NEEDS_NUMERIC },
    /* SUM(X) = SUM(PSUM(X)) */
    { "sum", 1, {"int2"},   "s:sum", 1, {"int8"},   {PSUM}, 0 },
    { "sum", 1, {"int4"},   "s:sum", 1, {"int8"},   {PSUM}, 0 },
    { "sum", 1, {"float4"}, "c:sum", 1, {"float4"}, {PSUM}, 0 },
    { "sum", 1, {"float8"}, "c:sum", 1, {"float8"}, {PSUM}, 0 },
    { "sum", 1, {"numeric"},"c:sum", 1, {"numeric"},{PSUM}, DEVFUNC_NEEDS_NUMERIC },
    /* STDDEV(X) = EX_STDDEV(NROWS(),PSUM(X),PSUM(X*X)) */
#define VARIANCE_ENTRY(name)                                                        \
    { name, 1, {"float4"}, "s:" name, 3, {"int4", "float8", "float8"}, {NROWS, PSUM, PSUM_X2}, 0 }, \
    { name, 1, {"float8"}, "s:" name, 3, {"int4", "float8", "float8"}, {NROWS, PSUM, PSUM_X2}, 0 }, \
    { name, 1, {"numeric"},"s:" name, 3, {"int4", "numeric", "numeric"}, {NROWS, PSUM, PSUM_X2}, DEVFUNC_NEEDS_NUMERIC }
    VARIANCE_ENTRY("stddev"),
    VARIANCE_ENTRY("stddev_pop"),
    VARIANCE_ENTRY("stddev_samp"),
    VARIANCE_ENTRY("variance"),
    VARIANCE_ENTRY("var_pop"),
    VARIANCE_ENTRY("var_samp"),
    /* CORR(X,Y) = PGSTROM.CORR(NROWS(X,Y), PCOV_X, PCOV_X2, PCOV_Y, PCOV_Y2, PCOV_XY) */
    { "corr", 2, {"float8", "float8"}, "s:corr", 6, {"int4", F8x5}, {NROWS, PCOV}, 0 },
    { "covar_pop", 2, {"float8", "float8"}, "s:covar_pop", 6, {"int4", F8x5}, {NROWS, PCOV}, 0 },
    { "covar_samp", 2, {"float8", "float8"}, "s:covar_samp", 6, {"int4", F8x5}, {NROWS, PCOV}, 0 },
};
#define lengthof(a) (sizeof(a) / sizeof((a)[0]))

/* aggregates and functions created by pg_strom--1.0.sql (what
 * SearchSysCache3(PROCNAMEARGSNSP) would find in schema pgstrom) and the
 * pg_catalog aggregates the "c:" entries resolve to */
static bool
sql_catalog_has_aggregate(const std::string &schema, const std::string &name,
                          const std::vector<std::string> &args)
{
    auto is = [&](std::initializer_list<const char *> a) {
        if (a.size() != args.size()) return false;
        size_t i = 0;
        for (const char *t : a)
            if (args[i++] != t) return false;
        return true;
    };
    if (schema == "pg_catalog")
        return true;    /* sum(int4), max/min(...), sum(float*) , sum(numeric) */
    if (name == "avg")
        return is({"int4", "int8"}) || is({"int4", "float8"}) || is({"int4", "numeric"});
    if (name == "sum")
        return is({"int8"});
    if (name == "avg_numeric")
        return is({"int4", "int8"});
    if (name == "stddev" || name == "stddev_samp" || name == "stddev_pop" ||
        name == "variance" || name == "var_samp" || name == "var_pop")
        return is({"int4", "float8", "float8"});
    if (name == "corr" || name == "covar_pop" || name == "covar_samp")
        return is({"int4", "float8", "float8", "float8", "float8", "float8"});
    return false;
}

/* return type of pgstrom.<func>(argtypes), "" when no such function
 * (pg_strom--1.0.sql:99-226) */
static std::string
sql_catalog_partial_func(const std::string &func, const std::vector<std::string> &args)
{
    if (func == "nrows")
    {
        if (args.size() > 4) return "";
        for (auto &a : args) if (a != "bool") return "";
        return "int4";
    }
    if (func == "pmax" || func == "pmin")
    {
        static const char *ok[] = {"int2", "int4", "int8", "float4", "float8", "numeric"};
        if (args.size() != 1) return "";
        for (const char *t : ok) if (args[0] == t) return t;
        return "";
    }
    if (func == "psum")
    {
        static const char *ok[] = {"int8", "float4", "float8", "numeric"};
        if (args.size() != 1) return "";
        for (const char *t : ok) if (args[0] == t) return t;
        return "";
    }
    if (func == "psum_x2")
    {
        if (args.size() == 1 && (args[0] == "float8" || args[0] == "numeric"))
            return args[0];
        return "";
    }
    if (func.compare(0, 5, "pcov_") == 0)
    {
        if (args.size() == 3 && args[0] == "bool" && args[1] == "float8" && args[2] == "float8")
            return "float8";
        return "";
    }
    return "";
}

static const aggfunc_catalog_t *
aggfunc_lookup(const JsonPtr &aggref)
{
    std::string name = aggref->s("aggname");
    std::vector<std::string> argtypes;
    const Json *jt = aggref->get("aggargtypes");
    if (jt)
        for (auto &t : jt->arr)
            argtypes.push_back(t->str);
    for (size_t i = 0; i < lengthof(aggfunc_catalog); i++)
    {
        const aggfunc_catalog_t *c = &aggfunc_catalog[i];
        if (name != c->aggfn_name || (size_t)c->aggfn_nargs != argtypes.size())
            continue;
        bool same = true;
        for (int k = 0; k < c->aggfn_nargs; k++)
            if (std::string(c->aggfn_argtypes[k]) != "any" &&
                argtypes[k] != c->aggfn_argtypes[k])
                same = false;
        if (same)
            return c;
    }
    return NULL;
}

/* ------------------------------------------------------------------ */
static JsonPtr
make_null_const(const std::string &type)
{
    JsonPtr c = Json::object();
    c->set("node", "Const");
    c->set("consttype", type);
    c->setb("constisnull", true);
    return c;
}

static JsonPtr
make_zero_const(const std::string &type)
{
    JsonPtr c = Json::object();
    c->set("node", "Const");
    c->set("consttype", type);
    c->setb("constisnull", false);
    c->set("constvalue", type == "bool" ? "f" : "0");
    return c;
}

static JsonPtr
make_bool_const(bool v)
{
    JsonPtr c = Json::object();
    c->set("node", "Const");
    c->set("consttype", "bool");
    c->setb("constisnull", false);
    c->set("constvalue", v ? "t" : "f");
    return c;
}

static JsonPtr
make_var(int varattno, const std::string &type)
{
    JsonPtr v = Json::object();
    v->set("node", "Var");
    v->set("varattno", varattno);
    v->set("vartype", type);
    return v;
}

/* make_expr_typecast (gpupreagg.c:557-609): pg_cast says int2/int4 -> int8
 * and float4 -> float8 are function casts named after the target type */
static JsonPtr
make_expr_typecast(const JsonPtr &expr, const std::string &target_type)
{
    std::string source_type = expr_type(expr);
    if (source_type == target_type)
        return expr;
    if (!devfunc_lookup(target_type, {source_type}, target_type))
        return JsonPtr();
    JsonPtr f = Json::object();
    f->set("node", "FuncExpr");
    f->set("funcname", target_type);
    f->set("funcresulttype", target_type);
    f->set("funcformat", "cast");
    JsonPtr args = Json::array();
    args->push(expr);
    f->set("args", args);
    return f;
}

/* make_expr_conditional (gpupreagg.c:611-643) */
static JsonPtr
make_expr_conditional(const JsonPtr &expr, const JsonPtr &filter, JsonPtr defresult)
{
    if (!defresult)
        defresult = make_null_const(expr_type(expr));
    JsonPtr cw = Json::object();
    cw->set("node", "CaseWhen");
    cw->set("expr", filter);
    cw->set("result", expr);
    JsonPtr c = Json::object();
    c->set("node", "CaseExpr");
    c->set("casetype", expr_type(expr));
    c->set("arg", Json::null());
    JsonPtr args = Json::array();
    args->push(cw);
    c->set("args", args);
    c->set("defresult", defresult);
    return c;
}

/* make_altfunc_expr (gpupreagg.c:645-678) */
static JsonPtr
make_altfunc_expr(const std::string &func_name, const std::vector<JsonPtr> &args)
{
    std::vector<std::string> argtypes;
    for (auto &a : args)
        argtypes.push_back(expr_type(a));
    std::string rettype = sql_catalog_partial_func(func_name, argtypes);
    if (rettype.empty())
        return JsonPtr();
    JsonPtr f = Json::object();
    f->set("node", "FuncExpr");
    f->set("funcschema", "pgstrom");
    f->set("funcname", func_name);
    f->set("funcresulttype", rettype);
    f->set("funcformat", "call");
    JsonPtr a = Json::array();
    for (auto &x : args)
        a->push(x);
    f->set("args", a);
    return f;
}

static JsonPtr
aggref_arg(const JsonPtr &aggref, size_t i)
{
    const Json *args = aggref->get("args");
    if (!args || i >= args->arr.size())
        return JsonPtr();
    JsonPtr a = args->arr[i];
    if (a->s("node") == "TargetEntry")
        a = a->getp("expr");
    return a;
}

static JsonPtr
aggref_filter(const JsonPtr &aggref)
{
    JsonPtr f = aggref->getp("aggfilter");
    if (f && f->is_null())
        return JsonPtr();
    return f;
}

/* make_altfunc_nrows_expr (gpupreagg.c:680-701) */
static JsonPtr
make_altfunc_nrows_expr(const JsonPtr &aggref)
{
    std::vector<JsonPtr> nrows_args;
    JsonPtr filter = aggref_filter(aggref);
    const Json *args = aggref->get("args");

    if (filter)
        nrows_args.push_back(filter);
    for (size_t i = 0; args && i < args->arr.size(); i++)
    {
        JsonPtr ntest = Json::object();
        ntest->set("node", "NullTest");
        ntest->set("arg", aggref_arg(aggref, i));
        ntest->set("nulltesttype", "IS_NOT_NULL");
        ntest->setb("argisrow", false);
        nrows_args.push_back(ntest);
    }
    return make_altfunc_expr("nrows", nrows_args);
}

/* make_altfunc_pcov_expr (gpupreagg.c:707-721) */
static JsonPtr
make_altfunc_pcov_expr(const JsonPtr &aggref, const char *func_name)
{
    JsonPtr filter = aggref_filter(aggref);
    if (!filter)
        filter = make_bool_const(true);
    return make_altfunc_expr(func_name, {filter, aggref_arg(aggref, 0), aggref_arg(aggref, 1)});
}

struct rewrite_context
{
    std::vector<int>   *grp_col_idx;
    std::vector<JsonPtr> pre_tlist;     /* TargetEntry nodes */
    int         extra_flags = 0;
    bool        invalid = false;
    std::string reason;
};

/* make_gpupreagg_refnode (gpupreagg.c:729-980) */
static JsonPtr
make_gpupreagg_refnode(const JsonPtr &aggref, rewrite_context &ctx)
{
    const aggfunc_catalog_t *aggfn_cat = aggfunc_lookup(aggref);
    const Json *args = aggref->get("args");

    /* Only aggregated functions listed on the catalog above is supported. */
    if (!aggfn_cat)
    {
        ctx.reason = "aggregate " + aggref->s("aggname") + " is not in aggfunc_catalog";
        return JsonPtr();
    }
    /* ordered-set / variadic / DISTINCT / ORDER BY aggregates are not */
    if (aggref->has("aggdirectargs") || aggref->flag("aggvariadic") ||
        aggref->has("aggorder") || aggref->has("aggdistinct") ||
        (args && args->arr.size() > 2))
    {
        ctx.reason = "unsupported aggregate modifiers";
        return JsonPtr();
    }
    ctx.extra_flags |= aggfn_cat->altfn_flags;

    /* Expression node that is executed in the device kernel has to be
     * supported by codegen */
    for (size_t i = 0; args && i < args->arr.size(); i++)
        if (!codegen_available_expression(aggref_arg(aggref, i)))
        {
            ctx.reason = "aggregate argument is not device runnable";
            return JsonPtr();
        }
    if (!codegen_available_expression(aggref_filter(aggref)))
    {
        ctx.reason = "aggregate filter is not device runnable";
        return JsonPtr();
    }

    std::string altfn_schema = (aggfn_cat->altfn_name[0] == 'c' ? "pg_catalog" : "pgstrom");
    std::string altfn_name = aggfn_cat->altfn_name + 2;
    std::vector<std::string> altfn_argtypes;
    for (int i = 0; i < aggfn_cat->altfn_nargs; i++)
        altfn_argtypes.push_back(aggfn_cat->altfn_argtypes[i]);
    if (!sql_catalog_has_aggregate(altfn_schema, altfn_name, altfn_argtypes))
    {
        ctx.reason = "no alternative aggregate function \"" + altfn_name + "\" exists";
        return JsonPtr();
    }

    JsonPtr altnode = Json::object();
    altnode->set("node", "Aggref");
    altnode->set("aggschema", altfn_schema == "pgstrom" ? "pgstrom" : "");
    altnode->set("aggname", altfn_name);
    altnode->set("aggtype", aggref->s("aggtype"));
    JsonPtr jt = Json::array();
    for (auto &t : altfn_argtypes)
        jt->push(Json::string(t));
    altnode->set("aggargtypes", jt);
    altnode->setb("aggstar", false);
    altnode->set("aggfilter", Json::null());    /* moved to GpuPreAgg */
    altnode->setb("args_are_subplan_outputs", true);
    /* remember what it replaces (used by tests / EXPLAIN of the host side) */
    altnode->set("orig_aggname", aggref->s("aggname"));
    altnode->set("orig_aggargtypes", aggref->getp("aggargtypes") ? aggref->getp("aggargtypes") : Json::array());
    JsonPtr altargs = Json::array();

    for (int i = 0; i < aggfn_cat->altfn_nargs; i++)
    {
        int         code = aggfn_cat->altfn_argexprs[i];
        std::string argtype = aggfn_cat->altfn_argtypes[i];
        JsonPtr     expr;
        JsonPtr     filter = aggref_filter(aggref);

        switch (code)
        {
            case ALTFUNC_EXPR_NROWS:
                expr = make_altfunc_nrows_expr(aggref);
                break;
            case ALTFUNC_EXPR_PMIN:
            case ALTFUNC_EXPR_PMAX:
                expr = aggref_arg(aggref, 0);
                if (filter)
                    expr = make_expr_conditional(expr, filter, JsonPtr());
                expr = make_altfunc_expr(code == ALTFUNC_EXPR_PMIN ? "pmin" : "pmax", {expr});
                break;
            case ALTFUNC_EXPR_PSUM:
            case ALTFUNC_EXPR_PSUM_X2:
                expr = aggref_arg(aggref, 0);
                if (expr_type(expr) != argtype)
                    expr = make_expr_typecast(expr, argtype);
                /* The reference makes the unmatched branch a zero constant
                 * (gpupreagg.c:883-892); sum(x) FILTER (WHERE ...) then yields
                 * 0 where PostgreSQL yields NULL for a group no row of which
                 * matches.  NULL is ignored by PSUM like any NULL input, and
                 * the final functions are strict on it. */
                if (expr && filter)
                    expr = make_expr_conditional(expr, filter, JsonPtr());
                if (expr)
                    expr = make_altfunc_expr(code == ALTFUNC_EXPR_PSUM ? "psum" : "psum_x2", {expr});
                break;
            case ALTFUNC_EXPR_PCOV_X:  expr = make_altfunc_pcov_expr(aggref, "pcov_x"); break;
            case ALTFUNC_EXPR_PCOV_Y:  expr = make_altfunc_pcov_expr(aggref, "pcov_y"); break;
            case ALTFUNC_EXPR_PCOV_X2: expr = make_altfunc_pcov_expr(aggref, "pcov_x2"); break;
            case ALTFUNC_EXPR_PCOV_Y2: expr = make_altfunc_pcov_expr(aggref, "pcov_y2"); break;
            case ALTFUNC_EXPR_PCOV_XY: expr = make_altfunc_pcov_expr(aggref, "pcov_xy"); break;
        }
        /* does aggregate function contained unsupported expression? */
        if (!expr)
        {
            ctx.reason = "no partial function for " + aggref->s("aggname");
            return JsonPtr();
        }
        /* check return type of the alternative functions */
        if (argtype != expr_type(expr))
        {
            ctx.reason = "Bug? result type is \"" + expr_type(expr) + "\", but \"" + argtype + "\" is expected";
            return JsonPtr();
        }
        /* add this expression node on the prep_tlist */
        int resno = 0;
        for (auto &tle : ctx.pre_tlist)
            if (expr_equal(tle->getp("expr"), expr))
            {
                resno = (int)tle->i("resno");
                break;
            }
        if (!resno)
        {
            JsonPtr tle = Json::object();
            resno = (int)ctx.pre_tlist.size() + 1;
            tle->set("node", "TargetEntry");
            tle->set("expr", expr);
            tle->set("resno", resno);
            tle->set("resname", Json::null());
            tle->setb("resjunk", false);
            ctx.pre_tlist.push_back(tle);
        }
        /* alternative aggregate function shall reference this resource. */
        JsonPtr varref = make_var(resno, expr_type(expr));
        varref->set("varno", "OUTER_VAR");
        altargs->push(varref);
    }
    altnode->set("args", altargs);
    return altnode;
}

/* gpupreagg_rewrite_mutator (gpupreagg.c:991-1031) */
static JsonPtr
gpupreagg_rewrite_mutator(const JsonPtr &node, rewrite_context &ctx)
{
    if (!node || node->is_null())
        return node;
    if (node->kind == Json::Array)
    {
        JsonPtr a = Json::array();
        for (auto &x : node->arr)
            a->push(gpupreagg_rewrite_mutator(x, ctx));
        return a;
    }
    if (node->kind != Json::Object)
        return node;
    std::string tag = node->s("node");
    if (tag == "Aggref")
    {
        JsonPtr alt = make_gpupreagg_refnode(node, ctx);
        if (!alt)
            ctx.invalid = true;
        return alt ? alt : Json::null();
    }
    if (tag == "Var")
    {
        int attno = (int)node->i("varattno");
        for (int g : *ctx.grp_col_idx)
            if (g == attno)
                return node;
        ctx.invalid = true;
        ctx.reason = "Var outside of the grouping keys";
        return Json::null();
    }
    /* expression_tree_mutator: copy with mutated children */
    JsonPtr copy = Json::object();
    for (auto &kv : node->obj)
    {
        if (kv.second && (kv.second->kind == Json::Object || kv.second->kind == Json::Array))
            copy->set(kv.first, gpupreagg_rewrite_mutator(kv.second, ctx));
        else
            copy->set(kv.first, kv.second);
    }
    return copy;
}

/* collect varattnos referenced by an expression */
static void
pull_varattnos(const JsonPtr &node, std::set<int> &attrs)
{
    if (!node)
        return;
    if (node->kind == Json::Array)
    {
        for (auto &x : node->arr)
            pull_varattnos(x, attrs);
        return;
    }
    if (node->kind != Json::Object)
        return;
    if (node->s("node") == "Var")
    {
        attrs.insert((int)node->i("varattno"));
        return;
    }
    for (auto &kv : node->obj)
        pull_varattnos(kv.second, attrs);
}

/* is psum's argument an int2/int4 value widened to int8?  Then a 64-bit
 * sum cannot overflow below 2^31 rows x 2^32 chunks and one cell suffices */
static bool
psum_arg_is_widened_int(const JsonPtr &arg)
{
    if (!arg || arg->is_null())
        return false;
    std::string tag = arg->s("node");
    if (tag == "FuncExpr" && arg->s("funcname") == "int8")
    {
        const Json *a = arg->get("args");
        if (a && a->arr.size() == 1)
        {
            std::string t = expr_type(a->arr[0]);
            return t == "int2" || t == "int4";
        }
        return false;
    }
    if (tag == "CaseExpr")
    {
        const Json *a = arg->get("args");
        for (size_t i = 0; a && i < a->arr.size(); i++)
            if (!psum_arg_is_widened_int(a->arr[i]->getp("result")))
                return false;
        JsonPtr def = arg->getp("defresult");
        if (def && !def->is_null())
            return psum_arg_is_widened_int(def);
        return true;
    }
    if (tag == "Const")
    {
        /* a literal counts only when its value fits 32 bits (the addend of
         * a LONGS cell is taken from the low word and sign-extended) */
        if (arg->flag("constisnull"))
            return true;
        const Json *jv = arg->get("constvalue");
        if (!jv || (jv->kind != Json::String && jv->kind != Json::Number))
            return false;
        char *end = NULL;
        long long v = strtoll(jv->str.c_str(), &end, 10);
        return (end && *end == '\0' && !jv->str.empty() &&
                v >= -2147483648LL && v <= 2147483647LL);
    }
    return false;
}

static const char *
aggcalc_method_of_type(const std::string &t)
{
    if (t == "int2") return "SHORT";
    if (t == "int4") return "INT";
    if (t == "int8") return "LONG";
    if (t == "float4") return "FLOAT";
    if (t == "float8") return "DOUBLE";
    if (t == "numeric") return "NUMERIC";
    return NULL;
}

/* ------------------------------------------------------------------
 * gpupreagg_codegen (gpupreagg.c:1902-1943)
 * ------------------------------------------------------------------ */
static bool
gpupreagg_codegen(GpuPreAggPlan &gp, const std::vector<JsonPtr> &pre_tlist,
                  const std::vector<JsonPtr> &outer_quals,
                  const std::vector<JsonPtr> &outer_tlist, std::string *err)
{
    CodegenContext context;
    std::ostringstream defs, fn_qual, fn_proj;
    std::ostringstream body, decl1;
    std::string gpagg_atts(pre_tlist.size(), (char)GPUPREAGG_FIELD_IS_NULL);
    std::set<int> attr_refs;
    std::set<int> qual_refs;        /* columns the qual reads */
    bool ok = true;
    bool use_temp_int4 = false, use_temp_float8x = false, use_temp_float8y = false;

    /* KPARAM_0 is an array of cl_char to inform which field is grouping
     * keys, or target of (partial) aggregate function. */
    JsonPtr kparam_0 = make_null_const("bytea");
    context.used_params.push_back(kparam_0);

    /* ---- gpupreagg_qual_eval (gpupreagg.c:1850-1900) ---- */
    fn_qual << "template <typename KDS>\n"
            << "DEVFN bool\n"
            << "gpupreagg_qual_eval(cl_int *errcode,\n"
            << "                    const kern_parambuf *kparams,\n"
            << "                    const KDS &kds,\n"
            << "                    const void *ktoast,\n"
            << "                    cl_uint kds_index)\n"
            << "{\n";
    if (!outer_quals.empty())
    {
        JsonPtr quals = Json::array();
        for (auto &q : outer_quals)
            quals->push(q);
        context.param_refs.clear();
        context.used_vars.clear();
        std::string expr_code = codegen_expression(quals, context, &ok);
        if (!ok)
        {
            *err = "outer qualifier is not device runnable";
            return false;
        }
        fn_qual << codegen_param_declarations(context, context.param_refs)
                << codegen_var_declarations(context)
                << "\n"
                << "  return EVAL(" << expr_code << ");\n";
        for (auto &q : outer_quals)
        {
            pull_varattnos(q, attr_refs);
            pull_varattnos(q, qual_refs);
        }
    }
    else
        fn_qual << "  return true;\n";
    fn_qual << "}\n";
    /* a qual that follows a varlena offset (numeric, text, bpchar) must not
     * be evaluated on the unused rows of a partial batch */
    const bool qual_derefs =
        (context.extra_flags & (DEVFUNC_NEEDS_NUMERIC | DEVFUNC_NEEDS_TEXTLIB)) != 0;

    /* ---- gpupreagg_projection (gpupreagg.c:1449-1837) ---- */
    context.param_refs.clear();
    context.used_vars.clear();
    std::set<int> proj_refs;
    int nkeys = 0, naggs = 0, ncells = 0;
    std::ostringstream key_list, agg_list, out_list, role_fn, index_fn;
    /* aggregates whose argument has the NULL-ness of the same outer column
     * share one "saw a non-NULL input" test per row */
    std::vector<std::pair<std::string, std::pair<int, unsigned> > > nn_classes;

    for (size_t i = 0; i < pre_tlist.size(); i++)
    {
        const JsonPtr &tle = pre_tlist[i];
        JsonPtr expr = tle->getp("expr");
        std::string tag = expr->s("node");
        int resno = (int)tle->i("resno");
        PartialColumn pc;

        pc.resno = resno;
        pc.expr = expr;
        pc.type = expr_type(expr);
        pc.agg_index = -1;
        pc.cell_index = -1;
        if (tag == "Var")
        {
            const DevType *dtype = devtype_lookup(pc.type);
            int attno = (int)expr->i("varattno");
            /* variable-length keys: text / bpchar travel as "kernel text"
             * (kern_textlib.cuh); numeric / bytea keys stay on the host */
            if (!dtype || ((dtype->type_flags & DEVTYPE_IS_VARLENA) &&
                           pc.type != "text" && pc.type != "bpchar"))
            {
                *err = "grouping key of type " + pc.type + " is not supported on the device yet";
                return false;
            }
            context.track_type(dtype);
            proj_refs.insert(attno);
            body << "  /* projection for resource " << (resno - 1) << " */\n";
            if (pc.type == "float4" || pc.type == "float8")
                body << "  KVAR_" << attno << ".value = (" << dtype->type_base
                     << ")pgs_f8_canon((double)KVAR_" << attno << ".value);\n";
            body << "  pg_" << dtype->type_name << "_vstore(kds_src,kds_in,errcode,"
                 << (resno - 1) << ",rowidx_out,KVAR_" << attno << ");\n";
            gpagg_atts[resno - 1] = (char)GPUPREAGG_FIELD_IS_GROUPKEY;
            pc.role = GPUPREAGG_FIELD_IS_GROUPKEY;
            pc.agg_index = nkeys;
            key_list << " _(" << nkeys << "," << (resno - 1) << "," << dtype->type_name << ")";
            out_list << " _(" << (resno - 1) << ",KEY," << nkeys << ",0,NONE,NONE)";
            nkeys++;
        }
        else if (tag == "Const")
        {
            body << "  /* projection for resource " << (resno - 1) << " */\n"
                 << "  pg_common_vstore(kds_src,kds_in,errcode," << (resno - 1)
                 << ",rowidx_out,true);\n";
            pc.role = GPUPREAGG_FIELD_IS_NULL;
            out_list << " _(" << (resno - 1) << ",NUL,0,0,NONE,NONE)";
        }
        else if (tag == "FuncExpr" && expr->s("funcschema") == "pgstrom")
        {
            std::string func_name = expr->s("funcname");
            const Json *fargs = expr->get("args");
            const DevType *dtype;

            body << "  /* projection for resource " << (resno - 1) << " */\n";
            pull_varattnos(expr, proj_refs);
            pc.role = GPUPREAGG_FIELD_IS_AGGFUNC;
            pc.func = func_name;
            if (func_name == "nrows")
            {
                dtype = devtype_lookup("int4");
                context.track_type(dtype);
                use_temp_int4 = true;
                body << "  temp_int4.isnull = false;\n";
                if (fargs && !fargs->arr.empty())
                {
                    body << "  if (";
                    for (size_t k = 0; k < fargs->arr.size(); k++)
                    {
                        std::string code = codegen_expression(fargs->arr[k], context, &ok);
                        if (!ok) { *err = "nrows() argument not device runnable"; return false; }
                        if (k) body << " &&\n      ";
                        body << "EVAL(" << code << ")";
                    }
                    body << ")\n    temp_int4.value = 1;\n  else\n    temp_int4.value = 0;\n";
                }
                else
                    body << "  temp_int4.value = 1;\n";
                body << "  pg_int4_vstore(kds_src,kds_in,errcode," << (resno - 1)
                     << ",rowidx_out,temp_int4);\n";
                pc.op = "PSUM";
                pc.cell_type = "INT";
            }
            else if (func_name == "pmax" || func_name == "pmin" || func_name == "psum")
            {
                JsonPtr clause = fargs->arr[0];
                std::string type_name = expr_type(clause);
                dtype = devtype_lookup(type_name);
                const char *method = aggcalc_method_of_type(type_name);
                if (!dtype || !method) { *err = "unexpected partial aggregate data-type"; return false; }
                context.track_type(dtype);
                std::string code = codegen_expression(clause, context, &ok);
                if (!ok) { *err = func_name + "() argument not device runnable"; return false; }
                body << "  pg_" << dtype->type_name << "_vstore(kds_src,kds_in,errcode,"
                     << (resno - 1) << ",rowidx_out," << code << ");\n";
                pc.op = (func_name == "pmax" ? "PMAX" : func_name == "pmin" ? "PMIN" : "PSUM");
                pc.cell_type = method;
                if (func_name == "psum" && type_name == "int8" && psum_arg_is_widened_int(clause))
                    pc.cell_type = "LONGS";
            }
            else if (func_name == "psum_x2")
            {
                JsonPtr clause = fargs->arr[0];
                if (expr_type(clause) != "float8")
                { *err = "psum_x2 on " + expr_type(clause) + " is not supported on the device"; return false; }
                dtype = devtype_lookup("float8");
                const DevFunc *dfunc = devfunc_lookup("float8mul", {"float8", "float8"}, "float8");
                context.track_func(dfunc);
                use_temp_float8x = true;
                std::string code = codegen_expression(clause, context, &ok);
                if (!ok) { *err = "psum_x2() argument not device runnable"; return false; }
                body << "  temp_float8x = " << code << ";\n"
                     << "  pg_float8_vstore(kds_src,kds_in,errcode," << (resno - 1) << ",rowidx_out,\n"
                     << "               pgfn_" << dfunc->func_alias << "(errcode, temp_float8x,\n"
                     << "                                temp_float8x));\n";
                pc.op = "PSUM";
                pc.cell_type = "DOUBLE";
            }
            else if (func_name.compare(0, 5, "pcov_") == 0)
            {
                JsonPtr filter = fargs->arr[0];
                JsonPtr x_clause = fargs->arr[1];
                JsonPtr y_clause = fargs->arr[2];
                const DevFunc *dfunc = devfunc_lookup("float8mul", {"float8", "float8"}, "float8");

                use_temp_float8x = use_temp_float8y = true;
                context.track_func(dfunc);
                if (filter->s("node") == "Const" && expr_type(filter) == "bool" &&
                    !filter->flag("constisnull") && filter->s("constvalue") == "t")
                    filter = JsonPtr();     /* no filter, actually */
                std::string xcode = codegen_expression(x_clause, context, &ok);
                if (!ok) { *err = "pcov argument not device runnable"; return false; }
                std::string ycode = codegen_expression(y_clause, context, &ok);
                if (!ok) { *err = "pcov argument not device runnable"; return false; }
                body << "  temp_float8x = " << xcode << ";\n"
                     << "  temp_float8y = " << ycode << ";\n"
                     << "  if (temp_float8x.isnull ||\n"
                     << "      temp_float8y.isnull";
                if (filter)
                {
                    std::string fcode = codegen_expression(filter, context, &ok);
                    if (!ok) { *err = "pcov filter not device runnable"; return false; }
                    body << " ||\n      !EVAL(" << fcode << ")";
                }
                body << ")\n  {\n    temp_float8x.isnull = true;\n    temp_float8x.value = 0.0;\n  }\n";
                if (func_name == "pcov_y")
                    body << "  else\n    temp_float8x = temp_float8y;\n";
                else if (func_name == "pcov_x2")
                    body << "  else\n    temp_float8x = pgfn_" << dfunc->func_alias
                         << "(errcode,\n                           temp_float8x,\n                           temp_float8x);\n";
                else if (func_name == "pcov_y2")
                    body << "  else\n    temp_float8x = pgfn_" << dfunc->func_alias
                         << "(errcode,\n                           temp_float8y,\n                           temp_float8y);\n";
                else if (func_name == "pcov_xy")
                    body << "  else\n    temp_float8x = pgfn_" << dfunc->func_alias
                         << "(errcode,\n                           temp_float8x,\n                           temp_float8y);\n";
                else if (func_name != "pcov_x")
                { *err = "unexpected partial covariance function: " + func_name; return false; }
                body << "  pg_float8_vstore(kds_src,kds_in,errcode," << (resno - 1)
                     << ",rowidx_out,temp_float8x);\n";
                pc.op = "PSUM";
                pc.cell_type = "DOUBLE";
            }
            else
            {
                *err = "Bug? unexpected partial aggregate function: " + func_name;
                return false;
            }
            /* NULL-ness class of this partial value */
            {
                std::string klass = "agg:" + std::to_string(naggs);
                if (func_name == "nrows")
                    klass = "never-null";
                else if (func_name == "pmax" || func_name == "pmin" || func_name == "psum")
                {
                    JsonPtr a = fargs->arr[0];
                    /* strict casts / relabels keep NULL-ness */
                    while (a && ((a->s("node") == "FuncExpr" && a->s("funcschema").empty() &&
                                  a->get("args") && a->get("args")->arr.size() == 1 &&
                                  (a->s("funcformat") == "cast" || a->s("funcformat") == "implicit")) ||
                                 a->s("node") == "RelabelType"))
                        a = (a->s("node") == "RelabelType") ? a->getp("arg") : a->get("args")->arr[0];
                    if (a && a->s("node") == "Var")
                        klass = "var:" + std::to_string(a->i("varattno"));
                }
                bool found = false;
                for (auto &kc : nn_classes)
                    if (kc.first == klass)
                    {
                        kc.second.second |= (1U << naggs);
                        found = true;
                    }
                if (!found)
                    nn_classes.push_back(std::make_pair(klass, std::make_pair(naggs, 1U << naggs)));
            }
            /* track usage of this field */
            gpagg_atts[resno - 1] = (char)GPUPREAGG_FIELD_IS_AGGFUNC;
            pc.agg_index = naggs;
            pc.cell_index = ncells;
            agg_list << " _(" << naggs << "," << ncells << "," << pc.op << "," << pc.cell_type << ")";
            out_list << " _(" << (resno - 1) << ",AGG," << naggs << "," << ncells << ","
                     << pc.op << "," << pc.cell_type << ")";
            naggs++;
            /* int8 sums are 128-bit (2 cells); numeric sums 128-bit at a fixed
             * scale + the display scale (3 cells, kern_numeric.cuh) */
            ncells += (pc.cell_type == "LONG" ? 2 :
                       (pc.cell_type == "NUMERIC" && pc.op == "PSUM") ? 3 : 1);
        }
        else
        {
            *err = "bug? unexpected node type in GpuPreAgg target list";
            return false;
        }
        gp.columns.push_back(pc);
    }
    if (naggs > 32)
    {
        *err = "too many partial aggregates for one kernel (max 32)";
        return false;
    }
    if (nkeys > 16)
    {
        *err = "too many grouping keys for one kernel (max 16)";
        return false;
    }
    /* declaration of referenced outer variables */
    for (auto &tle : outer_tlist)
    {
        int resno = (int)tle->i("resno");
        if (!proj_refs.count(resno))
            continue;
        const DevType *dtype = devtype_lookup(expr_type(tle));
        if (!dtype) { *err = "outer column type not supported"; return false; }
        context.track_type(dtype);
        decl1 << "  pg_" << dtype->type_name << "_t KVAR_" << resno
              << " = pg_" << dtype->type_name << "_vref(kds_in,ktoast,errcode,"
              << (resno - 1) << ",rowidx_in);\n";
        attr_refs.insert(resno);
    }
    std::string decl2 = codegen_param_declarations(context, context.param_refs);
    if (use_temp_int4) decl1 << "  pg_int4_t temp_int4;\n";
    if (use_temp_float8x) decl1 << "  pg_float8_t temp_float8x;\n";
    if (use_temp_float8y) decl1 << "  pg_float8_t temp_float8y;\n";
    fn_proj << "template <typename KDS>\n"
            << "DEVFN void\n"
            << "gpupreagg_projection(cl_int *errcode,\n"
            << "            const kern_parambuf *kparams,\n"
            << "            const KDS &kds_in,\n"
            << "            pagg_row &kds_src,\n"
            << "            const void *ktoast,\n"
            << "            cl_uint rowidx_in, cl_uint rowidx_out)\n"
            << "{\n"
            << decl2 << decl1.str() << "\n" << body.str()
            << "}\n";

    /* KPARAM_0 */
    {
        static const char hexd[] = "0123456789abcdef";
        std::string hex;
        for (unsigned char c : gpagg_atts)
        { hex += hexd[c >> 4]; hex += hexd[c & 15]; }
        kparam_0->setb("constisnull", false);
        kparam_0->set("constvalue", "");
        kparam_0->set("constbytes", hex);
    }

    /* ---- compile-time description of the query for the kernel templates ---- */
    int nincols = 0;
    std::ostringstream incol_list, slot_fn, attlen_fn, staged_fn;
    /* GROUP BY under a WHERE clause (PGSTROM_GATHER_PAYLOAD=0 turns it off):
     * only the qual's columns go through the TMA staging ring; the rows that
     * pass fetch their other columns from HBM by row number
     * (kern_gpupreagg.cuh, GPUPREAGG_GATHER_PAYLOAD).  Measured on B200, 50M
     * rows of where_agg, per cent of the HBM roofline with / without: qual
     * keeps 1% of the rows 93 / 53, 10% 54 / 43, 50% 20 / 20. */
    bool gather_payload = false;
    {
        const char *env = getenv("PGSTROM_GATHER_PAYLOAD");
        bool unstaged = false;
        for (int attno : attr_refs)
            if (!qual_refs.count(attno))
                unstaged = true;
        gather_payload = (!(env && atoi(env) == 0) && !outer_quals.empty() && nkeys > 0 &&
                          !(gp.num_groups >= 65536.0) && unstaged);
    }
    gp.row_bytes = 0;
    for (int attno : attr_refs)
    {
        const JsonPtr &tle = outer_tlist.at(attno - 1);
        const DevType *dtype = devtype_lookup(expr_type(tle));
        if (!dtype) { *err = "outer column type not supported"; return false; }
        int attlen = (dtype->type_length > 0 ? dtype->type_length : 4);
        incol_list << " _(" << nincols << "," << (attno - 1) << "," << attlen << ")";
        slot_fn << "    case " << (attno - 1) << ": return " << nincols << ";\n";
        attlen_fn << "    case " << nincols << ": return " << attlen << ";\n";
        if (gather_payload && !qual_refs.count(attno))
            staged_fn << "    case " << nincols << ": return 0;\n";
        gp.incol_index.push_back(attno - 1);
        gp.row_bytes += attlen;
        nincols++;
    }
    defs << "/* ---- generated by pg_strom_b200 (GpuPreAgg) ---- */\n"
         << "#define GPUPREAGG_NUM_INCOLS " << nincols << "\n"
         << "#define GPUPREAGG_INCOL_LIST(_)" << incol_list.str() << "\n"
         << "__host__ __device__ constexpr int\nGPUPREAGG_INCOL_SLOT(unsigned int colidx)\n{\n"
         << "  switch (colidx)\n  {\n" << slot_fn.str() << "    default: return 0;\n  }\n}\n"
         << "__host__ __device__ constexpr unsigned int\nGPUPREAGG_INCOL_ATTLEN(int slot)\n{\n"
         << "  switch (slot)\n  {\n" << attlen_fn.str() << "    default: return 0;\n  }\n}\n"
         /* 1: the column goes through the staging ring */
         << "__host__ __device__ constexpr unsigned int\nGPUPREAGG_INCOL_STAGED(int slot)\n{\n"
         << "  switch (slot)\n  {\n" << staged_fn.str() << "    default: return 1;\n  }\n}\n"
         << "#define GPUPREAGG_GATHER_PAYLOAD " << (gather_payload ? 1 : 0) << "\n"
         << "#define GPUPREAGG_NUM_KEYS " << nkeys << "\n"
         << "#define GPUPREAGG_KEY_LIST(_)" << key_list.str() << "\n"
         << "#define GPUPREAGG_NUM_AGGS " << naggs << "\n"
         << "#define GPUPREAGG_NUM_CELLS " << ncells << "\n"
         << "#define GPUPREAGG_AGG_LIST(_)" << agg_list.str() << "\n"
         << "#define GPUPREAGG_NNCLASS_LIST(_)" << [&]() {
                std::ostringstream o;
                for (auto &kc : nn_classes)
                    o << " _(" << kc.second.first << ",0x" << std::hex << kc.second.second << std::dec << "U)";
                return o.str(); }() << "\n"
         << "#define GPUPREAGG_HAS_QUAL " << (outer_quals.empty() ? 0 : 1) << "\n"
         << (qual_derefs ? "#define GPUPREAGG_QUAL_DEREFS 1\n" : "")
         /* partitioned aggregation is compiled in when the planner expects
          * very many groups (gpupreagg_partagg; the CUDA layer switches it on
          * with the same threshold) */
         << "#define GPUPREAGG_PARTITIONED "
         << ((nkeys > 0 && gp.num_groups >= 65536.0) ? 1 : 0) << "\n"
         << "#define GPUPREAGG_NUM_OUTCOLS " << pre_tlist.size() << "\n"
         << "#define GPUPREAGG_OUT_LIST(_)" << out_list.str() << "\n";
    role_fn << "__host__ __device__ constexpr int\nGPUPREAGG_FIELD_ROLE(unsigned int colidx)\n{\n  switch (colidx)\n  {\n";
    index_fn << "__host__ __device__ constexpr int\nGPUPREAGG_FIELD_INDEX(unsigned int colidx)\n{\n  switch (colidx)\n  {\n";
    for (auto &pc : gp.columns)
    {
        if (pc.role == GPUPREAGG_FIELD_IS_NULL)
            continue;
        role_fn << "    case " << (pc.resno - 1) << ": return " << pc.role << ";\n";
        index_fn << "    case " << (pc.resno - 1) << ": return " << pc.agg_index << ";\n";
    }
    role_fn << "    default: return 0;\n  }\n}\n";
    index_fn << "    default: return 0;\n  }\n}\n";
    defs << role_fn.str() << index_fn.str();

    std::ostringstream src;
    src << "#include \"pgstrom_kds.h\"\n"
        << defs.str()
        << "#include \"kern_common.cuh\"\n";
    if (context.extra_flags & DEVFUNC_NEEDS_NUMERIC)
        src << "#include \"kern_numeric.cuh\"\n";
    if (context.extra_flags & DEVFUNC_NEEDS_TIMELIB)
        src << "#include \"kern_timelib.cuh\"\n";
    if (context.extra_flags & DEVFUNC_NEEDS_TEXTLIB)
        src << "#include \"kern_textlib.cuh\"\n";
    src << "#include \"kern_gpupreagg.cuh\"\n"
        << "\n"
        << codegen_func_declarations(context) << "\n"
        << fn_qual.str() << "\n"
        << fn_proj.str() << "\n";
    gp.kern_source = src.str();
    gp.extra_flags = context.extra_flags | DEVKERNEL_NEEDS_GPUPREAGG;
    gp.used_params = context.used_params;
    gp.kparams = create_kern_parambuf(context.used_params);
    gp.num_cells = ncells;
    return true;
}

/* ------------------------------------------------------------------
 * pgstrom_try_insert_gpupreagg (gpupreagg.c:1987-2187)
 * ------------------------------------------------------------------ */
static std::vector<JsonPtr>
json_list(const Json *j)
{
    std::vector<JsonPtr> v;
    if (j && j->kind == Json::Array)
        v.assign(j->arr.begin(), j->arr.end());
    return v;
}

/* ------------------------------------------------------------------
 * cost_gpupreagg (gpupreagg.c:366-464), the part that is the extension's own:
 * cost / rows / width of the GpuPreAgg node from the outer plan's estimates.
 * The costs of the Sort and the Agg above it are PostgreSQL's cost_sort() /
 * cost_agg() over these numbers (:466-509) and the decision "cheaper than
 * the plain Agg, unless pg_strom.debug_force_gpupreagg" (:2105-2118) is
 * taken by the glue, which owns those functions.
 * Only done when the outer plan carries PostgreSQL's estimates.
 * ------------------------------------------------------------------ */
static void
count_operator_nodes(const Json *j, int *count)
{
    if (!j)
        return;
    if (j->kind == Json::Array)
    {
        for (auto &e : j->arr)
            count_operator_nodes(e.get(), count);
        return;
    }
    if (j->kind != Json::Object)
        return;
    /* cost_qual_eval_walker: procost (1 for every built-in and for the
     * pgstrom.* placeholders) x cpu_operator_cost per function call */
    std::string n = j->s("node");
    if (n == "OpExpr" || n == "FuncExpr" || n == "DistinctExpr" ||
        n == "NullIfExpr" || n == "CoerceViaIO")
        (*count)++;
    for (auto &kv : j->obj)
        count_operator_nodes(kv.second.get(), count);
}

static int
typavgwidth(const std::string &type, int typmod)
{
    /* get_typavgwidth(): typlen for fixed-length types, typmod-derived or 32
     * for varlena */
    static const struct { const char *name; int len; } fixed[] = {
        {"bool", 1}, {"int2", 2}, {"int4", 4}, {"int8", 8}, {"float4", 4},
        {"float8", 8}, {"date", 4}, {"time", 8}, {"timestamp", 8},
        {"timestamptz", 8},
    };
    for (auto &f : fixed)
        if (type == f.name)
            return f.len;
    if (type == "bpchar" && typmod > 4)
        return typmod - 4;
    return 32;
}

static void
cost_gpupreagg(const JsonPtr &agg, const JsonPtr &outer_plan,
               const std::vector<JsonPtr> &pre_tlist, JsonPtr &gpreagg)
{
    if (!outer_plan->has("total_cost") || !outer_plan->has("plan_rows"))
        return;
    const double BLCKSZ_ = 8192.0, page_header = 24.0, item_id = 4.0, htup_header = 24.0;
    double cpu_operator_cost = agg->d("cpu_operator_cost", 0.0025);
    double gpu_operator_cost = guc_real("gpu_operator_cost");
    double startup_cost = outer_plan->d("startup_cost", 0.0);
    double run_cost = outer_plan->d("total_cost") - startup_cost;
    double outer_rows = outer_plan->d("plan_rows");
    double outer_width = outer_plan->d("plan_width", 0.0);
    double num_groups = std::max(agg->d("plan_rows", agg->d("numGroups", 1.0)), 1.0);
    auto maxalign = [](double x) { return std::ceil(x / 8.0) * 8.0; };
    auto log2_ = [](double x) { return std::log(x) / 0.693147180559945; };

    startup_cost += guc_real("gpu_setup_cost");
    double rows_per_chunk =
        std::floor((double)(guc_int("pg_strom.chunk_size") << 20) / BLCKSZ_) *
        (BLCKSZ_ - maxalign(page_header)) /
        (item_id + maxalign(htup_header + outer_width));
    double num_chunks = std::max(outer_rows / rows_per_chunk, 1.0);
    double comparison_cost = 2.0 * gpu_operator_cost;
    startup_cost += comparison_cost * log2_(rows_per_chunk * rows_per_chunk) * num_chunks;
    run_cost += gpu_operator_cost * outer_rows;

    int noperators = 0, pagg_width = 0;
    for (auto &tle : pre_tlist)
    {
        JsonPtr e = tle->getp("expr");
        count_operator_nodes(e.get(), &noperators);
        int typmod = (e && e->has("vartypmod")) ? (int)e->i("vartypmod") : -1;
        pagg_width += typavgwidth(expr_type(tle), typmod);
    }
    double per_tuple = noperators * cpu_operator_cost;
    run_cost += per_tuple * gpu_operator_cost / cpu_operator_cost *
        log2_(rows_per_chunk) * num_chunks;

    gpreagg->set("startup_cost", Json::number(startup_cost));
    gpreagg->set("total_cost", Json::number(startup_cost + run_cost));
    gpreagg->set("plan_rows", Json::number(num_groups * num_chunks));
    gpreagg->set("plan_width", pagg_width);
}

GpuPreAggPlan
pgstrom_try_insert_gpupreagg(const JsonPtr &agg)
{
    GpuPreAggPlan gp;
    rewrite_context ctx;

    /* nothing to do, if feature is turned off */
    if (!pgstrom_enabled() || !guc_bool("enable_gpupreagg"))
    {
        gp.reject_reason = "pg_strom.enabled or enable_gpupreagg is off";
        return gp;
    }
    JsonPtr outer_plan = agg->getp("lefttree");
    JsonPtr sort_plan;
    std::string strategy = agg->s("aggstrategy", "plain");

    if (!outer_plan)
    {
        gp.reject_reason = "Agg without outer plan";
        return gp;
    }
    /* In case of sort-aggregate, GpuPreAgg is injected under the Sort */
    if (outer_plan->s("node") == "Sort")
    {
        sort_plan = outer_plan;
        outer_plan = outer_plan->getp("lefttree");
    }
    /* Any plan node can feed GpuPreAgg (gpupreagg.c:2031-2107): a SeqScan is
     * replaced by a GpuScan whose device quals move into this kernel
     * (gpuscan_try_replace_seqscan_plan, gpuscan.c:378-517); every other
     * node - a join, a Result, an index scan - stays what it is, keeps its
     * quals and hands its tuples over one by one (gpupreagg_load_next_outer,
     * gpupreagg.c:2418-2505: pgstrom_data_store_insert_tuple into ROW_FLAT
     * chunks). */
    std::string onode = outer_plan->s("node");
    bool outer_is_scan = (onode == "SeqScan" || onode == "GpuScan");
    if (onode.empty() || !outer_plan->getp("targetlist"))
    {
        gp.reject_reason = "outer plan " + onode + " cannot feed GpuPreAgg";
        return gp;
    }
    /* a chunk needs at least one column ("select sum(1E+48)" over a Result,
     * recheck_agg.sql, stays on the CPU) */
    if (json_list(outer_plan->get("targetlist")).empty())
    {
        gp.reject_reason = "outer plan " + onode + " has no output column";
        return gp;
    }
    for (auto &g : json_list(agg->get("grpColIdx")))
        gp.grp_col_idx.push_back((int)g->num);
    ctx.grp_col_idx = &gp.grp_col_idx;

    /* gpupreagg_rewrite_expr (gpupreagg.c:1033-1166): head of target-list
     * keeps the outer relation's column order; non-key columns become NULL */
    std::vector<JsonPtr> outer_tlist = json_list(outer_plan->get("targetlist"));
    for (auto &tle : outer_tlist)
    {
        int resno = (int)tle->i("resno");
        std::string type = expr_type(tle);
        bool is_key = false;
        JsonPtr tle_new = Json::object();

        gp.outer_colnames.push_back(tle->s("resname"));
        gp.outer_coltypes.push_back(type);
        for (int g : gp.grp_col_idx)
            if (g == resno)
                is_key = true;
        tle_new->set("node", "TargetEntry");
        if (is_key)
        {
            const DevType *dtype = devtype_lookup(type);
            /* grouping key must be a supported data type with comparison */
            if (!dtype || !dtype->type_cmpfunc)
            {
                gp.reject_reason = "grouping key type " + type + " is not supported";
                return gp;
            }
            JsonPtr keyvar = make_var(resno, type);
            /* character(n): the host pads the key back to n (kernel text) */
            JsonPtr src = tle->getp("expr");
            if (src && src->s("node") == "Var" && src->has("vartypmod"))
                keyvar->set("vartypmod", (int)src->i("vartypmod"));
            tle_new->set("expr", keyvar);
        }
        else
            tle_new->set("expr", make_null_const(type));
        tle_new->set("resno", (int)ctx.pre_tlist.size() + 1);
        tle_new->set("resname", tle->getp("resname") ? tle->getp("resname") : Json::null());
        tle_new->setb("resjunk", tle->flag("resjunk"));
        ctx.pre_tlist.push_back(tle_new);
    }
    /* replace aggregate functions in tlist / qual of the Agg node */
    JsonPtr agg_tlist = Json::array();
    for (auto &oldtle : json_list(agg->get("targetlist")))
    {
        JsonPtr newtle = Json::object();
        for (auto &kv : oldtle->obj)
            newtle->set(kv.first, kv.second);
        newtle->set("expr", gpupreagg_rewrite_mutator(oldtle->getp("expr"), ctx));
        if (ctx.invalid)
        {
            gp.reject_reason = ctx.reason;
            return gp;
        }
        agg_tlist->push(newtle);
    }
    JsonPtr agg_quals = Json::array();
    for (auto &q : json_list(agg->get("qual")))
    {
        agg_quals->push(gpupreagg_rewrite_mutator(q, ctx));
        if (ctx.invalid)
        {
            gp.reject_reason = ctx.reason;
            return gp;
        }
    }

    /* pull up device-runnable qualifiers of the scan
     * (gpuscan_try_replace_seqscan_plan, gpuscan.c:378-517) */
    std::vector<JsonPtr> outer_quals, host_quals;
    if (outer_is_scan)
    {
        for (auto &q : json_list(outer_plan->get("qual")))
        {
            if (codegen_available_expression(q))
                outer_quals.push_back(q);
            else
                host_quals.push_back(q);
        }
    }
    gp.outer_bulkload = outer_is_scan && host_quals.empty();
    gp.needs_grouping = !gp.grp_col_idx.empty();
    gp.num_groups = agg->d("numGroups", agg->d("plan_rows", 1.0));
    if (gp.num_groups < 1.0)
        gp.num_groups = 1.0;

    std::string err;
    if (!gpupreagg_codegen(gp, ctx.pre_tlist, outer_quals, outer_tlist, &err))
    {
        gp.reject_reason = err;
        gp.columns.clear();
        return gp;
    }
    gp.extra_flags |= ctx.extra_flags;

    /* ---- splice: Agg(alt aggregates) -> [Sort] -> GpuPreAgg -> GpuScan ---- */
    JsonPtr gpuscan = outer_plan;                /* not a scan: left alone */
    if (outer_is_scan)
    {
        gpuscan = Json::object();
        for (auto &kv : outer_plan->obj)
            gpuscan->set(kv.first, kv.second);
        gpuscan->set("node", "CustomPlan");
        gpuscan->set("custom_name", "GpuScan");
        JsonPtr hq = Json::array();
        for (auto &q : host_quals) hq->push(q);
        gpuscan->set("qual", hq);
        gpuscan->set("dev_quals", Json::array());   /* moved up */
    }

    JsonPtr gpreagg = Json::object();
    gpreagg->set("node", "CustomPlan");
    gpreagg->set("custom_name", "GpuPreAgg");
    JsonPtr ptl = Json::array();
    for (auto &t : ctx.pre_tlist) ptl->push(t);
    gpreagg->set("targetlist", ptl);
    JsonPtr oq = Json::array();
    for (auto &q : outer_quals) oq->push(q);
    gpreagg->set("outer_quals", oq);
    gpreagg->setb("outer_bulkload", gp.outer_bulkload);
    gpreagg->set("grpColIdx", agg->getp("grpColIdx") ? agg->getp("grpColIdx") : Json::array());
    gpreagg->set("num_groups", Json::number(gp.num_groups));
    gpreagg->set("extra_flags", gp.extra_flags);
    gpreagg->set("lefttree", gpuscan);
    cost_gpupreagg(agg, outer_plan, ctx.pre_tlist, gpreagg);

    JsonPtr new_agg = Json::object();
    for (auto &kv : agg->obj)
        new_agg->set(kv.first, kv.second);
    new_agg->set("targetlist", agg_tlist);
    new_agg->set("qual", agg_quals);
    if (sort_plan)
    {
        JsonPtr new_sort = Json::object();
        for (auto &kv : sort_plan->obj)
            new_sort->set(kv.first, kv.second);
        /* the Sort passes the GpuPreAgg target list through */
        new_sort->set("targetlist", ptl);
        new_sort->set("lefttree", gpreagg);
        new_agg->set("lefttree", new_sort);
    }
    else
        new_agg->set("lefttree", gpreagg);

    gp.plan = new_agg;
    gp.gpreagg = gpreagg;
    gp.valid = true;
    return gp;
}

/* grafter_try_replace_recurse (grafter.c:24-117) */
static JsonPtr
grafter_try_replace_recurse(const JsonPtr &plan, std::vector<GpuPreAggPlan> *plans)
{
    if (!plan || plan->is_null())
        return plan;
    JsonPtr newnode = plan;
    if (plan->s("node") == "Agg")
    {
        GpuPreAggPlan gp = pgstrom_try_insert_gpupreagg(plan);
        if (gp.valid)
        {
            newnode = gp.plan;
            if (plans)
            {
                /* the idx the per-node entry points of the C ABI take
                 * (pgs_plan_kernel_source(plan, idx) ...): the glue reads it
                 * when it turns the tree back into plan nodes */
                gp.gpreagg->set("gpupreagg_index", (int)plans->size());
                plans->push_back(gp);
            }
            /* a GpuScan below is this node's own; any other outer plan
             * (join, sub-query scan) may hold further aggregates */
            JsonPtr child = gp.gpreagg->getp("lefttree");
            if (child && !child->is_null() &&
                !(child->s("node") == "CustomPlan" && child->s("custom_name") == "GpuScan"))
                gp.gpreagg->set("lefttree", grafter_try_replace_recurse(child, plans));
            return newnode;
        }
    }
    JsonPtr copy = Json::object();
    for (auto &kv : newnode->obj)
    {
        if (kv.first == "lefttree" || kv.first == "righttree")
            copy->set(kv.first, grafter_try_replace_recurse(kv.second, plans));
        else
            copy->set(kv.first, kv.second);
    }
    return copy;
}

JsonPtr
pgstrom_grafter(const JsonPtr &plan_tree, std::vector<GpuPreAggPlan> *plans)
{
    if (!pgstrom_enabled())
        return plan_tree;
    return grafter_try_replace_recurse(plan_tree, plans);
}

/* ------------------------------------------------------------------
 * EXPLAIN
 * ------------------------------------------------------------------ */
static JsonPtr
child_of(const JsonPtr &plan)
{
    JsonPtr c = plan->getp("lefttree");
    return (c && !c->is_null()) ? c : JsonPtr();
}

/* deparse an expression that lives in `plan`'s target list / quals: Vars
 * that point to the child's outputs print the child's expression */
static std::string deparse_in_plan(const JsonPtr &expr, const JsonPtr &plan);

static std::vector<std::string>
scan_colnames(const JsonPtr &scan)
{
    std::vector<std::string> names;
    for (auto &tle : json_list(scan->get("targetlist")))
        names.push_back(tle->s("resname"));
    return names;
}

static bool
is_scan_node(const JsonPtr &plan)
{
    std::string n = plan->s("node");
    return n == "SeqScan" || (n == "CustomPlan" && plan->s("custom_name") == "GpuScan");
}

/* substitute OUTER Vars by the child's target entry expressions */
static JsonPtr
resolve_vars(const JsonPtr &node, const JsonPtr &child, bool *is_plain_var)
{
    if (!node || node->kind != Json::Object)
    {
        if (node && node->kind == Json::Array)
        {
            JsonPtr a = Json::array();
            for (auto &x : node->arr)
                a->push(resolve_vars(x, child, NULL));
            return a;
        }
        return node;
    }
    if (node->s("node") == "Var")
    {
        std::vector<JsonPtr> ctl = json_list(child->get("targetlist"));
        int attno = (int)node->i("varattno");
        if (attno >= 1 && (size_t)attno <= ctl.size())
        {
            JsonPtr ce = ctl[attno - 1]->getp("expr");
            if (is_plain_var)
                *is_plain_var = (ce->s("node") == "Var");
            /* mark: came from a sub-plan output */
            JsonPtr w = Json::object();
            w->set("node", "SubplanRef");
            w->set("expr", ce);
            return w;
        }
        return node;
    }
    JsonPtr copy = Json::object();
    for (auto &kv : node->obj)
    {
        if (kv.second && (kv.second->kind == Json::Object || kv.second->kind == Json::Array))
            copy->set(kv.first, resolve_vars(kv.second, child, NULL));
        else
            copy->set(kv.first, kv.second);
    }
    return copy;
}

static std::string
deparse_resolved(const JsonPtr &node, const JsonPtr &scope_plan);

static std::string
deparse_in_plan(const JsonPtr &expr, const JsonPtr &plan)
{
    if (is_scan_node(plan))
        return deparse_expression(expr, scan_colnames(plan), false);
    JsonPtr child = child_of(plan);
    if (!child)
        return deparse_expression(expr, std::vector<std::string>(), false);
    JsonPtr resolved = resolve_vars(expr, child, NULL);
    return deparse_resolved(resolved, child);
}

/* deparse where SubplanRef{expr} nodes print the child's expression in the
 * child's own scope, parenthesised unless it is a plain column */
static std::string
deparse_resolved(const JsonPtr &node, const JsonPtr &child)
{
    if (!node || node->is_null())
        return "";
    if (node->s("node") == "SubplanRef")
    {
        JsonPtr ce = node->getp("expr");
        std::string inner = deparse_in_plan(ce, child);
        if (ce->s("node") == "Var")
            return inner;
        return "(" + inner + ")";
    }
    if (node->s("node") == "Aggref")
    {
        std::string s = node->s("aggschema").empty() ? "" : node->s("aggschema") + ".";
        s += node->s("aggname") + "(";
        if (node->flag("aggstar"))
            s += "*";
        const Json *args = node->get("args");
        for (size_t i = 0; args && i < args->arr.size(); i++)
        {
            JsonPtr a = args->arr[i];
            if (a->s("node") == "TargetEntry")
                a = a->getp("expr");
            s += (i ? ", " : "") + deparse_resolved(a, child);
        }
        return s + ")";
    }
    /* generic expression above a sub-plan: rebuild text with placeholder
     * column names */
    std::vector<std::string> names;
    std::vector<JsonPtr> refs;
    /* collect SubplanRefs in order and replace them by pseudo Vars */
    struct Rewriter {
        std::vector<std::string> &names;
        const JsonPtr &child;
        JsonPtr rewrite(const JsonPtr &n)
        {
            if (!n || (n->kind != Json::Object && n->kind != Json::Array))
                return n;
            if (n->kind == Json::Array)
            {
                JsonPtr a = Json::array();
                for (auto &x : n->arr) a->push(rewrite(x));
                return a;
            }
            if (n->s("node") == "SubplanRef")
            {
                names.push_back(deparse_resolved(n, child));
                JsonPtr v = Json::object();
                v->set("node", "Var");
                v->set("varattno", (int)names.size());
                v->set("vartype", expr_type(n->getp("expr")));
                return v;
            }
            JsonPtr copy = Json::object();
            for (auto &kv : n->obj)
                copy->set(kv.first, rewrite(kv.second));
            return copy;
        }
    } rw{names, child};
    JsonPtr rewritten = rw.rewrite(node);
    return deparse_expression(rewritten, names, false);
}

static void
explain_node(const JsonPtr &plan, int depth, bool verbose, std::vector<std::string> &out)
{
    std::string node = plan->s("node");
    std::string label;
    std::string indent_label, indent_prop;

    if (depth == 0)
    {
        indent_label = "";
        indent_prop = "  ";
    }
    else
    {
        indent_label = std::string(2 + 6 * (depth - 1), ' ') + "->  ";
        indent_prop = std::string(8 + 6 * (depth - 1), ' ');
    }
    if (node == "Sort")
        label = "Sort";
    else if (node == "Agg")
    {
        std::string st = plan->s("aggstrategy", "plain");
        label = (st == "hashed" ? "HashAggregate" : st == "sorted" ? "GroupAggregate" : "Aggregate");
    }
    else if (node == "SeqScan")
        label = "Seq Scan on " + (verbose ? plan->s("schema", "public") + "." : std::string()) + plan->s("relname");
    else if (node == "CustomPlan")
    {
        label = "Custom (" + plan->s("custom_name") + ")";
        if (plan->s("custom_name") == "GpuScan")
            label += " on " + (verbose ? plan->s("schema", "public") + "." : std::string()) + plan->s("relname");
    }
    else
        label = node;
    out.push_back(indent_label + label);

    if (verbose)
    {
        std::string s = "Output: ";
        std::vector<JsonPtr> tl = json_list(plan->get("targetlist"));
        for (size_t i = 0; i < tl.size(); i++)
        {
            JsonPtr e = tl[i]->getp("expr");
            std::string t;
            if (is_scan_node(plan))
                t = deparse_expression(e, scan_colnames(plan), false);
            else
            {
                /* upper nodes refer to their child's outputs by OUTER Vars
                 * (set_plan_references): print the child's expression */
                t = deparse_in_plan(e, plan);
            }
            s += (i ? ", " : "") + t;
        }
        out.push_back(indent_prop + s);
    }
    if (node == "Sort")
    {
        std::string s = "Sort Key: ";
        std::vector<JsonPtr> keys = json_list(plan->get("sortkeys"));
        for (size_t i = 0; i < keys.size(); i++)
            s += (i ? ", " : "") + keys[i]->str;
        out.push_back(indent_prop + s);
    }
    if (node == "Agg" && plan->has("groupkeys") && !plan->get("groupkeys")->arr.empty())
    {
        std::string s = "Group Key: ";
        std::vector<JsonPtr> keys = json_list(plan->get("groupkeys"));
        for (size_t i = 0; i < keys.size(); i++)
            s += (i ? ", " : "") + keys[i]->str;
        out.push_back(indent_prop + s);
    }
    if (node == "CustomPlan" && plan->s("custom_name") == "GpuPreAgg")
    {
        /* gpupreagg_explain (gpupreagg.c:2859-2877) */
        out.push_back(indent_prop + "Bulkload: " + (plan->flag("outer_bulkload") ? "On" : "Off"));
        std::vector<JsonPtr> oq = json_list(plan->get("outer_quals"));
        if (!oq.empty())
        {
            JsonPtr child = child_of(plan);
            std::string s = "Device Filter: " +
                deparse_expression(make_andclause(oq), scan_colnames(child), oq.size() > 1 ? false : true);
            out.push_back(indent_prop + s);
        }
    }
    if ((node == "SeqScan" || (node == "CustomPlan" && plan->s("custom_name") == "GpuScan")))
    {
        std::vector<JsonPtr> q = json_list(plan->get("qual"));
        if (!q.empty())
            out.push_back(indent_prop + "Filter: " +
                          deparse_expression(make_andclause(q), scan_colnames(plan), q.size() > 1 ? false : true));
    }
    JsonPtr child = child_of(plan);
    if (child)
        explain_node(child, depth + 1, verbose, out);
}

std::vector<std::string>
explain_plan(const JsonPtr &plan_tree, bool verbose)
{
    std::vector<std::string> out;
    explain_node(plan_tree, 0, verbose, out);
    return out;
}

}   /* namespace pgs */
