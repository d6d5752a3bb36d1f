/*
 * gpupreagg_glue.c - the PostgreSQL side of the GpuPreAgg path over the C ABI
 * of libpgstrom_cuda.so (include/pgstrom_cuda.h).
 *
 * What a maintainer of andyrbm/pg_strom drops into the extension in place of
 *   main.c:104-281          (_PG_init, GUC registration)
 *   grafter.c:24-157        (planner hook, walk over the finished plan tree)
 *   gpupreagg.c:1987-2187   (pgstrom_try_insert_gpupreagg - now in the library;
 *                            here: plan tree -> JSON, rewritten JSON -> nodes)
 *   gpupreagg.c:2189-2979   (CustomPlanMethods of "GpuPreAgg")
 * against the 9.5devel tree with the CustomPlan interface the reference is
 * written for.  No PostgreSQL tree exists in the build image, so the file is
 * compiled and driven by tests/test_pg_glue_plan.py against a stand-in for the
 * handful of headers it includes (tests/native/pg_stub/); nothing in it is
 * specific to the stand-in except glue_type_name(), which reads the syscache
 * in a real build.
 *
 * The chunk producer below is the reference's non-bulk-load path
 * (gpupreagg_load_next_chunk, gpupreagg.c:2310-2420): the outer plan is pulled
 * tuple by tuple and its virtual tuples are packed into KDS_FORMAT_COLUMN
 * chunks (pgstrom_kds_column_build) of pg_strom.chunk_size bytes.  A GpuScan
 * child that hands over whole pgstrom_data_store blocks (KDS_FORMAT_ROW) plugs
 * into the same pgs_bulk_exec_fn callback.
 */
#include "postgres.h"
#include "fmgr.h"
#include "miscadmin.h"
#include "access/htup_details.h"
#include "catalog/pg_type.h"
#include "commands/explain.h"
#include "executor/executor.h"
#include "lib/stringinfo.h"
#include "nodes/makefuncs.h"
#include "nodes/plannodes.h"
#include "nodes/primnodes.h"
#include "optimizer/cost.h"
#include "optimizer/planner.h"
#include "parser/parse_func.h"
#include "utils/guc.h"
#include "utils/lsyscache.h"
#include "utils/resowner.h"
#include <ctype.h>
#include <math.h>

#include "pgstrom_cuda.h"

PG_MODULE_MAGIC;

void        _PG_init(void);
char       *pgstrom_plan_to_json(PlannedStmt *pstmt, Plan *plan);

/* ------------------------------------------------------------------
 * GUCs: the table stays the reference's (main.c:104-234,
 * gpupreagg.c:2946-2967); every assignment is forwarded to the library,
 * which owns the values the planner and executor halves read.
 * ------------------------------------------------------------------ */
static bool     guc_enabled;
static bool     guc_perfmon;
static bool     guc_enable_gpupreagg;
static bool     guc_debug_force_gpupreagg;
static bool     guc_devprog_optimization;
static int      guc_chunk_size;
static int      guc_max_async_chunks;
static int      guc_key_heap_size;
static double   guc_gpu_setup_cost;
static double   guc_gpu_operator_cost;
static double   guc_gpu_tuple_cost;
static bool     guc_show_device_kernel;

#define GLUE_ASSIGN_BOOL(fn, name) \
    static void fn(bool newval, void *extra) { (void) extra; pgstrom_guc_set(name, newval ? "on" : "off"); }
#define GLUE_ASSIGN_INT(fn, name) \
    static void fn(int newval, void *extra) \
    { char b[32]; (void) extra; snprintf(b, sizeof(b), "%d", newval); pgstrom_guc_set(name, b); }
#define GLUE_ASSIGN_REAL(fn, name) \
    static void fn(double newval, void *extra) \
    { char b[64]; (void) extra; snprintf(b, sizeof(b), "%.17g", newval); pgstrom_guc_set(name, b); }
GLUE_ASSIGN_BOOL(assign_enabled, "pg_strom.enabled")
GLUE_ASSIGN_BOOL(assign_perfmon, "pg_strom.perfmon")
GLUE_ASSIGN_BOOL(assign_enable_gpupreagg, "enable_gpupreagg")
GLUE_ASSIGN_BOOL(assign_debug_force_gpupreagg, "pg_strom.debug_force_gpupreagg")
GLUE_ASSIGN_BOOL(assign_devprog_optimization, "pg_strom.devprog_enable_optimization")
GLUE_ASSIGN_INT(assign_chunk_size, "pg_strom.chunk_size")
GLUE_ASSIGN_INT(assign_max_async_chunks, "pg_strom.max_async_chunks")
GLUE_ASSIGN_INT(assign_key_heap_size, "pg_strom.key_heap_size")
GLUE_ASSIGN_REAL(assign_gpu_setup_cost, "gpu_setup_cost")
GLUE_ASSIGN_REAL(assign_gpu_operator_cost, "gpu_operator_cost")
GLUE_ASSIGN_REAL(assign_gpu_tuple_cost, "gpu_tuple_cost")
GLUE_ASSIGN_BOOL(assign_show_device_kernel, "pg_strom.show_device_kernel")

static bool
boot_bool(const char *name)
{
    const char *v = pgstrom_guc_get(name);
    return v && (strcmp(v, "on") == 0 || strcmp(v, "true") == 0 || strcmp(v, "1") == 0);
}
static double
boot_num(const char *name, double fallback)
{
    const char *v = pgstrom_guc_get(name);
    return v ? atof(v) : fallback;
}

static void
pgstrom_init_gucs(void)
{
    DefineCustomBoolVariable("pg_strom.enabled", "Enables the planner's use of PG-Strom", NULL,
                             &guc_enabled, boot_bool("pg_strom.enabled"), PGC_USERSET,
                             GUC_NOT_IN_SAMPLE, NULL, assign_enabled, NULL);
    DefineCustomBoolVariable("pg_strom.perfmon", "Enables the performance monitor of PG-Strom", NULL,
                             &guc_perfmon, boot_bool("pg_strom.perfmon"), PGC_USERSET,
                             GUC_NOT_IN_SAMPLE, NULL, assign_perfmon, NULL);
    DefineCustomBoolVariable("enable_gpupreagg", "Enables the use of GPU preprocessed aggregate", NULL,
                             &guc_enable_gpupreagg, boot_bool("enable_gpupreagg"), PGC_USERSET,
                             GUC_NOT_IN_SAMPLE, NULL, assign_enable_gpupreagg, NULL);
    DefineCustomBoolVariable("pg_strom.debug_force_gpupreagg",
                             "Force GpuPreAgg regardless of the cost (debug)", NULL,
                             &guc_debug_force_gpupreagg, boot_bool("pg_strom.debug_force_gpupreagg"),
                             PGC_USERSET, GUC_NOT_IN_SAMPLE, NULL, assign_debug_force_gpupreagg, NULL);
    DefineCustomBoolVariable("pg_strom.devprog_enable_optimization",
                             "Enables optimization on device program build", NULL,
                             &guc_devprog_optimization, boot_bool("pg_strom.devprog_enable_optimization"),
                             PGC_USERSET, GUC_NOT_IN_SAMPLE, NULL, assign_devprog_optimization, NULL);
    DefineCustomIntVariable("pg_strom.chunk_size", "default size of pgstrom_data_store in MB", NULL,
                            &guc_chunk_size, (int) boot_num("pg_strom.chunk_size", 15), 4, 128,
                            PGC_USERSET, GUC_NOT_IN_SAMPLE, NULL, assign_chunk_size, NULL);
    DefineCustomIntVariable("pg_strom.max_async_chunks", "max number of chunks in flight", NULL,
                            &guc_max_async_chunks, (int) boot_num("pg_strom.max_async_chunks", 3), 1, 1024,
                            PGC_USERSET, GUC_NOT_IN_SAMPLE, NULL, assign_max_async_chunks, NULL);
    /* (not in the reference: its varlena grouping keys stay in the chunk's toast area) */
    DefineCustomIntVariable("pg_strom.key_heap_size",
                            "size of the device heap of long text grouping keys in MB", NULL,
                            &guc_key_heap_size, (int) boot_num("pg_strom.key_heap_size", 64), 0, 65536,
                            PGC_USERSET, GUC_NOT_IN_SAMPLE, NULL, assign_key_heap_size, NULL);
    /* main.c:158-186 */
    DefineCustomRealVariable("gpu_setup_cost", "Cost to setup GPU device to run", NULL,
                             &guc_gpu_setup_cost, boot_num("gpu_setup_cost", 500.0), 0, 1e30,
                             PGC_USERSET, GUC_NOT_IN_SAMPLE, NULL, assign_gpu_setup_cost, NULL);
    DefineCustomRealVariable("gpu_operator_cost", "Cost of processing each operators by GPU", NULL,
                             &guc_gpu_operator_cost, boot_num("gpu_operator_cost", 0.0025 / 100.0),
                             0, 1e30, PGC_USERSET, GUC_NOT_IN_SAMPLE, NULL, assign_gpu_operator_cost, NULL);
    DefineCustomRealVariable("gpu_tuple_cost", "Cost of processing each tuple for GPU", NULL,
                             &guc_gpu_tuple_cost, boot_num("gpu_tuple_cost", 0.01 / 32.0),
                             0, 1e30, PGC_USERSET, GUC_NOT_IN_SAMPLE, NULL, assign_gpu_tuple_cost, NULL);
    DefineCustomBoolVariable("pg_strom.show_device_kernel", "Enables to show device kernel on EXPLAIN", NULL,
                             &guc_show_device_kernel, boot_bool("pg_strom.show_device_kernel"), PGC_USERSET,
                             GUC_NOT_IN_SAMPLE, NULL, assign_show_device_kernel, NULL);
}

/* ------------------------------------------------------------------
 * names the JSON speaks: pg_type.typname, pg_proc.proname, pg_operator.oprname
 * ------------------------------------------------------------------ */
static const char *
glue_type_name(Oid typid)
{
#ifdef PG_NODES_STUB_H
    return pg_stub_type_name(typid);
#else
    HeapTuple   tp = SearchSysCache1(TYPEOID, ObjectIdGetDatum(typid));
    char       *name;

    if (!HeapTupleIsValid(tp))
        elog(ERROR, "cache lookup failed for type %u", typid);
    name = pstrdup(NameStr(((Form_pg_type) GETSTRUCT(tp))->typname));
    ReleaseSysCache(tp);
    return name;
#endif
}

static void
json_string(StringInfo str, const char *s)
{
    if (!s)
    {
        appendStringInfoString(str, "null");
        return;
    }
    appendStringInfoChar(str, '"');
    for (; *s; s++)
    {
        if (*s == '"' || *s == '\\')
        {
            appendStringInfoChar(str, '\\');
            appendStringInfoChar(str, *s);
        }
        else if ((unsigned char) *s < 0x20)
            appendStringInfo(str, "\\u%04x", (unsigned char) *s);
        else
            appendStringInfoChar(str, *s);
    }
    appendStringInfoChar(str, '"');
}

/* ------------------------------------------------------------------
 * plan tree -> JSON (the format of INTEGRATION.md section 3; the test
 * harness builds the same with pg_strom_b200/pgplan.py)
 * ------------------------------------------------------------------ */
static void expr_to_json(StringInfo str, Node *node);

static void
expr_list_to_json(StringInfo str, List *exprs)
{
    ListCell   *lc;
    bool        first = true;

    appendStringInfoChar(str, '[');
    foreach (lc, exprs)
    {
        if (!first)
            appendStringInfoChar(str, ',');
        first = false;
        expr_to_json(str, (Node *) lfirst(lc));
    }
    appendStringInfoChar(str, ']');
}

static const char *
collation_name(Oid collid)
{
    /* ordering operators on text are offloaded under C / POSIX only: the
     * default collation is resolved here, where the database is known */
    if (collid == DEFAULT_COLLATION_OID)
        return lc_collate_is_c(collid) ? "C" : "default";
    return get_collation_name(collid);
}

static void
expr_to_json(StringInfo str, Node *node)
{
    if (node == NULL)
    {
        appendStringInfoString(str, "null");
        return;
    }
    switch (nodeTag(node))
    {
        case T_Var:
        {
            Var *v = (Var *) node;

            appendStringInfo(str, "{\"node\":\"Var\",\"varattno\":%d,\"vartype\":", (int) v->varattno);
            json_string(str, glue_type_name(v->vartype));
            if (v->varno == OUTER_VAR)
                appendStringInfoString(str, ",\"varno\":\"OUTER\"");
            if (v->vartypmod >= 0)
                appendStringInfo(str, ",\"vartypmod\":%d", (int) v->vartypmod);
            appendStringInfoChar(str, '}');
            break;
        }
        case T_Const:
        {
            Const *c = (Const *) node;

            appendStringInfoString(str, "{\"node\":\"Const\",\"consttype\":");
            json_string(str, glue_type_name(c->consttype));
            if (c->constisnull)
                appendStringInfoString(str, ",\"constisnull\":true}");
            else
            {
                Oid     typoutput;
                bool    typIsVarlena;

                getTypeOutputInfo(c->consttype, &typoutput, &typIsVarlena);
                appendStringInfoString(str, ",\"constisnull\":false,\"constvalue\":");
                json_string(str, OidOutputFunctionCall(typoutput, c->constvalue));
                appendStringInfoChar(str, '}');
            }
            break;
        }
        case T_Param:
        {
            Param *p = (Param *) node;

            appendStringInfo(str, "{\"node\":\"Param\",\"paramkind\":%d,\"paramid\":%d,\"paramtype\":",
                             p->paramkind, p->paramid);
            json_string(str, glue_type_name(p->paramtype));
            appendStringInfoChar(str, '}');
            break;
        }
        case T_FuncExpr:
        {
            FuncExpr *f = (FuncExpr *) node;

            appendStringInfoString(str, "{\"node\":\"FuncExpr\",\"funcname\":");
            json_string(str, get_func_name(f->funcid));
            appendStringInfoString(str, ",\"funcresulttype\":");
            json_string(str, glue_type_name(f->funcresulttype));
            appendStringInfoString(str, ",\"funcformat\":");
            json_string(str, f->funcformat == COERCE_EXPLICIT_CALL ? "call" :
                        f->funcformat == COERCE_EXPLICIT_CAST ? "cast" : "implicit");
            if (OidIsValid(f->inputcollid))
            {
                appendStringInfoString(str, ",\"inputcollid\":");
                json_string(str, collation_name(f->inputcollid));
            }
            appendStringInfoString(str, ",\"args\":");
            expr_list_to_json(str, f->args);
            appendStringInfoChar(str, '}');
            break;
        }
        case T_OpExpr:
        case T_DistinctExpr:
        {
            OpExpr *op = (OpExpr *) node;

            appendStringInfo(str, "{\"node\":\"%s\",\"opname\":",
                             nodeTag(node) == T_OpExpr ? "OpExpr" : "DistinctExpr");
            json_string(str, get_opname(op->opno));
            appendStringInfoString(str, ",\"opfuncname\":");
            json_string(str, get_func_name(get_opcode(op->opno)));
            appendStringInfoString(str, ",\"opresulttype\":");
            json_string(str, glue_type_name(op->opresulttype));
            if (OidIsValid(op->inputcollid))
            {
                appendStringInfoString(str, ",\"inputcollid\":");
                json_string(str, collation_name(op->inputcollid));
            }
            appendStringInfoString(str, ",\"args\":");
            expr_list_to_json(str, op->args);
            appendStringInfoChar(str, '}');
            break;
        }
        case T_BoolExpr:
        {
            BoolExpr *b = (BoolExpr *) node;

            appendStringInfo(str, "{\"node\":\"BoolExpr\",\"boolop\":\"%s\",\"args\":",
                             b->boolop == AND_EXPR ? "AND" : b->boolop == OR_EXPR ? "OR" : "NOT");
            expr_list_to_json(str, b->args);
            appendStringInfoChar(str, '}');
            break;
        }
        case T_NullTest:
        {
            NullTest *nt = (NullTest *) node;

            appendStringInfoString(str, "{\"node\":\"NullTest\",\"arg\":");
            expr_to_json(str, (Node *) nt->arg);
            appendStringInfo(str, ",\"nulltesttype\":\"%s\",\"argisrow\":%s}",
                             nt->nulltesttype == IS_NULL ? "IS_NULL" : "IS_NOT_NULL",
                             nt->argisrow ? "true" : "false");
            break;
        }
        case T_BooleanTest:
        {
            static const char *names[] = { "IS_TRUE", "IS_NOT_TRUE", "IS_FALSE", "IS_NOT_FALSE",
                                           "IS_UNKNOWN", "IS_NOT_UNKNOWN" };
            BooleanTest *bt = (BooleanTest *) node;

            appendStringInfoString(str, "{\"node\":\"BooleanTest\",\"arg\":");
            expr_to_json(str, (Node *) bt->arg);
            appendStringInfo(str, ",\"booltesttype\":\"%s\"}", names[bt->booltesttype]);
            break;
        }
        case T_RelabelType:
        {
            RelabelType *r = (RelabelType *) node;

            appendStringInfoString(str, "{\"node\":\"RelabelType\",\"arg\":");
            expr_to_json(str, (Node *) r->arg);
            appendStringInfoString(str, ",\"resulttype\":");
            json_string(str, glue_type_name(r->resulttype));
            appendStringInfoChar(str, '}');
            break;
        }
        case T_CaseExpr:
        {
            CaseExpr *c = (CaseExpr *) node;

            appendStringInfoString(str, "{\"node\":\"CaseExpr\",\"casetype\":");
            json_string(str, glue_type_name(c->casetype));
            appendStringInfoString(str, ",\"arg\":");
            expr_to_json(str, (Node *) c->arg);
            appendStringInfoString(str, ",\"args\":");
            expr_list_to_json(str, c->args);
            appendStringInfoString(str, ",\"defresult\":");
            expr_to_json(str, (Node *) c->defresult);
            appendStringInfoChar(str, '}');
            break;
        }
        case T_CaseWhen:
        {
            CaseWhen *w = (CaseWhen *) node;

            appendStringInfoString(str, "{\"node\":\"CaseWhen\",\"expr\":");
            expr_to_json(str, (Node *) w->expr);
            appendStringInfoString(str, ",\"result\":");
            expr_to_json(str, (Node *) w->result);
            appendStringInfoChar(str, '}');
            break;
        }
        case T_Aggref:
        {
            Aggref     *a = (Aggref *) node;
            ListCell   *lc;
            bool        first = true;

            appendStringInfoString(str, "{\"node\":\"Aggref\",\"aggname\":");
            json_string(str, get_func_name(a->aggfnoid));
            appendStringInfoString(str, ",\"aggargtypes\":[");
            foreach (lc, a->args)
            {
                TargetEntry *tle = (TargetEntry *) lfirst(lc);
                Oid          t = InvalidOid;

                /* exprType(): the handful of node types an aggregate argument is */
                switch (nodeTag(tle->expr))
                {
                    case T_Var: t = ((Var *) tle->expr)->vartype; break;
                    case T_Const: t = ((Const *) tle->expr)->consttype; break;
                    case T_FuncExpr: t = ((FuncExpr *) tle->expr)->funcresulttype; break;
                    case T_OpExpr: t = ((OpExpr *) tle->expr)->opresulttype; break;
                    case T_RelabelType: t = ((RelabelType *) tle->expr)->resulttype; break;
                    case T_CaseExpr: t = ((CaseExpr *) tle->expr)->casetype; break;
                    default: t = BOOLOID; break;
                }
                if (!first)
                    appendStringInfoChar(str, ',');
                first = false;
                json_string(str, glue_type_name(t));
            }
            appendStringInfoString(str, "],\"aggtype\":");
            json_string(str, glue_type_name(a->aggtype));
            appendStringInfoString(str, ",\"args\":");
            expr_list_to_json(str, a->args);
            appendStringInfoString(str, ",\"aggfilter\":");
            expr_to_json(str, (Node *) a->aggfilter);
            appendStringInfo(str, ",\"aggstar\":%s", a->aggstar ? "true" : "false");
            /* DISTINCT / ORDER BY inside the aggregate: never rewritten
             * (gpupreagg.c:2047-2060); say so and the planner half declines */
            if (a->aggdistinct != NIL || a->aggorder != NIL)
                appendStringInfoString(str, ",\"aggdistinct\":true");
            appendStringInfoChar(str, '}');
            break;
        }
        case T_TargetEntry:
        {
            TargetEntry *tle = (TargetEntry *) node;

            appendStringInfoString(str, "{\"node\":\"TargetEntry\",\"expr\":");
            expr_to_json(str, (Node *) tle->expr);
            appendStringInfo(str, ",\"resno\":%d", (int) tle->resno);
            if (tle->resname)
            {
                appendStringInfoString(str, ",\"resname\":");
                json_string(str, tle->resname);
                appendStringInfo(str, ",\"resjunk\":%s", tle->resjunk ? "true" : "false");
            }
            appendStringInfoChar(str, '}');
            break;
        }
        default:
            /* something the device code generator has no word for: the
             * planner half sees an unknown node and leaves the Agg alone */
            appendStringInfo(str, "{\"node\":\"Unsupported\",\"tag\":%d}", (int) nodeTag(node));
            break;
    }
}

static void
plan_to_json(StringInfo str, PlannedStmt *pstmt, Plan *plan)
{
    if (plan == NULL)
    {
        appendStringInfoString(str, "null");
        return;
    }
    appendStringInfoString(str, "{\"node\":");
    switch (nodeTag(plan))
    {
        case T_SeqScan:
        {
            RangeTblEntry *rte = rt_fetch(((Scan *) plan)->scanrelid, pstmt->rtable);

            appendStringInfoString(str, "\"SeqScan\",\"relname\":");
            json_string(str, get_rel_name(rte->relid));
            appendStringInfoString(str, ",\"schema\":");
            json_string(str, get_namespace_name(get_rel_namespace(rte->relid)));
            appendStringInfoString(str, ",\"alias\":");
            json_string(str, rte->eref ? rte->eref->aliasname : get_rel_name(rte->relid));
            break;
        }
        case T_Agg:
        {
            Agg *agg = (Agg *) plan;

            appendStringInfo(str, "\"Agg\",\"aggstrategy\":\"%s\",\"grpColIdx\":[",
                             agg->aggstrategy == AGG_PLAIN ? "plain" :
                             agg->aggstrategy == AGG_SORTED ? "sorted" : "hashed");
            for (int i = 0; i < agg->numCols; i++)
                appendStringInfo(str, "%s%d", i ? "," : "", (int) agg->grpColIdx[i]);
            appendStringInfo(str, "],\"numGroups\":%.1f", (double) agg->numGroups);
            break;
        }
        case T_Sort:
        {
            Sort *sort = (Sort *) plan;

            appendStringInfoString(str, "\"Sort\",\"sortColIdx\":[");
            for (int i = 0; i < sort->numCols; i++)
                appendStringInfo(str, "%s%d", i ? "," : "", (int) sort->sortColIdx[i]);
            appendStringInfoChar(str, ']');
            break;
        }
        case T_Result:
            appendStringInfoString(str, "\"Result\"");
            break;
        case T_HashJoin:
            appendStringInfoString(str, "\"HashJoin\"");
            break;
        default:
            appendStringInfo(str, "\"Plan%d\"", (int) nodeTag(plan));
            break;
    }
    /* PostgreSQL's estimates: what cost_gpupreagg() works from */
    appendStringInfo(str, ",\"startup_cost\":%.2f,\"total_cost\":%.2f,\"plan_rows\":%.0f,\"plan_width\":%d",
                     plan->startup_cost, plan->total_cost, plan->plan_rows, plan->plan_width);
    appendStringInfoString(str, ",\"targetlist\":");
    expr_list_to_json(str, plan->targetlist);
    appendStringInfoString(str, ",\"qual\":");
    expr_list_to_json(str, plan->qual);
    if (plan->lefttree)
    {
        appendStringInfoString(str, ",\"lefttree\":");
        plan_to_json(str, pstmt, plan->lefttree);
    }
    if (plan->righttree)
    {
        appendStringInfoString(str, ",\"righttree\":");
        plan_to_json(str, pstmt, plan->righttree);
    }
    appendStringInfoChar(str, '}');
}

char *
pgstrom_plan_to_json(PlannedStmt *pstmt, Plan *plan)
{
    StringInfoData str;

    initStringInfo(&str);
    plan_to_json(&str, pstmt, plan);
    return str.data;
}

/* ------------------------------------------------------------------
 * a JSON reader just large enough for the trees the library returns
 * ------------------------------------------------------------------ */
typedef enum { J_NULL, J_BOOL, J_NUM, J_STR, J_ARR, J_OBJ } JKind;
typedef struct JNode
{
    JKind           kind;
    double          num;
    char           *str;
    int             n;          /* elements / members */
    struct JNode  **items;
    char          **keys;       /* J_OBJ */
} JNode;

static void
jskip(const char **p)
{
    while (**p && isspace((unsigned char) **p))
        (*p)++;
}

static char *
jparse_string(const char **p)
{
    StringInfoData s;

    initStringInfo(&s);
    (*p)++;                     /* opening quote */
    while (**p && **p != '"')
    {
        if (**p == '\\' && (*p)[1])
        {
            (*p)++;
            switch (**p)
            {
                case 'n': appendStringInfoChar(&s, '\n'); break;
                case 't': appendStringInfoChar(&s, '\t'); break;
                case 'r': appendStringInfoChar(&s, '\r'); break;
                case 'b': appendStringInfoChar(&s, '\b'); break;
                case 'f': appendStringInfoChar(&s, '\f'); break;
                case 'u':
                {
                    unsigned cp = 0;

                    for (int i = 1; i <= 4 && (*p)[i]; i++)
                        cp = cp * 16 + (unsigned) (isdigit((unsigned char) (*p)[i]) ? (*p)[i] - '0'
                                                   : (tolower((unsigned char) (*p)[i]) - 'a' + 10));
                    appendStringInfoChar(&s, (char) cp);    /* control characters only */
                    (*p) += 4;
                    break;
                }
                default: appendStringInfoChar(&s, **p); break;
            }
            (*p)++;
        }
        else
            appendStringInfoChar(&s, *(*p)++);
    }
    if (**p == '"')
        (*p)++;
    return s.data;
}

static JNode *
jparse(const char **p)
{
    JNode *j = (JNode *) palloc0(sizeof(JNode));

    jskip(p);
    if (**p == '{' || **p == '[')
    {
        bool    obj = (**p == '{');
        char    close = obj ? '}' : ']';
        int     cap = 8;

        j->kind = obj ? J_OBJ : J_ARR;
        j->items = (JNode **) palloc(sizeof(JNode *) * cap);
        if (obj)
            j->keys = (char **) palloc(sizeof(char *) * cap);
        (*p)++;
        jskip(p);
        while (**p && **p != close)
        {
            if (j->n == cap)
            {
                JNode **ni = (JNode **) palloc(sizeof(JNode *) * cap * 2);
                memcpy(ni, j->items, sizeof(JNode *) * cap);
                j->items = ni;
                if (obj)
                {
                    char **nk = (char **) palloc(sizeof(char *) * cap * 2);
                    memcpy(nk, j->keys, sizeof(char *) * cap);
                    j->keys = nk;
                }
                cap *= 2;
            }
            if (obj)
            {
                jskip(p);
                j->keys[j->n] = jparse_string(p);
                jskip(p);
                if (**p == ':')
                    (*p)++;
            }
            j->items[j->n++] = jparse(p);
            jskip(p);
            if (**p == ',')
                (*p)++;
            jskip(p);
        }
        if (**p == close)
            (*p)++;
    }
    else if (**p == '"')
    {
        j->kind = J_STR;
        j->str = jparse_string(p);
    }
    else if (strncmp(*p, "true", 4) == 0)  { j->kind = J_BOOL; j->num = 1; *p += 4; }
    else if (strncmp(*p, "false", 5) == 0) { j->kind = J_BOOL; j->num = 0; *p += 5; }
    else if (strncmp(*p, "null", 4) == 0)  { j->kind = J_NULL; *p += 4; }
    else
    {
        char *end;

        j->kind = J_NUM;
        j->num = strtod(*p, &end);
        *p = (end == *p ? *p + 1 : end);
    }
    return j;
}

static JNode *
jget(JNode *obj, const char *key)
{
    if (obj && obj->kind == J_OBJ)
        for (int i = 0; i < obj->n; i++)
            if (strcmp(obj->keys[i], key) == 0)
                return obj->items[i];
    return NULL;
}
static const char *
jstr(JNode *obj, const char *key)
{
    JNode *v = jget(obj, key);
    return (v && v->kind == J_STR) ? v->str : NULL;
}
static bool
jis(JNode *obj, const char *key, const char *value)
{
    const char *s = jstr(obj, key);
    return s && strcmp(s, value) == 0;
}
static int
jint(JNode *obj, const char *key, int fallback)
{
    JNode *v = jget(obj, key);
    return (v && (v->kind == J_NUM || v->kind == J_BOOL)) ? (int) v->num : fallback;
}

/* ------------------------------------------------------------------
 * rewritten JSON -> expression nodes.  Functions, operators and aggregates
 * are found by name and argument types, the way the reference looks up its
 * pgstrom.* placeholders (gpupreagg.c:729-980 make_gpupreagg_refnode).
 * ------------------------------------------------------------------ */
static Expr *expr_from_json(JNode *j);

static Oid
type_from_json(JNode *obj, const char *key)
{
    const char *name = jstr(obj, key);
    Oid         typid = name ? TypenameGetTypid(name) : InvalidOid;

    if (!OidIsValid(typid))
        elog(ERROR, "PG-Strom: type \"%s\" not found", name ? name : "(null)");
    return typid;
}

static List *
expr_list_from_json(JNode *arr)
{
    List *result = NIL;

    if (arr && arr->kind == J_ARR)
        for (int i = 0; i < arr->n; i++)
            result = lappend(result, expr_from_json(arr->items[i]));
    return result;
}

static Oid
expr_type(Expr *e)
{
    switch (nodeTag(e))
    {
        case T_Var: return ((Var *) e)->vartype;
        case T_Const: return ((Const *) e)->consttype;
        case T_Param: return ((Param *) e)->paramtype;
        case T_FuncExpr: return ((FuncExpr *) e)->funcresulttype;
        case T_OpExpr:
        case T_DistinctExpr: return ((OpExpr *) e)->opresulttype;
        case T_RelabelType: return ((RelabelType *) e)->resulttype;
        case T_CaseExpr: return ((CaseExpr *) e)->casetype;
        case T_Aggref: return ((Aggref *) e)->aggtype;
        case T_TargetEntry: return expr_type(((TargetEntry *) e)->expr);
        default: return BOOLOID;
    }
}

static Oid
lookup_function(const char *schema, const char *name, List *args)
{
    Oid         argtypes[FUNC_MAX_ARGS];
    int         nargs = 0;
    ListCell   *lc;
    List       *names;
    Oid         funcid;

    foreach (lc, args)
        argtypes[nargs++] = expr_type((Expr *) lfirst(lc));
    names = (schema && *schema) ? list_make2(makeString(pstrdup(schema)), makeString(pstrdup(name)))
                                : list_make1(makeString(pstrdup(name)));
    funcid = LookupFuncName(names, nargs, argtypes, true);
    if (!OidIsValid(funcid))
        elog(ERROR, "PG-Strom: function %s%s%s with %d arguments not found (is the pg_strom extension created?)",
             schema && *schema ? schema : "", schema && *schema ? "." : "", name, nargs);
    return funcid;
}

static Expr *
expr_from_json(JNode *j)
{
    const char *tag;

    if (j == NULL || j->kind != J_OBJ)
        return NULL;
    tag = jstr(j, "node");
    if (tag == NULL)
        return NULL;
    if (strcmp(tag, "TargetEntry") == 0)
    {
        const char *resname = jstr(j, "resname");

        return (Expr *) makeTargetEntry(expr_from_json(jget(j, "expr")), (AttrNumber) jint(j, "resno", 0),
                                        resname ? pstrdup(resname) : NULL, jint(j, "resjunk", 0) != 0);
    }
    if (strcmp(tag, "Var") == 0)
    {
        const char *varno = jstr(j, "varno");

        /* a scan-level Var keeps varno 1 until set_plan_references; Vars of
         * upper nodes point at their child's output */
        return (Expr *) makeVar(varno ? OUTER_VAR : 1, (AttrNumber) jint(j, "varattno", 0),
                                type_from_json(j, "vartype"), jint(j, "vartypmod", -1), InvalidOid, 0);
    }
    if (strcmp(tag, "Const") == 0)
    {
        Oid     typid = type_from_json(j, "consttype");
        Const  *c = makeNullConst(typid, -1, InvalidOid);

        if (!jint(j, "constisnull", 0))
        {
            Oid     typinput, typioparam;
            char    align;

            getTypeInputInfo(typid, &typinput, &typioparam);
            get_typlenbyvalalign(typid, (int16 *) &c->constlen, &c->constbyval, &align);
            c->constvalue = OidInputFunctionCall(typinput, pstrdup(jstr(j, "constvalue")), typioparam, -1);
            c->constisnull = false;
        }
        return (Expr *) c;
    }
    if (strcmp(tag, "FuncExpr") == 0)
    {
        FuncExpr   *f = makeNode(FuncExpr);
        const char *fmt = jstr(j, "funcformat");

        f->args = expr_list_from_json(jget(j, "args"));
        f->funcid = lookup_function(jstr(j, "funcschema"), jstr(j, "funcname"), f->args);
        f->funcresulttype = type_from_json(j, "funcresulttype");
        f->funcformat = (fmt && strcmp(fmt, "cast") == 0) ? COERCE_EXPLICIT_CAST :
                        (fmt && strcmp(fmt, "implicit") == 0) ? COERCE_IMPLICIT_CAST : COERCE_EXPLICIT_CALL;
        return (Expr *) f;
    }
    if (strcmp(tag, "OpExpr") == 0 || strcmp(tag, "DistinctExpr") == 0)
    {
        OpExpr *op = makeNode(OpExpr);

        if (strcmp(tag, "DistinctExpr") == 0)
            NodeSetTag(op, T_DistinctExpr);
        op->args = expr_list_from_json(jget(j, "args"));
        op->opno = OpernameGetOprid(list_make1(makeString(pstrdup(jstr(j, "opname")))),
                                    expr_type((Expr *) linitial(op->args)),
                                    expr_type((Expr *) list_nth(op->args, list_length(op->args) - 1)));
        if (!OidIsValid(op->opno))
            elog(ERROR, "PG-Strom: operator %s not found", jstr(j, "opname"));
        op->opfuncid = get_opcode(op->opno);
        op->opresulttype = type_from_json(j, "opresulttype");
        return (Expr *) op;
    }
    if (strcmp(tag, "BoolExpr") == 0)
    {
        BoolExpr *b = makeNode(BoolExpr);

        b->boolop = jis(j, "boolop", "AND") ? AND_EXPR : jis(j, "boolop", "OR") ? OR_EXPR : NOT_EXPR;
        b->args = expr_list_from_json(jget(j, "args"));
        return (Expr *) b;
    }
    if (strcmp(tag, "NullTest") == 0)
    {
        NullTest *nt = makeNode(NullTest);

        nt->arg = expr_from_json(jget(j, "arg"));
        nt->nulltesttype = jis(j, "nulltesttype", "IS_NULL") ? IS_NULL : IS_NOT_NULL;
        nt->argisrow = jint(j, "argisrow", 0) != 0;
        return (Expr *) nt;
    }
    if (strcmp(tag, "RelabelType") == 0)
    {
        RelabelType *r = makeNode(RelabelType);

        r->arg = expr_from_json(jget(j, "arg"));
        r->resulttype = type_from_json(j, "resulttype");
        r->resulttypmod = -1;
        r->relabelformat = COERCE_IMPLICIT_CAST;
        return (Expr *) r;
    }
    if (strcmp(tag, "CaseExpr") == 0)
    {
        CaseExpr *c = makeNode(CaseExpr);

        c->casetype = type_from_json(j, "casetype");
        c->arg = expr_from_json(jget(j, "arg"));
        c->args = expr_list_from_json(jget(j, "args"));
        c->defresult = expr_from_json(jget(j, "defresult"));
        return (Expr *) c;
    }
    if (strcmp(tag, "CaseWhen") == 0)
    {
        CaseWhen *w = makeNode(CaseWhen);

        w->expr = expr_from_json(jget(j, "expr"));
        w->result = expr_from_json(jget(j, "result"));
        return (Expr *) w;
    }
    if (strcmp(tag, "Aggref") == 0)
    {
        Aggref     *a = makeNode(Aggref);
        JNode      *args = jget(j, "args");
        List       *plain = NIL;

        /* the rewritten aggregates take the GpuPreAgg output columns: plain
         * expressions in the JSON, TargetEntry-wrapped in an Aggref */
        if (args && args->kind == J_ARR)
            for (int i = 0; i < args->n; i++)
            {
                Expr *e = expr_from_json(args->items[i]);

                plain = lappend(plain, e);
                if (!IsA(e, TargetEntry))
                    e = (Expr *) makeTargetEntry(e, (AttrNumber) (i + 1), NULL, false);
                a->args = lappend(a->args, e);
            }
        a->aggfnoid = lookup_function(jstr(j, "aggschema"), jstr(j, "aggname"), plain);
        a->aggtype = type_from_json(j, "aggtype");
        a->aggfilter = expr_from_json(jget(j, "aggfilter"));
        a->aggstar = jint(j, "aggstar", 0) != 0;
        a->aggkind = 'n';
        return (Expr *) a;
    }
    elog(ERROR, "PG-Strom: unexpected node \"%s\" in the rewritten plan", tag);
    return NULL;
}

/* ------------------------------------------------------------------
 * the GpuPreAgg plan node and its state (gpupreagg.c:47-132)
 * ------------------------------------------------------------------ */
static CustomPlanMethods    gpupreagg_plan_methods;

typedef struct
{
    CustomPlan  cplan;
    pgs_plan   *plan;           /* the planner half's result; lives as long as the plan */
    int         idx;            /* this node among the plan's GpuPreAgg nodes */
    List       *outer_quals;    /* evaluated inside the kernel; kept for EXPLAIN */
} GpuPreAggPlan;

typedef struct GpuPreAggState
{
    CustomPlanState         cps;
    struct GpuPreAggState  *next_live;  /* states with device work this backend holds */
    pgs_gpupreagg_state    *state;
    TupleDesc               scan_desc;
    bool                    outer_done;
    /* the chunk under construction: column arrays of the outer tuples */
    int                     ncols;
    kern_colmeta           *colmeta;
    uint32                  nrows, nrooms;
    char                  **values;     /* [ncols]: packed values, or varlena pointers */
    uint8_t               **isnull;     /* [ncols] */
    /* text / bpchar grouping keys of the result: the library returns them as
     * "kernel text" words, the slot needs varlena pointers */
    int                     nresult;
    int32                  *key_typmod;     /* [nresult]: KEY_NOT_TEXT or the atttypmod */
    char                  **key_buf;        /* [nresult]: varlena of the current row */
    size_t                 *key_buflen;
} GpuPreAggState;
#define KEY_NOT_TEXT    (-2)
#ifndef Max
#define Max(x, y)       ((x) > (y) ? (x) : (y))
#endif

/* Every state that holds a session is on this list until EndCustomPlan; a
 * transaction that aborts in between gets them closed by the resource-owner
 * callback below - in-flight chunks are waited for and handed back, nothing
 * of the device is leaked (restrack.c:180-254). */
static GpuPreAggState  *live_states = NULL;

static void
live_states_remove(GpuPreAggState *gpas)
{
    for (GpuPreAggState **p = &live_states; *p; p = &(*p)->next_live)
        if (*p == gpas)
        {
            *p = gpas->next_live;
            break;
        }
    gpas->next_live = NULL;
}

static void
pgstrom_release_callback(ResourceReleasePhase phase, bool isCommit, bool isTopLevel, void *arg)
{
    (void) isTopLevel;
    (void) arg;
    if (phase != RESOURCE_RELEASE_BEFORE_LOCKS || isCommit)
        return;
    while (live_states)
    {
        GpuPreAggState *gpas = live_states;

        live_states = gpas->next_live;
        gpas->next_live = NULL;
        if (gpas->state)
            (void) gpupreagg_end(gpas->state);      /* drains, releases, closes */
        gpas->state = NULL;
    }
}

/* rewritten JSON + original tree -> plan nodes.  The rewritten tree is the
 * original one with CustomPlan nodes spliced in above the outer plan of each
 * rewritten Agg (Agg -> [Sort ->] GpuPreAgg -> outer plan), so the two are
 * walked side by side. */
static Plan *
graft_plan(JNode *j, Plan *orig, pgs_plan *plan)
{
    if (j == NULL || j->kind != J_OBJ || orig == NULL)
        return orig;
    if (jis(j, "node", "CustomPlan") && jis(j, "custom_name", "GpuPreAgg"))
    {
        GpuPreAggPlan *gpreagg = (GpuPreAggPlan *) palloc0(sizeof(GpuPreAggPlan));

        NodeSetTag(gpreagg, T_CustomPlan);
        gpreagg->cplan.methods = &gpupreagg_plan_methods;
        gpreagg->plan = plan;
        gpreagg->idx = jint(j, "gpupreagg_index", 0);
        gpreagg->cplan.plan.targetlist = expr_list_from_json(jget(j, "targetlist"));
        gpreagg->cplan.plan.qual = NIL;
        gpreagg->outer_quals = expr_list_from_json(jget(j, "outer_quals"));
        gpreagg->cplan.plan.startup_cost = jget(j, "startup_cost") ? jget(j, "startup_cost")->num : orig->startup_cost;
        gpreagg->cplan.plan.total_cost = jget(j, "total_cost") ? jget(j, "total_cost")->num : orig->total_cost;
        gpreagg->cplan.plan.plan_rows = jget(j, "plan_rows") ? jget(j, "plan_rows")->num : orig->plan_rows;
        gpreagg->cplan.plan.plan_width = jint(j, "plan_width", orig->plan_width);
        outerPlan(gpreagg) = graft_plan(jget(j, "lefttree"), orig, plan);
        return &gpreagg->cplan.plan;
    }
    if (jis(j, "node", "CustomPlan") && jis(j, "custom_name", "GpuScan"))
    {
        /* the scan stays PostgreSQL's; the quals the device evaluates have
         * moved into the GpuPreAgg node, the rest stay here */
        orig->qual = expr_list_from_json(jget(j, "qual"));
        return orig;
    }
    if (IsA(orig, Agg) || IsA(orig, Sort))
    {
        /* target lists above a GpuPreAgg node refer to its output columns */
        JNode *child = jget(j, "lefttree");
        JNode *gpreagg_json = NULL;
        bool   sort_between = false;
        bool   spliced = false;

        for (JNode *c = child; c && c->kind == J_OBJ; c = jget(c, "lefttree"))
        {
            if (jis(c, "node", "CustomPlan") && jis(c, "custom_name", "GpuPreAgg"))
            {
                gpreagg_json = c;
                spliced = true;
            }
            if (!jis(c, "node", "Sort"))
                break;
            sort_between = true;
        }
        if (spliced && IsA(orig, Agg) && !guc_debug_force_gpupreagg &&
            jget(gpreagg_json, "total_cost") != NULL)
        {
            /* gpupreagg.c:2105-2118: the library has priced the GpuPreAgg node
             * (cost_gpupreagg, :366-464); what the Agg - and the Sort, if one
             * sits between - cost on top of it is PostgreSQL's own arithmetic.
             * The plan is only rewritten when that comes out cheaper. */
            Agg    *agg = (Agg *) orig;
            Path    dummy;
            Cost    startup = jget(gpreagg_json, "startup_cost") ? jget(gpreagg_json, "startup_cost")->num : 0.0;
            Cost    total = jget(gpreagg_json, "total_cost")->num;
            double  rows = jget(gpreagg_json, "plan_rows") ? jget(gpreagg_json, "plan_rows")->num : orig->plan_rows;
            int     width = jint(gpreagg_json, "plan_width", orig->plan_width);

            memset(&dummy, 0, sizeof(dummy));
            if (sort_between)
            {
                cost_sort(&dummy, NULL, NIL, total, rows, width, 0.0, work_mem, -1.0);
                startup = dummy.startup_cost;
                total = dummy.total_cost;
            }
            cost_agg(&dummy, NULL, agg->aggstrategy, NULL, agg->numCols, (double) agg->numGroups,
                     startup, total, rows);
            if (orig->total_cost <= dummy.total_cost)
                return orig;            /* PostgreSQL's plan stays */
        }
        if (spliced)
        {
            orig->targetlist = expr_list_from_json(jget(j, "targetlist"));
            if (IsA(orig, Agg))
                orig->qual = expr_list_from_json(jget(j, "qual"));
        }
    }
    if (orig->lefttree)
        orig->lefttree = graft_plan(jget(j, "lefttree"), orig->lefttree, plan);
    if (orig->righttree)
        orig->righttree = graft_plan(jget(j, "righttree"), orig->righttree, plan);
    return orig;
}

/* grafter.c:119-149 */
static planner_hook_type    planner_hook_next = NULL;

static PlannedStmt *
pgstrom_grafter_entrypoint(Query *parse, int cursorOptions, ParamListInfo boundParams)
{
    PlannedStmt *result;
    char        *json;
    const char  *cursor;
    pgs_plan    *plan;

    if (planner_hook_next)
        result = planner_hook_next(parse, cursorOptions, boundParams);
    else
        result = standard_planner(parse, cursorOptions, boundParams);
    if (!guc_enabled || !guc_enable_gpupreagg || result == NULL || result->planTree == NULL)
        return result;
    json = pgstrom_plan_to_json(result, result->planTree);
    plan = pgstrom_grafter_json(json);
    pfree(json);
    if (plan == NULL)
        return result;          /* not a tree the planner half understands */
    if (pgs_plan_num_gpupreagg(plan) == 0)
    {
        pgs_plan_free(plan);    /* reason: pgs_plan_reject_reason() */
        return result;
    }
    cursor = pgs_plan_tree_json(plan);
    result->planTree = graft_plan(jparse(&cursor), result->planTree, plan);
    return result;
}

/* grafter.c:131-146: the sub-plans (InitPlans / SubPlans) are walked like the
 * main tree, each with its own pgs_plan */
static Plan *
pgstrom_graft_one(PlannedStmt *pstmt, Plan *tree)
{
    char        *json = pgstrom_plan_to_json(pstmt, tree);
    pgs_plan    *plan = pgstrom_grafter_json(json);
    const char  *cursor;

    pfree(json);
    if (plan == NULL)
        return tree;
    if (pgs_plan_num_gpupreagg(plan) == 0)
    {
        pgs_plan_free(plan);
        return tree;
    }
    cursor = pgs_plan_tree_json(plan);
    return graft_plan(jparse(&cursor), tree, plan);
}

static PlannedStmt *
pgstrom_grafter_with_subplans(Query *parse, int cursorOptions, ParamListInfo boundParams)
{
    PlannedStmt *result = pgstrom_grafter_entrypoint(parse, cursorOptions, boundParams);
    ListCell    *cell;

    if (result == NULL || !guc_enabled || !guc_enable_gpupreagg)
        return result;
    foreach (cell, result->subplans)
        if (lfirst(cell) != NULL)
            lfirst(cell) = pgstrom_graft_one(result, (Plan *) lfirst(cell));
    return result;
}

/* ------------------------------------------------------------------
 * executor (CustomPlanMethods, gpupreagg.c:2189-2941)
 * ------------------------------------------------------------------ */
/* chunks live in pinned, 128-byte aligned host memory (pgs_chunk_alloc: the
 * reference's pgstrom_shmem_alloc of a pgstrom_data_store): the H2D copy is
 * asynchronous and the column arrays keep the alignment the bulk copies need */
static void
chunk_release(void *arg, const kern_data_store *kds)
{
    (void) arg;
    pgs_chunk_free((void *) kds);
}

static void
chunk_reset(GpuPreAggState *gpas)
{
    gpas->nrows = 0;
}

/* pgs_bulk_exec_fn: pulls outer tuples until a chunk is full
 * (gpupreagg_load_next_chunk, gpupreagg.c:2310-2420) */
static int
gpupreagg_next_chunk(void *child_state, pgs_bulkslot *slot)
{
    GpuPreAggState *gpas = (GpuPreAggState *) child_state;
    PlanState      *outer = outerPlanState(gpas);
    size_t          length;
    void           *kds;

    memset(slot, 0, sizeof(*slot));
    chunk_reset(gpas);
    while (!gpas->outer_done && gpas->nrows < gpas->nrooms)
    {
        TupleTableSlot *tts;

        CHECK_FOR_INTERRUPTS();
        tts = ExecProcNode(outer);
        if (TupIsNull(tts))
        {
            gpas->outer_done = true;
            break;
        }
        slot_getallattrs(tts);
        for (int c = 0; c < gpas->ncols; c++)
        {
            int16 attlen = gpas->colmeta[c].attlen;

            gpas->isnull[c][gpas->nrows] = tts->tts_isnull[c];
            if (attlen > 0)
            {
                /* pass-by-value datums are stored in the low bytes (little endian) */
                if (!tts->tts_isnull[c])
                    memcpy(gpas->values[c] + (size_t) gpas->nrows * attlen, &tts->tts_values[c], attlen);
                else
                    memset(gpas->values[c] + (size_t) gpas->nrows * attlen, 0, attlen);
            }
            else
                ((const void **) gpas->values[c])[gpas->nrows] =
                    tts->tts_isnull[c] ? NULL : DatumGetPointer(tts->tts_values[c]);
        }
        gpas->nrows++;
    }
    if (gpas->nrows == 0)
        return 0;               /* slot->kds == NULL: end of scan */
    length = pgstrom_kds_column_length(gpas->ncols, gpas->colmeta, gpas->nrows,
                                       (const void *const *) gpas->values,
                                       (const uint8_t *const *) gpas->isnull);
    kds = pgs_chunk_alloc(length);
    if (kds == NULL)
        return StromError_OutOfMemory;
    if (pgstrom_kds_column_build(kds, length, gpas->ncols, gpas->colmeta, gpas->nrows,
                                 (const void *const *) gpas->values,
                                 (const uint8_t *const *) gpas->isnull) != 0)
    {
        pgs_chunk_free(kds);
        return StromError_DataStoreCorruption;
    }
    slot->kds = (const kern_data_store *) kds;
    slot->release = chunk_release;
    return 0;
}

static CustomPlanState *
gpupreagg_begin_glue(CustomPlan *node, EState *estate, int eflags)
{
    GpuPreAggPlan  *gpreagg = (GpuPreAggPlan *) node;
    GpuPreAggState *gpas = (GpuPreAggState *) palloc0(sizeof(GpuPreAggState));
    size_t          row_bytes = 0;
    int             rc;

    NodeSetTag(gpas, T_CustomPlanState);
    gpas->cps.ps.plan = &node->plan;
    gpas->cps.ps.state = estate;
    gpas->cps.methods = &gpupreagg_plan_methods;
    outerPlanState(gpas) = ExecInitNode(outerPlan(gpreagg), estate, eflags);
    gpas->scan_desc = ExecGetResultType(outerPlanState(gpas));
    ExecInitResultTupleSlot(estate, &gpas->cps.ps);
    ExecAssignResultTypeFromTL(&gpas->cps.ps);

    /* which result columns are text / bpchar grouping keys (a GpuPreAgg
     * target list carries them as plain Vars of the outer plan) */
    gpas->nresult = list_length(node->plan.targetlist);
    gpas->key_typmod = (int32 *) palloc0(sizeof(int32) * (gpas->nresult + 1));
    gpas->key_buf = (char **) palloc0(sizeof(char *) * (gpas->nresult + 1));
    gpas->key_buflen = (size_t *) palloc0(sizeof(size_t) * (gpas->nresult + 1));
    {
        ListCell   *lc;
        int         c = 0;

        foreach(lc, node->plan.targetlist)
        {
            TargetEntry *tle = (TargetEntry *) lfirst(lc);
            Var        *var = (Var *) tle->expr;

            gpas->key_typmod[c] = KEY_NOT_TEXT;
            if (var != NULL && IsA(var, Var) &&
                (var->vartype == TEXTOID || var->vartype == BPCHAROID))
            {
                gpas->key_typmod[c] = var->vartypmod;
                /* allocated here, in the executor's per-query context;
                 * repalloc() keeps a chunk in the context it came from */
                gpas->key_buflen[c] = 256;
                gpas->key_buf[c] = (char *) palloc(gpas->key_buflen[c]);
            }
            c++;
        }
    }

    /* column arrays of one chunk (pg_strom.chunk_size MB of outer tuples) */
    gpas->ncols = gpas->scan_desc->natts;
    gpas->colmeta = (kern_colmeta *) palloc0(sizeof(kern_colmeta) * gpas->ncols);
    for (int c = 0; c < gpas->ncols; c++)
    {
        gpas->colmeta[c].attbyval = gpas->scan_desc->attbyval[c];
        gpas->colmeta[c].attalign = gpas->scan_desc->attalign[c];
        gpas->colmeta[c].attlen = gpas->scan_desc->attlen[c];
        gpas->colmeta[c].attnum = (int16) (c + 1);
        row_bytes += (gpas->scan_desc->attlen[c] > 0 ? gpas->scan_desc->attlen[c] : 32);
    }
    pgstrom_colmeta_set_cacheoff(gpas->ncols, gpas->colmeta);
    gpas->nrooms = (uint32) ((((size_t) guc_chunk_size) << 20) / (row_bytes ? row_bytes : 1));
    gpas->values = (char **) palloc0(sizeof(char *) * gpas->ncols);
    gpas->isnull = (uint8_t **) palloc0(sizeof(uint8_t *) * gpas->ncols);
    for (int c = 0; c < gpas->ncols; c++)
    {
        int16 attlen = gpas->colmeta[c].attlen;

        gpas->values[c] = (char *) palloc((size_t) gpas->nrooms * (attlen > 0 ? attlen : sizeof(void *)));
        gpas->isnull[c] = (uint8_t *) palloc(gpas->nrooms);
    }

    /* devices are opened by the first GpuPreAgg of the backend */
    {
        static bool cuda_ready = false;

        if (!cuda_ready)
        {
            rc = pgs_cuda_init(NULL, 0);
            if (rc != 0)
                elog(ERROR, "PG-Strom: %s (%s)", pgs_last_error(), pgstrom_strerror(rc));
            cuda_ready = true;
        }
    }
    /* device program + session (pgstrom_get_devprog_key / clserv_lookup_device_program,
     * gpupreagg.c:2281-2307) */
    rc = gpupreagg_begin(gpreagg->plan, gpreagg->idx, 0, gpupreagg_next_chunk, gpas, &gpas->state);
    if (rc != 0)
        elog(ERROR, "PG-Strom: GpuPreAgg: %s (%s)", pgs_last_error(), pgstrom_strerror(rc));
    gpas->next_live = live_states;
    live_states = gpas;
    return &gpas->cps;
}

static TupleTableSlot *
gpupreagg_exec_glue(CustomPlanState *node)
{
    GpuPreAggState *gpas = (GpuPreAggState *) node;
    TupleTableSlot *slot = gpas->cps.ps.ps_ResultTupleSlot;
    int             rc;

    ExecClearTuple(slot);
    /* one partial row per call, straight into the slot's arrays
     * (pgstrom_fetch_data_store of a TUPSLOT store, datastore.c:169-242) */
    rc = gpupreagg_exec(gpas->state, slot->tts_values, (char *) slot->tts_isnull);
    if (rc < 0)
        elog(ERROR, "PG-Strom: GpuPreAgg: %s (%s)", pgs_last_error(), pgstrom_strerror(-rc));
    if (rc == 0)
        return NULL;
    /* varlena grouping keys: word -> varlena that lives until the next call
     * (the reference's kernels fix the key pointers up to host addresses,
     * opencl_gpupreagg.h:326-366; here the strings of long keys come from the
     * session's key heap) */
    for (int c = 0; c < gpas->nresult; c++)
    {
        const void *heap = NULL;
        size_t      heap_len = 0, n;

        if (gpas->key_typmod[c] == KEY_NOT_TEXT || slot->tts_isnull[c])
            continue;
        if (gpupreagg_key_heap(gpas->state, &heap, &heap_len) != 0)
            elog(ERROR, "PG-Strom: GpuPreAgg: %s", pgs_last_error());
        for (;;)
        {
            n = pgstrom_fixup_kernel_text_heap(slot->tts_values[c], gpas->key_typmod[c],
                                               heap, heap_len,
                                               gpas->key_buf[c], gpas->key_buflen[c]);
            if (n > 0)
                break;
            /* the longest value a key heap of this size can hold, padded */
            if (gpas->key_buflen[c] >= heap_len + 4 * (size_t) Max(gpas->key_typmod[c], 0) + 64)
                elog(ERROR, "PG-Strom: GpuPreAgg: corrupted text grouping key");
            gpas->key_buflen[c] = heap_len + 4 * (size_t) Max(gpas->key_typmod[c], 0) + 64;
            gpas->key_buf[c] = (char *) repalloc(gpas->key_buf[c], gpas->key_buflen[c]);
        }
        slot->tts_values[c] = PointerGetDatum(gpas->key_buf[c]);
    }
    return ExecStoreVirtualTuple(slot);
}

static void
gpupreagg_end_glue(CustomPlanState *node)
{
    GpuPreAggState *gpas = (GpuPreAggState *) node;
    const char     *notice = NULL;

    live_states_remove(gpas);
    if (gpas->state)
        notice = gpupreagg_end(gpas->state);
    if (notice)
        elog(NOTICE, "%s", notice);     /* "GpuPreAgg: %u chunks were re-checked by CPU" */
    gpas->state = NULL;
    ExecEndNode(outerPlanState(node));
}

static void
gpupreagg_rescan_glue(CustomPlanState *node)
{
    GpuPreAggState *gpas = (GpuPreAggState *) node;
    int             rc = gpupreagg_rescan(gpas->state);

    if (rc != 0)
        elog(ERROR, "PG-Strom: GpuPreAgg rescan: %s", pgs_last_error());
    gpas->outer_done = false;
    ExecReScan(outerPlanState(node));
}

static void
gpupreagg_explain_glue(CustomPlanState *node, List *ancestors, ExplainState *es)
{
    GpuPreAggState *gpas = (GpuPreAggState *) node;
    const char     *text = gpupreagg_explain(gpas->state, es->verbose, es->analyze);

    (void) ancestors;
    /* "label: value" entries; the value of Kernel Source spans many lines, so
     * an entry ends where the next known label begins */
    while (text && *text)
    {
        static const char *labels[] = { "Bulkload: ", "Kernel Source: ", "Perfmon: " };
        const char *next = NULL;
        const char *sep = strstr(text, ": ");
        char       *label, *value;
        size_t      len;

        if (!sep)
            break;
        for (size_t i = 0; i < sizeof(labels) / sizeof(labels[0]); i++)
        {
            /* the next label at the start of a line behind this entry's own */
            const char *p = sep;

            while ((p = strstr(p, labels[i])) != NULL)
            {
                if (p > text && p[-1] == '\n' && (!next || p < next))
                    next = p;
                p += strlen(labels[i]);
            }
        }
        len = next ? (size_t) (next - text) : strlen(text);
        label = pstrdup(text);
        label[len] = '\0';
        while (len > 0 && label[len - 1] == '\n')
            label[--len] = '\0';
        label[sep - text] = '\0';
        value = label + (sep - text) + 2;
        ExplainPropertyText(label, value, es);
        text = next;
    }
}

static Bitmapset *
gpupreagg_get_relids(CustomPlanState *node)
{
    (void) node;
    return NULL;                /* gpupreagg.c:2880: no relation of its own */
}

static void
gpupreagg_textout_plan(StringInfo str, const CustomPlan *node)
{
    const GpuPreAggPlan *gpreagg = (const GpuPreAggPlan *) node;

    appendStringInfo(str, " :gpupreagg_index %d :extra_flags %d", gpreagg->idx,
                     pgs_plan_extra_flags(gpreagg->plan, gpreagg->idx));
}

static CustomPlan *
gpupreagg_copy_plan(const CustomPlan *from)
{
    GpuPreAggPlan *newnode = (GpuPreAggPlan *) palloc(sizeof(GpuPreAggPlan));

    memcpy(newnode, from, sizeof(GpuPreAggPlan));   /* pgs_plan is shared, read-only */
    return &newnode->cplan;
}

/* gpupreagg.c:2946-2979 */
static void
pgstrom_init_gpupreagg(void)
{
    memset(&gpupreagg_plan_methods, 0, sizeof(CustomPlanMethods));
    gpupreagg_plan_methods.CustomName          = "GpuPreAgg";
    gpupreagg_plan_methods.BeginCustomPlan     = gpupreagg_begin_glue;
    gpupreagg_plan_methods.ExecCustomPlan      = gpupreagg_exec_glue;
    gpupreagg_plan_methods.EndCustomPlan       = gpupreagg_end_glue;
    gpupreagg_plan_methods.ReScanCustomPlan    = gpupreagg_rescan_glue;
    gpupreagg_plan_methods.ExplainCustomPlan   = gpupreagg_explain_glue;
    gpupreagg_plan_methods.GetRelidsCustomPlan = gpupreagg_get_relids;
    gpupreagg_plan_methods.TextOutCustomPlan   = gpupreagg_textout_plan;
    gpupreagg_plan_methods.CopyCustomPlan      = gpupreagg_copy_plan;
}

/* main.c:237-281 */
void
_PG_init(void)
{
    if (!process_shared_preload_libraries_in_progress)
        ereport(ERROR,
                (errcode(ERRCODE_OBJECT_NOT_IN_PREREQUISITE_STATE),
                 errmsg("PG-Strom must be loaded via shared_preload_libraries")));
    /* was: OpenCL entry points, device info, program cache, message queues,
     * shared memory, the OpenCL background worker (main.c:248-266).  The CUDA
     * layer needs none of them; devices are opened by the first backend that
     * runs a GpuPreAgg (pgs_cuda_init is idempotent and cheap) */
    pgstrom_init_gucs();
    pgstrom_init_gpupreagg();
    RegisterResourceReleaseCallback(pgstrom_release_callback, NULL);
    planner_hook_next = planner_hook;
    planner_hook = pgstrom_grafter_with_subplans;
}
