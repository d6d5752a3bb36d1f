/*
 * Test driver for pg_glue/gpupreagg_glue.c under the stand-in backend
 * (pg_stub/): plays PostgreSQL for one query,
 *
 *   SELECT key, count(*), sum(w), avg(v), min(v), max(v)
 *     FROM bench_where WHERE f < :c GROUP BY key
 *
 * i.e. builds the plan tree standard_planner() would return (HashAggregate
 * over SeqScan), lets the glue's planner hook rewrite it, and runs the
 * rewritten tree's GpuPreAgg node through its CustomPlan callbacks.  Called
 * from tests/test_pg_glue_plan.py through ctypes.
 */
#include "pg_nodes_stub.h"
#include "pgstrom_cuda.h"

extern void _PG_init(void);
extern char *pgstrom_plan_to_json(PlannedStmt *pstmt, Plan *plan);
extern int  pg_stub_try(void (*fn)(void *), void *arg);
extern const char *pg_stub_error(void);
extern Oid  pg_stub_define_function(const char *nsp, const char *name, int nargs, const Oid *argtypes);
extern Oid  pg_stub_define_operator(const char *name, Oid left, Oid right, const char *funcname);
extern void pg_stub_define_table(const char *name, int ncols, const Oid *coltypes, int64 nrows,
                                 Datum *values, bool *isnull);
extern void pg_stub_set_standard_plan(PlannedStmt *pstmt);

#define F8 701
static Oid  op_int4lt, fn_count, fn_sum_int4, fn_avg_f8, fn_min_f8, fn_max_f8;
static PlannedStmt *the_stmt;

static TargetEntry *
tle(Expr *e, int resno, const char *name)
{
    return makeTargetEntry(e, (AttrNumber) resno, name ? pstrdup(name) : NULL, false);
}

static Expr *
aggref(Oid fn, Oid aggtype, Expr *arg, bool star)
{
    Aggref *a = makeNode(Aggref);

    a->aggfnoid = fn;
    a->aggtype = aggtype;
    a->aggstar = star;
    a->aggkind = 'n';
    if (arg)
        a->args = list_make1(tle(arg, 1, NULL));
    return (Expr *) a;
}

/* the plan of the query above, as set_plan_references() leaves it */
static PlannedStmt *
build_plan(int qual_const, long num_groups)
{
    static const char *colnames[] = { "f", "key", "v", "w" };
    static const Oid   coltypes[] = { INT4OID, INT4OID, F8, INT4OID };
    PlannedStmt    *pstmt = makeNode(PlannedStmt);
    RangeTblEntry  *rte = makeNode(RangeTblEntry);
    SeqScan        *scan = makeNode(SeqScan);
    Agg            *agg = makeNode(Agg);
    OpExpr         *qual = makeNode(OpExpr);
    Const          *c = makeNullConst(INT4OID, -1, InvalidOid);

    rte->relid = 50000;
    rte->eref = makeNode(Alias);
    rte->eref->aliasname = pstrdup("bench_where");
    pstmt->rtable = list_make1(rte);

    scan->scanrelid = 1;
    for (int i = 0; i < 4; i++)
        scan->plan.targetlist = lappend(scan->plan.targetlist,
                                        tle((Expr *) makeVar(1, (AttrNumber) (i + 1), coltypes[i], -1, InvalidOid, 0),
                                            i + 1, colnames[i]));
    c->constisnull = false;
    c->constbyval = true;
    c->constlen = 4;
    c->constvalue = (Datum) qual_const;
    qual->opno = op_int4lt;
    qual->opfuncid = get_opcode(op_int4lt);
    qual->opresulttype = BOOLOID;
    qual->args = list_make2(makeVar(1, 1, INT4OID, -1, InvalidOid, 0), c);
    scan->plan.qual = list_make1(qual);
    scan->plan.startup_cost = 0.0;
    scan->plan.total_cost = 2000000.0;
    scan->plan.plan_rows = 12500000;
    scan->plan.plan_width = 20;

    agg->aggstrategy = AGG_HASHED;
    agg->numCols = 1;
    agg->grpColIdx = (AttrNumber *) palloc(sizeof(AttrNumber));
    agg->grpColIdx[0] = 2;
    agg->numGroups = num_groups;
    agg->plan.targetlist = lappend(NIL, tle((Expr *) makeVar(OUTER_VAR, 2, INT4OID, -1, InvalidOid, 0), 1, "key"));
    agg->plan.targetlist = lappend(agg->plan.targetlist, tle(aggref(fn_count, 20, NULL, true), 2, "count"));
    agg->plan.targetlist = lappend(agg->plan.targetlist,
                                   tle(aggref(fn_sum_int4, 20, (Expr *) makeVar(OUTER_VAR, 4, INT4OID, -1, InvalidOid, 0), false),
                                       3, "sum"));
    agg->plan.targetlist = lappend(agg->plan.targetlist,
                                   tle(aggref(fn_avg_f8, F8, (Expr *) makeVar(OUTER_VAR, 3, F8, -1, InvalidOid, 0), false),
                                       4, "avg"));
    agg->plan.targetlist = lappend(agg->plan.targetlist,
                                   tle(aggref(fn_min_f8, F8, (Expr *) makeVar(OUTER_VAR, 3, F8, -1, InvalidOid, 0), false),
                                       5, "min"));
    agg->plan.targetlist = lappend(agg->plan.targetlist,
                                   tle(aggref(fn_max_f8, F8, (Expr *) makeVar(OUTER_VAR, 3, F8, -1, InvalidOid, 0), false),
                                       6, "max"));
    agg->plan.startup_cost = 2100000.0;
    agg->plan.total_cost = 2100010.0;
    agg->plan.plan_rows = (double) num_groups;
    agg->plan.plan_width = 44;
    outerPlan(agg) = &scan->plan;
    pstmt->planTree = &agg->plan;
    return pstmt;
}

static void
do_init(void *arg)
{
    Oid i4[2] = { INT4OID, INT4OID }, f8[1] = { F8 };

    (void) arg;
    op_int4lt = pg_stub_define_operator("<", INT4OID, INT4OID, "int4lt");
    fn_count = pg_stub_define_function("pg_catalog", "count", 0, NULL);
    fn_sum_int4 = pg_stub_define_function("pg_catalog", "sum", 1, i4);
    fn_avg_f8 = pg_stub_define_function("pg_catalog", "avg", 1, f8);
    fn_min_f8 = pg_stub_define_function("pg_catalog", "min", 1, f8);
    fn_max_f8 = pg_stub_define_function("pg_catalog", "max", 1, f8);
    pg_stub_define_function("pg_catalog", "int8", 1, i4);      /* the int4 -> int8 cast */
    {
        static const Oid coltypes[] = { INT4OID, INT4OID, F8, INT4OID };

        pg_stub_define_table("bench_where", 4, coltypes, 0, NULL, NULL);
    }
    _PG_init();
}
int driver_init(void) { return pg_stub_try(do_init, NULL); }

/* JSON of the plain plan (what the glue hands to pgstrom_grafter_json) */
const char *
driver_plan_json(int qual_const, long num_groups)
{
    PlannedStmt *pstmt = build_plan(qual_const, num_groups);

    return pgstrom_plan_to_json(pstmt, pstmt->planTree);
}

static void
do_plan(void *arg)
{
    (void) arg;
    the_stmt = planner_hook(NULL, 0, NULL);
}

/* runs the planner hook; returns a one-line description of the resulting
 * tree, e.g. "Agg[6] -> Custom(GpuPreAgg)[10] -> SeqScan[4] quals=0" */
const char *
driver_run_planner(int qual_const, long num_groups)
{
    static char buf[512];
    int         n = 0;

    pg_stub_set_standard_plan(build_plan(qual_const, num_groups));
    if (pg_stub_try(do_plan, NULL) != 0)
    {
        snprintf(buf, sizeof(buf), "ERROR: %s", pg_stub_error());
        return buf;
    }
    for (Plan *p = the_stmt->planTree; p; p = outerPlan(p))
    {
        const char *name = IsA(p, Agg) ? "Agg" : IsA(p, Sort) ? "Sort" : IsA(p, SeqScan) ? "SeqScan" : "?";
        char        custom[64];

        if (IsA(p, CustomPlan))
        {
            snprintf(custom, sizeof(custom), "Custom(%s)", ((CustomPlan *) p)->methods->CustomName);
            name = custom;
        }
        n += snprintf(buf + n, sizeof(buf) - n, "%s%s[%d]", n ? " -> " : "", name, list_length(p->targetlist));
        if (IsA(p, SeqScan))
            n += snprintf(buf + n, sizeof(buf) - n, " quals=%d", list_length(p->qual));
    }
    return buf;
}

/* executes the GpuPreAgg node of the rewritten plan over `nrows` rows of
 * (f, key, v, w) and stores its partial rows, ncols Datums + ncols null flags
 * each.  Returns the number of rows, or -1 (message: driver_error()). */
typedef struct { Datum *out; bool *out_null; long max_rows; long nrows; int ncols; } ExecArg;

static void
do_exec(void *p)
{
    ExecArg        *a = (ExecArg *) p;
    EState         *estate = (EState *) palloc0(sizeof(EState));
    Plan           *plan = the_stmt->planTree;
    PlanState      *ps;
    TupleTableSlot *slot;

    while (plan && !IsA(plan, CustomPlan))
        plan = outerPlan(plan);
    if (plan == NULL)
        elog(ERROR, "no GpuPreAgg node in the plan");
    estate->es_plannedstmt = the_stmt;
    estate->es_range_table = the_stmt->rtable;
    ps = ExecInitNode(plan, estate, 0);
    a->ncols = list_length(plan->targetlist);
    for (int pass = 0; pass < 2; pass++)
    {
        /* twice: the second pass goes through ReScanCustomPlan */
        a->nrows = 0;
        if (pass == 1)
            ExecReScan(ps);
        while ((slot = ExecProcNode(ps)) != NULL)
        {
            if (a->nrows >= a->max_rows)
                elog(ERROR, "result buffer too small");
            memcpy(a->out + a->nrows * a->ncols, slot->tts_values, sizeof(Datum) * a->ncols);
            memcpy(a->out_null + a->nrows * a->ncols, slot->tts_isnull, sizeof(bool) * a->ncols);
            a->nrows++;
        }
    }
    {
        ExplainState es;

        memset(&es, 0, sizeof(es));
        es.verbose = true;
        ((CustomPlanState *) ps)->methods->ExplainCustomPlan((CustomPlanState *) ps, NIL, &es);
    }
    ExecEndNode(ps);
}

long
driver_execute(long nrows, Datum *table_values, bool *table_isnull,
               Datum *out, bool *out_null, long max_rows, int *ncols)
{
    static const Oid coltypes[] = { INT4OID, INT4OID, F8, INT4OID };
    ExecArg a;

    memset(&a, 0, sizeof(a));
    a.out = out;
    a.out_null = out_null;
    a.max_rows = max_rows;
    pg_stub_define_table("bench_where", 4, coltypes, nrows, table_values, table_isnull);
    if (pg_stub_try(do_exec, &a) != 0)
        return -1;
    *ncols = a.ncols;
    return a.nrows;
}
const char *driver_error(void) { return pg_stub_error(); }
