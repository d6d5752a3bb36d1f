/*
 * TEST STAND-IN (see pg_nodes_stub.h): just enough of a PostgreSQL backend to
 * load pg_glue/gpupreagg_glue.c, run its planner hook over a hand-built plan
 * tree and drive its CustomPlan callbacks - memory, lists, StringInfo, node
 * constructors, a catalog of a dozen types / functions / operators, a GUC
 * registry, an executor whose SeqScan reads rows the test handed over.
 * elog(ERROR) longjmps to the test's entry point (pg_stub_try).
 */
#include "pg_nodes_stub.h"
#include <setjmp.h>

int  pg_stub_module_magic;
bool process_shared_preload_libraries_in_progress = true;
planner_hook_type planner_hook = NULL;

/* ---- errors ---- */
static jmp_buf  stub_jmp;
static bool     stub_jmp_armed = false;
char            pg_stub_last_notice[256];

void
pg_stub_raise(int level, const char *fmt, ...)
{
    char    buf[512];
    va_list ap;

    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (level < ERROR)
    {
        snprintf(pg_stub_last_notice, sizeof(pg_stub_last_notice), "%.255s", buf);
        return;
    }
    snprintf(pg_stub_error_message, sizeof(pg_stub_error_message), "%.255s", buf);
    if (!stub_jmp_armed)
    {
        fprintf(stderr, "pg_stub: ERROR outside pg_stub_try: %s\n", buf);
        abort();
    }
    longjmp(stub_jmp, 1);
}

/* runs fn(arg); returns 0, or 1 when it raised an ERROR (message in
 * pg_stub_error_message) */
int
pg_stub_try(void (*fn)(void *), void *arg)
{
    int rc = 0;

    pg_stub_error_message[0] = '\0';
    stub_jmp_armed = true;
    if (setjmp(stub_jmp) == 0)
        fn(arg);
    else
        rc = 1;
    stub_jmp_armed = false;
    return rc;
}
const char *pg_stub_error(void) { return pg_stub_error_message; }
const char *pg_stub_notice(void) { return pg_stub_last_notice; }

/* ---- memory, lists, StringInfo, nodes ---- */
void *palloc(Size n) { void *p = malloc(n ? n : 1); if (!p) abort(); return p; }
void *palloc0(Size n) { void *p = calloc(1, n ? n : 1); if (!p) abort(); return p; }
char *pstrdup(const char *s) { char *p = palloc(strlen(s) + 1); strcpy(p, s); return p; }
void  pfree(void *p) { free(p); }
void *repalloc(void *p, Size n) { p = realloc(p, n ? n : 1); if (!p) abort(); return p; }
Node *pg_stub_new_node(Size size, NodeTag tag) { Node *n = palloc0(size); n->type = tag; return n; }

static List *
list_append_cell(List *list, ListCell *cell, NodeTag tag)
{
    if (list == NIL)
    {
        list = palloc0(sizeof(List));
        list->type = tag;
    }
    cell->next = NULL;
    if (list->tail)
        list->tail->next = cell;
    else
        list->head = cell;
    list->tail = cell;
    list->length++;
    return list;
}
List *lappend(List *list, void *datum)
{ ListCell *c = palloc0(sizeof(ListCell)); c->data.ptr_value = datum; return list_append_cell(list, c, T_List); }
List *lappend_int(List *list, int datum)
{ ListCell *c = palloc0(sizeof(ListCell)); c->data.int_value = datum; return list_append_cell(list, c, T_IntList); }
int list_length(const List *list) { return list ? list->length : 0; }
void *list_nth(const List *list, int n)
{
    ListCell *c = list ? list->head : NULL;
    while (c && n-- > 0)
        c = c->next;
    return c ? c->data.ptr_value : NULL;
}
Value *makeString(char *str) { Value *v = palloc0(sizeof(Value)); v->type = T_String; v->str = str; return v; }

void initStringInfo(StringInfo str)
{ str->maxlen = 1024; str->data = palloc(str->maxlen); str->len = 0; str->data[0] = '\0'; }
static void
sinfo_room(StringInfo str, int need)
{
    if (str->len + need + 1 > str->maxlen)
    {
        while (str->len + need + 1 > str->maxlen)
            str->maxlen *= 2;
        str->data = realloc(str->data, str->maxlen);
    }
}
void appendStringInfoString(StringInfo str, const char *s)
{ int n = (int) strlen(s); sinfo_room(str, n); memcpy(str->data + str->len, s, n + 1); str->len += n; }
void appendStringInfoChar(StringInfo str, char ch)
{ sinfo_room(str, 1); str->data[str->len++] = ch; str->data[str->len] = '\0'; }
void appendStringInfo(StringInfo str, const char *fmt, ...)
{
    char    buf[1024];
    va_list ap;

    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    appendStringInfoString(str, buf);
}

TargetEntry *makeTargetEntry(Expr *expr, AttrNumber resno, char *resname, bool resjunk)
{ TargetEntry *t = makeNode(TargetEntry); t->expr = expr; t->resno = resno; t->resname = resname; t->resjunk = resjunk; return t; }
Var *makeVar(Index varno, AttrNumber varattno, Oid vartype, int32 vartypmod, Oid varcollid, Index varlevelsup)
{ Var *v = makeNode(Var); (void) varlevelsup; v->varno = varno; v->varattno = varattno; v->vartype = vartype;
  v->vartypmod = vartypmod; v->varcollid = varcollid; return v; }
Const *makeNullConst(Oid consttype, int32 consttypmod, Oid constcollid)
{ Const *c = makeNode(Const); c->consttype = consttype; c->consttypmod = consttypmod; c->constcollid = constcollid;
  c->constisnull = true; return c; }

/* ---- catalog ---- */
typedef struct { Oid oid; const char *name; int16 len; bool byval; char align; } StubType;
static const StubType stub_types[] = {
    { 16, "bool", 1, true, 1 },   { 21, "int2", 2, true, 2 },     { 23, "int4", 4, true, 4 },
    { 20, "int8", 8, true, 8 },   { 700, "float4", 4, true, 4 },  { 701, "float8", 8, true, 8 },
    { 1700, "numeric", -1, false, 4 }, { 25, "text", -1, false, 4 }, { 1082, "date", 4, true, 4 },
    { 1114, "timestamp", 8, true, 8 }, { 1042, "bpchar", -1, false, 4 },
};
static const StubType *
stub_type(Oid typid)
{
    for (size_t i = 0; i < sizeof(stub_types) / sizeof(stub_types[0]); i++)
        if (stub_types[i].oid == typid)
            return &stub_types[i];
    pg_stub_raise(ERROR, "cache lookup failed for type %u", typid);
    return NULL;
}
char *pg_stub_type_name(Oid typid) { return pstrdup(stub_type(typid)->name); }
Oid TypenameGetTypid(const char *typname)
{
    for (size_t i = 0; i < sizeof(stub_types) / sizeof(stub_types[0]); i++)
        if (strcmp(stub_types[i].name, typname) == 0)
            return stub_types[i].oid;
    return InvalidOid;
}
void get_typlenbyvalalign(Oid typid, int16 *typlen, bool *typbyval, char *typalign)
{ const StubType *t = stub_type(typid); *typlen = t->len; *typbyval = t->byval; *typalign = t->align; }
void getTypeOutputInfo(Oid type, Oid *typOutput, bool *typIsVarlena)
{ *typOutput = type; *typIsVarlena = stub_type(type)->len < 0; }
void getTypeInputInfo(Oid type, Oid *typInput, Oid *typIOParam) { *typInput = type; *typIOParam = type; }
char *OidOutputFunctionCall(Oid functionId, Datum val)
{
    char buf[64];

    switch (functionId)
    {
        case 16: snprintf(buf, sizeof(buf), "%s", val ? "t" : "f"); break;
        case 21: snprintf(buf, sizeof(buf), "%d", (int) (int16) val); break;
        case 23: case 1082: snprintf(buf, sizeof(buf), "%d", (int) (int32) val); break;
        case 20: case 1114: snprintf(buf, sizeof(buf), "%lld", (long long) (int64) val); break;
        case 701: snprintf(buf, sizeof(buf), "%.17g", DatumGetFloat8(val)); break;
        case 25: case 1700: snprintf(buf, sizeof(buf), "%s", (const char *) DatumGetPointer(val)); break;
        default: pg_stub_raise(ERROR, "no output function for type %u", functionId);
    }
    return pstrdup(buf);
}
Datum OidInputFunctionCall(Oid functionId, char *str, Oid typioparam, int32 typmod)
{
    (void) typioparam; (void) typmod;
    switch (functionId)
    {
        case 16: return (Datum) (str[0] == 't');
        case 21: case 23: case 1082: return (Datum) (uint32_t) atoi(str);
        case 20: case 1114: return (Datum) atoll(str);
        case 701: return Float8GetDatum(atof(str));
        default: return PointerGetDatum(str);
    }
}
bool lc_collate_is_c(Oid collation) { (void) collation; return true; }
char *get_collation_name(Oid colloid) { return pstrdup(colloid == 950 ? "C" : "default"); }

/* functions and aggregates: (namespace, name, argument types).  pg_catalog's
 * are preloaded as the tests need them; pgstrom.* (pg_strom--1.0.sql) are
 * registered on first look-up and remembered, so that the test can check
 * WHICH catalog entries the glue asked for */
#define STUB_MAX_FUNCS  256
typedef struct { Oid oid; char nsp[16]; char name[48]; int nargs; Oid argtypes[8]; } StubFunc;
static StubFunc stub_funcs[STUB_MAX_FUNCS];
static int      stub_nfuncs = 0;

static Oid
stub_func_add(const char *nsp, const char *name, int nargs, const Oid *argtypes)
{
    StubFunc *f = &stub_funcs[stub_nfuncs];

    if (stub_nfuncs == STUB_MAX_FUNCS)
        pg_stub_raise(ERROR, "stub catalog full");
    f->oid = 16384 + stub_nfuncs;
    snprintf(f->nsp, sizeof(f->nsp), "%s", nsp);
    snprintf(f->name, sizeof(f->name), "%s", name);
    f->nargs = nargs;
    for (int i = 0; i < nargs && i < 8; i++)
        f->argtypes[i] = argtypes[i];
    stub_nfuncs++;
    return f->oid;
}
/* test entry point: pre-register a pg_catalog function / aggregate */
Oid pg_stub_define_function(const char *nsp, const char *name, int nargs, const Oid *argtypes)
{ return stub_func_add(nsp, name, nargs, argtypes); }
/* test entry point: "nsp.name(type,type)" of entry i, NULL past the end */
const char *
pg_stub_function_signature(int i)
{
    static char buf[256];
    int         n;

    if (i < 0 || i >= stub_nfuncs)
        return NULL;
    n = snprintf(buf, sizeof(buf), "%s.%s(", stub_funcs[i].nsp, stub_funcs[i].name);
    for (int a = 0; a < stub_funcs[i].nargs; a++)
        n += snprintf(buf + n, sizeof(buf) - n, "%s%s", a ? "," : "", stub_type(stub_funcs[i].argtypes[a])->name);
    snprintf(buf + n, sizeof(buf) - n, ")");
    return buf;
}
Oid
LookupFuncName(List *funcname, int nargs, const Oid *argtypes, bool noError)
{
    const char *nsp = list_length(funcname) == 2 ? strVal(linitial(funcname)) : "pg_catalog";
    const char *name = strVal(list_nth(funcname, list_length(funcname) - 1));

    for (int i = 0; i < stub_nfuncs; i++)
    {
        StubFunc *f = &stub_funcs[i];

        if (strcmp(f->nsp, nsp) == 0 && strcmp(f->name, name) == 0 && f->nargs == nargs &&
            memcmp(f->argtypes, argtypes, sizeof(Oid) * (nargs < 8 ? nargs : 8)) == 0)
            return f->oid;
    }
    if (strcmp(nsp, "pgstrom") == 0)
        return stub_func_add(nsp, name, nargs, argtypes);   /* CREATE EXTENSION pg_strom */
    if (!noError)
        pg_stub_raise(ERROR, "function %s.%s does not exist", nsp, name);
    return InvalidOid;
}
static StubFunc *
stub_func(Oid funcid)
{
    for (int i = 0; i < stub_nfuncs; i++)
        if (stub_funcs[i].oid == funcid)
            return &stub_funcs[i];
    pg_stub_raise(ERROR, "cache lookup failed for function %u", funcid);
    return NULL;
}
char *get_func_name(Oid funcid) { return pstrdup(stub_func(funcid)->name); }
Oid get_func_namespace(Oid funcid) { return strcmp(stub_func(funcid)->nsp, "pgstrom") == 0 ? 2200 : 11; }
char *get_namespace_name(Oid nspid) { return pstrdup(nspid == 11 ? "pg_catalog" : nspid == 2200 ? "pgstrom" : "public"); }

/* operators: (name, left, right) -> oid, implementing function */
#define STUB_MAX_OPERS  64
typedef struct { Oid oid; char name[8]; Oid left, right; Oid opcode; } StubOper;
static StubOper stub_opers[STUB_MAX_OPERS];
static int      stub_nopers = 0;
Oid
pg_stub_define_operator(const char *name, Oid left, Oid right, const char *funcname)
{
    StubOper *o = &stub_opers[stub_nopers];
    Oid       args[2] = { left, right };

    o->oid = 90000 + stub_nopers;
    snprintf(o->name, sizeof(o->name), "%s", name);
    o->left = left;
    o->right = right;
    o->opcode = stub_func_add("pg_catalog", funcname, 2, args);
    stub_nopers++;
    return o->oid;
}
static StubOper *
stub_oper(Oid opno)
{
    for (int i = 0; i < stub_nopers; i++)
        if (stub_opers[i].oid == opno)
            return &stub_opers[i];
    pg_stub_raise(ERROR, "cache lookup failed for operator %u", opno);
    return NULL;
}
char *get_opname(Oid opno) { return pstrdup(stub_oper(opno)->name); }
Oid get_opcode(Oid opno) { return stub_oper(opno)->opcode; }
Oid
OpernameGetOprid(List *names, Oid oprleft, Oid oprright)
{
    const char *name = strVal(list_nth(names, list_length(names) - 1));

    for (int i = 0; i < stub_nopers; i++)
        if (strcmp(stub_opers[i].name, name) == 0 && stub_opers[i].left == oprleft &&
            stub_opers[i].right == oprright)
            return stub_opers[i].oid;
    return InvalidOid;
}

/* relations: one table, handed over by the test */
typedef struct { Oid relid; char name[48]; int ncols; Oid coltypes[16]; int64 nrows; Datum *values; bool *isnull; } StubRel;
static StubRel stub_rel;
void
pg_stub_define_table(const char *name, int ncols, const Oid *coltypes, int64 nrows, Datum *values, bool *isnull)
{
    stub_rel.relid = 50000;
    snprintf(stub_rel.name, sizeof(stub_rel.name), "%s", name);
    stub_rel.ncols = ncols;
    memcpy(stub_rel.coltypes, coltypes, sizeof(Oid) * ncols);
    stub_rel.nrows = nrows;
    stub_rel.values = values;       /* [nrows][ncols] */
    stub_rel.isnull = isnull;
}
char *get_rel_name(Oid relid) { (void) relid; return pstrdup(stub_rel.name); }
Oid get_rel_namespace(Oid relid) { (void) relid; return 1; }

/* ---- GUC registry ---- */
typedef struct { char name[64]; int kind; void *addr; void *assign; } StubGuc;
static StubGuc stub_gucs[64];
static int     stub_ngucs = 0;
static void
guc_add(const char *name, int kind, void *addr, void *assign)
{
    StubGuc *g = &stub_gucs[stub_ngucs++];
    snprintf(g->name, sizeof(g->name), "%s", name);
    g->kind = kind; g->addr = addr; g->assign = assign;
}
void DefineCustomBoolVariable(const char *name, const char *s, const char *l, bool *valueAddr, bool bootValue,
                              GucContext c, int flags, void *chk, GucBoolAssignHook assign_hook, void *show)
{ (void) s; (void) l; (void) c; (void) flags; (void) chk; (void) show; *valueAddr = bootValue; guc_add(name, 0, valueAddr, (void *) assign_hook); }
void DefineCustomIntVariable(const char *name, const char *s, const char *l, int *valueAddr, int bootValue,
                             int minValue, int maxValue, GucContext c, int flags, void *chk,
                             GucIntAssignHook assign_hook, void *show)
{ (void) s; (void) l; (void) c; (void) flags; (void) chk; (void) show; (void) minValue; (void) maxValue;
  *valueAddr = bootValue; guc_add(name, 1, valueAddr, (void *) assign_hook); }
void DefineCustomRealVariable(const char *name, const char *s, const char *l, double *valueAddr, double bootValue,
                              double minValue, double maxValue, GucContext c, int flags, void *chk,
                              GucRealAssignHook assign_hook, void *show)
{ (void) s; (void) l; (void) c; (void) flags; (void) chk; (void) show; (void) minValue; (void) maxValue;
  *valueAddr = bootValue; guc_add(name, 2, valueAddr, (void *) assign_hook); }
int pg_stub_num_gucs(void) { return stub_ngucs; }
const char *pg_stub_guc_name(int i) { return i < stub_ngucs ? stub_gucs[i].name : NULL; }
/* SET name = value */
int
pg_stub_set_guc(const char *name, const char *value)
{
    for (int i = 0; i < stub_ngucs; i++)
    {
        StubGuc *g = &stub_gucs[i];

        if (strcmp(g->name, name) != 0)
            continue;
        if (g->kind == 0)
        {
            bool v = (strcmp(value, "on") == 0 || strcmp(value, "true") == 0);
            *((bool *) g->addr) = v;
            if (g->assign) ((GucBoolAssignHook) g->assign)(v, NULL);
        }
        else if (g->kind == 1)
        {
            *((int *) g->addr) = atoi(value);
            if (g->assign) ((GucIntAssignHook) g->assign)(atoi(value), NULL);
        }
        else
        {
            *((double *) g->addr) = atof(value);
            if (g->assign) ((GucRealAssignHook) g->assign)(atof(value), NULL);
        }
        return 0;
    }
    return -1;
}

/* ---- resource owner: AbortTransaction runs the release callbacks ---- */
static ResourceReleaseCallback stub_release_cb[8];
static void *stub_release_arg[8];
static int   stub_nrelease = 0;
void RegisterResourceReleaseCallback(ResourceReleaseCallback callback, void *arg)
{ stub_release_cb[stub_nrelease] = callback; stub_release_arg[stub_nrelease++] = arg; }
int
pg_stub_abort_transaction(void)
{
    for (int i = 0; i < stub_nrelease; i++)
        for (int phase = 0; phase < 3; phase++)
            stub_release_cb[i]((ResourceReleasePhase) phase, false, true, stub_release_arg[i]);
    return stub_nrelease;
}

/* ---- costs: the shape of costsize.c's cost_sort / cost_agg (in-memory sort,
 * no aggregate transition costs) with PostgreSQL's default constants ---- */
int    work_mem = 4096;
double cpu_tuple_cost = 0.01, cpu_operator_cost = 0.0025;
void
cost_sort(Path *path, PlannerInfo *root, List *pathkeys, Cost input_cost, double tuples, int width,
          Cost comparison_cost, int sort_mem, double limit_tuples)
{
    double  n = tuples < 2.0 ? 2.0 : tuples;
    double  lg = 0.0;

    (void) root; (void) pathkeys; (void) width; (void) sort_mem; (void) limit_tuples;
    for (double x = n; x > 1.0; x /= 2.0)
        lg += 1.0;
    path->startup_cost = input_cost + (comparison_cost + 2.0 * cpu_operator_cost) * n * lg;
    path->total_cost = path->startup_cost + cpu_operator_cost * n;
}
void
cost_agg(Path *path, PlannerInfo *root, AggStrategy aggstrategy, const AggClauseCosts *aggcosts,
         int numGroupCols, double numGroups, Cost input_startup_cost, Cost input_total_cost,
         double input_tuples)
{
    (void) root; (void) aggcosts;
    if (aggstrategy == AGG_PLAIN)
    {
        path->startup_cost = input_total_cost;
        path->total_cost = path->startup_cost + cpu_tuple_cost;
    }
    else if (aggstrategy == AGG_SORTED)
    {
        path->startup_cost = input_startup_cost;
        path->total_cost = input_total_cost + cpu_operator_cost * numGroupCols * input_tuples +
            cpu_tuple_cost * numGroups;
    }
    else
    {
        path->startup_cost = input_total_cost + cpu_operator_cost * numGroupCols * input_tuples;
        path->total_cost = path->startup_cost + cpu_tuple_cost * numGroups;
    }
    path->rows = numGroups;
}

/* ---- planner ---- */
static PlannedStmt *stub_next_plan = NULL;
void pg_stub_set_standard_plan(PlannedStmt *pstmt) { stub_next_plan = pstmt; }
PlannedStmt *standard_planner(Query *parse, int cursorOptions, ParamListInfo boundParams)
{ (void) parse; (void) cursorOptions; (void) boundParams; return stub_next_plan; }

/* ---- executor ---- */
typedef struct { PlanState ps; int64 next; } StubScanState;

static TupleDesc
tupdesc_from_types(int natts, const Oid *types)
{
    TupleDesc d = palloc0(sizeof(*d));

    d->natts = natts;
    d->atttypid = palloc(sizeof(Oid) * natts);
    d->attlen = palloc(sizeof(int16) * natts);
    d->attalign = palloc(natts);
    d->attbyval = palloc(sizeof(bool) * natts);
    for (int i = 0; i < natts; i++)
    {
        d->atttypid[i] = types[i];
        get_typlenbyvalalign(types[i], &d->attlen[i], &d->attbyval[i], &d->attalign[i]);
    }
    return d;
}
static TupleTableSlot *
make_slot(TupleDesc desc)
{
    TupleTableSlot *slot = palloc0(sizeof(TupleTableSlot));

    slot->tts_isempty = true;
    slot->tts_tupleDescriptor = desc;
    slot->tts_values = palloc0(sizeof(Datum) * (desc->natts ? desc->natts : 1));
    slot->tts_isnull = palloc0(sizeof(bool) * (desc->natts ? desc->natts : 1));
    return slot;
}
PlanState *
ExecInitNode(Plan *node, EState *estate, int eflags)
{
    if (node == NULL)
        return NULL;
    if (IsA(node, CustomPlan))
        return &((CustomPlan *) node)->methods->BeginCustomPlan((CustomPlan *) node, estate, eflags)->ps;
    if (IsA(node, SeqScan))
    {
        StubScanState *ss = palloc0(sizeof(StubScanState));

        ss->ps.type = T_PlanState;
        ss->ps.plan = node;
        ss->ps.state = estate;
        ss->ps.ps_ResultTupleSlot = make_slot(tupdesc_from_types(stub_rel.ncols, stub_rel.coltypes));
        return &ss->ps;
    }
    pg_stub_raise(ERROR, "stub executor: node %d not supported", (int) nodeTag(node));
    return NULL;
}
TupleTableSlot *
ExecProcNode(PlanState *node)
{
    if (node->type == T_CustomPlanState)
        return ((CustomPlanState *) node)->methods->ExecCustomPlan((CustomPlanState *) node);
    {
        StubScanState  *ss = (StubScanState *) node;
        TupleTableSlot *slot = ss->ps.ps_ResultTupleSlot;

        /* (host quals of the scan are not evaluated here: the tests move
         * every qual to the device or have none) */
        if (ss->next >= stub_rel.nrows)
        {
            slot->tts_isempty = true;
            return NULL;
        }
        memcpy(slot->tts_values, stub_rel.values + ss->next * stub_rel.ncols, sizeof(Datum) * stub_rel.ncols);
        memcpy(slot->tts_isnull, stub_rel.isnull + ss->next * stub_rel.ncols, sizeof(bool) * stub_rel.ncols);
        slot->tts_isempty = false;
        ss->next++;
        return slot;
    }
}
void ExecEndNode(PlanState *node)
{
    if (node && node->type == T_CustomPlanState)
        ((CustomPlanState *) node)->methods->EndCustomPlan((CustomPlanState *) node);
}
void ExecReScan(PlanState *node)
{
    if (node->type == T_CustomPlanState)
        ((CustomPlanState *) node)->methods->ReScanCustomPlan((CustomPlanState *) node);
    else
        ((StubScanState *) node)->next = 0;
}
TupleDesc ExecGetResultType(PlanState *planstate) { return planstate->ps_ResultTupleSlot->tts_tupleDescriptor; }
static Oid
stub_expr_type(Expr *e)
{
    switch (nodeTag(e))
    {
        case T_Var: return ((Var *) e)->vartype;
        case T_Const: return ((Const *) e)->consttype;
        case T_FuncExpr: return ((FuncExpr *) e)->funcresulttype;
        case T_OpExpr: return ((OpExpr *) e)->opresulttype;
        case T_Aggref: return ((Aggref *) e)->aggtype;
        case T_CaseExpr: return ((CaseExpr *) e)->casetype;
        case T_RelabelType: return ((RelabelType *) e)->resulttype;
        default: return BOOLOID;
    }
}
void ExecInitResultTupleSlot(EState *estate, PlanState *planstate) { (void) estate; (void) planstate; }
void ExecAssignResultTypeFromTL(PlanState *planstate)
{
    int         natts = list_length(planstate->plan->targetlist), i = 0;
    Oid        *types = palloc(sizeof(Oid) * (natts ? natts : 1));
    ListCell   *lc;

    foreach (lc, planstate->plan->targetlist)
        types[i++] = stub_expr_type(((TargetEntry *) lfirst(lc))->expr);
    planstate->ps_ResultTupleSlot = make_slot(tupdesc_from_types(natts, types));
}
TupleTableSlot *ExecClearTuple(TupleTableSlot *slot) { slot->tts_isempty = true; return slot; }
TupleTableSlot *ExecStoreVirtualTuple(TupleTableSlot *slot) { slot->tts_isempty = false; return slot; }
void slot_getallattrs(TupleTableSlot *slot) { (void) slot; }

static char stub_explain[8192];
void ExplainPropertyText(const char *qlabel, const char *value, ExplainState *es)
{
    size_t n = strlen(stub_explain);
    (void) es;
    snprintf(stub_explain + n, sizeof(stub_explain) - n, "%s: %s\n", qlabel, value);
}
const char *pg_stub_explain_text(void) { return stub_explain; }
void pg_stub_explain_reset(void) { stub_explain[0] = '\0'; }
