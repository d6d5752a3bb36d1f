/*
 * TEST STAND-IN for the PostgreSQL headers pg_glue/gpupreagg_glue.c includes
 * (nodes/, executor/, optimizer/planner.h, utils/guc.h, utils/lsyscache.h,
 * commands/explain.h ... of the 9.5devel tree with the CustomPlan interface
 * the reference is written against).  Only what that file uses; the one-line
 * headers beside this one all include it.  Not PostgreSQL code, not shipped:
 * the real extension is built against the real tree.  Catalog look-ups are
 * answered from the small tables of pg_glue_stub.c.
 */
#ifndef PG_NODES_STUB_H
#define PG_NODES_STUB_H
#include "postgres.h"
#include <stdarg.h>

typedef int16_t         int16;
typedef int16_t         AttrNumber;
typedef unsigned int    Index;
typedef uint32_t        uint32;
typedef size_t          Size;
typedef double          Cost;
#define InvalidOid      ((Oid) 0)
#define OidIsValid(o)   ((o) != InvalidOid)
#define MAXALIGN(x)     (((uintptr_t) (x) + 7) & ~(uintptr_t) 7)
#define Assert(x)       ((void) 0)
#define NOTICE          18
#define PG_MODULE_MAGIC extern int pg_stub_module_magic
#define ERRCODE_OBJECT_NOT_IN_PREREQUISITE_STATE 0x55000
#define ERRCODE_INTERNAL_ERROR                   0x58000

/* ereport outside fmgr functions: record and longjmp-free "raise" */
extern void pg_stub_raise(int level, const char *fmt, ...);
#undef elog
#undef ereport
#define elog(lev, ...)      pg_stub_raise((lev), __VA_ARGS__)
#define ereport(lev, rest)  do { rest; pg_stub_raise((lev), "%s", pg_stub_error_message); } while (0)

extern void *palloc(Size n);
extern void *palloc0(Size n);
extern char *pstrdup(const char *s);
extern void  pfree(void *p);
extern void *repalloc(void *p, Size n);

/* ---- nodes ---- */
typedef enum NodeTag
{
    T_Invalid = 0, T_List, T_IntList, T_String,
    T_Plan, T_Result, T_SeqScan, T_Agg, T_Sort, T_HashJoin, T_CustomPlan,
    T_PlanState, T_CustomPlanState, T_PlannedStmt, T_RangeTblEntry, T_Alias,
    T_TargetEntry, T_Var, T_Const, T_Param, T_FuncExpr, T_OpExpr, T_DistinctExpr,
    T_BoolExpr, T_NullTest, T_BooleanTest, T_RelabelType, T_CaseExpr, T_CaseWhen,
    T_Aggref
} NodeTag;
typedef struct Node { NodeTag type; } Node;
#define nodeTag(n)      (((const Node *) (n))->type)
#define IsA(n, t)       (nodeTag(n) == T_##t)
#define NodeSetTag(n,t) (((Node *) (n))->type = (t))
extern Node *pg_stub_new_node(Size size, NodeTag tag);
#define makeNode(t)     ((t *) pg_stub_new_node(sizeof(t), T_##t))

typedef struct ListCell { union { void *ptr_value; int int_value; } data; struct ListCell *next; } ListCell;
typedef struct List { NodeTag type; int length; ListCell *head, *tail; } List;
#define NIL                 ((List *) NULL)
#define lfirst(lc)          ((lc)->data.ptr_value)
#define lfirst_int(lc)      ((lc)->data.int_value)
#define list_head(l)        ((l) ? (l)->head : NULL)
#define lnext(lc)           ((lc)->next)
#define foreach(cell, l)    for ((cell) = list_head(l); (cell) != NULL; (cell) = lnext(cell))
#define linitial(l)         lfirst(list_head(l))
extern List *lappend(List *list, void *datum);
extern List *lappend_int(List *list, int datum);
extern int   list_length(const List *list);
extern void *list_nth(const List *list, int n);
#define list_make1(x)       lappend(NIL, (x))
#define list_make2(x, y)    lappend(list_make1(x), (y))
typedef struct Value { NodeTag type; char *str; } Value;
extern Value *makeString(char *str);
#define strVal(v)           (((Value *) (v))->str)

typedef struct StringInfoData { char *data; int len; int maxlen; } StringInfoData, *StringInfo;
extern void initStringInfo(StringInfo str);
extern void appendStringInfo(StringInfo str, const char *fmt, ...);
extern void appendStringInfoString(StringInfo str, const char *s);
extern void appendStringInfoChar(StringInfo str, char ch);

/* ---- expressions ---- */
#define OUTER_VAR   65001
typedef struct Expr { NodeTag type; } Expr;
typedef struct Var { Expr xpr; Index varno; AttrNumber varattno; Oid vartype; int32 vartypmod; Oid varcollid; } Var;
typedef struct Const { Expr xpr; Oid consttype; int32 consttypmod; Oid constcollid; int constlen;
                       Datum constvalue; bool constisnull; bool constbyval; } Const;
typedef struct Param { Expr xpr; int paramkind; int paramid; Oid paramtype; } Param;
typedef enum CoercionForm { COERCE_EXPLICIT_CALL, COERCE_EXPLICIT_CAST, COERCE_IMPLICIT_CAST } CoercionForm;
typedef struct FuncExpr { Expr xpr; Oid funcid; Oid funcresulttype; bool funcretset; CoercionForm funcformat;
                          Oid funccollid; Oid inputcollid; List *args; } FuncExpr;
typedef struct OpExpr { Expr xpr; Oid opno; Oid opfuncid; Oid opresulttype; bool opretset; Oid opcollid;
                        Oid inputcollid; List *args; } OpExpr;
typedef OpExpr DistinctExpr;
typedef enum BoolExprType { AND_EXPR, OR_EXPR, NOT_EXPR } BoolExprType;
typedef struct BoolExpr { Expr xpr; BoolExprType boolop; List *args; } BoolExpr;
typedef enum NullTestType { IS_NULL, IS_NOT_NULL } NullTestType;
typedef struct NullTest { Expr xpr; Expr *arg; NullTestType nulltesttype; bool argisrow; } NullTest;
typedef enum BoolTestType { IS_TRUE, IS_NOT_TRUE, IS_FALSE, IS_NOT_FALSE, IS_UNKNOWN, IS_NOT_UNKNOWN } BoolTestType;
typedef struct BooleanTest { Expr xpr; Expr *arg; BoolTestType booltesttype; } BooleanTest;
typedef struct RelabelType { Expr xpr; Expr *arg; Oid resulttype; int32 resulttypmod; Oid resultcollid;
                             CoercionForm relabelformat; } RelabelType;
typedef struct CaseExpr { Expr xpr; Oid casetype; Oid casecollid; Expr *arg; List *args; Expr *defresult; } CaseExpr;
typedef struct CaseWhen { Expr xpr; Expr *expr; Expr *result; } CaseWhen;
typedef struct Aggref { Expr xpr; Oid aggfnoid; Oid aggtype; Oid aggcollid; Oid inputcollid; List *aggdirectargs;
                        List *args; List *aggorder; List *aggdistinct; Expr *aggfilter; bool aggstar;
                        bool aggvariadic; char aggkind; Index agglevelsup; } Aggref;
typedef struct TargetEntry { Expr xpr; Expr *expr; AttrNumber resno; char *resname; Index ressortgroupref;
                             Oid resorigtbl; AttrNumber resorigcol; bool resjunk; } TargetEntry;
extern TargetEntry *makeTargetEntry(Expr *expr, AttrNumber resno, char *resname, bool resjunk);
extern Var *makeVar(Index varno, AttrNumber varattno, Oid vartype, int32 vartypmod, Oid varcollid, Index varlevelsup);
extern Const *makeNullConst(Oid consttype, int32 consttypmod, Oid constcollid);

/* ---- plans ---- */
typedef struct Plan { NodeTag type; Cost startup_cost, total_cost; double plan_rows; int plan_width;
                      List *targetlist; List *qual; struct Plan *lefttree, *righttree; } Plan;
#define outerPlan(node)     (((Plan *) (node))->lefttree)
#define innerPlan(node)     (((Plan *) (node))->righttree)
typedef struct Scan { Plan plan; Index scanrelid; } Scan;
typedef Scan SeqScan;
typedef enum AggStrategy { AGG_PLAIN, AGG_SORTED, AGG_HASHED } AggStrategy;
typedef struct Agg { Plan plan; AggStrategy aggstrategy; int numCols; AttrNumber *grpColIdx; Oid *grpOperators;
                     long numGroups; } Agg;
typedef struct Sort { Plan plan; int numCols; AttrNumber *sortColIdx; Oid *sortOperators; Oid *collations;
                      bool *nullsFirst; } Sort;
typedef struct Alias { NodeTag type; char *aliasname; List *colnames; } Alias;
typedef struct RangeTblEntry { NodeTag type; int rtekind; Oid relid; Alias *eref; } RangeTblEntry;
typedef struct PlannedStmt { NodeTag type; Plan *planTree; List *rtable; List *subplans; } PlannedStmt;
#define rt_fetch(idx, rtable)   ((RangeTblEntry *) list_nth(rtable, (idx) - 1))

typedef struct tupleDesc { int natts; Oid *atttypid; int16 *attlen; char *attalign; bool *attbyval; } *TupleDesc;
typedef struct TupleTableSlot { NodeTag type; bool tts_isempty; TupleDesc tts_tupleDescriptor;
                                Datum *tts_values; bool *tts_isnull; } TupleTableSlot;
typedef struct EState { NodeTag type; List *es_range_table; PlannedStmt *es_plannedstmt; } EState;
typedef struct PlanState { NodeTag type; Plan *plan; EState *state; List *targetlist; List *qual;
                           struct PlanState *lefttree, *righttree; TupleTableSlot *ps_ResultTupleSlot; } PlanState;
#define outerPlanState(node)    (((PlanState *) (node))->lefttree)
#define TupIsNull(slot)         ((slot) == NULL || (slot)->tts_isempty)
extern PlanState *ExecInitNode(Plan *node, EState *estate, int eflags);
extern TupleTableSlot *ExecProcNode(PlanState *node);
extern void ExecEndNode(PlanState *node);
extern void ExecReScan(PlanState *node);
extern void ExecInitResultTupleSlot(EState *estate, PlanState *planstate);
extern void ExecAssignResultTypeFromTL(PlanState *planstate);
extern TupleTableSlot *ExecClearTuple(TupleTableSlot *slot);
extern TupleTableSlot *ExecStoreVirtualTuple(TupleTableSlot *slot);
extern TupleDesc ExecGetResultType(PlanState *planstate);
extern void slot_getallattrs(TupleTableSlot *slot);

typedef struct Bitmapset Bitmapset;
typedef struct ExplainState { StringInfo str; bool verbose; bool analyze; int indent; } ExplainState;
extern void ExplainPropertyText(const char *qlabel, const char *value, ExplainState *es);

struct CustomPlan;
struct CustomPlanState;
typedef struct CustomPlanMethods
{
    const char *CustomName;
    struct CustomPlanState *(*BeginCustomPlan)(struct CustomPlan *cplan, EState *estate, int eflags);
    TupleTableSlot *(*ExecCustomPlan)(struct CustomPlanState *node);
    void        (*EndCustomPlan)(struct CustomPlanState *node);
    void        (*ReScanCustomPlan)(struct CustomPlanState *node);
    void        (*ExplainCustomPlan)(struct CustomPlanState *node, List *ancestors, ExplainState *es);
    Bitmapset  *(*GetRelidsCustomPlan)(struct CustomPlanState *node);
    void        (*TextOutCustomPlan)(StringInfo str, const struct CustomPlan *node);
    struct CustomPlan *(*CopyCustomPlan)(const struct CustomPlan *from);
} CustomPlanMethods;
typedef struct CustomPlan { Plan plan; const CustomPlanMethods *methods; } CustomPlan;
typedef struct CustomPlanState { PlanState ps; const CustomPlanMethods *methods; } CustomPlanState;

/* ---- optimizer/cost.h ---- */
typedef struct Path { NodeTag type; Cost startup_cost, total_cost; double rows; } Path;
typedef struct PlannerInfo PlannerInfo;
typedef struct AggClauseCosts AggClauseCosts;
extern int    work_mem;
extern double cpu_tuple_cost, cpu_operator_cost;
extern void cost_sort(Path *path, PlannerInfo *root, List *pathkeys, Cost input_cost, double tuples, int width,
                      Cost comparison_cost, int sort_mem, double limit_tuples);
extern void cost_agg(Path *path, PlannerInfo *root, AggStrategy aggstrategy, const AggClauseCosts *aggcosts,
                     int numGroupCols, double numGroups, Cost input_startup_cost, Cost input_total_cost,
                     double input_tuples);

/* ---- planner hook, GUC, miscadmin ---- */
typedef struct Query Query;
typedef struct ParamListInfoData *ParamListInfo;
typedef PlannedStmt *(*planner_hook_type)(Query *parse, int cursorOptions, ParamListInfo boundParams);
extern planner_hook_type planner_hook;
extern PlannedStmt *standard_planner(Query *parse, int cursorOptions, ParamListInfo boundParams);
extern bool process_shared_preload_libraries_in_progress;
typedef enum { PGC_POSTMASTER, PGC_SIGHUP, PGC_SUSET, PGC_USERSET } GucContext;
#define GUC_NOT_IN_SAMPLE   0x0020
typedef void (*GucBoolAssignHook)(bool newval, void *extra);
typedef void (*GucIntAssignHook)(int newval, void *extra);
typedef void (*GucRealAssignHook)(double newval, void *extra);
extern void DefineCustomBoolVariable(const char *name, const char *short_desc, const char *long_desc,
                                     bool *valueAddr, bool bootValue, GucContext context, int flags,
                                     void *check_hook, GucBoolAssignHook assign_hook, void *show_hook);
extern void DefineCustomIntVariable(const char *name, const char *short_desc, const char *long_desc,
                                    int *valueAddr, int bootValue, int minValue, int maxValue,
                                    GucContext context, int flags,
                                    void *check_hook, GucIntAssignHook assign_hook, void *show_hook);
extern void DefineCustomRealVariable(const char *name, const char *short_desc, const char *long_desc,
                                     double *valueAddr, double bootValue, double minValue, double maxValue,
                                     GucContext context, int flags,
                                     void *check_hook, GucRealAssignHook assign_hook, void *show_hook);
#define CHECK_FOR_INTERRUPTS()  ((void) 0)
typedef enum { RESOURCE_RELEASE_BEFORE_LOCKS, RESOURCE_RELEASE_LOCKS, RESOURCE_RELEASE_AFTER_LOCKS } ResourceReleasePhase;
typedef void (*ResourceReleaseCallback)(ResourceReleasePhase phase, bool isCommit, bool isTopLevel, void *arg);
extern void RegisterResourceReleaseCallback(ResourceReleaseCallback callback, void *arg);

/* ---- catalog look-ups (utils/lsyscache.h, parser/parse_func.h, catalog/namespace.h) ---- */
extern char *get_rel_name(Oid relid);
extern Oid   get_rel_namespace(Oid relid);
extern char *get_namespace_name(Oid nspid);
extern char *get_func_name(Oid funcid);
extern Oid   get_func_namespace(Oid funcid);
extern char *get_opname(Oid opno);
extern Oid   get_opcode(Oid opno);
extern char *get_collation_name(Oid colloid);
extern char *pg_stub_type_name(Oid typid);      /* pg_type.typname (the real glue reads the syscache) */
extern void  get_typlenbyvalalign(Oid typid, int16 *typlen, bool *typbyval, char *typalign);
extern void  getTypeOutputInfo(Oid type, Oid *typOutput, bool *typIsVarlena);
extern char *OidOutputFunctionCall(Oid functionId, Datum val);
extern void  getTypeInputInfo(Oid type, Oid *typInput, Oid *typIOParam);
extern Datum OidInputFunctionCall(Oid functionId, char *str, Oid typioparam, int32 typmod);
extern Oid   TypenameGetTypid(const char *typname);
extern Oid   LookupFuncName(List *funcname, int nargs, const Oid *argtypes, bool noError);
extern Oid   OpernameGetOprid(List *names, Oid oprleft, Oid oprright);
extern bool  lc_collate_is_c(Oid collation);
#define DEFAULT_COLLATION_OID   100
#define BOOLOID     16
#define INT4OID     23
#endif
