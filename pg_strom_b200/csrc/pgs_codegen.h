/*
 * pgs_codegen.h - expression tree -> CUDA device function text.
 *
 * Host-side mirror of the reference's codegen.c: same catalogues (device
 * types codegen.c:46-78, device functions codegen.c:211-630), same walker
 * (codegen.c:1065-1392), same declaration emitters (codegen.c:1435-1623) and
 * the same "is this expression device-runnable" test (codegen.c:1631-1759,
 * extended to BoolExpr / CaseExpr, see SURVEY.md section 2 defects), but the
 * emitted text is CUDA C++ for NVRTC instead of OpenCL C.
 *
 * Expression nodes are JSON objects {"node": "<NodeTag>", ...}; field names
 * follow PostgreSQL's primnodes.h (see INTEGRATION.md for the full list).
 */
#ifndef PGS_CODEGEN_H
#define PGS_CODEGEN_H

#include <set>
#include <string>
#include <vector>
#include "pgs_json.h"

namespace pgs {

/* flags of types / functions: which device library the program needs
 * (pg_strom.h:263-274 DEVFUNC_NEEDS_*) */
#define DEVFUNC_NEEDS_TIMELIB       0x0001
#define DEVFUNC_NEEDS_TEXTLIB       0x0002
#define DEVFUNC_NEEDS_NUMERIC       0x0004
#define DEVFUNC_NEEDS_MATHLIB       0x0008
#define DEVFUNC_INCL_FLAGS          0x000f
#define DEVKERNEL_NEEDS_GPUPREAGG   0x0100
#define DEVTYPE_IS_VARLENA          0x1000

struct DevType
{
    int         type_oid;
    const char *type_name;      /* pg_type.typname: pg_<type_name>_t */
    const char *type_base;      /* C type on the device */
    int         type_length;    /* typlen, -1 = varlena */
    bool        type_byval;
    int         type_align;     /* bytes */
    const char *sql_name;       /* format_type() spelling, for EXPLAIN */
    const char *type_eqfunc;
    const char *type_cmpfunc;
    int         type_flags;
};

const DevType *devtype_lookup(const std::string &type_name);
const DevType *devtype_lookup_by_oid(int type_oid);

struct DevFunc
{
    std::string func_name;
    std::vector<const DevType *> func_args;
    const DevType *func_rettype;
    std::string func_alias;     /* pgfn_<func_alias>(errcode, ...) */
    std::string func_decl;      /* emitted declaration, empty if built in */
    int         func_flags;
};

const DevFunc *devfunc_lookup(const std::string &func_name,
                              const std::vector<std::string> &argtypes,
                              const std::string &rettype);

struct CodegenContext
{
    std::vector<const DevType *>  type_defs;
    std::vector<const DevFunc *>  func_defs;
    std::vector<JsonPtr>          used_params;  /* Const / Param nodes */
    std::vector<JsonPtr>          used_vars;
    std::set<int>                 param_refs;
    const char *var_label = "KVAR";
    const char *kds_label = "kds";
    const char *ktoast_label = "ktoast";
    const char *kds_index_label = "kds_index";
    int         extra_flags = 0;

    void track_type(const DevType *t);
    void track_func(const DevFunc *f);
};

/* exprType() */
std::string expr_type(const JsonPtr &node);
/* equal() */
bool expr_equal(const JsonPtr &a, const JsonPtr &b);

/* returns "" and sets *ok=false if the expression is not device runnable */
std::string codegen_expression(const JsonPtr &expr, CodegenContext &ctx, bool *ok);
/* a List of quals is AND-ed (codegen.c:1411-1417) */
JsonPtr make_andclause(const std::vector<JsonPtr> &quals);
bool codegen_available_expression(const JsonPtr &expr);

std::string codegen_func_declarations(const CodegenContext &ctx);
std::string codegen_param_declarations(const CodegenContext &ctx,
                                       const std::set<int> &param_refs);
std::string codegen_var_declarations(const CodegenContext &ctx);

/* kern_parambuf image from used_params (datastore.c:41-148) */
std::vector<unsigned char> create_kern_parambuf(const std::vector<JsonPtr> &used_params);

/* SQL-ish text of an expression as EXPLAIN VERBOSE prints it (ruleutils.c
 * behaviour restated for the node kinds above); colnames[varattno-1] */
std::string deparse_expression(const JsonPtr &node,
                               const std::vector<std::string> &colnames,
                               bool toplevel_parens);

}   /* namespace pgs */
#endif  /* PGS_CODEGEN_H */
