/*
 * codegen.cpp - expression tree -> CUDA device function text.
 * See pgs_codegen.h for what it mirrors in the reference (codegen.c).
 */
#include <cmath>
#include <cstdint>
#include <sstream>
#include "pgs_codegen.h"
#include "../../include/pgstrom_kds.h"

namespace pgs {

/* ------------------------------------------------------------------
 * Catalog of data types supported by device code (codegen.c:46-78).
 * naming convention of types: pg_<type_name>_t
 * ------------------------------------------------------------------ */
static const DevType devtype_catalog[] = {
    /* basic datatypes */
    { BOOLOID,      "bool",      "cl_bool",   1, true, 1, "boolean",
      "booleq",     "btboolcmp",   0 },
    { INT2OID,      "int2",      "cl_short",  2, true, 2, "smallint",
      "int2eq",     "btint2cmp",   0 },
    { INT4OID,      "int4",      "cl_int",    4, true, 4, "integer",
      "int4eq",     "btint4cmp",   0 },
    { INT8OID,      "int8",      "cl_long",   8, true, 8, "bigint",
      "int8eq",     "btint8cmp",   0 },
    { FLOAT4OID,    "float4",    "cl_float",  4, true, 4, "real",
      "float4eq",   "btfloat4cmp", 0 },
    { FLOAT8OID,    "float8",    "cl_double", 8, true, 8, "double precision",
      "float8eq",   "btfloat8cmp", 0 },
    /* date and time datatypes (HAVE_INT64_TIMESTAMP) */
    { DATEOID,      "date",      "cl_int",    4, true, 4, "date",
      "date_eq",    "date_cmp",    DEVFUNC_NEEDS_TIMELIB },
    { TIMEOID,      "time",      "cl_long",   8, true, 8, "time without time zone",
      "time_eq",    "time_cmp",    DEVFUNC_NEEDS_TIMELIB },
    { TIMESTAMPOID, "timestamp", "cl_long",   8, true, 8, "timestamp without time zone",
      "timestamp_eq", "timestamp_cmp", DEVFUNC_NEEDS_TIMELIB },
    /* variable length datatypes */
    { BPCHAROID,    "bpchar",    "varlena",  -1, false, 4, "character",
      "bpchareq",   "bpcharcmp",   DEVFUNC_NEEDS_TEXTLIB | DEVTYPE_IS_VARLENA },
    { NUMERICOID,   "numeric",   "varlena",  -1, false, 4, "numeric",
      "numeric_eq", "numeric_cmp", DEVFUNC_NEEDS_NUMERIC | DEVTYPE_IS_VARLENA },
    { BYTEAOID,     "bytea",     "varlena",  -1, false, 4, "bytea",
      "byteaeq",    "byteacmp",    DEVTYPE_IS_VARLENA },
    { TEXTOID,      "text",      "varlena",  -1, false, 4, "text",
      "texteq",     "bttextcmp",   DEVFUNC_NEEDS_TEXTLIB | DEVTYPE_IS_VARLENA },
};
#define lengthof(a) (sizeof(a) / sizeof((a)[0]))

const DevType *
devtype_lookup(const std::string &type_name)
{
    for (size_t i = 0; i < lengthof(devtype_catalog); i++)
        if (type_name == devtype_catalog[i].type_name)
            return &devtype_catalog[i];
    /* SQL spellings are accepted too */
    for (size_t i = 0; i < lengthof(devtype_catalog); i++)
        if (type_name == devtype_catalog[i].sql_name)
            return &devtype_catalog[i];
    return NULL;
}

const DevType *
devtype_lookup_by_oid(int type_oid)
{
    for (size_t i = 0; i < lengthof(devtype_catalog); i++)
        if (type_oid == devtype_catalog[i].type_oid)
            return &devtype_catalog[i];
    return NULL;
}

/* ------------------------------------------------------------------
 * Catalog of functions supported by device code (codegen.c:211-630).
 *
 * func_template: [<attributes>/](c|l|b|B|f|F):<extra>
 *  'c' : type cast, implemented in kern_common.cuh as pgfn_<arg>_<ret>
 *        (with PostgreSQL's range checks)
 *  'l' : left unary operator, emitted inline
 *  'b' : binary operator, emitted inline
 *  'B' : binary comparison through devfunc_float_comp (NaN ordering)
 *  'f' : wraps a CUDA built-in function, emitted inline
 *  'F' : implemented in the device runtime under the name <extra>
 * attributes: a = needs an alias, n/m/s/t = needs numeric/math/text/time lib
 * ------------------------------------------------------------------ */
struct devfunc_catalog_t {
    std::string func_name;
    std::vector<std::string> func_argtypes;
    std::string func_template;
};

static std::vector<devfunc_catalog_t> devfunc_catalog;
static std::vector<DevFunc *> devfunc_cache;

static void
cat(const std::string &name, std::vector<std::string> args, const std::string &tmpl)
{
    devfunc_catalog_t e;
    e.func_name = name;
    e.func_argtypes = args;
    e.func_template = tmpl;
    devfunc_catalog.push_back(e);
}

static void
build_devfunc_catalog(void)
{
    if (!devfunc_catalog.empty())
        return;
    static const char *ints[] = {"int2", "int4", "int8"};
    static const char *nums[] = {"int2", "int4", "int8", "float4", "float8"};

    /* Type cast functions: <rettype>(<argtype>) */
    for (const char *r : nums)
        for (const char *a : nums)
            if (std::string(r) != a)
                cat(r, {a}, "a/c:");
    cat("int4", {"bool"}, "a/c:");

    /* arithmetic operators; PostgreSQL names: int4pl, int24pl, float48pl ... */
    struct { const char *sfx; } ops[] = {{"pl"}, {"mi"}, {"mul"}, {"div"}};
    for (auto &op : ops)
    {
        for (const char *a : ints)
            for (const char *b : ints)
            {
                std::string n = std::string("int") + (a + 3) +
                    (std::string(a) == b ? "" : std::string(b + 3)) + op.sfx;
                cat(n, {a, b}, "m/F:" + n);
            }
        cat(std::string("float4") + op.sfx, {"float4", "float4"}, std::string("m/F:float4") + op.sfx);
        cat(std::string("float48") + op.sfx, {"float4", "float8"}, std::string("m/F:float48") + op.sfx);
        cat(std::string("float84") + op.sfx, {"float8", "float4"}, std::string("m/F:float84") + op.sfx);
        cat(std::string("float8") + op.sfx, {"float8", "float8"}, std::string("m/F:float8") + op.sfx);
    }
    /* '%' : reminder operators */
    cat("int2mod", {"int2", "int2"}, "m/F:int2mod");
    cat("int4mod", {"int4", "int4"}, "m/F:int4mod");
    cat("int8mod", {"int8", "int8"}, "m/F:int8mod");
    /* unary plus / minus / abs */
    for (const char *a : nums)
        cat(std::string(a) + "up", {a}, "l:+");
    for (const char *a : ints)
        cat(std::string(a) + "um", {a}, std::string("m/F:") + a + "um");
    cat("float4um", {"float4"}, "l:-");
    cat("float8um", {"float8"}, "l:-");
    for (const char *a : ints)
    {
        cat(std::string(a) + "abs", {a}, std::string("m/F:") + a + "abs");
        cat("abs", {a}, std::string("m/F:") + a + "abs");
    }
    cat("float4abs", {"float4"}, "f:fabsf");
    cat("float8abs", {"float8"}, "f:fabs");
    cat("abs", {"float4"}, "af:fabsf");
    cat("abs", {"float8"}, "af:fabs");

    /* comparison operators */
    struct { const char *sfx; const char *op; } cmps[] = {
        {"eq", "=="}, {"ne", "!="}, {"gt", ">"}, {"lt", "<"}, {"ge", ">="}, {"le", "<="}};
    for (auto &c : cmps)
    {
        for (const char *a : ints)
            for (const char *b : ints)
            {
                std::string n = std::string("int") + (a + 3) +
                    (std::string(a) == b ? "" : std::string(b + 3)) + c.sfx;
                cat(n, {a, b}, std::string("b:") + c.op);
            }
        cat(std::string("float4") + c.sfx, {"float4", "float4"}, std::string("B:") + c.op);
        cat(std::string("float48") + c.sfx, {"float4", "float8"}, std::string("B:") + c.op);
        cat(std::string("float84") + c.sfx, {"float8", "float4"}, std::string("B:") + c.op);
        cat(std::string("float8") + c.sfx, {"float8", "float8"}, std::string("B:") + c.op);
        cat(std::string("bool") + c.sfx, {"bool", "bool"}, std::string("b:") + c.op);
        cat(std::string("date_") + c.sfx, {"date", "date"}, std::string("t/b:") + c.op);
        cat(std::string("time_") + c.sfx, {"time", "time"}, std::string("t/b:") + c.op);
        cat(std::string("timestamp_") + c.sfx, {"timestamp", "timestamp"}, std::string("t/b:") + c.op);
    }
    /* bitwise operators */
    for (const char *a : ints)
    {
        cat(std::string(a) + "and", {a, a}, "b:&");
        cat(std::string(a) + "or", {a, a}, "b:|");
        cat(std::string(a) + "xor", {a, a}, "b:^");
        cat(std::string(a) + "not", {a}, "l:~");
        cat(std::string(a) + "shr", {a, "int4"}, "b:>>");
        cat(std::string(a) + "shl", {a, "int4"}, "b:<<");
    }
    /* comparison functions */
    cat("btboolcmp", {"bool", "bool"}, "f:devfunc_int_comp");
    for (const char *a : ints)
        for (const char *b : ints)
        {
            std::string n = std::string("btint") + (a + 3) +
                (std::string(a) == b ? "" : std::string(b + 3)) + "cmp";
            cat(n, {a, b}, "f:devfunc_int_comp");
        }
    cat("btfloat4cmp", {"float4", "float4"}, "f:devfunc_float_comp");
    cat("btfloat48cmp", {"float4", "float8"}, "f:devfunc_float_comp");
    cat("btfloat84cmp", {"float8", "float4"}, "f:devfunc_float_comp");
    cat("btfloat8cmp", {"float8", "float8"}, "f:devfunc_float_comp");
    cat("date_cmp", {"date", "date"}, "t/f:devfunc_int_comp");
    cat("time_cmp", {"time", "time"}, "t/f:devfunc_int_comp");
    cat("timestamp_cmp", {"timestamp", "timestamp"}, "t/f:devfunc_int_comp");

    /* Mathmatical functions */
    cat("cbrt", {"float8"}, "f:cbrt");
    cat("dcbrt", {"float8"}, "f:cbrt");
    cat("ceil", {"float8"}, "f:ceil");
    cat("ceiling", {"float8"}, "f:ceil");
    cat("exp", {"float8"}, "f:exp");
    cat("dexp", {"float8"}, "f:exp");
    cat("floor", {"float8"}, "f:floor");
    cat("ln", {"float8"}, "f:log");
    cat("dlog1", {"float8"}, "f:log");
    cat("log", {"float8"}, "f:log10");
    cat("dlog10", {"float8"}, "f:log10");
    cat("pi", {}, "m/F:dpi");
    cat("sign", {"float8"}, "m/F:dsign");
    cat("degrees", {"float8"}, "m/F:degrees");
    cat("radians", {"float8"}, "m/F:radians");
    cat("power", {"float8", "float8"}, "m/F:dpow");
    cat("pow", {"float8", "float8"}, "m/F:dpow");
    cat("dpow", {"float8", "float8"}, "m/F:dpow");
    cat("round", {"float8"}, "f:rint");
    cat("dround", {"float8"}, "f:rint");
    cat("sqrt", {"float8"}, "f:sqrt");
    cat("dsqrt", {"float8"}, "f:sqrt");
    cat("trunc", {"float8"}, "f:trunc");
    cat("dtrunc", {"float8"}, "f:trunc");
    /* Trigonometric function */
    cat("acos", {"float8"}, "f:acos");
    cat("asin", {"float8"}, "f:asin");
    cat("atan", {"float8"}, "f:atan");
    cat("atan2", {"float8", "float8"}, "f:atan2");
    cat("cos", {"float8"}, "f:cos");
    cat("sin", {"float8"}, "f:sin");
    cat("tan", {"float8"}, "f:tan");

    /* Numeric functions (kern_numeric.cuh) */
    cat("int2", {"numeric"}, "n/F:numeric_int2");
    cat("int4", {"numeric"}, "n/F:numeric_int4");
    cat("int8", {"numeric"}, "n/F:numeric_int8");
    cat("float4", {"numeric"}, "n/F:numeric_float4");
    cat("float8", {"numeric"}, "n/F:numeric_float8");
    cat("numeric", {"int2"}, "n/F:int2_numeric");
    cat("numeric", {"int4"}, "n/F:int4_numeric");
    cat("numeric", {"int8"}, "n/F:int8_numeric");
    cat("numeric", {"float4"}, "n/F:float4_numeric");
    cat("numeric", {"float8"}, "n/F:float8_numeric");
    cat("numeric_add", {"numeric", "numeric"}, "n/F:numeric_add");
    cat("numeric_sub", {"numeric", "numeric"}, "n/F:numeric_sub");
    cat("numeric_mul", {"numeric", "numeric"}, "n/F:numeric_mul");
    cat("numeric_uplus", {"numeric"}, "n/F:numeric_uplus");
    cat("numeric_uminus", {"numeric"}, "n/F:numeric_uminus");
    cat("numeric_abs", {"numeric"}, "n/F:numeric_abs");
    cat("abs", {"numeric"}, "n/F:numeric_abs");
    cat("numeric_eq", {"numeric", "numeric"}, "n/F:numeric_eq");
    cat("numeric_ne", {"numeric", "numeric"}, "n/F:numeric_ne");
    cat("numeric_lt", {"numeric", "numeric"}, "n/F:numeric_lt");
    cat("numeric_le", {"numeric", "numeric"}, "n/F:numeric_le");
    cat("numeric_gt", {"numeric", "numeric"}, "n/F:numeric_gt");
    cat("numeric_ge", {"numeric", "numeric"}, "n/F:numeric_ge");
    cat("numeric_cmp", {"numeric", "numeric"}, "n/F:numeric_cmp");

    /* Date and time functions (kern_timelib.cuh) */
    cat("date", {"date"}, "ta/c:");
    cat("date", {"timestamp"}, "t/F:timestamp_date");
    cat("time", {"timestamp"}, "t/F:timestamp_time");
    cat("time", {"time"}, "ta/c:");
    cat("timestamp", {"timestamp"}, "ta/c:");
    cat("timestamp", {"date"}, "t/F:date_timestamp");
    cat("date_pli", {"date", "int4"}, "t/F:date_pli");
    cat("date_mii", {"date", "int4"}, "t/F:date_mii");
    cat("date_mi", {"date", "date"}, "t/F:date_mi");
    cat("datetime_pl", {"date", "time"}, "t/F:datetime_pl");
    cat("integer_pl_date", {"int4", "date"}, "t/F:integer_pl_date");
    cat("timedate_pl", {"time", "date"}, "t/F:timedata_pl");
    for (const char *sfx : {"eq", "ne", "lt", "le", "gt", "ge", "cmp"})
    {
        cat(std::string("date_") + sfx + "_timestamp", {"date", "timestamp"},
            std::string("t/F:date_") + sfx + "_timestamp");
        cat(std::string("timestamp_") + sfx + "_date", {"timestamp", "date"},
            std::string("t/F:timestamp_") + sfx + "_date");
    }

    /* Text functions (kern_textlib.cuh) */
    for (const char *sfx : {"eq", "ne", "lt", "le", "gt", "ge", "cmp"})
        cat(std::string("bpchar") + sfx, {"bpchar", "bpchar"}, std::string("s/F:bpchar") + sfx);
    cat("texteq", {"text", "text"}, "s/F:texteq");
    cat("textne", {"text", "text"}, "s/F:textne");
    cat("text_lt", {"text", "text"}, "s/F:text_lt");
    cat("text_le", {"text", "text"}, "s/F:text_le");
    cat("text_gt", {"text", "text"}, "s/F:text_gt");
    cat("text_ge", {"text", "text"}, "s/F:text_ge");
    cat("bttextcmp", {"text", "text"}, "s/F:text_cmp");
}

/* text ordering on the device is bytewise = the "C" collation; equality does
 * not depend on the collation (kern_textlib.cuh).  "inputcollid" is the name
 * of the operator's input collation as the PostgreSQL glue resolves it
 * ("default" already replaced by the database's lc_collate). */
static bool
collation_is_device_compatible(const std::string &func_name, const JsonPtr &node)
{
    static const char *ordered[] = {
        "text_lt", "text_le", "text_gt", "text_ge", "bttextcmp",
        "bpcharlt", "bpcharle", "bpchargt", "bpcharge", "bpcharcmp" };
    bool is_ordered = false;
    for (const char *n : ordered)
        if (func_name == n)
            is_ordered = true;
    if (!is_ordered)
        return true;
    std::string coll = node->s("inputcollid", "C");
    return coll == "C" || coll == "POSIX";
}

static std::string
argname_suffix(const std::vector<const DevType *> &args)
{
    std::string s;
    for (auto *t : args)
        s += std::string("_") + t->type_name;
    return s;
}

static void
devfunc_setup_oper_both(DevFunc *entry, const std::string &oper, bool has_alias,
                        bool float_comp)
{
    const DevType *a = entry->func_args[0];
    const DevType *b = entry->func_args[1];
    std::ostringstream s;

    entry->func_alias = has_alias
        ? entry->func_name + "_" + a->type_name + "_" + b->type_name
        : entry->func_name;
    s << "DEVFN pg_" << entry->func_rettype->type_name << "_t pgfn_" << entry->func_alias
      << "(cl_int *errcode, pg_" << a->type_name << "_t arg1, pg_" << b->type_name << "_t arg2)\n"
      << "{\n"
      << "    pg_" << entry->func_rettype->type_name << "_t result;\n";
    if (float_comp)
        s << "    result.value = (" << entry->func_rettype->type_base
          << ")(devfunc_float_comp(arg1.value, arg2.value) " << oper << " 0);\n";
    else
        s << "    result.value = (" << entry->func_rettype->type_base
          << ")(arg1.value " << oper << " arg2.value);\n";
    s << "    result.isnull = arg1.isnull | arg2.isnull;\n"
      << "    return result;\n"
      << "}\n";
    entry->func_decl = s.str();
}

static void
devfunc_setup_oper_left(DevFunc *entry, const std::string &oper, bool has_alias)
{
    const DevType *a = entry->func_args[0];
    std::ostringstream s;

    entry->func_alias = has_alias ? entry->func_name + "_" + a->type_name
                                  : entry->func_name;
    s << "DEVFN pg_" << entry->func_rettype->type_name << "_t pgfn_" << entry->func_alias
      << "(cl_int *errcode, pg_" << a->type_name << "_t arg)\n"
      << "{\n"
      << "    pg_" << entry->func_rettype->type_name << "_t result;\n"
      << "    result.value = (" << entry->func_rettype->type_base << ")(" << oper << "arg.value);\n"
      << "    result.isnull = arg.isnull;\n"
      << "    return result;\n"
      << "}\n";
    entry->func_decl = s.str();
}

static void
devfunc_setup_func_decl(DevFunc *entry, const std::string &builtin, bool has_alias)
{
    std::ostringstream s;

    entry->func_alias = has_alias ? entry->func_name + argname_suffix(entry->func_args)
                                  : entry->func_name;
    s << "DEVFN pg_" << entry->func_rettype->type_name << "_t pgfn_" << entry->func_alias
      << "(cl_int *errcode";
    for (size_t i = 0; i < entry->func_args.size(); i++)
        s << ", pg_" << entry->func_args[i]->type_name << "_t arg" << (i + 1);
    s << ")\n{\n"
      << "    pg_" << entry->func_rettype->type_name << "_t result;\n"
      << "    result.value = 0;\n"
      << "    result.isnull = ";
    if (entry->func_args.empty())
        s << "false";
    for (size_t i = 0; i < entry->func_args.size(); i++)
        s << (i ? " | " : "") << "arg" << (i + 1) << ".isnull";
    s << ";\n"
      << "    if (!result.isnull)\n";
    /* PostgreSQL's float.c raises an error outside the function's domain
     * (dsqrt, dlog1, dlog10, dacos, dasin: "cannot take ... / input is out of
     * range"; dcos, dsin, dtan: errno on an infinite input) and on overflow /
     * underflow of dexp (CHECKFLOATVAL).  The CUDA built-ins return NaN / Inf /
     * 0 instead: such rows are left to the host (StromError_CpuReCheck),
     * which raises the error PostgreSQL would have raised. */
    static const struct { const char *builtin, *pre, *post; } domain[] = {
        { "sqrt",  "arg1.value < 0.0", NULL },
        { "log",   "arg1.value <= 0.0", NULL },
        { "log10", "arg1.value <= 0.0", NULL },
        { "acos",  "fabs(arg1.value) > 1.0", NULL },
        { "asin",  "fabs(arg1.value) > 1.0", NULL },
        { "cos",   "isinf(arg1.value)", NULL },
        { "sin",   "isinf(arg1.value)", NULL },
        { "tan",   "isinf(arg1.value)", NULL },
        { "exp",   NULL, "CHECKFLOATVAL(result.value, isinf(arg1.value), false)" },
        { "cbrt",  NULL, "CHECKFLOATVAL(result.value, isinf(arg1.value), arg1.value == 0.0)" },
    };
    const char *pre = NULL, *post = NULL;
    for (auto &d : domain)
        if (builtin == d.builtin)
        { pre = d.pre; post = d.post; }
    s << "    {\n";
    if (pre)
        s << "        if (" << pre << ")\n"
          << "        {\n"
          << "            PGS_MATH_FAIL(result);\n"
          << "            return result;\n"
          << "        }\n";
    s << "        result.value = (" << entry->func_rettype->type_base << ") " << builtin << "(";
    for (size_t i = 0; i < entry->func_args.size(); i++)
        s << (i ? ", " : "") << "arg" << (i + 1) << ".value";
    s << ");\n";
    if (post)
        s << "        if (" << post << ")\n"
          << "            PGS_MATH_FAIL(result);\n";
    s << "    }\n"
      << "    return result;\n"
      << "}\n";
    entry->func_decl = s.str();
}

static DevFunc *
devfunc_setup_boolop(bool is_and, const std::string &fn_name, int fn_nargs)
{
    DevFunc *entry = new DevFunc;
    const DevType *dtype = devtype_lookup("bool");
    std::ostringstream s;

    for (int i = 0; i < fn_nargs; i++)
        entry->func_args.push_back(dtype);
    entry->func_rettype = dtype;
    entry->func_name = fn_name;
    entry->func_alias = fn_name;
    entry->func_flags = 0;
    /* three-valued logic: FALSE AND NULL = FALSE, TRUE OR NULL = TRUE.
     * (The reference ORs the null flags; a NULL result only matters where
     * EVAL() would be false anyway for AND, but not for OR / NOT, so the
     * exact rule is emitted here.) */
    s << "DEVFN pg_bool_t pgfn_" << fn_name << "(cl_int *errcode";
    for (int i = 0; i < fn_nargs; i++)
        s << ", pg_bool_t arg" << (i + 1);
    s << ")\n{\n  pg_bool_t result;\n  bool anynull = ";
    for (int i = 0; i < fn_nargs; i++)
        s << (i ? " | " : "") << "arg" << (i + 1) << ".isnull";
    s << ";\n  bool decided = ";
    for (int i = 0; i < fn_nargs; i++)
        s << (i ? " | " : "") << "(!arg" << (i + 1) << ".isnull && "
          << (is_and ? "!" : "") << "arg" << (i + 1) << ".value)";
    s << ";\n"
      << "  result.value = (cl_bool)(decided ? " << (is_and ? "0" : "1") << " : "
      << (is_and ? "1" : "0") << ");\n"
      << "  result.isnull = (!decided && anynull);\n"
      << "  return result;\n}\n";
    entry->func_decl = s.str();
    return entry;
}

const DevFunc *
devfunc_lookup(const std::string &func_name,
               const std::vector<std::string> &argtypes,
               const std::string &rettype)
{
    build_devfunc_catalog();
    for (DevFunc *f : devfunc_cache)
    {
        if (f->func_name != func_name || f->func_args.size() != argtypes.size())
            continue;
        bool same = true;
        for (size_t i = 0; i < argtypes.size(); i++)
            if (argtypes[i] != f->func_args[i]->type_name)
                same = false;
        if (same)
            return f;
    }
    for (const devfunc_catalog_t &procat : devfunc_catalog)
    {
        if (procat.func_name != func_name ||
            procat.func_argtypes != argtypes)
            continue;
        DevFunc *entry = new DevFunc;
        std::string tmpl = procat.func_template;
        bool has_alias = false;
        int flags = 0;

        entry->func_name = func_name;
        entry->func_rettype = devtype_lookup(rettype);
        if (!entry->func_rettype)
        { delete entry; return NULL; }
        for (const std::string &a : argtypes)
        {
            const DevType *t = devtype_lookup(a);
            if (!t) { delete entry; return NULL; }
            entry->func_args.push_back(t);
        }
        size_t slash = tmpl.find('/');
        if (slash != std::string::npos)
        {
            for (size_t i = 0; i < slash; i++)
            {
                switch (tmpl[i])
                {
                    case 'a': has_alias = true; break;
                    case 'n': flags |= DEVFUNC_NEEDS_NUMERIC; break;
                    case 'm': flags |= DEVFUNC_NEEDS_MATHLIB; break;
                    case 's': flags |= DEVFUNC_NEEDS_TEXTLIB; break;
                    case 't': flags |= DEVFUNC_NEEDS_TIMELIB; break;
                }
            }
            tmpl = tmpl.substr(slash + 1);
        }
        /* "af:" style (attribute without slash) */
        if (tmpl.size() > 2 && tmpl[0] == 'a' && tmpl[2] == ':')
        { has_alias = true; tmpl = tmpl.substr(1); }
        entry->func_flags = flags;
        std::string extra = tmpl.substr(2);
        if (tmpl.compare(0, 2, "c:") == 0)
            entry->func_alias = std::string(entry->func_args[0]->type_name) + "_" +
                entry->func_rettype->type_name;
        else if (tmpl.compare(0, 2, "b:") == 0)
            devfunc_setup_oper_both(entry, extra, has_alias, false);
        else if (tmpl.compare(0, 2, "B:") == 0)
            devfunc_setup_oper_both(entry, extra, has_alias, true);
        else if (tmpl.compare(0, 2, "l:") == 0)
            devfunc_setup_oper_left(entry, extra, has_alias);
        else if (tmpl.compare(0, 2, "f:") == 0)
            devfunc_setup_func_decl(entry, extra, has_alias);
        else if (tmpl.compare(0, 2, "F:") == 0)
            entry->func_alias = extra;
        else
        { delete entry; return NULL; }
        devfunc_cache.push_back(entry);
        return entry;
    }
    return NULL;
}

void
CodegenContext::track_type(const DevType *t)
{
    for (auto *x : type_defs)
        if (x == t)
            return;
    type_defs.push_back(t);
    extra_flags |= (t->type_flags & DEVFUNC_INCL_FLAGS);
}

void
CodegenContext::track_func(const DevFunc *f)
{
    for (auto *x : func_defs)
        if (x == f)
            return;
    func_defs.push_back(f);
    extra_flags |= (f->func_flags & DEVFUNC_INCL_FLAGS);
    track_type(f->func_rettype);
    for (auto *t : f->func_args)
        track_type(t);
}

/* ------------------------------------------------------------------ */
std::string
expr_type(const JsonPtr &node)
{
    if (!node || node->is_null())
        return "";
    std::string tag = node->s("node");
    if (tag == "Const") return node->s("consttype");
    if (tag == "Param") return node->s("paramtype");
    if (tag == "Var") return node->s("vartype");
    if (tag == "FuncExpr") return node->s("funcresulttype");
    if (tag == "OpExpr" || tag == "DistinctExpr") return node->s("opresulttype", "bool");
    if (tag == "NullTest" || tag == "BooleanTest" || tag == "BoolExpr") return "bool";
    if (tag == "RelabelType") return node->s("resulttype");
    if (tag == "CaseExpr") return node->s("casetype");
    if (tag == "Aggref") return node->s("aggtype");
    if (tag == "TargetEntry") return expr_type(node->getp("expr"));
    return "";
}

bool
expr_equal(const JsonPtr &a, const JsonPtr &b)
{
    if (!a || !b)
        return !a && !b;
    return a->dump() == b->dump();
}

static std::vector<std::string>
arg_types(const Json *args)
{
    std::vector<std::string> v;
    if (args)
        for (auto &a : args->arr)
            v.push_back(expr_type(a));
    return v;
}

/* device support of a type is complete (kern_*.cuh has its vref / param /
 * operators)?  Catalogued types without it are treated like unknown types:
 * the expression stays on the host. */
static bool
devtype_is_runnable(const DevType *dtype)
{
    if (!dtype)
        return false;
    if (dtype->type_flags & DEVTYPE_IS_VARLENA)
    {
        std::string n = dtype->type_name;
        /* kern_numeric.cuh, kern_textlib.cuh */
        return n == "numeric" || n == "text" || n == "bpchar";
    }
    return true;
}

static bool
codegen_expression_walker(const JsonPtr &node, CodegenContext &ctx, std::string &out)
{
    if (!node || node->is_null())
        return true;
    std::string tag = node->s("node");
    if (tag != "BoolExpr" && tag != "NullTest" && tag != "BooleanTest" &&
        tag != "CaseWhen" && !devtype_is_runnable(devtype_lookup(expr_type(node))))
        return false;

    if (tag == "Const" || tag == "Param")
    {
        const DevType *dtype = devtype_lookup(expr_type(node));
        if (!dtype)
            return false;
        if (tag == "Param" && node->s("paramkind", "extern") != "extern")
            return false;
        ctx.track_type(dtype);
        size_t index;
        for (index = 0; index < ctx.used_params.size(); index++)
            if (expr_equal(node, ctx.used_params[index]))
                break;
        if (index == ctx.used_params.size())
            ctx.used_params.push_back(node);
        out += "KPARAM_" + std::to_string(index);
        ctx.param_refs.insert((int)index);
        return true;
    }
    if (tag == "Var")
    {
        const DevType *dtype = devtype_lookup(expr_type(node));
        if (!dtype)
            return false;
        ctx.track_type(dtype);
        out += std::string(ctx.var_label) + "_" + std::to_string(node->i("varattno"));
        bool found = false;
        for (auto &v : ctx.used_vars)
            if (expr_equal(v, node))
                found = true;
        if (!found)
            ctx.used_vars.push_back(node);
        return true;
    }
    if (tag == "FuncExpr" || tag == "OpExpr" || tag == "DistinctExpr")
    {
        std::string fname = (tag == "FuncExpr" ? node->s("funcname")
                                               : node->s("opfuncname"));
        const Json *args = node->get("args");
        const DevFunc *dfunc = devfunc_lookup(fname, arg_types(args), expr_type(node));
        if (!dfunc || !collation_is_device_compatible(fname, node))
            return false;
        ctx.track_func(dfunc);
        if (tag == "DistinctExpr")
        {
            /* a IS DISTINCT FROM b: the node carries the equality operator,
             * but the result is its NULL-safe negation - never NULL; two
             * NULLs are not distinct, one NULL is (execQual.c
             * ExecEvalDistinct).  The reference emits the bare operator call
             * here (codegen.c:1171-1195), i.e. evaluates `a = b`. */
            if (!args || args->arr.size() != 2 ||
                std::string(dfunc->func_rettype->type_name) != "bool")
                return false;
            std::string dname = "distinct_" + dfunc->func_alias;
            const DevFunc *wrap = NULL;
            for (DevFunc *f : devfunc_cache)
                if (f->func_name == dname && f->func_args == dfunc->func_args)
                    wrap = f;
            if (!wrap)
            {
                DevFunc *f = new DevFunc;
                std::ostringstream s;

                f->func_name = dname;
                f->func_alias = dname + argname_suffix(dfunc->func_args);
                f->func_args = dfunc->func_args;
                f->func_rettype = dfunc->func_rettype;
                f->func_flags = dfunc->func_flags;
                s << "DEVFN pg_bool_t pgfn_" << f->func_alias << "(cl_int *errcode, pg_"
                  << f->func_args[0]->type_name << "_t arg1, pg_"
                  << f->func_args[1]->type_name << "_t arg2)\n"
                  << "{\n"
                  << "    pg_bool_t result;\n"
                  << "    result.isnull = false;\n"
                  << "    if (arg1.isnull | arg2.isnull)\n"
                  << "        result.value = (cl_bool)(arg1.isnull != arg2.isnull);\n"
                  << "    else\n"
                  << "        result.value = (cl_bool)!EVAL(pgfn_" << dfunc->func_alias
                  << "(errcode, arg1, arg2));\n"
                  << "    return result;\n"
                  << "}\n";
                f->func_decl = s.str();
                devfunc_cache.push_back(f);
                wrap = f;
            }
            ctx.track_func(wrap);       /* after the operator it calls */
            dfunc = wrap;
        }
        out += "pgfn_" + dfunc->func_alias + "(errcode";
        if (args)
            for (auto &a : args->arr)
            {
                out += ", ";
                if (!codegen_expression_walker(a, ctx, out))
                    return false;
            }
        out += ")";
        return true;
    }
    if (tag == "NullTest")
    {
        JsonPtr arg = node->getp("arg");
        const DevType *dtype = devtype_lookup(expr_type(arg));
        if (!dtype || node->flag("argisrow"))
            return false;
        ctx.track_type(dtype);
        std::string t = node->s("nulltesttype");
        const char *func_name = (t == "IS_NULL" ? "isnull" : "isnotnull");
        out += std::string("pgfn_") + dtype->type_name + "_" + func_name + "(errcode, ";
        if (!codegen_expression_walker(arg, ctx, out))
            return false;
        out += ")";
        return true;
    }
    if (tag == "BooleanTest")
    {
        static const char *names[][2] = {
            {"IS_TRUE", "bool_is_true"}, {"IS_NOT_TRUE", "bool_is_not_true"},
            {"IS_FALSE", "bool_is_false"}, {"IS_NOT_FALSE", "bool_is_not_false"},
            {"IS_UNKNOWN", "bool_is_unknown"}, {"IS_NOT_UNKNOWN", "bool_is_not_unknown"}};
        std::string t = node->s("booltesttype");
        const char *func_name = NULL;
        for (auto &n : names)
            if (t == n[0])
                func_name = n[1];
        if (!func_name || expr_type(node->getp("arg")) != "bool")
            return false;
        out += std::string("pgfn_") + func_name + "(errcode, ";
        if (!codegen_expression_walker(node->getp("arg"), ctx, out))
            return false;
        out += ")";
        return true;
    }
    if (tag == "BoolExpr")
    {
        std::string op = node->s("boolop");
        const Json *args = node->get("args");
        if (!args)
            return false;
        if (op == "NOT")
        {
            if (args->arr.size() != 1)
                return false;
            out += "pgfn_boolop_not(errcode, ";
            if (!codegen_expression_walker(args->arr[0], ctx, out))
                return false;
            out += ")";
            return true;
        }
        if (op != "AND" && op != "OR")
            return false;
        int nargs = (int)args->arr.size();
        std::string namebuf = std::string(op == "AND" ? "boolop_and_" : "boolop_or_") +
            std::to_string(nargs);
        /* AND/OR are device only functions without catalog entries */
        const DevFunc *dfunc = NULL;
        for (DevFunc *f : devfunc_cache)
            if (f->func_name == namebuf)
                dfunc = f;
        if (!dfunc)
        {
            DevFunc *f = devfunc_setup_boolop(op == "AND", namebuf, nargs);
            devfunc_cache.push_back(f);
            dfunc = f;
        }
        ctx.track_func(dfunc);
        out += "pgfn_" + dfunc->func_alias + "(errcode";
        for (auto &a : args->arr)
        {
            if (expr_type(a) != "bool")
                return false;
            out += ", ";
            if (!codegen_expression_walker(a, ctx, out))
                return false;
        }
        out += ")";
        return true;
    }
    if (tag == "RelabelType")
    {
        /* both types share the binary form: nothing to do */
        return codegen_expression_walker(node->getp("arg"), ctx, out);
    }
    if (tag == "CaseExpr")
    {
        const Json *args = node->get("args");
        JsonPtr carg = node->getp("arg");
        bool has_arg = (carg && !carg->is_null());
        const DevType *rtype = devtype_lookup(expr_type(node));
        size_t nwhen = args ? args->arr.size() : 0;

        if (!rtype)
            return false;
        ctx.track_type(rtype);
        for (size_t i = 0; i < nwhen; i++)
        {
            const JsonPtr &cw = args->arr[i];
            if (has_arg)
            {
                const DevType *dtype = devtype_lookup(expr_type(carg));
                if (!dtype)
                    return false;
                ctx.track_type(dtype);
                const DevFunc *dfunc = devfunc_lookup(dtype->type_eqfunc,
                    {dtype->type_name, dtype->type_name}, "bool");
                if (!dfunc)
                    return false;
                ctx.track_func(dfunc);
                out += "EVAL(pgfn_" + dfunc->func_alias + "(errcode, ";
                if (!codegen_expression_walker(carg, ctx, out)) return false;
                out += ", ";
                if (!codegen_expression_walker(cw->getp("expr"), ctx, out)) return false;
                out += ")) ? (";
            }
            else
            {
                out += "EVAL(";
                if (!codegen_expression_walker(cw->getp("expr"), ctx, out)) return false;
                out += ") ? (";
            }
            if (!codegen_expression_walker(cw->getp("result"), ctx, out)) return false;
            out += ") : (";
        }
        JsonPtr def = node->getp("defresult");
        if (!def || def->is_null())
            out += std::string("pg_") + rtype->type_name + "_null()";
        else if (!codegen_expression_walker(def, ctx, out))
            return false;
        for (size_t i = 0; i < nwhen; i++)
            out += ")";
        return true;
    }
    return false;
}

JsonPtr
make_andclause(const std::vector<JsonPtr> &quals)
{
    if (quals.size() == 1)
        return quals[0];
    JsonPtr n = Json::object();
    n->set("node", "BoolExpr");
    n->set("boolop", "AND");
    JsonPtr a = Json::array();
    for (auto &q : quals)
        a->push(q);
    n->set("args", a);
    return n;
}

std::string
codegen_expression(const JsonPtr &expr, CodegenContext &ctx, bool *ok)
{
    CodegenContext walker = ctx;    /* commit only on success */
    std::string out;
    JsonPtr node = expr;

    if (node && node->kind == Json::Array)
    {
        std::vector<JsonPtr> quals(node->arr.begin(), node->arr.end());
        node = make_andclause(quals);
    }
    if (!codegen_expression_walker(node, walker, out))
    {
        if (ok) *ok = false;
        return "";
    }
    ctx = walker;
    if (ok) *ok = true;
    return out;
}

bool
codegen_available_expression(const JsonPtr &expr)
{
    if (!expr || expr->is_null())
        return true;
    if (expr->kind == Json::Array)
    {
        for (auto &e : expr->arr)
            if (!codegen_available_expression(e))
                return false;
        return true;
    }
    CodegenContext scratch;
    bool ok = false;
    scratch.used_params.push_back(Json::null());    /* KPARAM_0 placeholder */
    codegen_expression(expr, scratch, &ok);
    return ok;
}

std::string
codegen_func_declarations(const CodegenContext &ctx)
{
    std::string s;
    for (auto *f : ctx.func_defs)
        if (!f->func_decl.empty())
            s += f->func_decl + "\n";
    return s;
}

std::string
codegen_param_declarations(const CodegenContext &ctx, const std::set<int> &param_refs)
{
    std::string s;
    for (size_t index = 0; index < ctx.used_params.size(); index++)
    {
        if (!param_refs.count((int)index))
            continue;
        const DevType *dtype = devtype_lookup(expr_type(ctx.used_params[index]));
        if (!dtype)
            continue;
        s += std::string("  pg_") + dtype->type_name + "_t KPARAM_" + std::to_string(index) +
            " = pg_" + dtype->type_name + "_param(kparams,errcode," + std::to_string(index) + ");\n";
    }
    return s;
}

std::string
codegen_var_declarations(const CodegenContext &ctx)
{
    std::string s;
    for (auto &var : ctx.used_vars)
    {
        const DevType *dtype = devtype_lookup(expr_type(var));
        long long attno = var->i("varattno");
        s += std::string("  pg_") + dtype->type_name + "_t " + ctx.var_label + "_" +
            std::to_string(attno) + " = pg_" + dtype->type_name + "_vref(" +
            ctx.kds_label + "," + ctx.ktoast_label + ",errcode," +
            std::to_string(attno - 1) + "," + ctx.kds_index_label + ");\n";
    }
    return s;
}

/* ------------------------------------------------------------------
 * kern_parambuf
 * ------------------------------------------------------------------ */
static bool
parse_bool_text(const std::string &s)
{
    return s == "t" || s == "true" || s == "TRUE" || s == "1";
}

static double
parse_float_text(const std::string &s)
{
    if (s == "NaN") return NAN;
    if (s == "Infinity" || s == "inf") return INFINITY;
    if (s == "-Infinity" || s == "-inf") return -INFINITY;
    return strtod(s.c_str(), NULL);
}

/* defined in numeric_host.cpp */
bool pgs_numeric_from_text(const std::string &text, std::vector<unsigned char> &varlena);

static void
append_datum(std::vector<unsigned char> &buf, const JsonPtr &node, bool *isnull)
{
    std::string tag = node->s("node");
    const DevType *dtype = devtype_lookup(expr_type(node));
    const Json *jv = node->get(tag == "Const" ? "constvalue" : "value");
    bool null_flag = (tag == "Const" ? node->flag("constisnull") : node->flag("isnull"));

    *isnull = true;
    if (!dtype || null_flag || !jv || jv->is_null())
        return;
    std::string text = (jv->kind == Json::Bool ? (jv->b ? "t" : "f") : jv->str);
    std::string tn = dtype->type_name;
    union { int16_t i2; int32_t i4; int64_t i8; float f4; double f8; unsigned char b[8]; } u;
    memset(&u, 0, sizeof(u));
    if (tn == "bool")       { u.b[0] = parse_bool_text(text) ? 1 : 0; buf.insert(buf.end(), u.b, u.b + 1); }
    else if (tn == "int2")  { u.i2 = (int16_t)strtoll(text.c_str(), NULL, 10); buf.insert(buf.end(), u.b, u.b + 2); }
    else if (tn == "int4" || tn == "date")
                            { u.i4 = (int32_t)strtoll(text.c_str(), NULL, 10); buf.insert(buf.end(), u.b, u.b + 4); }
    else if (tn == "int8" || tn == "time" || tn == "timestamp")
                            { u.i8 = (int64_t)strtoll(text.c_str(), NULL, 10); buf.insert(buf.end(), u.b, u.b + 8); }
    else if (tn == "float4"){ u.f4 = (float)parse_float_text(text); buf.insert(buf.end(), u.b, u.b + 4); }
    else if (tn == "float8"){ u.f8 = parse_float_text(text); buf.insert(buf.end(), u.b, u.b + 8); }
    else if (tn == "numeric")
    {
        std::vector<unsigned char> vl;
        if (!pgs_numeric_from_text(text, vl))
            return;
        buf.insert(buf.end(), vl.begin(), vl.end());
    }
    else if (tn == "bytea" || tn == "text" || tn == "bpchar")
    {
        /* "constbytes": hex image of the payload, else the text itself */
        std::string payload;
        if (node->has("constbytes"))
        {
            std::string hex = node->s("constbytes");
            for (size_t i = 0; i + 1 < hex.size(); i += 2)
                payload += (char)strtoul(hex.substr(i, 2).c_str(), NULL, 16);
        }
        else
            payload = text;
        uint32_t hdr = (uint32_t)((payload.size() + 4) << 2);  /* SET_VARSIZE */
        unsigned char h[4];
        memcpy(h, &hdr, 4);
        buf.insert(buf.end(), h, h + 4);
        buf.insert(buf.end(), payload.begin(), payload.end());
    }
    else
        return;
    *isnull = false;
}

std::vector<unsigned char>
create_kern_parambuf(const std::vector<JsonPtr> &used_params)
{
    size_t nparams = used_params.size();
    size_t offset = STROMALIGN(offsetof(kern_parambuf, poffset) + sizeof(cl_uint) * nparams);
    std::vector<unsigned char> buf(offset, 0);
    std::vector<cl_uint> poffset(nparams, 0);

    for (size_t index = 0; index < nparams; index++)
    {
        bool isnull = true;
        size_t pos = buf.size();

        if (used_params[index] && !used_params[index]->is_null())
            append_datum(buf, used_params[index], &isnull);
        if (isnull)
        {
            buf.resize(pos);
            poffset[index] = 0;
        }
        else
            poffset[index] = (cl_uint)pos;
        buf.resize(STROMALIGN(buf.size()), 0);
    }
    kern_parambuf *kpbuf = (kern_parambuf *)buf.data();
    kpbuf->length = (cl_uint)buf.size();
    kpbuf->nparams = (cl_uint)nparams;
    if (nparams)
        memcpy(kpbuf->poffset, poffset.data(), sizeof(cl_uint) * nparams);
    return buf;
}

/* ------------------------------------------------------------------
 * deparse (what EXPLAIN VERBOSE prints; ruleutils.c get_rule_expr)
 * ------------------------------------------------------------------ */
static std::string
deparse_const(const JsonPtr &node)
{
    const DevType *dtype = devtype_lookup(expr_type(node));
    std::string tname = dtype ? dtype->sql_name : expr_type(node);
    if (node->flag("constisnull"))
        return "NULL::" + tname;
    const Json *jv = node->get("constvalue");
    std::string text = jv ? (jv->kind == Json::Bool ? (jv->b ? "t" : "f") : jv->str) : "";
    std::string tn = dtype ? dtype->type_name : "";
    if (tn == "bool")
        return parse_bool_text(text) ? "true" : "false";
    if (tn == "int4")
        return (text.size() && text[0] == '-') ? "(" + text + ")" : text;
    if (tn == "numeric" || tn == "float8")
    {
        /* numeric literals print bare, float8 as '...'::double precision */
        if (tn == "numeric")
            return text;
        return "'" + text + "'::" + tname;
    }
    if (tn == "int2" || tn == "int8" || tn == "float4")
        return "'" + text + "'::" + tname;
    return "'" + text + "'::" + tname;
}

static bool
is_simple_node(const JsonPtr &n)
{
    std::string t = n->s("node");
    return t == "Var" || t == "Const" || t == "Param";
}

std::string
deparse_expression(const JsonPtr &node, const std::vector<std::string> &colnames,
                   bool toplevel_parens)
{
    if (!node || node->is_null())
        return "";
    std::string tag = node->s("node");
    if (tag == "Var")
    {
        long long a = node->i("varattno");
        if (a >= 1 && (size_t)a <= colnames.size())
            return colnames[a - 1];
        return "?column" + std::to_string(a) + "?";
    }
    if (tag == "Const")
        return deparse_const(node);
    if (tag == "Param")
        return "$" + std::to_string(node->i("paramid"));
    if (tag == "FuncExpr")
    {
        std::string fmt = node->s("funcformat", "call");
        const Json *args = node->get("args");
        if ((fmt == "cast" || fmt == "implicit") && args && args->arr.size() == 1)
        {
            const DevType *dtype = devtype_lookup(expr_type(node));
            std::string inner = deparse_expression(args->arr[0], colnames, true);
            return "(" + inner + ")::" + (dtype ? dtype->sql_name : expr_type(node));
        }
        std::string s = node->s("funcschema").empty() ? "" : node->s("funcschema") + ".";
        s += node->s("funcname") + "(";
        if (args)
            for (size_t i = 0; i < args->arr.size(); i++)
                s += (i ? ", " : "") + deparse_expression(args->arr[i], colnames, true);
        return s + ")";
    }
    if (tag == "OpExpr")
    {
        const Json *args = node->get("args");
        std::string s;
        if (args && args->arr.size() == 2)
            s = deparse_expression(args->arr[0], colnames, true) + " " +
                node->s("opname") + " " +
                deparse_expression(args->arr[1], colnames, true);
        else if (args && args->arr.size() == 1)
            s = node->s("opname") + " " + deparse_expression(args->arr[0], colnames, true);
        return toplevel_parens ? "(" + s + ")" : s;
    }
    if (tag == "DistinctExpr")
    {
        const Json *args = node->get("args");
        std::string s;
        if (args && args->arr.size() == 2)
            s = deparse_expression(args->arr[0], colnames, true) + " IS DISTINCT FROM " +
                deparse_expression(args->arr[1], colnames, true);
        return toplevel_parens ? "(" + s + ")" : s;
    }
    if (tag == "NullTest")
    {
        std::string s = deparse_expression(node->getp("arg"), colnames, true) +
            (node->s("nulltesttype") == "IS_NULL" ? " IS NULL" : " IS NOT NULL");
        return toplevel_parens ? "(" + s + ")" : s;
    }
    if (tag == "BooleanTest")
    {
        std::string t = node->s("booltesttype");
        std::string txt = t.substr(3);
        for (auto &c : txt) if (c == '_') c = ' ';
        std::string s = deparse_expression(node->getp("arg"), colnames, true) + " IS " + txt;
        return toplevel_parens ? "(" + s + ")" : s;
    }
    if (tag == "BoolExpr")
    {
        std::string op = node->s("boolop");
        const Json *args = node->get("args");
        std::string s;
        if (op == "NOT")
            s = "NOT " + deparse_expression(args->arr[0], colnames, true);
        else
            for (size_t i = 0; args && i < args->arr.size(); i++)
                s += (i ? " " + op + " " : "") + deparse_expression(args->arr[i], colnames, true);
        return toplevel_parens ? "(" + s + ")" : s;
    }
    if (tag == "RelabelType")
    {
        const DevType *dtype = devtype_lookup(expr_type(node));
        return "(" + deparse_expression(node->getp("arg"), colnames, true) + ")::" +
            (dtype ? dtype->sql_name : expr_type(node));
    }
    if (tag == "CaseExpr")
    {
        std::string s = "CASE";
        JsonPtr carg = node->getp("arg");
        if (carg && !carg->is_null())
            s += " " + deparse_expression(carg, colnames, true);
        const Json *args = node->get("args");
        for (size_t i = 0; args && i < args->arr.size(); i++)
            s += " WHEN " + deparse_expression(args->arr[i]->getp("expr"), colnames, false) +
                " THEN " + deparse_expression(args->arr[i]->getp("result"), colnames, true);
        JsonPtr def = node->getp("defresult");
        if (def && !def->is_null())
            s += " ELSE " + deparse_expression(def, colnames, true);
        return s + " END";
    }
    if (tag == "Aggref")
    {
        std::string s = node->s("aggschema").empty() ? "" : node->s("aggschema") + ".";
        s += node->s("aggname") + "(";
        const Json *args = node->get("args");
        if (node->flag("aggstar"))
            s += "*";
        else if (args)
            for (size_t i = 0; i < args->arr.size(); i++)
            {
                JsonPtr a = args->arr[i];
                if (a->s("node") == "TargetEntry")
                    a = a->getp("expr");
                /* arguments that are themselves function calls are shown in
                 * parentheses: they come from a lower plan's target list */
                std::string t = deparse_expression(a, colnames, true);
                if (node->flag("args_are_subplan_outputs") && !is_simple_node(a))
                    t = "(" + t + ")";
                s += (i ? ", " : "") + t;
            }
        s += ")";
        JsonPtr f = node->getp("aggfilter");
        if (f && !f->is_null())
            s += " FILTER (WHERE " + deparse_expression(f, colnames, false) + ")";
        return s;
    }
    return "???";
}

}   /* namespace pgs */
