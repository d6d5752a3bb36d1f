"""CPU: the code generator, end to end.  The CUDA source the planner half
emits for a query (gpupreagg_qual_eval + gpupreagg_projection, the counterpart
of what /root/reference/codegen.c:1065-1430 and gpupreagg.c:1449-1900 emit) is
compiled with g++ against host stand-ins of the two CUDA-only headers - cut,
at test time, out of the real kern_common.cuh / kern_gpupreagg.cuh, so the
type templates, EVAL, the boolean tests, pagg_row and the vstore functions
are the product's own text, and kern_mathlib / numeric / timelib / textlib are
the real files - and run row by row against the oracle's evaluation of the
same expression trees (oracle/pg_expr.py).  Hand-written queries plus a
seeded fuzzer over typed random expression trees: NULL propagation,
three-valued AND / OR / NOT, CASE, IS [NOT] NULL, casts, overflow / division
by zero (-> StromError_CpuReCheck), cross-type comparison, date arithmetic,
text comparison."""
import ctypes as C
import math
import os
import random
import struct
import subprocess

import pytest
from decimal import Decimal

from oracle import pg_expr, pg_typelib as T
from oracle.pg_agg import PgError, f4
from pg_strom_b200 import gpupreagg as gp
from pg_strom_b200 import pgplan as P

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "pg_strom_b200", "csrc")
SIM = os.path.join(HERE, "native", "_hostsim")
CPU_RECHECK = 2
GUCS = {"pg_strom.enabled": "on", "pg_strom.debug_force_gpupreagg": "on"}

HOST_PREAMBLE = r'''
#include <cstdint>
#include <cstring>
#include <cmath>
using std::isnan; using std::isinf;
#define DEVFN static inline
#ifndef INT_MAX
#define SHRT_MAX    32767
#define SHRT_MIN    (-32767-1)
#define INT_MAX     2147483647
#define INT_MIN     (-INT_MAX-1)
#endif
#undef LONG_MAX
#undef LONG_MIN
#define LONG_MAX    9223372036854775807LL
#define LONG_MIN    (-LONG_MAX-1LL)
static inline long long __mul64hi(long long a, long long b)
{ return (long long)(((__int128)a * (__int128)b) >> 64); }
static inline double __longlong_as_double(long long v)
{ double d; memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d)
{ long long v; memcpy(&v, &d, 8); return v; }
static inline unsigned int __float_as_uint(float f)
{ unsigned int v; memcpy(&v, &f, 4); return v; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned int min(unsigned int a, unsigned int b) { return a < b ? a : b; }
static inline unsigned int max(unsigned int a, unsigned int b) { return a > b ? a : b; }
/* single-threaded stand-ins of the atomics the ATOMIC / SHARED flavours use */
template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicCAS(T *p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
template <typename T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
static inline void STROM_SET_ERROR(cl_int *p_error, cl_int errcode)
{
    cl_int oldcode = *p_error;
    if (StromErrorIsSignificant(errcode))
    {
        if (!StromErrorIsSignificant(oldcode))
            *p_error = errcode;
    }
    else if (errcode > oldcode)
        *p_error = errcode;
}
'''

# CUDA decorations of the generated text itself (it starts with constexpr
# helpers marked __host__ __device__)
PRE = r'''
#define __device__
#define __host__
#define __forceinline__ inline
#define __constant__
#define __align__(x)
'''

WRAPPER = r'''
struct host_kds
{
    const unsigned long long *vals;     /* one 8-byte word per staged slot */
    unsigned int valid;                 /* bit per slot: NOT NULL */
    template <typename T> bool fetch(int slot, cl_uint rowidx, T &out) const
    {
        union { unsigned long long u; T t; } cv;
        cv.u = vals[slot];
        out = cv.t;
        return ((valid >> slot) & 1U) != 0;
    }
};
extern "C" int sim_row(const unsigned long long *vals, unsigned int valid,
                       const void *kparams, const void *ktoast,
                       unsigned long long *key_out, unsigned char *key_null,
                       unsigned long long *agg_out, unsigned char *agg_null, int *passed)
{
    cl_int e1 = 0, e2 = 0;
    host_kds k = { vals, valid };
    pagg_row prow;
    bool v = gpupreagg_qual_eval(&e1, (const kern_parambuf *)kparams, k, ktoast, 0);
    memset(&prow, 0, sizeof(prow));
    *passed = v ? 1 : 0;
    if (v && e1 == 0)
        gpupreagg_projection(&e2, (const kern_parambuf *)kparams, k, prow, ktoast, 0, 0);
    for (int i = 0; i < GPUPREAGG_NUM_KEYS; i++)
    { key_out[i] = prow.key[i].ulong_val; key_null[i] = prow.key[i].isnull; }
    for (int i = 0; i < GPUPREAGG_NUM_AGGS; i++)
    { agg_out[i] = prow.agg[i].ulong_val; agg_null[i] = prow.agg[i].isnull; }
    return e1 | (e2 << 8);
}

#ifdef KERN_TEXTLIB_CUH
/* the session's key heap (kern_textlib.cuh) in caller memory; nslots = 0: none */
extern "C" void sim_keyheap_set(void *slots, unsigned int nslots, void *heap,
                                unsigned long long heap_bytes, void *used)
{
    pgs_keyheap.slots = (cl_ulong *)slots;
    pgs_keyheap.nslots = nslots;
    pgs_keyheap.heap = (unsigned char *)heap;
    pgs_keyheap.heap_bytes = heap_bytes;
    pgs_keyheap.heap_used = (cl_ulong *)used;
    pgs_keyheap.max_probe = 512;
}
#endif

/* ---- aggregation: the same rows through every flavour of the merge rules ----
 * The cell formats of the register flavours (PLAIN / THREAD, no GROUP BY) and
 * of the table flavours (ATOMIC / SHARED, GROUP BY) differ (e.g. int4 min /
 * max: zero- vs sign-extended), so each query runs the flavours its kernels
 * use:   no GROUP BY: 0 PLAIN, 1 THREAD, 2 PLAIN(even rows) <- PLAIN(odd rows)
 *        GROUP BY   : 0 ATOMIC, 1 SHARED, 2 ATOMIC(even) <- SHARED(odd), the
 *                     spill of a CTA-local table into the global one
 * state 3 holds the odd rows until agg_finish merges them into state 2. */
#define SIM_NFLAV   4
#define SIM_MAXGRP  1024
struct sim_group
{
    bool        used;
    cl_ulong    keys[PGS_MAX(GPUPREAGG_NUM_KEYS, 1)];
    cl_uint     knull;
    cl_ulong    cells[PGS_MAX(GPUPREAGG_NUM_CELLS, 1)];
    cl_uint     nn;
};
static sim_group sim_state[SIM_NFLAV][SIM_MAXGRP];
static unsigned int sim_nrows;

extern "C" void agg_reset(void)
{
    memset(sim_state, 0, sizeof(sim_state));
    sim_nrows = 0;
}
static sim_group *sim_find(int f, const pagg_row &prow)
{
    cl_uint knull = 0;
    cl_ulong keys[PGS_MAX(GPUPREAGG_NUM_KEYS, 1)] = {0};
    for (int k = 0; k < GPUPREAGG_NUM_KEYS; k++)
    {
        if (prow.key[k].isnull) knull |= (1U << k);
        else keys[k] = prow.key[k].ulong_val;
    }
    for (int g = 0; g < SIM_MAXGRP; g++)
    {
        sim_group *s = &sim_state[f][g];
        if (!s->used)
        {
            s->used = true;
            s->knull = knull;
            memcpy(s->keys, keys, sizeof(keys));
            pgs_cells_init(s->cells);
            s->nn = 0;
            return s;
        }
        if (s->knull == knull && memcmp(s->keys, keys, sizeof(keys)) == 0)
            return s;
    }
    return NULL;
}
/* returns the row's error code; 0x100 = filtered by the qual */
extern "C" int agg_row(const unsigned long long *vals, unsigned int valid,
                       const void *kparams, const void *ktoast)
{
    cl_int e = 0;
    host_kds k = { vals, valid };
    pagg_row prow;
    memset(&prow, 0, sizeof(prow));
    bool v = gpupreagg_qual_eval(&e, (const kern_parambuf *)kparams, k, ktoast, 0);
    if (e != 0) return e;
    if (!v) return 0x100;
    gpupreagg_projection(&e, (const kern_parambuf *)kparams, k, prow, ktoast, 0, 0);
    gpupreagg_aggcheck(&e, prow);
    if (e != 0) return e;
    unsigned int odd = (sim_nrows++ & 1U);
    for (int f = 0; f < SIM_NFLAV; f++)
    {
        if ((f == 2 && odd) || (f == 3 && !odd))
            continue;
        sim_group *s = sim_find(f, prow);
        if (!s) return -1;
#if GPUPREAGG_NUM_KEYS == 0
        if (f == 1)
        {
            bool nnflag[PGS_MAX(PGS_NUM_NNCLASSES, 1)];
            for (int i = 0; i < PGS_MAX(PGS_NUM_NNCLASSES, 1); i++) nnflag[i] = false;
            gpupreagg_aggcalc_thread(s->cells, prow, true, nnflag);
            s->nn |= pgs_nnflags_to_mask(nnflag);
        }
        else
            s->nn |= gpupreagg_aggcalc_plain(s->cells, prow, true);
#else
        if (f == 1 || f == 3)
            s->nn |= gpupreagg_aggcalc_shared(s->cells, prow);
        else
            s->nn |= gpupreagg_aggcalc_atomic(s->cells, prow);
#endif
    }
    return 0;
}
/* state 3 -> state 2 with the state -> state rules */
extern "C" int agg_finish(void)
{
    for (int g = 0; g < SIM_MAXGRP; g++)
    {
        sim_group *src = &sim_state[3][g];
        if (!src->used) continue;
        pagg_row fake;
        memset(&fake, 0, sizeof(fake));
        for (int k = 0; k < GPUPREAGG_NUM_KEYS; k++)
        { fake.key[k].isnull = (src->knull >> k) & 1U; fake.key[k].ulong_val = src->keys[k]; }
        sim_group *dst = sim_find(2, fake);
        if (!dst) return -1;
#if GPUPREAGG_NUM_KEYS == 0
        gpupreagg_aggmerge_plain(dst->cells, src->cells, src->nn);
#else
        gpupreagg_aggmerge_atomic(dst->cells, src->cells, src->nn);
#endif
        dst->nn |= src->nn;
    }
    return 0;
}
/* overwrite one cell of group 0 (tests of the flush: values no test input reaches) */
extern "C" void agg_poke(int f, int cell, unsigned long long value, unsigned int nn)
{
    pagg_row none;
    memset(&none, 0, sizeof(none));
    sim_group *s = sim_find(f, none);
    s->cells[cell] = value;
    s->nn |= nn;
}
/* partial rows of one flavour into a TUPSLOT store prepared by the host library */
extern "C" int agg_flush(int f, kern_data_store *kds_dst)
{
    kern_gpupreagg kg;
    memset(&kg, 0, sizeof(kg));
    if (GPUPREAGG_NUM_KEYS == 0 && !sim_state[f][0].used)
    {
        pagg_row none;              /* no row at all: the identity state */
        memset(&none, 0, sizeof(none));
        sim_find(f, none);
    }
    for (int g = 0; g < SIM_MAXGRP; g++)
    {
        sim_group *s = &sim_state[f][g];
        if (!s->used) continue;
        cl_uint nsplit = pgs_flush_nsplit(s->cells);
        cl_uint base = kds_dst->nitems;
        kds_dst->nitems += nsplit;
        pgs_flush_rows(kds_dst, &kg, base, nsplit, s->keys, s->knull, s->cells, s->nn);
    }
    return kg.status;
}
'''


def install_device_numeric_range(setattr_fn):
    """The oracle evaluates NUMERIC without limits; the device keeps 57 bits
    of mantissa and a display scale <= 32 and re-checks what does not fit
    (opencl_numeric.h:141-162), and pgs_float_to_numeric declines tiny / huge
    floats.  Inside these tests the oracle raises where the device declines,
    so that "PgError <=> CpuReCheck" holds for numeric intermediates too."""
    orig = pg_expr._call

    def call(name, argtypes, rettype, args):
        if name == "numeric" and argtypes and argtypes[0] in ("float4", "float8"):
            v = args[0]
            if math.isfinite(v) and v != 0:
                digits = 15 if argtypes[0] == "float8" else 6
                if digits - 1 - Decimal(v).adjusted() > 32 or abs(v) >= 2.0 ** 127:
                    raise PgError("float -> numeric: left to the host")
        if name in ("numeric_add", "numeric_sub"):
            # operands whose display scales are more than 20 apart would need
            # more than 128 bits to align: the device declines
            sa, sb = (max(0, -a.as_tuple().exponent) for a in args)
            if abs(sa - sb) > 20:
                raise PgError("numeric add: scales too far apart for the device")
        r = orig(name, argtypes, rettype, args)
        if isinstance(r, Decimal) and not _numeric_fits(r):
            raise PgError("beyond the 64-bit device numeric")
        return r
    setattr_fn(pg_expr, "_call", call)


@pytest.fixture(autouse=True)
def device_numeric_range(monkeypatch):
    install_device_numeric_range(monkeypatch.setattr)


def _cut(path, start, end):
    text = open(path).read()
    i = text.index(start)
    return text[i:text.index(end, i)]


@pytest.fixture(scope="module")
def simdir(lib):
    os.makedirs(SIM, exist_ok=True)
    common = _cut(os.path.join(CSRC, "kern_common.cuh"),
                  "#define STROMCL_SIMPLE_DATATYPE_TEMPLATE", "#endif  /* KERN_COMMON_CUH */")
    with open(os.path.join(SIM, "kern_common.cuh"), "w") as f:
        f.write("#pragma once\n" + HOST_PREAMBLE + common)
    real = os.path.join(CSRC, "kern_gpupreagg.cuh")
    # pagg_row + vstore, the state cells and the PLAIN / THREAD / ATOMIC merge
    # rules ... (the typed register accumulators in between are inline PTX)
    part1 = _cut(real, "#define PGS_MAX(a,b)", "/* ---- typed thread accumulators")
    # ... the SHARED flavours, 128-bit and numeric sums, row -> state and
    # state -> state functions (without struct pgs_tacc) ...
    part2 = _cut(real, "/* ---- SHARED flavours", "/* typed per-thread state of the no-group fast path")
    part3 = _cut(real, "DEVFN void\npgs_row_special", "/* ------------------------------------------------------------------\n * device-resident session state")
    # ... and the flush: state -> TUPSLOT partial rows, splitting big counts / sums
    part4 = _cut(real, "#define PGS_INT4_MAX", "/* launched with whole warps.")
    with open(os.path.join(SIM, "kern_gpupreagg.cuh"), "w") as f:
        f.write("#pragma once\n"
                "static inline double pgs_f8_canon(double v)\n"
                "{ if (isnan(v)) return __longlong_as_double(0x7FF8000000000000LL);"
                " return (v == 0.0) ? 0.0 : v; }\n" + part1 + part2 + part3 + part4)
    return SIM


_nbuilt = [0]


def build_sim(plan, simdir):
    _nbuilt[0] += 1
    src = os.path.join(simdir, "q%d.cpp" % _nbuilt[0])
    out = os.path.join(simdir, "q%d.so" % _nbuilt[0])
    with open(src, "w") as f:
        f.write(PRE + plan.kernel_source() + WRAPPER)
    r = subprocess.run(["g++", "-std=c++17", "-O0", "-fPIC", "-shared", "-w", "-ffp-contract=off",
                        "-I", simdir, "-I", os.path.join(ROOT, "include"), "-I", CSRC,
                        "-o", out, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[:3000]
    so = C.CDLL(out)
    so.sim_row.argtypes = [C.POINTER(C.c_uint64), C.c_uint, C.c_char_p, C.c_char_p,
                           C.POINTER(C.c_uint64), C.c_char_p, C.POINTER(C.c_uint64), C.c_char_p,
                           C.POINTER(C.c_int)]
    return so


def pack(value, typ):
    if typ == "bool":
        return int(bool(value))
    if typ in ("int2", "int4", "int8", "date", "time", "timestamp"):
        return value & 0xFFFFFFFFFFFFFFFF if typ in ("int8", "time", "timestamp") else \
            value & (0xFFFF if typ == "int2" else 0xFFFFFFFF)
    if typ == "float4":
        return struct.unpack("<I", struct.pack("<f", value))[0]
    if typ == "float8":
        return struct.unpack("<Q", struct.pack("<d", value))[0]
    raise KeyError(typ)


def fill_vals(row, incols, coltypes, vals):
    """One input row as the staged words of its referenced columns: by-value
    types in place, varlena columns as the offset of the datum in a toast
    buffer (what a chunk holds).  -> (validity bits, toast bytes)"""
    toast = bytearray(b"\0" * 8)
    valid = 0
    for slot, c in enumerate(incols):
        v = row[c]
        vals[slot] = 0
        if v is None:
            continue
        valid |= 1 << slot
        if coltypes[c] in ("text", "bpchar", "numeric"):
            while len(toast) % 4:
                toast.append(0)
            vals[slot] = len(toast)
            toast += (T.varlena(v) if coltypes[c] != "numeric"
                      else gp.numeric_datum(format(v, "f")))
        else:
            vals[slot] = pack(v, coltypes[c])
    return valid, bytes(toast)


def same(a, b, typ):
    if a is None or b is None:
        return a is None and b is None
    if typ in ("float4", "float8"):
        a, b = float(a), float(b)
        if math.isnan(a) or math.isnan(b):
            return math.isnan(a) and math.isnan(b)
        return a == b
    return a == b


def find_node(tree):
    n = tree
    while n is not None:
        if n.get("node") == "CustomPlan" and n.get("custom_name") == "GpuPreAgg":
            return n
        n = n.get("lefttree")
    raise AssertionError("no GpuPreAgg node")


def check_query(table, tree, rows, simdir, key_heap_bytes=0):
    """Runs every row through the compiled generated code and through the
    oracle; returns (#passed, #errors) for the caller's sanity checks.
    key_heap_bytes > 0: the program gets a key heap like a session's (text
    keys of more than 7 bytes are grouped on the device)."""
    plan = gp.Plan(tree, gucs=GUCS)
    try:
        assert plan.num_gpupreagg == 1, plan.reject_reason
        desc = plan.describe()
        node = find_node(plan.tree())
        so = build_sim(plan, simdir)
        kparams = plan.kparams()
        incols = desc["incol_index"]
        coltypes = [t for _, t in table.columns]
        cols = desc["columns"]
        keycols = [c for c in cols if c["role"] == 1]
        aggcols = [c for c in cols if c["role"] == 2]
        quals = node.get("outer_quals") or []
        tlist = node["targetlist"]
        npass = nerr = 0
        vals = (C.c_uint64 * max(1, len(incols)))()
        key_out = (C.c_uint64 * 16)()
        agg_out = (C.c_uint64 * 32)()
        key_null = C.create_string_buffer(16)
        agg_null = C.create_string_buffer(32)
        passed = C.c_int()
        kh_heap = kh_used = None
        if key_heap_bytes:
            kh_slots = (C.c_uint64 * (2 * 1024))()
            kh_heap = C.create_string_buffer(key_heap_bytes)
            kh_used = C.c_uint64(0)
            so.sim_keyheap_set.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_uint64, C.c_void_p]
            so.sim_keyheap_set(kh_slots, 1024, kh_heap, key_heap_bytes, C.byref(kh_used))
        words = {}
        for row in rows:
          try:
            valid, toast = fill_vals(row, incols, coltypes, vals)
            rc = so.sim_row(vals, valid, kparams, toast, key_out, key_null,
                            agg_out, agg_null, C.byref(passed))
            e1, e2 = rc & 0xff, rc >> 8
            # the device evaluates every qual (no short circuit across the
            # implicit AND): an error anywhere flags the row
            qvals, qerr = [], False
            for q in quals:
                try:
                    qvals.append(pg_expr.evaluate(q, row))
                except PgError:
                    qerr = True
            if qerr:
                assert e1 == CPU_RECHECK, (row, rc)
                nerr += 1
                continue
            assert e1 == 0, (row, rc)
            want = all(v is True for v in qvals)
            assert bool(passed.value) == want, (row, qvals)
            if not want:
                continue
            npass += 1
            exp, perr = {}, False
            for c in cols:
                if c["role"] == 0:
                    continue
                try:
                    exp[c["resno"]] = pg_expr.evaluate(tlist[c["resno"] - 1]["expr"], row)
                except PgError:
                    perr = True
            # a text key longer than 7 bytes does not fit the key word
            # ("kernel text"): that row is the host's
            for c in keycols:
                v = exp.get(c["resno"])
                if c["type"] in ("text", "bpchar") and v is not None and \
                        len(v.rstrip(b" ") if c["type"] == "bpchar" else v) > 7 and \
                        not key_heap_bytes:
                    perr = True
            if perr:
                assert e2 == CPU_RECHECK, (row, rc)
                nerr += 1
                continue
            assert e2 == 0, (row, rc)
            for i, c in enumerate(keycols):
                heap = (C.addressof(kh_heap), kh_used.value) if key_heap_bytes else (None, 0)
                got = None if key_null.raw[i] != b"\0"[0] else \
                    gp.decode_datum(key_out[i], False, c["type"], c.get("typmod", -1), heap)
                if got is not None and c["type"] in ("text", "bpchar"):
                    # equal strings <=> equal key words
                    assert words.setdefault((i, got), key_out[i]) == key_out[i]
                e = exp[c["resno"]]
                if c["type"] in ("float4", "float8") and e is not None and e == 0:
                    e = 0.0                                     # -0 groups with +0
                assert same(got, e, c["type"]), (row, c["text"], got, e)
            for i, c in enumerate(aggcols):
                got = None if agg_null.raw[i] != b"\0"[0] else \
                    gp.decode_datum(agg_out[i], False, c["type"])
                e = exp[c["resno"]]
                if c["type"] == "float4" and e is not None:
                    e = f4(e)
                assert same(got, e, c["type"]), (row, c["text"], got, e)
          except AssertionError as exc:
            raise AssertionError("%s\n%s" % (exc, "\n".join(plan.explain()[:10]))) from None
        return npass, nerr
    finally:
        plan.free()


# ---- data ------------------------------------------------------------------
TBL = P.Table("fz", [("b", "bool"), ("s2", "int2"), ("i4", "int4"), ("i8", "int8"),
                     ("f4", "float4"), ("f8", "float8"), ("d", "date"), ("ts", "timestamp"),
                     ("tx", "text"), ("n", "numeric"), ("k", "int4")])
EDGE = {
    "bool": [True, False],
    "int2": [0, 1, -1, 32767, -32768, 100, -100, 7],
    "int4": [0, 1, -1, 2147483647, -2147483648, 65536, -65536, 10, 3],
    "int8": [0, 1, -1, 2 ** 63 - 1, -2 ** 63, 2 ** 32, -2 ** 32, 3037000500, 5],
    "float4": [0.0, -0.0, 1.0, -1.5, 2.5, 3.4e38, -3.4e38, 1e-38, 16777216.0, float("inf"),
               float("nan")],
    "float8": [0.0, -0.0, 1.0, -1.5, 2.5, 0.5, 1e308, -1e308, 1e-308, 2147483647.5,
               9.3e18, float("inf"), float("-inf"), float("nan"), 32767.5],
    "date": [0, 1, -1, 7305, T.DATE_NOBEGIN, T.DATE_NOEND, 106751992, -2451545, 2147483000],
    "timestamp": [0, 1, -1, T.DT_NOBEGIN, T.DT_NOEND, 631152000000000, 86400000000,
                  -86400000001, 2 ** 62],
    "text": [b"", b"a", b"abc", b"abd", b"ab", b"b", b"\xc3\xa9", b"abc ", b"zzzzzzzzzzzz"],
    "numeric": [Decimal(x) for x in ("0", "1", "-1", "0.5", "1.50", "-2.25", "100", "0.001",
                                     "123456789012345", "-99999.99999", "3.0000")],
}


def rand_value(typ, rng):
    if rng.random() < 0.12:
        return None
    if rng.random() < 0.55:
        v = rng.choice(EDGE[typ])
    elif typ == "bool":
        v = rng.random() < 0.5
    elif typ in ("int2", "int4", "int8"):
        bits = {"int2": 15, "int4": 31, "int8": 63}[typ]
        v = rng.randrange(-2 ** bits, 2 ** bits) >> rng.choice([0, bits // 2, bits - 3])
    elif typ in ("float4", "float8"):
        v = rng.uniform(-1, 1) * 10.0 ** rng.randrange(-6, 12)
    elif typ == "date":
        v = rng.randrange(-20000, 20000)
    elif typ == "timestamp":
        v = rng.randrange(-10 ** 15, 10 ** 15)
    elif typ == "numeric":
        v = Decimal(rng.randrange(-10 ** rng.choice([2, 6, 12]), 10 ** rng.choice([2, 6, 12]))) \
            .scaleb(-rng.choice([0, 0, 1, 2, 4, 8]))
    else:
        v = bytes(rng.choice(b"abz ") for _ in range(rng.randrange(0, 5)))
    if typ == "float4":
        v = f4(v)
    return v


def rand_rows(n, rng):
    types = [t for _, t in TBL.columns]
    return [tuple(rand_value(t, rng) for t in types[:-1]) + (rng.randrange(0, 5),)
            for _ in range(n)]


# ---- random typed expressions ----------------------------------------------
NUM = ("int2", "int4", "int8", "float4", "float8")
COLS = {"bool": ["b"], "int2": ["s2"], "int4": ["i4", "k"], "int8": ["i8"], "float4": ["f4"],
        "float8": ["f8"], "date": ["d"], "timestamp": ["ts"], "text": ["tx"], "numeric": ["n"]}
CMP = ["=", "<>", "<", "<=", ">", ">="]


def const_of(typ, rng):
    v = rng.choice([x for x in EDGE[typ] if not (isinstance(x, float) and (x != x or abs(x) == math.inf))])
    if typ == "text":
        return P.Const("text", v.decode("utf-8", "replace").rstrip("�") or "a")
    if typ == "bool":
        return P.Const("bool", v)
    if typ in ("float4", "float8"):
        return P.Const(typ, repr(float(v)))
    if typ == "numeric":
        return P.Const("numeric", format(v, "f"))
    return P.Const(typ, v)


def gen(typ, depth, rng):
    """Random expression of SQL type `typ`."""
    if depth <= 0 or rng.random() < 0.25:
        if rng.random() < 0.7 or typ == "text":
            return TBL.col(rng.choice(COLS[typ]))
        return const_of(typ, rng)
    if typ == "bool":
        kind = rng.random()
        if kind < 0.35:
            t = rng.choice(NUM)
            u = rng.choice(NUM) if rng.random() < 0.4 else t
            try:
                return P.Op(rng.choice(CMP), gen(t, depth - 1, rng), gen(u, depth - 1, rng))
            except TypeError:
                return P.Op(rng.choice(CMP), gen(t, depth - 1, rng), gen(t, depth - 1, rng))
        if kind < 0.42:
            a, b = rng.choice([("date", "date"), ("timestamp", "timestamp"),
                               ("date", "timestamp"), ("timestamp", "date")])
            return P.Op(rng.choice(CMP), gen(a, depth - 1, rng), gen(b, depth - 1, rng))
        if kind < 0.47:
            return P.Op(rng.choice(CMP), gen("numeric", depth - 1, rng),
                        gen("numeric", depth - 1, rng))
        if kind < 0.55:
            return P.Op(rng.choice(CMP), gen("text", 0, rng), const_of("text", rng), collation="C")
        if kind < 0.70:
            n = rng.choice([2, 2, 3])
            mk = P.And if rng.random() < 0.5 else P.Or
            return mk(*[gen("bool", depth - 1, rng) for _ in range(n)])
        if kind < 0.74:
            return P.Not(gen("bool", depth - 1, rng))
        if kind < 0.77:
            return {"node": "BooleanTest", "arg": gen("bool", depth - 1, rng),
                    "booltesttype": rng.choice(["IS_TRUE", "IS_NOT_TRUE", "IS_FALSE",
                                                "IS_NOT_FALSE", "IS_UNKNOWN", "IS_NOT_UNKNOWN"])}
        if kind < 0.80:
            ty = rng.choice(list(NUM) + ["date", "bool"])
            return P.Distinct(gen(ty, depth - 1, rng), gen(ty, depth - 1, rng))
        if kind < 0.91:
            t = rng.choice(list(COLS))
            return P.IsNull(gen(t, depth - 1, rng) if t != "text" else TBL.col("tx"),
                            notnull=rng.random() < 0.5)
        return P.Case([(gen("bool", depth - 1, rng), gen("bool", depth - 1, rng))],
                      gen("bool", depth - 1, rng) if rng.random() < 0.7 else None, "bool")
    if typ in NUM:
        kind = rng.random()
        if kind < 0.45:
            op = rng.choice(["+", "-", "*", "/"] + (["%"] if typ in ("int2", "int4", "int8") else []))
            return P.Op(op, gen(typ, depth - 1, rng), gen(typ, depth - 1, rng))
        if kind < 0.70:
            src = rng.choice([t for t in NUM if t != typ])
            return P.Cast(gen(src, depth - 1, rng), typ)
        if kind < 0.75:
            # numeric -> int rounds half away from zero; -> float only from a
            # column / literal (<= 15 digits: exact in binary64 on both sides)
            return P.Cast(gen("numeric", depth - 1 if typ in ("int2", "int4", "int8") else 0, rng),
                          typ)
        if kind < 0.85:
            return P.Case([(gen("bool", depth - 1, rng), gen(typ, depth - 1, rng))],
                          gen(typ, depth - 1, rng) if rng.random() < 0.7 else None, typ)
        if kind < 0.9:
            # CASE <arg> WHEN <value> THEN ... : compares with the type's "=" operator
            at = rng.choice(["int4", "int2", "float8", "date"])
            whens = [{"node": "CaseWhen", "expr": const_of(at, rng),
                      "result": gen(typ, depth - 1, rng)} for _ in range(rng.choice([1, 2]))]
            return {"node": "CaseExpr", "casetype": typ, "arg": gen(at, depth - 1, rng),
                    "args": whens,
                    "defresult": gen(typ, depth - 1, rng) if rng.random() < 0.7 else None}
        if typ == "int4" and rng.random() < 0.5:
            return P.Op("-", gen("date", depth - 1, rng), gen("date", depth - 1, rng))
        return gen(typ, 0, rng)
    if typ == "numeric":
        kind = rng.random()
        if kind < 0.4:
            return P.Op(rng.choice(["+", "-", "*"]), gen("numeric", depth - 1, rng),
                        gen("numeric", depth - 1, rng))
        if kind < 0.7:
            return P.Cast(gen(rng.choice(NUM), depth - 1, rng), "numeric")
        if kind < 0.8:
            return P.Case([(gen("bool", depth - 1, rng), gen("numeric", depth - 1, rng))],
                          gen("numeric", depth - 1, rng) if rng.random() < 0.7 else None, "numeric")
        return gen("numeric", 0, rng)
    if typ == "date":
        kind = rng.random()
        if kind < 0.4:
            return P.Op(rng.choice(["+", "-"]), gen("date", depth - 1, rng), gen("int4", depth - 1, rng))
        if kind < 0.6:
            return P.Cast(gen("timestamp", depth - 1, rng), "date")
        return gen("date", 0, rng)
    if typ == "timestamp":
        if rng.random() < 0.5:
            return P.Cast(gen("date", depth - 1, rng), "timestamp")
        return gen("timestamp", 0, rng)
    return gen(typ, 0, rng)


def test_hand_written_queries(simdir):
    rng = random.Random(100)
    rows = rand_rows(400, rng)
    t = TBL
    cnt = (P.Agg("count", star=True), "count")
    q1 = P.make_agg_plan(
        t, [(t.col("k"), "k"), cnt, (P.Agg("sum", [t.col("i4")]), "sum"),
            (P.Agg("avg", [t.col("f8")]), "avg"), (P.Agg("max", [t.col("s2")]), "max"),
            (P.Agg("count", [t.col("tx")]), "count")],
        group_by=["k"], num_groups=8,
        where=[P.Or(P.Op("<", t.col("i4"), P.Const("int4", 100)), P.IsNull(t.col("f8")))])
    npass, nerr = check_query(t, q1, rows, simdir)
    assert npass > 50
    q2 = P.make_agg_plan(
        t, [(t.col("f8"), "f8"), cnt,
            (P.Agg("sum", [P.Op("*", t.col("i4"), t.col("s2"))]), "sum"),
            (P.Agg("min", [P.Op("/", t.col("i8"), P.Cast(t.col("i4"), "int8"))]), "min"),
            (P.Agg("variance", [t.col("f8")]), "variance"),
            (P.Agg("corr", [t.col("f8"), t.col("f4")]), "corr")],
        group_by=["f8"], num_groups=8,
        where=[P.Not(P.And(t.col("b"), P.Op(">", t.col("d"), P.Const("date", 0))))])
    npass, nerr = check_query(t, q2, rows, simdir)
    assert npass > 50 and nerr > 5          # int4 * int2 overflow, division by zero


def test_long_text_keys_go_through_the_key_heap(simdir):
    """GROUP BY a text column (and a text CASE expression's input) whose
    values are longer than the 7 bytes a key word holds: with a key heap the
    generated projection interns the string (pgs_text_keybits ->
    pgs_keyheap_intern) and the host gets it back from the heap copy; no row
    is left to the host.  Without one the same rows are re-checked (the
    default of every other test in this file)."""
    rng = random.Random(77)
    pool = [b"", b"abc", b"abcdefg", b"abcdefgh", b"abcdefgh ", b"category-with-a-long-name",
            b"x" * 300, b"caf\xc3\xa9 au lait", b"zzzzzzzzzzzz"]
    types = [t for _, t in TBL.columns]
    rows = []
    for _ in range(400):
        r = list(rand_rows(1, rng)[0])
        r[8] = None if rng.random() < 0.1 else rng.choice(pool)
        rows.append(tuple(r))
    assert types[8] == "text"
    tree = P.make_agg_plan(TBL, [(TBL.col("tx"), "tx"), (P.Agg("count", star=True), "n"),
                                 (P.Agg("max", [TBL.col("i4")]), "m")],
                           group_by=["tx"], num_groups=16,
                           where=[P.Op(">", TBL.col("k"), P.Const("int4", 0))])
    npass, nerr = check_query(TBL, tree, rows, simdir, key_heap_bytes=1 << 16)
    assert npass > 200 and nerr == 0
    npass0, nerr0 = check_query(TBL, tree, rows, simdir)
    assert npass0 == npass and nerr0 > 100         # long keys: rows for the host


def test_fuzz_generated_code(simdir):
    rng = random.Random(2024)
    rows = rand_rows(250, rng)
    total_pass = total_err = built = 0
    for _ in range(60):
        quals = [gen("bool", 4, rng) for _ in range(rng.choice([1, 1, 2]))]
        aggs = []
        for _ in range(rng.choice([1, 2, 3])):
            typ = rng.choice(NUM)
            fn = rng.choice(["min", "max", "sum", "avg", "count"] if typ != "int8"
                            else ["min", "max", "avg", "count"])
            aggs.append((P.Agg(fn, [gen(typ, 3, rng)]), fn))
        keyed = rng.random() < 0.6
        keycol = rng.choice(["k", "f8", "d", "b", "tx", "f4", "i8", "ts", "s2"])
        targets = ([(TBL.col(keycol), keycol)] if keyed else []) + \
            [(P.Agg("count", star=True), "count")] + aggs
        tree = P.make_agg_plan(TBL, targets, group_by=[keycol] if keyed else [],
                               where=quals, num_groups=8)
        plan = gp.Plan(tree, gucs=GUCS)
        ok = plan.num_gpupreagg == 1 and "#define GPUPREAGG_HAS_QUAL 1" in plan.kernel_source()
        plan.free()
        if not ok:
            continue            # e.g. an aggregate the catalogue does not offload
        npass, nerr = check_query(TBL, tree, rows, simdir)
        total_pass += npass
        total_err += nerr
        built += 1
    assert built >= 40 and total_pass > 500 and total_err > 200


# ---- aggregation semantics ---------------------------------------------------
PSUM_LIMIT = {"DOUBLE": 2.0 ** 960, "FLOAT": 2.0 ** 88}


def _numeric_fits(d):
    """the 64-bit device numeric holds it: 57-bit mantissa, display scale <= 32"""
    if d is None:
        return True
    if not d.is_finite():
        return False
    scale = max(0, -d.as_tuple().exponent)
    return int(abs(d).scaleb(scale)) <= (1 << 57) - 2 and scale <= 32


def _var_attnos(e, out):
    if isinstance(e, dict):
        if e.get("node") == "Var":
            out.add(e["varattno"])
        for v in e.values():
            _var_attnos(v, out)
    elif isinstance(e, list):
        for v in e:
            _var_attnos(v, out)
    return out


def _canon(v):
    from oracle.partial import _canon_key
    return _canon_key(v)


def check_aggregation(table, tree, rows, simdir, lib):
    """All rows through qual + projection + the flavours of the merge rules
    the query's kernels use (no GROUP BY: PLAIN, THREAD, two PLAIN halves
    merged state -> state; GROUP BY: ATOMIC, SHARED, a SHARED half spilled
    into an ATOMIC half) and the flush, compiled for the host; the partial
    rows must equal the oracle's (oracle/partial.py).  Returns (#aggregated,
    #rechecked)."""
    from oracle import bench_oracle, partial
    plan = gp.Plan(tree, gucs=GUCS)
    try:
        assert plan.num_gpupreagg == 1, plan.reject_reason
        desc = plan.describe()
        node = find_node(plan.tree())
        so = build_sim(plan, simdir)
        so.agg_row.argtypes = [C.POINTER(C.c_uint64), C.c_uint, C.c_char_p, C.c_char_p]
        so.agg_flush.argtypes = [C.c_int, C.c_void_p]
        kparams = plan.kparams()
        incols = desc["incol_index"]
        coltypes = [t for _, t in table.columns]
        cols = desc["columns"]
        quals = node.get("outer_quals") or []
        tlist = node["targetlist"]
        vals = (C.c_uint64 * max(1, len(incols)))()
        qual_vars = _var_attnos(quals, set())
        tlist_vars = _var_attnos([x["expr"] for x in tlist], set())
        so.agg_reset()
        ok_rows, nrecheck = [], 0
        for row in rows:
            valid, toast = fill_vals(row, incols, coltypes, vals)
            rc = so.agg_row(vals, valid, kparams, toast)
            # what the row must have done
            want = 0
            try:
                qv = [pg_expr.evaluate(q, row) for q in quals]
            except PgError:
                qv, want = [], CPU_RECHECK
            # a numeric column value beyond the device format fails where it is read
            for a in qual_vars:
                if coltypes[a - 1] == "numeric" and not _numeric_fits(row[a - 1]):
                    want = CPU_RECHECK
            if want == 0 and not all(v is True for v in qv):
                want = 0x100
            if want == 0:
                for a in tlist_vars:
                    if coltypes[a - 1] == "numeric" and not _numeric_fits(row[a - 1]):
                        want = CPU_RECHECK
            if want == 0:
                try:
                    for c in cols:
                        if c["role"] == 0:
                            continue
                        v = pg_expr.evaluate(tlist[c["resno"] - 1]["expr"], row)
                        if c["role"] == 1 and c["type"] in ("text", "bpchar") and \
                                v is not None and len(v) > 7:
                            want = CPU_RECHECK      # key does not fit the key word
                        if c["type"] == "numeric" and v is not None:
                            # the 64-bit device numeric: 57-bit mantissa, scale <= 32;
                            # a sum cell takes scale <= 16 and < 2^96 at that scale
                            from decimal import Decimal
                            d = Decimal(v)
                            if not d.is_finite():           # numeric NaN: host only
                                want = CPU_RECHECK
                                continue
                            scale = max(0, -d.as_tuple().exponent)
                            mant = int(abs(d).scaleb(scale))
                            if mant > (1 << 57) - 2 or scale > 32 or \
                                    (c["op"] == "PSUM" and
                                     (scale > 16 or (mant * 10 ** (16 - scale)) >> 96)):
                                want = CPU_RECHECK
                        if c["role"] == 2 and c["op"] == "PSUM" and v is not None and \
                                c["cell_type"] in PSUM_LIMIT and \
                                abs(float(v)) > PSUM_LIMIT[c["cell_type"]]:
                            want = CPU_RECHECK      # could overflow a sum: host's row
                except PgError:
                    want = CPU_RECHECK
            assert rc == want, (row, rc, want, "\n".join(plan.explain()[:8]))
            if want == 0:
                ok_rows.append(row)
            elif want == CPU_RECHECK:
                nrecheck += 1
        assert so.agg_finish() == 0
        exp, _ = partial.partial_rows(node, ok_rows, len(table.columns))
        exp = {tuple(_canon(k) for k in key): v for key, v in exp.items()}
        ncols = len(cols)
        colmeta = (gp.kern_colmeta * ncols)()
        lib.pgs_plan_result_colmeta(plan.handle, 0, colmeta, ncols)
        length = lib.pgstrom_kds_tupslot_length(ncols, 4096)
        values = (C.c_uint64 * ncols)()
        isnull = C.create_string_buffer(ncols)
        for flavour in range(3):
            buf = C.create_string_buffer(length)
            gp.check(lib.pgstrom_kds_tupslot_init(buf, length, ncols, colmeta, 4096))
            assert so.agg_flush(flavour, buf) == 0
            kds = gp.kern_data_store.from_buffer(buf)
            drows = []
            for r in range(kds.nitems):
                gp.check(lib.pgstrom_fetch_data_store(buf, r, values, isnull))
                drows.append(tuple(gp.decode_datum(values[i], isnull.raw[i] != 0, cols[i]["type"],
                                                   cols[i].get("typmod", -1))
                                   for i in range(ncols)))
            key_idx = [i for i, c in enumerate(cols) if c["role"] == 1]
            drows = [tuple(_canon(v) if i in key_idx else v for i, v in enumerate(r))
                     for r in drows]
            got = bench_oracle.combine_device_rows(desc, drows)
            assert set(got) == set(exp), (flavour, sorted(map(repr, got))[:4],
                                          sorted(map(repr, exp))[:4], "\n".join(plan.explain()[:8]))
            for key, erow in exp.items():
                for i, c in enumerate(cols):
                    if c["role"] != 2:
                        continue
                    e = erow[i]
                    if c["type"] == "float4" and e is not None:
                        try:
                            e = f4(e)       # a float4 sum lives in a double cell, rounded once
                        except OverflowError:
                            e = math.copysign(math.inf, e)
                    g = got[key][i]
                    if flavour == 2 and c["op"] == "PSUM" and c["type"] in ("float4", "float8") \
                            and g is not None and e is not None \
                            and math.isfinite(g) and math.isfinite(e):
                        # two halves added up: another summation order (the
                        # north star's float tolerance)
                        tol = 1e-6 if c["type"] == "float4" else 1e-12
                        if abs(g - e) > tol * max(abs(g), abs(e)):
                            # cancellation (3.4e38 - 3.4e38 + 1): the error bound
                            # of a re-ordered sum is relative to the sum of |x|
                            expr = tlist[c["resno"] - 1]["expr"]
                            kexprs = [tlist[k["resno"] - 1]["expr"] for k in cols if k["role"] == 1]
                            mag = 0.0
                            for r in ok_rows:
                                if tuple(_canon(pg_expr.evaluate(k, r)) for k in kexprs) == key:
                                    v = pg_expr.evaluate(expr, r)
                                    mag += abs(float(v)) if v is not None else 0.0
                            assert abs(g - e) <= tol * mag, (key, c["text"], g, e, mag)
                        continue
                    assert same(g, e, c["type"]), \
                        (flavour, key, c["text"], g, e, "\n".join(plan.explain()[:8]))
                    if c["type"] == "numeric" and g is not None:
                        # the display scale too: sums carry the largest input
                        # scale, min / max return an input as it is
                        from oracle.pg_agg import numeric_out
                        assert numeric_out(g) == numeric_out(e), (flavour, key, c["text"], g, e)
        return len(ok_rows), nrecheck
    finally:
        plan.free()


def test_aggregation_flavours_hand_written(simdir, lib):
    rng = random.Random(7)
    rows = rand_rows(500, rng)
    t = TBL
    cnt = (P.Agg("count", star=True), "count")

    def aggs_of(col):
        return [(P.Agg(f, [t.col(col)]), f) for f in ("count", "min", "max", "sum", "avg")
                if not (f == "sum" and col == "i8")]
    for keycol in (None, "k", "f8", "tx"):
        for col in ("s2", "i4", "i8", "f4", "f8"):
            targets = ([(t.col(keycol), keycol)] if keycol else []) + [cnt] + aggs_of(col)
            if col in ("f4", "f8"):
                targets += [(P.Agg("stddev", [t.col(col)]), "stddev"),
                            (P.Agg("variance", [t.col(col)]), "variance")]
            tree = P.make_agg_plan(t, targets, group_by=[keycol] if keycol else [], num_groups=16)
            nok, nre = check_aggregation(t, tree, rows, simdir, lib)
            assert nok > 100
    tree = P.make_agg_plan(
        t, [cnt, (P.Agg("corr", [t.col("f8"), t.col("f4")]), "corr"),
            (P.Agg("covar_pop", [t.col("f8"), t.col("i4")]), "covar_pop"),
            (P.Agg("avg", [P.Cast(t.col("i4"), "numeric")]), "avg"),
            (P.Agg("max", [P.Cast(t.col("f8"), "numeric")]), "max")],
        where=[P.IsNull(t.col("b"), notnull=True)])
    nok, nre = check_aggregation(t, tree, rows, simdir, lib)
    assert nok > 100 and nre > 10


def test_aggregation_flavours_fuzz(simdir, lib):
    rng = random.Random(99)
    rows = rand_rows(250, rng)
    built = total = 0
    for _ in range(30):
        quals = [gen("bool", 3, rng)] if rng.random() < 0.7 else []
        aggs = []
        for _ in range(rng.choice([1, 2, 3, 4])):
            typ = rng.choice(NUM)
            fn = rng.choice(["min", "max", "sum", "avg", "count"] if typ != "int8"
                            else ["min", "max", "avg", "count"])
            aggs.append((P.Agg(fn, [gen(typ, 2, rng)]), fn))
        keyed = rng.random() < 0.6
        keycol = rng.choice(["k", "f8", "d", "b", "tx", "f4", "i8", "s2"])
        targets = ([(TBL.col(keycol), keycol)] if keyed else []) + \
            [(P.Agg("count", star=True), "count")] + aggs
        tree = P.make_agg_plan(TBL, targets, group_by=[keycol] if keyed else [],
                               where=quals, num_groups=8)
        plan = gp.Plan(tree, gucs=GUCS)
        ok = plan.num_gpupreagg == 1 and \
            ("#define GPUPREAGG_HAS_QUAL 1" in plan.kernel_source()) == bool(quals)
        plan.free()
        if not ok:
            continue
        nok, nre = check_aggregation(TBL, tree, rows, simdir, lib)
        built += 1
        total += nok
    assert built >= 20 and total > 500


# ---- end to end on the CPU: original query vs rewritten plan -----------------
def _final_values(desc, partial_rows):
    """PostgreSQL's final Agg over the partial rows (the pgstrom.* final
    aggregates of pg_strom--1.0.sql:247-401, oracle/pg_agg.FinalAgg):
    {group key: [value per Aggref of the Agg target list]}."""
    from oracle import pg_agg
    cols = desc["columns"]
    key_idx = [i for i, c in enumerate(cols) if c["role"] == 1]
    aggrefs = [t["expr"] for t in desc["agg_targetlist"] if t["expr"]["node"] == "Aggref"]
    groups = {}
    for pr in partial_rows:
        groups.setdefault(tuple(_canon(pr[i]) for i in key_idx), []).append(pr)
    if not key_idx and not groups:
        groups[()] = []
    out = {}
    for k, prs in groups.items():
        vals = []
        for e in aggrefs:
            fa = pg_agg.FinalAgg(e["orig_aggname"], e.get("orig_aggargtypes") or [])
            argcols = [a["varattno"] - 1 for a in e["args"]]
            for pr in prs:
                fa.accum([pr[c] for c in argcols])
            import harness
            harness.check_extension_accum(e, argcols, prs, fa)   # the library's pgstrom_*_accum
            vals.append((fa.final(), fa.rettype))
        out[k] = vals
    return out


def _own_values(tree, rows):
    """The same query on PostgreSQL's own executor (pg_strom.enabled = off):
    SeqScan qual, grouping, the aggregates' own transition / final functions."""
    from oracle import pg_agg
    agg = tree
    scan = agg["lefttree"]
    keys = [t["expr"] for t in agg["targetlist"] if t["expr"]["node"] == "Var"]
    aggrefs = [t["expr"] for t in agg["targetlist"] if t["expr"]["node"] == "Aggref"]
    groups = {}
    for r in rows:
        if not all(pg_expr.evaluate(q, r) is True for q in scan["qual"]):
            continue
        k = tuple(_canon(pg_expr.evaluate(e, r)) for e in keys)
        st = groups.get(k)
        if st is None:
            st = groups[k] = [pg_agg.make_pg_agg(e["aggname"], e["aggargtypes"]) for e in aggrefs]
        for e, a in zip(aggrefs, st):
            if e.get("aggfilter") is not None and pg_expr.evaluate(e["aggfilter"], r) is not True:
                continue
            vals = [pg_expr.evaluate(t["expr"], r) for t in e["args"]]
            if vals and any(v is None for v in vals):
                continue                    # strict transition functions
            a.accum(*vals)
    if not keys and not groups:
        groups[()] = [pg_agg.make_pg_agg(e["aggname"], e["aggargtypes"]) for e in aggrefs]
    return {k: [(a.final(), a.rettype) for a in st] for k, st in groups.items()}


def test_end_to_end_against_postgres_own_aggregates(simdir, lib):
    """Original query on PostgreSQL's own executor == rewritten plan: device
    partial rows (host simulation) + host partial rows of re-checked input
    rows, merged by the pgstrom.* final aggregates.  Covers the rewrite of
    every aggregate of the catalogue (gpupreagg.c:134-333) including FILTER
    clauses, which the regression suite does not use."""
    from oracle import partial
    from oracle.pg_agg import PgError as AggError
    rng = random.Random(4242)
    t = TBL
    # values that cannot make PostgreSQL itself raise: the comparison is about
    # results, not errors
    safe = {"s2": lambda: rng.randrange(-300, 300), "i4": lambda: rng.randrange(-10 ** 6, 10 ** 6),
            "i8": lambda: rng.choice([rng.randrange(-10 ** 12, 10 ** 12), 2 ** 62, -2 ** 62]),
            "f4": lambda: f4(rng.randrange(-4000, 4000) / 8.0),
            "f8": lambda: rng.choice([rng.randrange(-10 ** 6, 10 ** 6) / 64.0, float("nan"), 2.5])}
    rows = []
    for _ in range(400):
        r = list(rand_rows(1, rng)[0])
        for name, mk in safe.items():
            i = t.colnames().index(name)
            r[i] = None if rng.random() < 0.1 else mk()
        rows.append(tuple(r))
    nqueries = 0
    for _ in range(40):
        aggs = [(P.Agg("count", star=True,
                       filter=gen("bool", 1, rng) if rng.random() < 0.5 else None), "count")]
        for _ in range(rng.choice([1, 2, 3])):
            col = rng.choice(["s2", "i4", "i8", "f4", "f8"])
            fns = ["count", "min", "max", "avg"] + ([] if col == "i8" else ["sum"]) + \
                (["stddev", "variance", "var_pop", "stddev_pop"] if col in ("f4", "f8") else [])
            fn = rng.choice(fns)
            flt = None
            if rng.random() < 0.5:
                flt = P.Op(rng.choice(CMP), t.col(rng.choice(["i4", "s2", "k"])),
                           P.Const("int4", rng.choice([0, 3, -100])))
            aggs.append((P.Agg(fn, [t.col(col)], filter=flt), fn))
        if rng.random() < 0.3:
            aggs.append((P.Agg(rng.choice(["corr", "covar_pop", "covar_samp"]),
                               [t.col("f8"), t.col("f4")]), "corr"))
        keycol = rng.choice([None, "k", "b", "s2"])
        where = [P.Op(">", t.col("i4"), P.Const("int4", -500000))] if rng.random() < 0.5 else []
        tree = P.make_agg_plan(t, ([(t.col(keycol), keycol)] if keycol else []) + aggs,
                               group_by=[keycol] if keycol else [], where=where, num_groups=8)
        plan = gp.Plan(tree, gucs=GUCS)
        try:
            if plan.num_gpupreagg != 1:
                continue
            desc = plan.describe()
            node = find_node(plan.tree())
            so = build_sim(plan, simdir)
            so.agg_row.argtypes = [C.POINTER(C.c_uint64), C.c_uint, C.c_char_p, C.c_char_p]
            so.agg_flush.argtypes = [C.c_int, C.c_void_p]
            kparams = plan.kparams()
            incols = desc["incol_index"]
            coltypes = [ty for _, ty in t.columns]
            cols = desc["columns"]
            vals = (C.c_uint64 * max(1, len(incols)))()
            so.agg_reset()
            host_rows = []
            for row in rows:
                valid, toast = fill_vals(row, incols, coltypes, vals)
                rc = so.agg_row(vals, valid, kparams, toast)
                assert rc in (0, 0x100, CPU_RECHECK), rc
                if rc == CPU_RECHECK:       # gpupreagg_next_tuple_fallback
                    host_rows.append(row)
            ncols = len(cols)
            colmeta = (gp.kern_colmeta * ncols)()
            lib.pgs_plan_result_colmeta(plan.handle, 0, colmeta, ncols)
            length = lib.pgstrom_kds_tupslot_length(ncols, 4096)
            buf = C.create_string_buffer(length)
            gp.check(lib.pgstrom_kds_tupslot_init(buf, length, ncols, colmeta, 4096))
            assert so.agg_flush(0 if not keycol else 1, buf) == 0
            kds = gp.kern_data_store.from_buffer(buf)
            values = (C.c_uint64 * ncols)()
            isnull = C.create_string_buffer(ncols)
            prow_list = []
            for r in range(kds.nitems):
                gp.check(lib.pgstrom_fetch_data_store(buf, r, values, isnull))
                prow_list.append(tuple(gp.decode_datum(values[i], isnull.raw[i] != 0,
                                                       cols[i]["type"]) for i in range(ncols)))
            try:
                hp, _ = partial.partial_rows(node, host_rows, len(t.columns))
                own = _own_values(tree, rows)
            except AggError:
                continue        # PostgreSQL itself raises (float overflow, date range ...)
            if not (not keycol and not host_rows):
                prow_list += [tuple(v) for v in hp.values()]
            got = _final_values(desc, prow_list)
            assert set(got) == set(own), (sorted(map(repr, got)), sorted(map(repr, own)))
            for k, ovals in own.items():
                for (gv, gt), (ov, ot), (e, name) in zip(got[k], ovals, aggs):
                    what = (k, name, gv, ov, "\n".join(plan.explain()[:4]))
                    if gv is None or ov is None:
                        assert gv is None and ov is None, what
                    elif isinstance(ov, float) or isinstance(gv, float):
                        gv, ov = float(gv), float(ov)
                        if math.isnan(ov) or math.isnan(gv):
                            assert math.isnan(ov) and math.isnan(gv), what
                        else:
                            tol = 2e-3 if "float4" in (gt, ot) else 1e-9
                            assert abs(gv - ov) <= tol * max(abs(gv), abs(ov), 1e-30), what
                    else:
                        assert gv == ov, what
            nqueries += 1
        finally:
            plan.free()
    assert nqueries >= 25


def test_flush_splits_what_does_not_fit_a_column(simdir, lib):
    """nrows is int4 in the catalogue and psum(int8) is int8: a count above
    2^31-1 and a 128-bit sum beyond int8 leave the device as several partial
    rows that add up to the exact value (PostgreSQL's final aggregates sum
    partial rows).  No test input reaches such values, so the cells are set
    directly."""
    from oracle import bench_oracle
    t = TBL
    tree = P.make_agg_plan(t, [(P.Agg("count", star=True), "count"),
                               (P.Agg("avg", [t.col("i8")]), "avg"),
                               (P.Agg("max", [t.col("i4")]), "max")])
    plan = gp.Plan(tree, gucs=GUCS)
    try:
        desc = plan.describe()
        cols = desc["columns"]
        so = build_sim(plan, simdir)
        so.agg_flush.argtypes = [C.c_int, C.c_void_p]
        so.agg_poke.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint]
        by_text = {c["text"]: c for c in cols if c["role"] == 2}
        nrows = by_text["pgstrom.nrows()"]
        nrows_i8 = by_text["pgstrom.nrows((i8 IS NOT NULL))"]
        psum = by_text["pgstrom.psum(i8)"]
        pmax = by_text["pgstrom.pmax(i4)"]
        assert psum["cell_type"] == "LONG"
        ncols = len(cols)
        colmeta = (gp.kern_colmeta * ncols)()
        lib.pgs_plan_result_colmeta(plan.handle, 0, colmeta, ncols)
        length = lib.pgstrom_kds_tupslot_length(ncols, 64)
        values = (C.c_uint64 * ncols)()
        isnull = C.create_string_buffer(ncols)
        for count, total in ((5_000_000_000, 3 * (2 ** 63 - 1) + 12345),
                             (2 ** 31 - 1, 2 ** 63 - 1),
                             (2 ** 31, -(2 ** 63) - 1),
                             (7, -5 * 2 ** 63)):
            so.agg_reset()
            allbits = 0xffffffff
            so.agg_poke(0, nrows["cell_index"], count, allbits)
            so.agg_poke(0, nrows_i8["cell_index"], count, allbits)
            so.agg_poke(0, psum["cell_index"], total & (2 ** 64 - 1), allbits)
            so.agg_poke(0, psum["cell_index"] + 1, (total >> 64) & (2 ** 64 - 1), allbits)
            so.agg_poke(0, pmax["cell_index"], 42, allbits)
            buf = C.create_string_buffer(length)
            gp.check(lib.pgstrom_kds_tupslot_init(buf, length, ncols, colmeta, 64))
            assert so.agg_flush(0, buf) == 0
            kds = gp.kern_data_store.from_buffer(buf)
            rows = []
            for r in range(kds.nitems):
                gp.check(lib.pgstrom_fetch_data_store(buf, r, values, isnull))
                rows.append(tuple(gp.decode_datum(values[i], isnull.raw[i] != 0, cols[i]["type"])
                                  for i in range(ncols)))
            want_rows = max(-(-count // (2 ** 31 - 1)), -(-abs(total) // (2 ** 63 - 1 if total >= 0 else 2 ** 63)), 1)
            assert len(rows) == want_rows, (count, total, len(rows), want_rows)
            for r in rows:          # every piece fits its column
                assert 0 <= r[nrows["resno"] - 1] <= 2 ** 31 - 1
                assert -(2 ** 63) <= r[psum["resno"] - 1] <= 2 ** 63 - 1
            merged = bench_oracle.combine_device_rows(desc, rows)[()]
            assert merged[nrows["resno"] - 1] == count
            assert merged[nrows_i8["resno"] - 1] == count
            assert merged[psum["resno"] - 1] == total
            assert merged[pmax["resno"] - 1] == 42
    finally:
        plan.free()


def test_aggregation_numeric(simdir, lib):
    """NUMERIC through the merge rules and the flush: 128-bit sums at a fixed
    scale with the largest display scale seen, 57-bit min / max, sums that
    leave the device as several pieces; what does not fit (more than 17
    digits, scale beyond 16 in a sum / 32 anywhere) is the host's row."""
    from decimal import Decimal
    rng = random.Random(31)
    t = P.Table("nm", [("k", "int4"), ("n", "numeric"), ("m", "numeric"), ("i", "int4")])
    rows = []
    for _ in range(600):
        def num():
            if rng.random() < 0.08:
                return None
            scale = rng.choice([0, 0, 1, 2, 2, 4, 6, 10, 16, 17, 20])
            digits = rng.choice([1, 3, 6, 9, 12, 15, 17, 18])
            mant = rng.randrange(0, 10 ** digits)
            d = Decimal(mant).scaleb(-scale)
            return -d if rng.random() < 0.4 else d
        rows.append((rng.randrange(0, 6), num(), num(), rng.randrange(-1000, 1000)))
    n, m, i = t.col("n"), t.col("m"), t.col("i")
    cnt = (P.Agg("count", star=True), "count")
    total_ok = total_re = 0
    for keycol in (None, "k"):
        for targets, where in (
                ([cnt, (P.Agg("sum", [n]), "sum"), (P.Agg("avg", [m]), "avg")], []),
                ([cnt, (P.Agg("min", [n]), "min"), (P.Agg("max", [n]), "max"),
                  (P.Agg("count", [m]), "count")], [P.Op(">", n, m)]),
                ([cnt, (P.Agg("sum", [P.Op("+", n, P.Cast(i, "numeric"))]), "sum"),
                  (P.Agg("max", [P.Op("*", m, P.Const("numeric", "1.5"))]), "max")],
                 [P.Op("<>", m, P.Const("numeric", "0"))])):
            tree = P.make_agg_plan(t, ([(t.col(keycol), keycol)] if keycol else []) + targets,
                                   group_by=[keycol] if keycol else [], where=where, num_groups=8)
            nok, nre = check_aggregation(t, tree, rows, simdir, lib)
            total_ok += nok
            total_re += nre
    assert total_ok > 800 and total_re > 300


def test_flush_numeric_sum_pieces(simdir, lib):
    """A 128-bit numeric sum leaves as at most three device numerics (base
    10^17) that add up exactly and keep the display scale - also a sum like
    10^12 + 10^-16, whose 29 digits no single 57-bit mantissa holds."""
    from decimal import Decimal
    from oracle import bench_oracle
    from oracle.pg_agg import numeric_add, numeric_out
    t = P.Table("nm", [("n", "numeric")])
    tree = P.make_agg_plan(t, [(P.Agg("sum", [t.col("n")]), "sum")])
    plan = gp.Plan(tree, gucs=GUCS)
    try:
        desc = plan.describe()
        cols = desc["columns"]
        so = build_sim(plan, simdir)
        so.agg_flush.argtypes = [C.c_int, C.c_void_p]
        so.agg_poke.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_uint]
        psum = [c for c in cols if c["role"] == 2][0]
        assert psum["cell_type"] == "NUMERIC" and psum["op"] == "PSUM"
        ncols = len(cols)
        colmeta = (gp.kern_colmeta * ncols)()
        lib.pgs_plan_result_colmeta(plan.handle, 0, colmeta, ncols)
        length = lib.pgstrom_kds_tupslot_length(ncols, 16)
        values = (C.c_uint64 * ncols)()
        isnull = C.create_string_buffer(ncols)
        for text, npieces in (("0", 1), ("12.50", 1), ("-0.0001", 1), ("99999999999999999", 1),
                              ("100000000000000000", 2), ("1000000000000.0000000000000001", 2),
                              ("-1000000000000.0000000000000001", 2),
                              ("12345678901234567890.12", 2),
                              ("1234567890123456789012.3456789012345678", 3),
                              ("-9999999999999999999999.9999999999999999", 3)):
            d = Decimal(text)
            ds = max(0, -d.as_tuple().exponent)
            v = int(d.scaleb(16))                   # the cell keeps the sum at scale 16
            so.agg_reset()
            so.agg_poke(0, psum["cell_index"], v & (2 ** 64 - 1), 0xffffffff)
            so.agg_poke(0, psum["cell_index"] + 1, (v >> 64) & (2 ** 64 - 1), 0xffffffff)
            so.agg_poke(0, psum["cell_index"] + 2, ds, 0xffffffff)
            buf = C.create_string_buffer(length)
            gp.check(lib.pgstrom_kds_tupslot_init(buf, length, ncols, colmeta, 16))
            assert so.agg_flush(0, buf) == 0
            kds = gp.kern_data_store.from_buffer(buf)
            assert kds.nitems == npieces, (text, kds.nitems)
            total = None
            for r in range(kds.nitems):
                gp.check(lib.pgstrom_fetch_data_store(buf, r, values, isnull))
                piece = gp.decode_datum(values[psum["resno"] - 1], isnull.raw[psum["resno"] - 1] != 0,
                                        "numeric")
                # what PostgreSQL's numeric_in makes of the piece, then its sum
                piece = Decimal(piece) if piece.as_tuple().exponent <= 0 else \
                    Decimal(int(piece))
                total = piece if total is None else numeric_add(total, piece)
            assert total == d and numeric_out(total) == numeric_out(d), (text, total)
    finally:
        plan.free()
