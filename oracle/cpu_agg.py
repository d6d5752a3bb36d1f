"""ctypes wrapper of oracle/cpu_agg.c (PostgreSQL-style Agg over SeqScan in C).

TEST INFRASTRUCTURE (oracle): checker and reported CPU baseline only.
"""
import ctypes as C
import os
import subprocess
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libcpu_agg.so")

KINDS = {"count_star": 0, "count": 1, "sum_int4": 2, "avg_int4": 3, "avg_int8": 4,
         "min_int4": 5, "max_int4": 6, "sum_float8": 7, "avg_float8": 8,
         "min_float8": 9, "max_float8": 10, "var_float8": 11}
COLTYPE = {"int4": 4, "int8": 8, "float8": 108}


class cpu_table(C.Structure):
    _fields_ = [("ncols", C.c_int32), ("coltype", C.c_int32 * 8),
                ("values", C.c_void_p * 8), ("nulls", C.c_void_p * 8),
                ("nrows", C.c_int64)]


class cpu_query(C.Structure):
    _fields_ = [("qual_col", C.c_int32), ("qual_const", C.c_int32),
                ("key_col", C.c_int32), ("naggs", C.c_int32),
                ("agg_kind", C.c_int32 * 16), ("agg_col", C.c_int32 * 16),
                ("partitioned", C.c_int32), ("pad", C.c_int32)]


class cpu_heap(C.Structure):
    _fields_ = [("ncols", C.c_int32), ("coltype", C.c_int32 * 8),
                ("pages", C.c_void_p), ("npages", C.c_int64)]


class agg_state(C.Structure):
    _fields_ = [("n", C.c_int64), ("isum_lo", C.c_int64), ("isum_hi", C.c_int64),
                ("fsum", C.c_double), ("fsum2", C.c_double),
                ("imin", C.c_int64), ("imax", C.c_int64),
                ("fmin", C.c_double), ("fmax", C.c_double),
                ("has_value", C.c_int32), ("pad", C.c_int32)]


_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        _lib = C.CDLL(LIB)
        _lib.cpu_agg_run.restype = C.c_int64
        _lib.cpu_agg_run.argtypes = [C.POINTER(cpu_table), C.POINTER(cpu_query), C.c_int,
                                     C.POINTER(C.c_int64), C.POINTER(agg_state), C.c_int64]
        _lib.cpu_agg_run_heap.restype = C.c_int64
        _lib.cpu_agg_run_heap.argtypes = [C.POINTER(cpu_heap), C.POINTER(cpu_query), C.c_int,
                                          C.POINTER(C.c_int64), C.POINTER(agg_state), C.c_int64]
        _lib.cpu_heap_form_pages.restype = C.c_int64
        _lib.cpu_heap_form_pages.argtypes = [C.POINTER(cpu_table), C.c_void_p, C.c_int64, C.c_int64,
                                             C.POINTER(C.c_int64)]
        assert _lib.cpu_agg_state_size() == C.sizeof(agg_state)
    return _lib


# the bench queries (pg_strom_b200/workloads.py) as cpu_query
QUERIES = {
    "nogrp_agg": dict(coltypes=["int4", "float8"], qual=None, key=None, aggs=[
        ("count_star", 0), ("count", 0), ("sum_int4", 0), ("avg_int4", 0),
        ("min_int4", 0), ("max_int4", 0), ("sum_float8", 1), ("avg_float8", 1),
        ("min_float8", 1), ("max_float8", 1)]),
    "where_agg": dict(coltypes=["int4", "int4", "float8", "int4"], qual=(0, 10), key=1, aggs=[
        ("count_star", 0), ("sum_int4", 3), ("avg_float8", 2), ("min_float8", 2),
        ("max_float8", 2)]),
    "high_cardinality": dict(coltypes=["int8", "int8", "float8"], qual=None, key=0, partitioned=True, aggs=[
        ("count_star", 0), ("avg_int8", 1), ("avg_float8", 2), ("var_float8", 2)]),
}


def _table(qd, cols):
    t = cpu_table()
    keep = []
    t.ncols = len(cols)
    t.nrows = len(cols[0][0])
    for i, (v, m) in enumerate(cols):
        a = np.ascontiguousarray(v)
        keep.append(a)
        t.coltype[i] = COLTYPE[qd["coltypes"][i]]
        t.values[i] = a.ctypes.data
        if m is not None:
            mm = np.ascontiguousarray(m, dtype=np.uint8)
            keep.append(mm)
            t.nulls[i] = mm.ctypes.data
    return t, keep


def _query(qd, qual_const):
    q = cpu_query()
    q.qual_col = -1 if qd["qual"] is None else qd["qual"][0]
    q.qual_const = 0 if qd["qual"] is None else (qual_const if qual_const is not None else qd["qual"][1])
    q.key_col = -1 if qd["key"] is None else qd["key"]
    q.naggs = len(qd["aggs"])
    q.partitioned = 1 if qd.get("partitioned") else 0
    for j, (k, c) in enumerate(qd["aggs"]):
        q.agg_kind[j] = KINDS[k]
        q.agg_col[j] = c
    return q


def _result_buffers(max_groups, naggs):
    """keys[max_groups], states[max_groups * naggs] as ctypes views of
    UNTOUCHED memory (the C side writes the groups it finds): a zero-filled
    ctypes array of 2^24 groups x 10 aggregates is 13 GB of page faults."""
    kbuf = np.empty(max_groups, dtype=np.int64)
    sbuf = np.empty(max_groups * naggs * C.sizeof(agg_state), dtype=np.uint8)
    keys = (C.c_int64 * max_groups).from_buffer(kbuf)
    states = (agg_state * (max_groups * naggs)).from_buffer(sbuf)
    return keys, states


def run(name, cols, nthreads=1, max_groups=1 << 24, qual_const=None):
    """cols: [(values ndarray, nullmask|None)...].  Returns (seconds, keys,
    states ndarray-of-struct, ngroups)."""
    lib = load()
    qd = QUERIES[name]
    t, keep = _table(qd, cols)
    q = _query(qd, qual_const)
    keys, states = _result_buffers(max_groups, q.naggs)
    t0 = time.perf_counter()
    ng = lib.cpu_agg_run(C.byref(t), C.byref(q), nthreads, keys, states, max_groups)
    dt = time.perf_counter() - t0
    return dt, keys, states, int(ng)


def form_heap_pages(name, cols):
    """The table as PostgreSQL heap pages (heap_form_tuple + PageAddItem in C):
    returns (uint8 ndarray of npages * 8192 bytes, npages)."""
    lib = load()
    qd = QUERIES[name]
    t, keep = _table(qd, cols)
    width = 24 + 8 + sum(8 if c != "int4" else 4 for c in qd["coltypes"]) + 8 + 4
    per_page = max(1, (8192 - 24) // width)
    maxpages = int(t.nrows) // per_page + 2
    pages = np.zeros(maxpages * 8192, dtype=np.uint8)
    done = C.c_int64()
    npages = int(lib.cpu_heap_form_pages(C.byref(t), pages.ctypes.data, maxpages, 0, C.byref(done)))
    assert done.value == t.nrows, "pages estimate too small: %d of %d rows" % (done.value, t.nrows)
    return pages, npages


def run_heap(name, pages, npages, nthreads=1, max_groups=1 << 24, qual_const=None):
    """The same query over heap pages (address or uint8 ndarray): every tuple
    is de-formed like slot_deform_tuple.  Returns like run()."""
    lib = load()
    qd = QUERIES[name]
    h = cpu_heap()
    h.ncols = len(qd["coltypes"])
    for i, c in enumerate(qd["coltypes"]):
        h.coltype[i] = COLTYPE[c]
    h.pages = pages.ctypes.data if hasattr(pages, "ctypes") else int(pages)
    h.npages = npages
    q = _query(qd, qual_const)
    keys, states = _result_buffers(max_groups, q.naggs)
    t0 = time.perf_counter()
    ng = lib.cpu_agg_run_heap(C.byref(h), C.byref(q), nthreads, keys, states, max_groups)
    dt = time.perf_counter() - t0
    return dt, keys, states, int(ng)
