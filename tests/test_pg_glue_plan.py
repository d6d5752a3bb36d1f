"""pg_glue/gpupreagg_glue.c - _PG_init, GUC registration, planner hook
(plan tree -> JSON -> pgstrom_grafter_json -> plan nodes) and the CustomPlan
callbacks of "GpuPreAgg" (reference: main.c:237-281, grafter.c:119-157,
gpupreagg.c:2189-2979) - compiled against a stand-in for the PostgreSQL
headers (tests/native/pg_stub/) and driven by tests/native/pg_glue_driver.c,
which plays PostgreSQL for one query:

    SELECT key, count(*), sum(w), avg(v), min(v), max(v)
      FROM bench_where WHERE f < 10 GROUP BY key

CPU: the serialised plan is what pgplan.py builds for the same query (so the
C serialiser speaks the format every planner test is written in), the hook
splices a Custom(GpuPreAgg) node with the reference's target list shape, looks
up the pgstrom.* catalog entries of pg_strom--1.0.sql, GUC assignments reach
the library, and BeginCustomPlan reports the missing device.  GPU: the
rewritten plan is executed (twice: ReScan) and its partial rows are checked
against the oracle."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from pg_strom_b200 import gpupreagg as gp
from pg_strom_b200 import workloads as W

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STUB = os.path.join(HERE, "native", "pg_stub")


@pytest.fixture(scope="module")
def glue(lib):
    out = os.path.join(HERE, "native", "_pg_glue_plan.so")
    subprocess.run(["gcc", "-std=gnu11", "-Wall", "-Werror", "-O1", "-fPIC", "-shared",
                    "-I", STUB, "-I", os.path.join(ROOT, "include"), "-o", out,
                    os.path.join(ROOT, "pg_glue", "gpupreagg_glue.c"),
                    os.path.join(HERE, "native", "pg_glue_driver.c"),
                    os.path.join(STUB, "pg_glue_stub.c"), os.path.join(STUB, "pg_stub.c"),
                    "-L", os.path.join(ROOT, "pg_strom_b200"), "-lpgstrom_cuda",
                    "-Wl,-rpath," + os.path.join(ROOT, "pg_strom_b200")], check=True)
    so = C.CDLL(out)
    so.driver_plan_json.restype = C.c_char_p
    so.driver_plan_json.argtypes = [C.c_int, C.c_long]
    so.driver_run_planner.restype = C.c_char_p
    so.driver_run_planner.argtypes = [C.c_int, C.c_long]
    so.driver_error.restype = C.c_char_p
    so.pg_stub_function_signature.restype = C.c_char_p
    so.pg_stub_guc_name.restype = C.c_char_p
    so.pg_stub_set_guc.argtypes = [C.c_char_p, C.c_char_p]
    so.pg_stub_explain_text.restype = C.c_char_p
    so.pg_stub_notice.restype = C.c_char_p
    so.driver_execute.restype = C.c_long
    so.driver_execute.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_long, C.POINTER(C.c_int)]
    lib.pgstrom_guc_reset_all()
    assert so.driver_init() == 0, so.driver_error()
    yield so
    lib.pgstrom_guc_reset_all()


def _strip(node):
    """The keys the serialiser adds for the cost model, and the ones pgplan.py
    adds for EXPLAIN's benefit, are not part of the comparison."""
    if isinstance(node, dict):
        return {k: _strip(v) for k, v in node.items()
                if k not in ("startup_cost", "total_cost", "plan_rows", "plan_width",
                             "groupkeys", "sortkeys", "varno")}
    if isinstance(node, list):
        return [_strip(v) for v in node]
    return node


def test_plan_to_json_is_the_harness_format(glue):
    got = json.loads(glue.driver_plan_json(10, 1000).decode())
    want = W.where_plan()
    assert got["plan_rows"] == 1000 and got["lefttree"]["total_cost"] == 2000000.0
    # (the Aggref arguments are TargetEntries without a name in both; after
    # set_plan_references() the Vars of the Agg are OUTER_VAR Vars - "varno":
    # "OUTER" - where the harness writes scan-level Vars: same column either way)
    assert got["targetlist"][0]["expr"]["varno"] == "OUTER"
    assert _strip(got) == _strip(json.loads(json.dumps(want)))


def test_guc_registration_and_assignment(glue, lib):
    names = {glue.pg_stub_guc_name(i).decode() for i in range(glue.pg_stub_num_gucs())}
    assert {"pg_strom.enabled", "pg_strom.perfmon", "enable_gpupreagg",
            "pg_strom.debug_force_gpupreagg", "pg_strom.chunk_size",
            "pg_strom.max_async_chunks", "gpu_setup_cost", "gpu_operator_cost",
            "gpu_tuple_cost", "pg_strom.show_device_kernel",
            "pg_strom.devprog_enable_optimization"} <= names
    # every one of them is a GUC of the library's table (main.c:104-234)
    for n in names:
        assert lib.pgstrom_guc_get(n.encode()) is not None, n
    assert glue.pg_stub_set_guc(b"pg_strom.chunk_size", b"32") == 0
    assert lib.pgstrom_guc_get(b"pg_strom.chunk_size") == b"32"
    assert glue.pg_stub_set_guc(b"pg_strom.chunk_size", b"15") == 0


def test_planner_hook_splices_gpupreagg(glue):
    assert glue.pg_stub_set_guc(b"pg_strom.enabled", b"on") == 0
    assert glue.pg_stub_set_guc(b"pg_strom.debug_force_gpupreagg", b"on") == 0
    shape = glue.driver_run_planner(10, 1000).decode()
    # Agg (6 columns) -> GpuPreAgg (4 table columns as NULL / key + 8 partial
    # columns) -> the scan, whose qual moved into the kernel
    plan = gp.Plan(W.where_plan(), gucs={"pg_strom.enabled": "on",
                                         "pg_strom.debug_force_gpupreagg": "on"})
    npartial = len(plan.tree()["lefttree"]["targetlist"])
    plan.free()
    assert shape == "Agg[6] -> Custom(GpuPreAgg)[%d] -> SeqScan[4] quals=0" % npartial, shape
    sigs = []
    i = 0
    while glue.pg_stub_function_signature(i):
        sigs.append(glue.pg_stub_function_signature(i).decode())
        i += 1
    asked = sorted(s for s in sigs if s.startswith("pgstrom."))
    # the placeholders and final aggregates of pg_strom--1.0.sql this query needs
    assert asked == sorted(["pgstrom.nrows()", "pgstrom.nrows(bool)", "pgstrom.psum(int8)",
                            "pgstrom.psum(float8)", "pgstrom.pmin(float8)", "pgstrom.pmax(float8)",
                            "pgstrom.sum(int8)", "pgstrom.avg(int4,float8)"]), asked
    # pg_strom.enabled = off: the plan stays PostgreSQL's
    assert glue.pg_stub_set_guc(b"pg_strom.enabled", b"off") == 0
    assert glue.driver_run_planner(10, 1000).decode() == "Agg[6] -> SeqScan[4] quals=1"
    assert glue.pg_stub_set_guc(b"pg_strom.enabled", b"on") == 0


def test_cost_decision(glue):
    """gpupreagg.c:2105-2118: unless pg_strom.debug_force_gpupreagg is on, the
    plan is only rewritten when the Agg over GpuPreAgg (the library's
    cost_gpupreagg numbers + PostgreSQL's cost_agg / cost_sort on top) comes
    out cheaper than PostgreSQL's own plan."""
    assert glue.pg_stub_set_guc(b"pg_strom.enabled", b"on") == 0
    assert glue.pg_stub_set_guc(b"pg_strom.debug_force_gpupreagg", b"off") == 0
    try:
        # 12.5 M rows into 1000 groups: the device plan is cheaper
        assert "GpuPreAgg" in glue.driver_run_planner(10, 1000).decode()
        # an absurd set-up cost: PostgreSQL's plan stays, the scan keeps its qual
        assert glue.pg_stub_set_guc(b"gpu_setup_cost", b"5e9") == 0
        assert glue.driver_run_planner(10, 1000).decode() == "Agg[6] -> SeqScan[4] quals=1"
        # ... unless forced
        assert glue.pg_stub_set_guc(b"pg_strom.debug_force_gpupreagg", b"on") == 0
        assert "GpuPreAgg" in glue.driver_run_planner(10, 1000).decode()
    finally:
        assert glue.pg_stub_set_guc(b"gpu_setup_cost", b"500") == 0
        assert glue.pg_stub_set_guc(b"pg_strom.debug_force_gpupreagg", b"on") == 0


def test_abort_callback_is_registered(glue):
    """restrack.c:180-254: a transaction that aborts while GpuPreAgg states are
    open must not leak device work.  The glue registers a resource-owner
    release callback in _PG_init that ends every open state; with none open
    (no device here) an abort is a no-op."""
    assert glue.pg_stub_abort_transaction() == 1        # one callback: the glue's
    assert "GpuPreAgg" in glue.driver_run_planner(10, 1000).decode()    # and the backend lives on


def _table(nrows):
    cols = W.where_columns(0, nrows)
    vals = np.zeros((nrows, 4), np.uint64)
    for c, (v, m) in enumerate(cols):
        if v.dtype == np.float64:
            vals[:, c] = v.view(np.uint64)
        else:
            vals[:, c] = v.astype(np.int64).view(np.uint64) & np.uint64(0xffffffff)
    nulls = np.zeros((nrows, 4), np.bool_)
    return cols, np.ascontiguousarray(vals), np.ascontiguousarray(nulls)


def test_begin_without_a_device_raises(glue):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    assert glue.pg_stub_set_guc(b"pg_strom.enabled", b"on") == 0
    assert "GpuPreAgg" in glue.driver_run_planner(10, 1000).decode()
    cols, vals, nulls = _table(1000)
    out = np.zeros((2000, 16), np.uint64)
    outn = np.zeros((2000, 16), np.bool_)
    ncols = C.c_int()
    n = glue.driver_execute(1000, vals.ctypes.data, nulls.ctypes.data, out.ctypes.data,
                            outn.ctypes.data, 2000, C.byref(ncols))
    assert n == -1
    assert b"PG-Strom" in glue.driver_error()       # ereport(ERROR), not a crash or a CPU result


@pytest.mark.gpu
def test_glue_executes_on_the_device(glue, cuda):
    """The whole PostgreSQL-side path without PostgreSQL: planner hook, then
    BeginCustomPlan / ExecCustomPlan until NULL / ReScanCustomPlan / again /
    ExplainCustomPlan / EndCustomPlan over a SeqScan of 300 K rows."""
    from oracle import bench_oracle
    assert glue.pg_stub_set_guc(b"pg_strom.enabled", b"on") == 0
    assert glue.pg_stub_set_guc(b"pg_strom.debug_force_gpupreagg", b"on") == 0
    assert glue.pg_stub_set_guc(b"pg_strom.chunk_size", b"4") == 0      # several chunks
    assert "GpuPreAgg" in glue.driver_run_planner(10, 1000).decode()
    nrows = 300_000
    cols, vals, nulls = _table(nrows)
    plan = gp.Plan(W.where_plan(), gucs={"pg_strom.enabled": "on",
                                         "pg_strom.debug_force_gpupreagg": "on"})
    desc = plan.describe()
    node = plan.tree()["lefttree"]
    width = len(node["targetlist"])
    out = np.zeros((4000, width), np.uint64)
    outn = np.zeros((4000, width), np.bool_)
    ncols = C.c_int()
    glue.pg_stub_explain_reset()
    n = glue.driver_execute(nrows, vals.ctypes.data, nulls.ctypes.data, out.ctypes.data,
                            outn.ctypes.data, 4000, C.byref(ncols))
    assert n > 0, glue.driver_error()
    assert ncols.value == width
    coltypes = [c["type"] for c in desc["columns"]]
    rows = [tuple(gp.decode_datum(int(out[r, c]), bool(outn[r, c]), coltypes[c], -1)
                  for c in range(width)) for r in range(n)]
    ng = bench_oracle.assert_partial_equal_node(desc, node, rows, cols)
    assert ng == 1000
    assert b"Kernel Source" in glue.pg_stub_explain_text() or b"Bulkload" in glue.pg_stub_explain_text()
    plan.free()
    assert glue.pg_stub_set_guc(b"pg_strom.chunk_size", b"15") == 0
