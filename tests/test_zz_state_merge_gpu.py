"""-m gpu: the state export / import that carries the multi-GPU merge
(gpupreagg_export / gpupreagg_import, what pgs_preagg_merge_nccl sends over
NVLink) exercised on ONE device: two sessions of the same
query scan different shards, the second session's state is exported as
records and imported into the first, which then returns the partial rows of
both shards - exactly what rank 0 does with the records of rank 1.  Also the
admin JSON views and abort."""
import ctypes as C
import json

import numpy as np
import pytest

from oracle import bench_oracle
from pg_strom_b200 import gpupreagg as gp
from pg_strom_b200 import workloads as W

pytestmark = pytest.mark.gpu
GUCS = {"pg_strom.enabled": "on", "pg_strom.debug_force_gpupreagg": "on"}


def _concat(parts):
    out = []
    for c in range(len(parts[0])):
        v = np.concatenate([p[c][0] for p in parts])
        if all(p[c][1] is None for p in parts):
            m = None
        else:
            m = np.concatenate([np.zeros(len(p[c][0]), np.uint8) if p[c][1] is None else p[c][1]
                                for p in parts])
        out.append((v, m))
    return out


@pytest.mark.parametrize("name,plan_kw,col_kw", [
    ("nogrp_agg", {}, {}),
    ("where_agg", {}, {}),
    ("where_agg", {}, {"with_nulls": True}),
    ("high_cardinality", {"num_groups": 100_000}, {"num_groups": 100_000}),
])
def test_export_import_merges_two_shards(cuda, name, plan_kw, col_kw):
    lib = gp._capi.load()
    plan = gp.Plan(W.WORKLOADS[name]["plan"](**plan_kw), gucs=GUCS)
    a = gp.Session(plan)
    b = gp.Session(plan)
    dbuf = None
    try:
        n_a, n_b = 400_000, 300_004
        ds_a, cols_a = W.make_chunk(name, 0, n_a, **col_kw)
        ds_b, cols_b = W.make_chunk(name, 1_000_000, n_b, **col_kw)
        assert a.wait(a.submit(ds_a)) == 0
        assert b.wait(b.submit(ds_b)) == 0
        # size the record buffer from a first (too small) call
        nrec = C.c_uint32()
        recb = C.c_size_t()
        probe = lib.pgs_device_alloc(0, 4096)
        rc = lib.pgs_preagg_state_export(b.handle, probe, 0, C.byref(nrec), C.byref(recb))
        lib.pgs_device_free(0, probe)
        assert rc in (0, 301) and recb.value > 0
        nbytes = (nrec.value + 64) * recb.value
        dbuf = lib.pgs_device_alloc(0, nbytes)
        assert dbuf
        gp.check(lib.pgs_preagg_state_export(b.handle, dbuf, nbytes, C.byref(nrec), C.byref(recb)))
        assert nrec.value >= 1
        gp.check(lib.pgs_preagg_state_import(a.handle, dbuf, nrec.value))
        rows = a.finish()
        node = plan.tree()["lefttree"]
        bench_oracle.assert_partial_equal_node(plan.describe(), node, rows,
                                               _concat([cols_a, cols_b]))
        # the exporting session keeps its own state
        rows_b = b.finish()
        bench_oracle.assert_partial_equal_node(plan.describe(), node, rows_b, cols_b)
        ds_a.free()
        ds_b.free()
    finally:
        if dbuf:
            lib.pgs_device_free(0, dbuf)
        a.close()
        b.close()
        plan.free()


def test_admin_views_and_abort(cuda):
    """pgstrom_opencl_device_info() / pgstrom_opencl_program_info()
    (pg_strom--1.0.sql:47-72) as JSON; abort waits for the in-flight chunks
    (restrack.c:180-254: nothing the device still reads may be freed) and
    retires the session."""
    lib = gp._capi.load()
    devs = json.loads(lib.pgs_cuda_device_info_json().decode())
    assert len(devs) >= 1 and "B200" in json.dumps(devs[0])
    plan = gp.Plan(W.where_plan(), gucs=GUCS)
    sess = gp.Session(plan)
    try:
        progs = json.loads(lib.pgs_program_info_json().decode())
        assert len(progs) >= 1
        ds, cols = W.make_chunk("where_agg", 0, 200_000)
        sess.submit(ds)
        sess.submit(ds)
        lib.pgs_preagg_abort(sess.handle)
        # abort returns only when the device no longer reads the chunks: the
        # caller may free them at once; the session takes no more work
        ds.free()
        ds2, _ = W.make_chunk("where_agg", 0, 1000)
        t = C.c_int64()
        rc = lib.pgs_preagg_submit(sess.handle, ds2.ptr, None, C.byref(t))
        assert rc != 0
        assert b"aborted" in lib.pgs_last_error()
        ds2.free()
    finally:
        sess.close()
        plan.free()
