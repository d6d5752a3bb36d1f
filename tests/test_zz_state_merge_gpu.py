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


@pytest.mark.parametrize("name,plan_kw,col_kw", [
    ("nogrp_agg", {}, {}),
    ("where_agg", {}, {}),
    ("where_agg", {}, {"with_nulls": True}),
])
def test_peer_merge_three_ranks_one_device(cuda, name, plan_kw, col_kw):
    """pgs_preagg_merge_peer (gpupreagg_peer_push / _pull): three sessions act
    as the ranks of one merge on ONE device - the exchange area is plain device
    memory there, over NVLink it is the root's HBM mapped into the peers.  Four
    scans in a row: the areas' two buffers per rank are re-used, the epochs
    keep them apart, and the root returns the partial rows of all shards."""
    lib = gp._capi.load()
    plan = gp.Plan(W.WORKLOADS[name]["plan"](**plan_kw), gucs=GUCS)
    node = plan.tree()["lefttree"]
    ranks = [gp.Session(plan) for _ in range(3)]
    held = []
    try:
        for r, s in enumerate(ranks):
            gp.check(lib.pgs_preagg_peer_setup(s.handle, r, 3, 0, None))
        for s in ranks[1:]:
            gp.check(lib.pgs_preagg_peer_attach_session(s.handle, ranks[0].handle))
        for scan in range(4):
            shards = []
            for r, s in enumerate(ranks):
                n = 200_000 + 40 * r + 4 * scan
                ds, cols = W.make_chunk(name, 4_000_000 * scan + 1_000_000 * r, n, **col_kw)
                held.append(ds)
                shards.append(cols)
                s.submit(ds)
            # the pushes are queued before the root's pull: on one device the
            # pull would otherwise wait for kernels that cannot start
            for s in ranks[1:]:
                gp.check(lib.pgs_preagg_merge_peer(s.handle))
            gp.check(lib.pgs_preagg_merge_peer(ranks[0].handle))
            rows = ranks[0].finish()
            bench_oracle.assert_partial_equal_node(plan.describe(), node, rows, _concat(shards))
            for s in ranks[1:]:
                # the whole state went to the root and was reset by the push
                # kernel: nothing to flush (no flush kernel, no read-back)
                launches = s.launch_count()
                assert s.finish() == []
                assert s.launch_count() == launches
    finally:
        for s in ranks:
            s.close()
        for ds in held:
            ds.free()
        plan.free()


def _without_rows(cols, drop):
    keep = np.ones(len(cols[0][0]), bool)
    keep[np.asarray(drop, dtype=np.int64)] = False
    return [(v[keep], None if m is None else m[keep]) for v, m in cols]


@pytest.mark.parametrize("name,plan_groups,data_groups", [
    ("where_agg", 1000, 60_000),            # CTA-local tables + a global table for 1000
    ("where_agg", 20, 3_000),
    ("high_cardinality", 70_000, 900_000),  # partitions and images sized for 70 K
])
def test_group_estimate_far_too_low(cuda, name, plan_groups, data_groups):
    """The planner's numGroups is an estimate (PostgreSQL's default for an
    expression is 200); the reference never depends on it - every chunk is
    reduced on its own (gpupreagg.c:2169-2186).  Here the tables are sized
    from it, so a scan that meets 45-60 times the groups must still finish:
    what finds no room goes to the overflow log, the table is replaced by a
    larger one between chunks (session_grow_table), and rows that find no room
    at all are handed to the host as re-check rows - never an error."""
    plan = gp.Plan(W.WORKLOADS[name]["plan"](num_groups=plan_groups), gucs=GUCS)
    node = plan.tree()["lefttree"]
    sess = gp.Session(plan)
    held, shards = [], []
    try:
        for i in range(4):
            ds, cols = W.make_chunk(name, 2_000_000 * i, 400_000, num_groups=data_groups)
            held.append(ds)
            t = sess.submit(ds)
            st = sess.wait(t)
            assert st in (0, 2), st            # StromError_CpuReCheck = 2
            drop = sess.recheck_rows(t) if st == 2 else []
            shards.append(_without_rows(cols, drop) if drop else cols)
        rows = sess.finish()
        pm = sess.perfmon()
        assert pm["num_table_grown"] >= 1
        ng = bench_oracle.assert_partial_equal_node(plan.describe(), node, rows, _concat(shards))
        assert ng > 5 * plan_groups
    finally:
        sess.close()
        for ds in held:
            ds.free()
        plan.free()
