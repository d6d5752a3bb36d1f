"""-m gpu: the bench workloads through the session API (C ABI), device
partial rows == oracle partial rows, bit-exact (the float columns sit on a
dyadic grid, so their sums do not depend on the summation order)."""
import os

import numpy as np
import pytest

from oracle import bench_oracle
from pg_strom_b200 import gpupreagg as gp
from pg_strom_b200 import workloads as W

pytestmark = pytest.mark.gpu
GUCS = {"pg_strom.enabled": "on", "pg_strom.debug_force_gpupreagg": "on"}


def _run(name, nrows, chunk_rows, plan_kw=None, col_kw=None, session_kw=None):
    plan = gp.Plan(W.WORKLOADS[name]["plan"](**(plan_kw or {})), gucs=GUCS)
    assert plan.num_gpupreagg == 1, plan.reject_reason
    sess = gp.Session(plan, **(session_kw or {}))
    allcols = None
    held = []
    try:
        for r0 in range(0, nrows, chunk_rows):
            n = min(chunk_rows, nrows - r0)
            ds, cols = W.make_chunk(name, r0, n, **(col_kw or {}))
            held.append(ds)
            t = sess.submit(ds)
            assert sess.wait(t) == 0
            if allcols is None:
                allcols = [[c[0]], [c[1]]] if False else [([c[0]], [c[1]]) for c in cols]
            else:
                for a, c in zip(allcols, cols):
                    a[0].append(c[0])
                    a[1].append(c[1])
        rows = sess.finish()
        merged = []
        for vals, masks in allcols:
            v = np.concatenate(vals)
            m = None if all(x is None for x in masks) else np.concatenate(
                [np.zeros(len(vv), np.uint8) if mm is None else mm for vv, mm in zip(vals, masks)])
            merged.append((v, m))
        node = plan.tree()["lefttree"]
        ngroups = bench_oracle.assert_partial_equal_node(plan.describe(), node, rows, merged)
        return ngroups, rows, sess.perfmon()
    finally:
        sess.close()
        for ds in held:
            ds.free()
        plan.free()


def test_nogrp_single_chunk(cuda):
    ng, rows, pm = _run("nogrp_agg", 1_000_000, 1_000_000)
    assert ng == 1 and len(rows) == 1


def test_nogrp_ragged_chunks(cuda):
    # chunk sizes that are not multiples of the tile (2048 rows) nor of 128
    _run("nogrp_agg", 1_234_568, 333_332)


def test_nogrp_no_nulls(cuda):
    _run("nogrp_agg", 500_000, 500_000, col_kw={"with_nulls": False})


def test_nogrp_tiny(cuda):
    for n in (4, 100, 2048, 2052):
        _run("nogrp_agg", n, n)


def test_where_1k_groups(cuda):
    ng, rows, pm = _run("where_agg", 2_000_000, 1_000_000)
    assert ng == 1000
    assert pm["sh_nslots"] > 0          # CTA-local table in use


def test_where_with_nulls(cuda):
    _run("where_agg", 700_000, 700_000, col_kw={"with_nulls": True})


def test_where_selectivity(cuda):
    for pct in (1, 50, 100):
        _run("where_agg", 400_000, 400_000, plan_kw={"selectivity_pct": pct})


def test_where_global_table_only(cuda, monkeypatch):
    monkeypatch.setenv("PGSTROM_SH_SLOTS", "0")
    ng, rows, pm = _run("where_agg", 600_000, 300_000)
    assert pm["sh_nslots"] == 0 and ng == 1000


def test_where_small_local_table_spills(cuda, monkeypatch):
    monkeypatch.setenv("PGSTROM_SH_SLOTS", "256")
    ng, rows, pm = _run("where_agg", 600_000, 300_000)
    assert pm["sh_nslots"] == 256 and ng == 1000


def test_high_cardinality(cuda):
    """Many groups: rows are dealt into partitions, every partition is
    aggregated in shared memory into its persistent table image."""
    ng, rows, pm = _run("high_cardinality", 1_000_000, 500_000,
                        plan_kw={"num_groups": 200_000}, col_kw={"num_groups": 200_000})
    assert ng > 190_000
    assert pm["part_nparts"] > 0


def test_high_cardinality_ragged_chunks(cuda):
    ng, rows, pm = _run("high_cardinality", 1_300_001, 299_996,
                        plan_kw={"num_groups": 100_000}, col_kw={"num_groups": 100_000})
    assert pm["part_nparts"] > 0


def test_high_cardinality_global_table(cuda, monkeypatch):
    """The same through the global open-addressing table alone."""
    monkeypatch.setenv("PGSTROM_NO_PARTITION", "1")
    ng, rows, pm = _run("high_cardinality", 1_000_000, 500_000,
                        plan_kw={"num_groups": 200_000}, col_kw={"num_groups": 200_000})
    assert ng > 190_000 and pm["part_nparts"] == 0


def test_high_cardinality_partition_overflow(cuda):
    """Record areas sized for 50 K-row chunks receive 500 K-row chunks: what
    the partitions refuse goes to the global table, groups then live in both
    and PostgreSQL's final Agg (here: the checker) merges them."""
    ng, rows, pm = _run("high_cardinality", 1_000_000, 500_000,
                        plan_kw={"num_groups": 200_000}, col_kw={"num_groups": 200_000},
                        session_kw={"max_chunk_rows": 50_000})
    assert ng > 190_000 and pm["part_nparts"] > 0
    assert len(rows) > ng           # some groups came back as two partial rows


def test_high_cardinality_segment_mode(cuda, monkeypatch):
    """The deal pass with one record segment per CTA and its cursors in shared
    memory (PGSTROM_SEGMENTS=1; off by default: measured slower than the global
    cursors at 10 M groups): full chunks, ragged chunks and segments far too
    small for their chunk (the rest goes to the global table)."""
    monkeypatch.setenv("PGSTROM_SEGMENTS", "1")
    ng, rows, pm = _run("high_cardinality", 1_000_000, 500_000,
                        plan_kw={"num_groups": 200_000}, col_kw={"num_groups": 200_000})
    assert ng > 190_000 and pm["part_seg_cap"] > 0 and pm["part_seg_max"] >= 1
    _run("high_cardinality", 1_300_001, 299_996,
         plan_kw={"num_groups": 100_000}, col_kw={"num_groups": 100_000})
    ng, rows, pm = _run("high_cardinality", 1_000_000, 500_000,
                        plan_kw={"num_groups": 200_000}, col_kw={"num_groups": 200_000},
                        session_kw={"max_chunk_rows": 50_000})
    assert pm["part_seg_cap"] > 0 and len(rows) > ng


def test_high_cardinality_small_images(cuda, monkeypatch):
    """Images of 64 slots fill up (75% limit): the rest spills to the global table."""
    monkeypatch.setenv("PGSTROM_PART_SLOTS", "64")
    ng, rows, pm = _run("high_cardinality", 600_000, 300_000,
                        plan_kw={"num_groups": 150_000}, col_kw={"num_groups": 150_000})
    assert pm["part_slots"] == 64


def test_where_many_groups(cuda):
    """WHERE + GROUP BY with more groups than a CTA-local table takes: the
    qual's survivors are dealt into partitions."""
    ng, rows, pm = _run("where_agg", 1_500_000, 500_000,
                        plan_kw={"num_groups": 120_000}, col_kw={"num_groups": 120_000})
    assert pm["sh_nslots"] == 0 and pm["part_nparts"] > 0


@pytest.mark.parametrize("gather", ["1", "0"])
def test_where_gather_payload_variant(cuda, monkeypatch, gather):
    """GROUP BY under WHERE, both flavours of the scan: the ring carries only
    the qual's column and survivors gather key / v / w from HBM by row number
    (the default), or every column is staged (PGSTROM_GATHER_PAYLOAD=0).
    Same results."""
    monkeypatch.setenv("PGSTROM_GATHER_PAYLOAD", gather)
    ng, rows, pm = _run("where_agg", 2_000_000, 1_000_000)
    assert ng == 1000 and (pm["tile_rows"] > 4096) == (gather == "1")
    _run("where_agg", 700_002, 233_332, col_kw={"with_nulls": True})
    for pct in (1, 50, 100):
        _run("where_agg", 400_000, 400_000, plan_kw={"selectivity_pct": pct})
    for n in (4, 100, 2048, 8193):
        _run("where_agg", n, n)
    monkeypatch.setenv("PGSTROM_SH_SLOTS", "256")
    _run("where_agg", 600_000, 300_000)
    monkeypatch.setenv("PGSTROM_SH_SLOTS", "0")
    _run("where_agg", 600_000, 300_000)
