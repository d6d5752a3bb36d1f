"""CPU, world_size 2 over gloo: the host logic of the multi-GPU path.

Chunks of the regression table are dealt round-robin to the ranks
(pg_strom_b200/multigpu.py), every rank produces the partial rows of ITS
chunks (here with the oracle's restatement of the partial aggregation - there
is no GPU in this test), the root collects them and PostgreSQL's final Agg
(oracle) over the union must print exactly the golden result: partial
aggregation is associative, so how the table is cut over ranks must not show.
Also covers the id broadcast and the max-over-ranks helper.
"""
import json
import os
import socket
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STATEMENTS = [
    "select count(integer_x) from gpupreagg_test;",
    "select avg(bigint_x) from gpupreagg_test;",
    "select avg(float_x) from gpupreagg_test;",
    "select key,max(integer_x) from gpupreagg_test group by key order by key;",
    "select key,stddev(float_x) from gpupreagg_test group by key order by key;",
    "select key,avg(smlint_x) from gpupreagg_test group by key order by key;",
]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, chunk_rows, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import harness
    from oracle import partial
    from pg_strom_b200 import gpupreagg as gp
    from pg_strom_b200 import multigpu
    from pg_strom_b200 import pgplan as P

    # plumbing helpers
    blob = multigpu.broadcast_bytes(dist, bytes(range(128)) if rank == 0 else b"", 128)
    assert blob == bytes(range(128))
    assert multigpu.max_over_ranks(dist, 1.0 + rank) == float(world)
    # the CUDA IPC handles of the exchange areas travel like this (64 bytes per rank)
    handles = multigpu.allgather_bytes(dist, bytes([rank + 1]) * 64, 64)
    assert handles == [bytes([r + 1]) * 64 for r in range(world)]

    results = {}
    for sql in STATEMENTS:
        q = P.parse_regression_sql(sql)
        table, rows = harness.fixture_table(q["table"])
        tuples = harness.rows_as_tuples(table, rows)
        plan = gp.Plan(P.plan_regression_sql(sql, table), gucs=harness.GUCS)
        assert plan.num_gpupreagg == 1, plan.reject_reason
        desc = plan.describe()
        node = harness.find_gpreagg_node(plan.tree())
        mine = []
        shards = multigpu.shard_rows(len(tuples), chunk_rows, rank, world)
        for row0, n in shards:
            groups, order = partial.partial_rows(node, tuples[row0:row0 + n], len(table.columns))
            mine.extend(tuple(groups[k]) for k in order)
        gathered = [None] * world if rank == 0 else None
        dist.gather_object((shards, mine), gathered, dst=0)
        if rank == 0:
            covered = sorted(s for sh, _ in gathered for s in sh)
            assert covered == multigpu.shard_rows(len(tuples), chunk_rows, 0, 1)
            allrows = [r for _, part in gathered for r in part]
            out, types, _ = harness.final_aggregate(desc, allrows, q)
            results[sql] = out
        plan.free()
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(results, f)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_merge_like_one(tmp_path, lib):
    import torch.multiprocessing as mp
    out_path = str(tmp_path / "merged.json")
    mp.spawn(_worker, args=(2, _free_port(), 7001, out_path), nprocs=2, join=True)
    with open(out_path) as f:
        merged = json.load(f)
    golden = {}
    for name in ("nogrp_agg", "group_agg"):
        with open(os.path.join(HERE, "golden", name + ".json")) as f:
            for s in json.load(f):
                golden[" ".join(s["sql"].split())] = s["rows"]
    sys.path.insert(0, HERE)
    import harness
    for sql in STATEMENTS:
        key = next(k for k in golden if k.replace(" ", "") == sql.replace(" ", ""))
        got, exp = merged[sql], golden[key]
        assert len(got) == len(exp), sql
        for g, e in zip(got, exp):
            for a, b in zip(g, e):
                assert a == b or harness.cells_match(a, b, "float8"), (sql, g, e)


def test_choose_merge():
    """Which merge a session takes: one record or a small
    table goes to the root over NVLink peer memory, millions of groups are
    partitioned over the ranks."""
    from pg_strom_b200 import multigpu
    assert multigpu.choose_merge(False, 0, 0) == "peer"
    assert multigpu.choose_merge(True, 4096, 0) == "peer"
    assert multigpu.choose_merge(True, 65536, 0) == "peer"
    assert multigpu.choose_merge(True, 131072, 0) == "exchange"
    assert multigpu.choose_merge(True, 2048, 19532) == "exchange"
    # text keys of a session's key heap are session-local words: no merge, every
    # rank returns its own partial rows to the final Agg (gpupreagg.c:2169-2186)
    assert multigpu.choose_merge(True, 4096, 0, key_heap_nslots=65536) == "none"
    assert multigpu.choose_merge(True, 2048, 19532, key_heap_nslots=65536) == "none"


def test_deal_chunks():
    from pg_strom_b200 import multigpu
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            seen += multigpu.deal_chunks(37, r, world)
        assert sorted(seen) == list(range(37))
    assert multigpu.shard_rows(10, 4, 1, 2) == [(4, 4)]
    with pytest.raises(ValueError):
        multigpu.deal_chunks(4, 2, 2)
