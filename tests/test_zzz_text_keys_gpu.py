"""-m gpu: GROUP BY text / character(n) columns through the executor half of
the C ABI, on column chunks and on heap-page chunks.  Keys of at most 7 bytes
travel by value ("kernel text"), longer ones through the session's key heap in
HBM (kern_textlib.cuh: pgs_keyheap_intern; the reference moves varlena keys as
toast offsets and fixes the pointers up for the host,
/root/reference/opencl_gpupreagg.h:326-366).  The checker is the oracle's
restatement of the partial aggregation (oracle/partial.py)."""
import random

import numpy as np
import pytest

from oracle import bench_oracle, partial, pg_typelib as T
from pg_strom_b200 import gpupreagg as gp
from pg_strom_b200 import pgplan as P
from test_typelib_device_code import GUCS
from test_typelib_gpu import find_node

pytestmark = pytest.mark.gpu


CATS = P.Table("cats", [("cat", "text"), ("code", "bpchar"), ("f", "int4"), ("v", "int8")],
               typmods={"code": 4 + 5})


def make_cats(n, seed):
    rng = random.Random(seed)
    cats = [bytes([97 + i]) * 3 for i in range(26)] + [b"", b"a", b"abcdefg", b"caf\xc3\xa9"]
    long_cats = [b"abcdefgh", b"category-with-a-long-name", b"x" * 300]
    codes = [c.ljust(5) for c in (b"ab", b"abcde", b"x", b"", b"a b")] + [b"caf\xc3\xa9 "]
    rows = []
    for _ in range(n):
        cat = rng.choice(long_cats) if rng.random() < 0.03 else rng.choice(cats)
        rows.append((None if rng.random() < 0.04 else cat,
                     None if rng.random() < 0.04 else rng.choice(codes),
                     rng.randrange(0, 100), rng.randrange(-10 ** 15, 10 ** 15)))
    return rows


@pytest.mark.parametrize("fmt", ["column", "row"])
@pytest.mark.parametrize("with_qual", [False, True])
@pytest.mark.parametrize("key_heap_mb", [64, 0])
def test_text_and_bpchar_group_keys(cuda, fmt, with_qual, key_heap_mb):
    """GROUP BY a text and a character(5) column: keys of at most 7 bytes are
    their own 8-byte key word ("kernel text"), longer ones are stored once in
    the session's key heap and grouped by their heap word - the device
    returns every group, the host turns the words back into varlenas
    (pgstrom_fixup_kernel_text_heap; the reference's varlena key move and
    pointer fix-up, opencl_gpupreagg.h:326-366).  With pg_strom.key_heap_size
    = 0 rows with a long key come back for the host instead (CpuReCheck) and
    PostgreSQL's final Agg - here the checker - merges both."""
    t = CATS
    rows = make_cats(8000, seed=31)
    tree = P.make_agg_plan(
        t, [(t.col("cat"), "cat"), (t.col("code"), "code"), (P.Agg("count", star=True), "count"),
            (P.Agg("sum", [t.col("f")]), "sum"), (P.Agg("min", [t.col("v")]), "min")],
        group_by=["cat", "code"], num_groups=200,
        where=[P.Op("<", t.col("f"), P.Const("int4", 50))] if with_qual else [])
    plan = gp.Plan(tree, gucs=dict(GUCS, **{"pg_strom.key_heap_size": key_heap_mb}))
    try:
        assert plan.num_gpupreagg == 1, plan.reject_reason
        desc = plan.describe()
        node = find_node(plan.tree())
        coltypes = [c for _, c in t.columns]
        chunk_rows = 3000
        chunks = []
        for lo in range(0, len(rows), chunk_rows):
            part = rows[lo:lo + chunk_rows]
            cols = []
            for c, typ in enumerate(coltypes):
                raw = [r[c] for r in part]
                if gp.PGTYPES[typ][0] > 0:
                    cols.append((np.array(raw, dtype=gp.PGTYPES[typ][3]), None))
                else:
                    cols.append(([None if v is None else T.varlena(v) for v in raw], None))
            chunks.append(gp.DataStore(coltypes, cols, nrows=len(part)) if fmt == "column"
                          else gp.HeapDataStore(coltypes, cols, nrows=len(part)))
        st = gp.GpuPreAggState(plan, chunks)
        try:
            device_rows = st.fetch_all()
            recheck = sorted(s * chunk_rows + r for s, r in st.recheck_rows())
        finally:
            st.end()
        for ds in chunks:
            ds.free()
    finally:
        plan.free()
        gp._capi.load().pgstrom_guc_set(b"pg_strom.key_heap_size", b"64")
    long_rows = [i for i, r in enumerate(rows)
                 if r[0] is not None and len(r[0]) > 7 and (not with_qual or r[2] < 50)]
    assert len(long_rows) > 50
    assert recheck == ([] if key_heap_mb else long_rows)
    # what gpupreagg_next_tuple_fallback produces for the flagged rows
    host, _ = partial.partial_rows(node, [rows[i] for i in recheck], len(t.columns))
    exp, _ = partial.partial_rows(node, rows, len(t.columns))
    got = bench_oracle.combine_device_rows(desc, list(device_rows) + [tuple(v) for v in host.values()])
    assert len(exp) > 100 and set(got) == set(exp)
    for key, erow in exp.items():
        assert list(got[key]) == list(erow), (key, got[key], erow)
    # character(5) keys come back padded; long keys only ever from the key heap
    longest = max((len(r[0]) for r in device_rows if r[0] is not None), default=0)
    assert longest == (300 if key_heap_mb else 7)
    for r in device_rows:
        assert r[1] is None or len(r[1].decode("utf-8")) == 5
    if key_heap_mb:
        # one partial row per group: no group was split between device and host
        assert len(device_rows) == len(exp)
