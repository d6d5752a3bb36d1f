"""CPU: the heap-page leg of the CPU baseline (oracle/cpu_agg.c:
cpu_heap_form_pages / cpu_agg_run_heap - heapgettup_pagemode +
slot_deform_tuple + the same Agg loop as the columnar leg; SURVEY.md 8d asks
for per-tuple deform in the baseline).  Two independent implementations of
PostgreSQL's page / tuple layout meet here: the oracle's reader must get the
columnar leg's answers from the pages the oracle forms AND from the pages the
product's own builder forms (pgstrom_heap_form_pages, what the heap-page
kernels scan)."""
import numpy as np
import pytest

from oracle import cpu_agg
from pg_strom_b200 import gpupreagg as gp
from pg_strom_b200 import workloads as W


def _states(name, keys, states, ng):
    """{key: per-aggregate state}; the high word of the integer sum only means
    something for the 128-bit sum of avg(int8) (the combine step of the other
    kinds leaves carries of its unsigned addition in it, nobody reads them)."""
    aggs = cpu_agg.QUERIES[name]["aggs"]
    naggs = len(aggs)
    out = {}
    for g in range(ng):
        out[keys[g]] = [(s.n, s.isum_lo, s.isum_hi if aggs[j][0] == "avg_int8" else 0, s.fsum,
                         s.fsum2, s.imin, s.imax, s.fmin, s.fmax, s.has_value)
                        for j, s in enumerate(states[g * naggs:(g + 1) * naggs])]
    return out


@pytest.mark.parametrize("name,nrows,kw", [
    ("nogrp_agg", 300_001, {}),                         # 5 % NULLs in both columns
    ("where_agg", 200_000, {"with_nulls": True}),
    ("high_cardinality", 120_000, {}),
])
def test_heap_scan_equals_columnar_scan(lib, name, nrows, kw):
    cols = W.WORKLOADS[name]["columns"](0, nrows, **kw)
    _, keys, states, ng = cpu_agg.run(name, cols, nthreads=3, max_groups=1 << 18)
    want = _states(name, keys, states, ng)
    # pages formed by the oracle
    pages, npages = cpu_agg.form_heap_pages(name, cols)
    for nthreads in (1, 4):
        _, k2, s2, ng2 = cpu_agg.run_heap(name, pages, npages, nthreads=nthreads, max_groups=1 << 18)
        assert ng2 == ng and _states(name, k2, s2, ng2) == want
    # pages formed by the product's builder
    coltypes = [t for _, t in W.WORKLOADS[name]["table"].columns]
    ds = gp.HeapDataStore(coltypes, cols, nrows=nrows)
    try:
        _, k3, s3, ng3 = cpu_agg.run_heap(name, ds._pages, ds.npages, nthreads=2, max_groups=1 << 18)
        assert ng3 == ng and _states(name, k3, s3, ng3) == want
        # and the two builders agree on how many tuples fit a page
        assert abs(ds.npages - npages) <= 1 + npages // 100
    finally:
        ds.free()


def test_heap_pages_look_like_postgres_pages(lib):
    """Spot checks of the layout against the numbers PostgreSQL's headers
    give: 24-byte page header, 4-byte line pointers, t_hoff 24 without and 24
    (one bitmap byte) with NULLs for a two-column table, int4 + float8 tuple =
    24 + 4 + pad 4 + 8 = 40 bytes."""
    x = np.arange(10, dtype=np.int32)
    y = np.arange(10, dtype=np.float64) / 4
    nx = np.zeros(10, np.uint8)
    nx[3] = 1
    pages, npages = cpu_agg.form_heap_pages("nogrp_agg", [(x, nx), (y, None)])
    assert npages == 1
    page = bytes(pages[:8192])
    lower, upper = int.from_bytes(page[12:14], "little"), int.from_bytes(page[14:16], "little")
    assert lower == 24 + 4 * 10
    lps = [int.from_bytes(page[24 + 4 * i:28 + 4 * i], "little") for i in range(10)]
    assert all((lp >> 15) & 3 == 1 for lp in lps)
    assert (lps[0] >> 17) == 40 and (lps[0] & 0x7fff) == 8192 - 40
    # row 3: x is NULL -> HEAP_HASNULL, bitmap 0b10, only y stored at t_hoff
    off3, len3 = lps[3] & 0x7fff, lps[3] >> 17
    assert page[off3 + 20] & 1 and page[off3 + 23] == 0b10 and page[off3 + 22] == 24 and len3 == 32
    assert upper == min(lp & 0x7fff for lp in lps)


def test_key_partitioned_plan_equals_one_core(lib):
    """GROUP BY with very many groups on several cores: every worker owns the
    groups of one hash partition (no combine step).  Same groups, same states
    as one core (the float columns sit on a dyadic grid: sums are exact)."""
    cols = W.WORKLOADS["high_cardinality"]["columns"](0, 400_000)
    _, k1, s1, n1 = cpu_agg.run("high_cardinality", cols, nthreads=1, max_groups=1 << 19)
    want = _states("high_cardinality", k1, s1, n1)
    for nthreads in (2, 7):
        _, k2, s2, n2 = cpu_agg.run("high_cardinality", cols, nthreads=nthreads, max_groups=1 << 19)
        assert n2 == n1 and _states("high_cardinality", k2, s2, n2) == want
